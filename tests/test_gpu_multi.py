"""2-GPU checks of the data-parallel engine (needs >= 2 visible GPUs; skipped otherwise).

The peer-memory exchange (stores into every rank's comm buffer + epoch flags inside the GAE / loss /
Adam kernels, include/b200ppo.h) must give the same parameters as the NCCL all-reduce path: with two
ranks the rank-ordered sum a + b is the same float as NCCL's, so the comparison is bit-exact; and the
replicas must stay in sync (identical parameters on both ranks).  Also with clip_by_global_norm (the
norm is taken from the summed gradient inside the exchange launch) and with the gradient norm logged:
the norm is identical on every rank and must not be multiplied by the world size when the metrics are
summed over ranks.

The driver's round-end GPU box has one GPU (this file is skipped there); the log of a 2-GPU run of this
file is kept under profiles/ (r2_gputest_2gpu.log, r2b_gputest_2gpu.log).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, p2p, out_dir, clip=None, twohop=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), B200PPO_P2P="1" if p2p else "0",
                      B200PPO_P2P_2HOP="2" if twohop else "99")      # smallest world that uses the two-hop exchange
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    from nnx_ppo_b200 import Rngs
    from nnx_ppo_b200.algorithms import ppo
    from nnx_ppo_b200.envs import SyntheticEnv
    from nnx_ppo_b200.networks.factories import make_mlp_actor_critic
    env = SyntheticEnv(24, 4, 16, 2048)
    nets = make_mlp_actor_critic(24, 4, [64, 64], [128, 128], Rngs(0))
    from nnx_ppo_b200.algorithms.types import LoggingLevel
    ts = ppo.new_training_state(env, nets, 256, 17, gradient_clipping=clip)
    hyper = (256, 8, 0.95, 0.99, 0.2, True, False, 2, 4)
    lvl = LoggingLevel.LOSSES | (LoggingLevel.GRAD_NORM if clip is not None else LoggingLevel.NONE)
    metrics = None
    for _ in range(3):                       # iteration 0 eager, 1 captures the graph, 2 replays it
        ts, metrics = ppo.ppo_step(env, ts, *hyper, logging_level=lvl)
    eng = ppo._engine_for(env, ts, 256, 8, 0.95, 0.99, 0.2, True, 2, 4, 1.0)
    assert eng.p2p == bool(p2p)
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"r{rank}_{int(p2p)}.npz"), params=eng.net.arena.cpu().numpy(),
             mu=eng.opt.mu.cpu().numpy(), m=np.array([float(v) for v in metrics.values()], np.float64),
             gn_logged=eng.metrics_host[:, 3].numpy().copy(), gn_last=float(eng.grad.double().norm()))
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)                              # NCCL kernels captured in a graph: skip the slow teardown


@pytest.mark.timeout(600)
@pytest.mark.parametrize("clip,twohop", [(None, False), (0.05, False), (None, True), (0.05, True)])
def test_peer_exchange_matches_nccl(tmp_path, clip, twohop):
    """twohop: the reduce-scatter + all-gather form of the gradient exchange (default from 4 ranks up), forced at
    world size 2 - owner-rank sums are rank ordered, so the result must still be NCCL's a + b bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    res = {}
    base = 29611 + (2 if clip else 0) + (4 if twohop else 0)
    for p2p, port in ((True, base), (False, base + 1)):
        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=_worker, args=(r, 2, port, p2p, str(tmp_path), clip, twohop)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(240)
            assert p.exitcode == 0
        for r in range(2):
            res[(r, p2p)] = np.load(tmp_path / f"r{r}_{int(p2p)}.npz")
    for p2p in (True, False):                # replicas in sync
        np.testing.assert_array_equal(res[(0, p2p)]["params"], res[(1, p2p)]["params"])
    np.testing.assert_array_equal(res[(0, True)]["params"], res[(0, False)]["params"])
    np.testing.assert_array_equal(res[(0, True)]["mu"], res[(0, False)]["mu"])
    np.testing.assert_allclose(res[(0, True)]["m"], res[(0, False)]["m"], rtol=1e-6)
    assert np.all(np.isfinite(res[(0, True)]["params"]))
    if clip is not None:
        for p2p in (True, False):
            gn = res[(0, p2p)]["gn_logged"]
            # the logged norm of the last update is the norm of the summed gradient, not world_size times it
            assert abs(gn[-1] - res[(0, p2p)]["gn_last"]) < 1e-5 * max(1.0, gn[-1]), (p2p, gn[-1], res[(0, p2p)]["gn_last"])
            np.testing.assert_array_equal(gn, res[(1, p2p)]["gn_logged"])
            assert gn.min() > clip                      # the threshold really clips in this test
        np.testing.assert_allclose(res[(0, True)]["gn_logged"], res[(0, False)]["gn_logged"], rtol=1e-6)


def _recurrent_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), B200PPO_P2P="1")
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    from nnx_ppo_b200 import Rngs
    from nnx_ppo_b200.algorithms import ppo
    from nnx_ppo_b200.envs import SyntheticEnv
    from nnx_ppo_b200.networks.factories import make_recurrent_actor_critic
    from nnx_ppo_b200.networks.plan import compile_network
    O, A, B, T, E, M, P, H = 16, 4, 256, 12, 2, 2, 32, 64
    nets = make_recurrent_actor_critic(O, A, P, H, [48], Rngs(1), trainable_initial_state=True)
    env = SyntheticEnv(O, A, max_len=10, term_thresh16=3000)
    ts = ppo.new_training_state(env, nets, B, 17)
    net = compile_network(nets)
    p0 = net.arena.cpu().numpy().copy()
    ms = []
    for _ in range(3):                       # eager, graph capture + replay, replay
        ts, m = ppo.ppo_step(env, ts, B, T, 0.95, 0.99, 0.2, True, False, E, M)
        ms.append([m[k] for k in ("losses/actor/mean", "losses/critic/mean", "losses/regularization/mean")])
    eng = next(iter(net.engines.values()))
    assert eng.p2p and eng.world == world and eng.r_graph is not None
    torch.cuda.synchronize()
    c, h = net.get_carry(ts.network_states)
    np.savez(os.path.join(out_dir, f"rec{rank}.npz"), params=net.arena.cpu().numpy(), p0=p0, m=np.array(ms, np.float64),
             obs=ts.env_states.obs.cpu().numpy(), mean=net.normalizer.mean.numpy(), h=h.cpu().numpy())
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


@pytest.mark.timeout(600)
def test_recurrent_data_parallel_replicas_stay_in_sync(tmp_path):
    """Recurrent (LSTM, trainable initial carry) actor on 2 ranks: every rank steps its own envs, the advantage
    moments and the gradient (recurrent actor + critic) are exchanged inside the kernels, so parameters,
    normaliser statistics and the (global) loss metrics are identical on both ranks while the env shards differ."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_recurrent_worker, args=(r, 2, 29641, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    a, b = np.load(tmp_path / "rec0.npz"), np.load(tmp_path / "rec1.npz")
    np.testing.assert_array_equal(a["params"], b["params"])
    np.testing.assert_array_equal(a["mean"], b["mean"])
    np.testing.assert_allclose(a["m"], b["m"], rtol=1e-6)
    assert np.all(np.isfinite(a["params"])) and np.abs(a["params"] - a["p0"]).max() > 1e-5
    assert not np.array_equal(a["obs"], b["obs"]) and not np.array_equal(a["h"], b["h"])       # different env shards


def _distill_worker(rank, world, port, p2p, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), B200PPO_P2P="1" if p2p else "0")
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    from nnx_ppo_b200 import Rngs
    from nnx_ppo_b200.algorithms import distillation
    from nnx_ppo_b200.envs import SyntheticEnv
    from nnx_ppo_b200.networks.factories import make_mlp_actor_critic
    from nnx_ppo_b200.networks.plan import compile_network
    O, A, B, T, E, M = 24, 4, 256, 8, 2, 4
    env = SyntheticEnv(O, A, 16, 2048)
    student = make_mlp_actor_critic(O, A, [64, 64], [32], Rngs(1))
    teacher = make_mlp_actor_critic(O, A, [48], [16], Rngs(2))
    teacher.eval()
    ds = distillation.new_distillation_state(env, teacher, student, B, 17, learning_rate=3e-4)
    net = compile_network(student)
    ms = []
    for _ in range(3):                       # eager, graph capture + replay, replay
        ds, m = distillation.distillation_step(env, teacher, ds, B, T, E, M)
        ms.append([float(m["losses/distillation_nll/mean"]), float(m["losses/regularization/mean"])])
    eng = next(iter(net.engines.values()))
    assert eng.p2p == bool(p2p) and eng.world == world and eng.nll
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"d{rank}_{int(p2p)}.npz"), params=net.arena.cpu().numpy(), m=np.array(ms, np.float64),
             obs=ds.env_states.obs.cpu().numpy())
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


@pytest.mark.timeout(600)
def test_distillation_data_parallel(tmp_path):
    """Policy distillation on 2 ranks (NLL loss head: no advantage-moment exchange, the gradient exchange of the PPO
    path): replicas stay in sync, the peer-memory path equals the NCCL path bit for bit, the env shards differ."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    res = {}
    for p2p, port in ((True, 29671), (False, 29672)):
        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=_distill_worker, args=(r, 2, port, p2p, str(tmp_path))) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(240)
            assert p.exitcode == 0
        for r in range(2):
            res[(r, p2p)] = np.load(tmp_path / f"d{r}_{int(p2p)}.npz")
    for p2p in (True, False):
        np.testing.assert_array_equal(res[(0, p2p)]["params"], res[(1, p2p)]["params"])
        np.testing.assert_allclose(res[(0, p2p)]["m"], res[(1, p2p)]["m"], rtol=1e-6)
        assert not np.array_equal(res[(0, p2p)]["obs"], res[(1, p2p)]["obs"])
    np.testing.assert_array_equal(res[(0, True)]["params"], res[(0, False)]["params"])
    assert np.all(np.isfinite(res[(0, True)]["params"]))
