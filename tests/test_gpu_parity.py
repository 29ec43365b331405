"""GPU parity tests: every CUDA kernel of the hot path, called through the C ABI, against the CPU
oracle on identical seeded inputs.  Integer / mask / index results must be bit-exact; float
results must agree to the float32 tolerances written next to each assert.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from nnx_ppo_b200 import _lib, prng as hprng                                     # noqa: E402
from nnx_ppo_b200 import Rngs                                                      # noqa: E402
from nnx_ppo_b200.algorithms import ppo, rollout                                   # noqa: E402
from nnx_ppo_b200.algorithms.config import PPOConfig, TrainConfig, EvalConfig      # noqa: E402
from nnx_ppo_b200.algorithms.engine import PPOEngine, AdamOptimizer                # noqa: E402
from nnx_ppo_b200.envs import SyntheticEnv                                         # noqa: E402
from nnx_ppo_b200.networks.factories import make_mlp_actor_critic                  # noqa: E402
from nnx_ppo_b200.networks.normalizer import Normalizer                            # noqa: E402
from nnx_ppo_b200.networks.plan import compile_network                             # noqa: E402
from oracle import env as oenv, nets as onets, ppo as oppo, prng as oprng          # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def dev(cuda_device):
    from nnx_ppo_b200 import build
    build.build()
    _lib.load()
    return cuda_device


def u32(t):
    return t.cpu().numpy().view(np.uint32)


# ------------------------------------------------------------------------------------------
# K7 threefry
# ------------------------------------------------------------------------------------------
def test_random_bits_bit_exact(dev):
    lib = _lib.load()
    for seed, n in ((0, 1), (7, 1000), (123456789, 100003)):
        k = oprng.key(seed)
        out = torch.empty(n, dtype=torch.int32, device=dev)
        _lib.check(lib.b200ppo_random_bits(_lib.current_stream(), int(k[0]), int(k[1]), n, out.data_ptr()))
        assert np.array_equal(u32(out), oprng.random_bits(k, (n,)))
    assert lib.b200ppo_random_bits(_lib.current_stream(), 0, 0, 0, 0) == 0       # empty input


def test_random_normal_matches_oracle(dev):
    lib = _lib.load()
    k = oprng.fold_in(oprng.key(42), 5)
    n = 1 << 18
    out = torch.empty(n, device=dev)
    _lib.check(lib.b200ppo_random_normal(_lib.current_stream(), int(k[0]), int(k[1]), n, out.data_ptr()))
    ref = oprng.normal(k, (n,))
    got = out.cpu().numpy()
    # same bits, same polynomial; only log1p / sqrt rounding differs (float32 tolerance 2e-6 rel)
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) < 2e-6


# ------------------------------------------------------------------------------------------
# K6 permutation indices — bit exact, incl. n = 1, non powers of two and the global-scratch path
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,epochs", [(1, 2), (7, 3), (256, 4), (1024, 4), (4096, 4), (5000, 2),
                                       (20000, 2)])
def test_permutation_bit_exact(dev, n, epochs):
    lib = _lib.load()
    new_key = oprng.fold_in(oprng.key(17), n)
    kd = torch.from_numpy(new_key.view(np.int32).copy()).to(dev)
    out = torch.zeros(epochs, n, dtype=torch.int32, device=dev)
    scratch = torch.zeros(int(lib.b200ppo_permutation_scratch_bytes(n, epochs)) // 4 + 2,
                          dtype=torch.int32, device=dev)
    _lib.check(lib.b200ppo_permutation(_lib.current_stream(), kd.data_ptr(), n, epochs, out.data_ptr(),
                                       scratch.data_ptr()))
    got = out.cpu().numpy()
    for e in range(epochs):
        ref = oprng.permutation(oprng.fold_in(new_key, e), n)
        assert np.array_equal(got[e], ref)
        assert np.array_equal(np.sort(got[e]), np.arange(n))


def test_minibatch_indices_match_oracle(dev):
    lib = _lib.load()
    new_key = oprng.key(99)
    ref = oppo.minibatch_indices(new_key, 512, 4, 8)
    kd = torch.from_numpy(new_key.view(np.int32).copy()).to(dev)
    out = torch.zeros(4, 512, dtype=torch.int32, device=dev)
    scratch = torch.zeros(int(lib.b200ppo_permutation_scratch_bytes(512, 4)) // 4 + 2, dtype=torch.int32, device=dev)
    _lib.check(lib.b200ppo_permutation(_lib.current_stream(), kd.data_ptr(), 512, 4, out.data_ptr(), scratch.data_ptr()))
    assert np.array_equal(out.cpu().numpy().reshape(32, 64), ref)


# ------------------------------------------------------------------------------------------
# K2 GAE
# ------------------------------------------------------------------------------------------
def test_gae_reference_known_answer(dev):
    """The reference's own test_gae (ppo_test.py:229-264): max |diff| < 1e-6 vs its float64 loop."""
    g = np.load(os.path.join(GOLDEN, "gae_kat.npz"))
    T, B = 100, 512
    done = np.unpackbits(g["done_bits"])[: T * B].reshape(T, B).astype(bool)
    trunc = np.unpackbits(g["trunc_bits"])[: T * B].reshape(T, B).astype(bool)
    r = torch.from_numpy(g["rewards_f32"]).to(dev)
    v = torch.from_numpy(g["values_f32"]).to(dev)
    adv = ppo.gae(r, v[:-1], v[-1], torch.from_numpy(done).to(dev), torch.from_numpy(trunc).to(dev), 0.95, 0.8)
    got = adv.cpu().numpy()
    assert np.abs(got - g["adv_f64"]).max() < 1e-6
    ref32 = oppo.gae(g["rewards_f32"], g["values_f32"][:-1], g["values_f32"][-1], done, trunc, 0.95, 0.8)
    assert np.array_equal(got, ref32)          # same op order, no FMA contraction -> bit exact


def test_gae_edge_cases(dev):
    # all-done, all-truncated, T = 1, ragged B (not a multiple of the block size)
    rs = np.random.default_rng(0)
    for T, B in ((1, 1), (1, 130), (5, 33), (64, 1000)):
        r = rs.standard_normal((T, B)).astype(np.float32)
        v = rs.standard_normal((T + 1, B)).astype(np.float32)
        for pd in (0.0, 0.3, 1.0):
            done = rs.random((T, B)) < pd
            trunc = done & (rs.random((T, B)) < 0.5)
            ref = oppo.gae(r, v[:-1], v[-1], done, trunc, 0.9, 0.97)
            got = ppo.gae(torch.from_numpy(r).to(dev), torch.from_numpy(v[:-1]).to(dev),
                          torch.from_numpy(v[-1]).to(dev), torch.from_numpy(done).to(dev),
                          torch.from_numpy(trunc).to(dev), 0.9, 0.97).cpu().numpy()
            assert np.array_equal(got, ref)


# ------------------------------------------------------------------------------------------
# K5 Normalizer
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("O,T,B", [(8, 5, 16), (5, 30, 1024), (64, 32, 4096), (300, 3, 77)])
def test_normalizer_statistics(dev, O, T, B):
    """normalizer_test.py:42-65 / factories_test.py:76-119: moments within 1e-5; counter == T*B;
    a second merge equals the oracle's Chan merge."""
    g = np.random.default_rng(O)
    data = (g.standard_normal(O) * 3 + g.standard_normal((T, B, O)) * (1 + g.random(O))).astype(np.float32)
    nz = Normalizer(O)
    nz.update_statistics(torch.from_numpy(data).to(dev))
    onet = onets.make_mlp_actor_critic(O, 1, [4], [4], seed=0)
    onet.update_statistics(data)
    assert float(nz.counter.numpy()[0]) == T * B == float(onet.counter)
    mean, M2 = nz.mean.numpy(), nz.M2.numpy()
    d64 = data.astype(np.float64)
    # float64 truth at the reference tests' 1e-5 gate (NumPy's own float32 mean over the two
    # leading axes accumulates sequentially and is ~1e-4 off at 131072 rows, so it is not the truth)
    assert np.abs(mean - d64.mean(axis=(0, 1))).max() < 1e-5
    assert np.abs(np.sqrt(M2 / (T * B)) - d64.std(axis=(0, 1))).max() < 1e-5
    # the float32 oracle (sequential NumPy sums) agrees to its own accuracy
    assert np.allclose(mean, onet.mean, rtol=1e-4, atol=2e-4) and np.allclose(M2, onet.M2, rtol=5e-4)
    data2 = (1.5 + 0.5 * g.standard_normal((T, B, O))).astype(np.float32)
    nz.update_statistics(torch.from_numpy(data2).to(dev))
    onet.update_statistics(data2)
    assert float(nz.counter.numpy()[0]) == 2 * T * B
    both = np.concatenate([d64.reshape(-1, O), data2.astype(np.float64).reshape(-1, O)])
    assert np.abs(nz.mean.numpy() - both.mean(0)).max() < 1e-5
    assert np.abs(np.sqrt(nz.M2.numpy() / (2 * T * B)) - both.std(0)).max() < 2e-5
    assert np.allclose(nz.mean.numpy(), onet.mean, rtol=1e-4, atol=2e-4)
    assert np.allclose(nz.M2.numpy(), onet.M2, rtol=5e-4)
    nz.prepare()
    torch.cuda.synchronize()
    assert np.allclose(nz._std.cpu().numpy(), np.sqrt(np.maximum(nz.M2.numpy() / (2 * T * B), 1e-6)), rtol=1e-6)


def test_normalizer_default_std_is_ten(dev):
    """normalizer_test.py:33-40: before any update the forward divides by 10."""
    nets = make_mlp_actor_critic(4, 2, [8], [8], Rngs(0))
    net = compile_network(nets)
    net.normalizer.prepare()
    torch.cuda.synchronize()
    assert np.array_equal(net.normalizer._std.cpu().numpy(), np.full(4, 10.0, np.float32))


# ------------------------------------------------------------------------------------------
# K1 policy step (sample / replay / deterministic)
# ------------------------------------------------------------------------------------------
def _trunk_pair(obs_dim, act_dim, trunk, ah, ch, seed, act="relu"):
    from nnx_ppo_b200.networks.factories import make_shared_trunk_actor_critic
    from oracle import sharednet
    nets = make_shared_trunk_actor_critic(obs_dim, act_dim, trunk, ah, ch, Rngs(seed), activation=act)
    onet = sharednet.make_shared_trunk_actor_critic(obs_dim, act_dim, trunk, ah, ch, seed=seed, activation=act)
    return nets, onet


def _pair(obs_dim, act_dim, ah, ch, seed, act="relu", **kw):
    nets = make_mlp_actor_critic(obs_dim, act_dim, ah, ch, Rngs(seed), activation=act, **kw)
    onet = onets.make_mlp_actor_critic(obs_dim, act_dim, ah, ch, seed=seed, activation=act, **kw)
    return nets, onet


@pytest.mark.parametrize("cfg", [dict(O=64, A=8, ah=[64] * 4, ch=[256] * 2, B=300, act="relu"),
                                  dict(O=5, A=1, ah=[64] * 4, ch=[256] * 2, B=33, act="tanh"),
                                  dict(O=24, A=5, ah=[48, 40], ch=[72], B=64, act="swish")])
def test_policy_step_matches_oracle(dev, cfg):
    nets, onet = _pair(cfg["O"], cfg["A"], cfg["ah"], cfg["ch"], 3, cfg["act"])
    g = np.random.default_rng(1)
    # give the normalizer non-trivial statistics on both sides
    hist = (2 + 3 * g.standard_normal((4, 50, cfg["O"]))).astype(np.float32)
    nets.layers[0].update_statistics(torch.from_numpy(hist).to(dev))
    onet.update_statistics(hist)
    obs = (2 + 3 * g.standard_normal((cfg["B"], cfg["O"]))).astype(np.float32)
    state = nets.initialize_state(cfg["B"])
    out = nets(state, torch.from_numpy(obs).to(dev))
    ref = onets.policy_forward(onet, obs)
    po = out.output
    tol = dict(rtol=2e-5, atol=2e-5)       # float32: summation order + libm differences
    assert np.allclose(out.rollout_extras[1]["action"][-1].cpu().numpy(), ref["raw_action"], **tol)
    assert np.allclose(po.actions.cpu().numpy(), ref["action"], **tol)
    assert np.allclose(po.loglikelihoods.cpu().numpy(), ref["loglik"], rtol=1e-4, atol=1e-4)
    assert np.allclose(po.value_estimates.cpu().numpy(), ref["value"], **tol)
    assert np.allclose(out.regularization_loss.cpu().numpy(), ref["reg"], rtol=1e-4, atol=1e-5)
    assert nets.layers[-1].action.layers[-1].rng.count == onet.rng_count
    # replay (adapter_test.py:61-75): same weights + stored extras -> same actions and log-probs
    out2 = nets(state, torch.from_numpy(obs).to(dev), out.rollout_extras)
    assert torch.equal(out2.output.actions, po.actions)
    assert torch.allclose(out2.output.loglikelihoods, po.loglikelihoods, rtol=0, atol=0)
    onets.policy_forward(onet, obs, raw_action=ref["raw_action"])
    # deterministic (eval mode): action = tanh(mean); one RNG count per call
    nets.eval()
    c0 = nets.layers[1].action.layers[-1].rng.count
    out3 = nets(state, torch.from_numpy(obs).to(dev))
    nets.train()
    ref3 = onets.policy_forward(onet, obs, deterministic=True)
    assert np.allclose(out3.output.actions.cpu().numpy(), np.tanh(ref3["mu"]), **tol)
    assert nets.layers[1].action.layers[-1].rng.count == c0 + 1


# ------------------------------------------------------------------------------------------
# synthetic env reset + fused rollout
# ------------------------------------------------------------------------------------------
def _oracle_state(oe, onet, B, seed):
    return oppo.new_training_state(oe, onet, B, seed)


def test_split_rows_and_episode_wrapper_key_flow(dev):
    """`b200ppo_split_rows`: element j of jax.random.split(key_i) per row key, bit exact; EpisodeWrapper.reset
    (episode_wrapper.py:25) hands split(key)[0] to the wrapped env and draws the start counter from split(key)[1]
    - entirely on the device (the generic rollout resets all envs every step, rollout.py:39)."""
    import dataclasses as dc
    from nnx_ppo_b200.wrappers import EpisodeWrapper
    lib = _lib.load()
    keys = oprng.split(oprng.key(21), 777)
    kd = torch.from_numpy(keys.view(np.int32).copy()).to(dev)
    for j in (0, 1):
        out = torch.empty_like(kd)
        _lib.check(lib.b200ppo_split_rows(_lib.current_stream(), kd.data_ptr(), 777, j, out.data_ptr()))
        ref = np.stack([oprng.split(k)[j] for k in keys])
        assert np.array_equal(u32(out), ref)
    assert lib.b200ppo_split_rows(_lib.current_stream(), 0, 0, 0, 0) == 0

    @dc.dataclass
    class St:
        obs: torch.Tensor
        reward: torch.Tensor
        done: torch.Tensor
        info: dict

    class Inner:
        observation_size, action_size = 2, 1
        def reset(self, k):
            self.seen = k.clone()
            z = torch.zeros(k.shape[0], device=k.device)
            return St(torch.zeros(k.shape[0], 2, device=k.device), z, z, {})
        def step(self, st, a):
            return dc.replace(st, reward=st.reward + 1)

    inner = Inner()
    env = EpisodeWrapper(inner, max_len=40)
    st = env.reset(kd)
    assert np.array_equal(u32(inner.seen), np.stack([oprng.split(k)[0] for k in keys]))
    oe = oenv.SyntheticEnv(2, 1, max_len=40)
    assert np.array_equal(st.info["step_counter"].cpu().numpy(), oe.reset_fast(keys).step_counter)
    st2 = env.step(st, torch.zeros(777, 1, device=dev))
    assert torch.equal(st2.info["step_counter"], st.info["step_counter"] + 1)


def test_env_reset_matches_oracle(dev):
    env = SyntheticEnv(64, 8, max_len=64)
    oe = oenv.SyntheticEnv(64, 8, max_len=64)
    k = hprng.split(hprng.key(5))[0]
    st = env.reset_from_split(k, 1000, dev)
    keys = oprng.split(np.array(k, np.uint32), 1000)
    ref = oe.reset(keys)
    ref_fast = oe.reset_fast(keys)
    assert np.array_equal(ref.step_counter, ref_fast.step_counter)
    assert np.array_equal(st.step_counter.cpu().numpy(), ref.step_counter)            # bit exact
    assert np.array_equal(u32(st.term_state), ref.term_state)                          # bit exact
    assert np.allclose(st.obs.cpu().numpy(), ref.obs, rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("cfg", [dict(O=64, A=8, ah=[64] * 4, ch=[256] * 2, B=200, T=40, max_len=24, thr=1500),
                                  dict(O=5, A=1, ah=[32, 32], ch=[32], B=64, T=30, max_len=16, thr=3000),
                                  dict(O=12, A=3, ah=[48], ch=[16, 16], B=31, T=9, max_len=8, thr=0),
                                  dict(O=21, A=5, ah=[100, 37, 64], ch=[16], B=50, T=12, max_len=8, thr=1500, act="tanh"),
                                  # tiles too wide for the k-split scratch tile: the unsplit path
                                  dict(O=900, A=4, ah=[32], ch=[16], B=20, T=5, max_len=8, thr=1500)])
@pytest.mark.parametrize("engine", [1, 0, 2, 3, 4])
def test_fused_rollout_matches_oracle(dev, cfg, engine):
    """engine = 1: warp-level tensor-core tiles (mma.sync, 3xTF32; default), 0: fp32 FFMA tiles, 2: the batched
    per-step path (one tcgen05 tile GEMM per layer and step over all envs; the default beyond shared-memory-sized
    networks, forced here for every shape), 3 / 4: the one-tile-per-CTA / two-groups-per-CTA tensor-core kernel forced
    (mode 1 picks between them by the number of env tiles)."""
    prev = _lib.load().b200ppo_set_rollout_mode(engine)
    try:
        _fused_rollout(dev, cfg)
    finally:
        _lib.load().b200ppo_set_rollout_mode(prev)


def _fused_rollout(dev, cfg):
    nets, onet = _pair(cfg["O"], cfg["A"], cfg["ah"], cfg["ch"], 11, cfg.get("act", "relu"))
    env = SyntheticEnv(cfg["O"], cfg["A"], cfg["max_len"], cfg["thr"])
    oe = oenv.SyntheticEnv(cfg["O"], cfg["A"], cfg["max_len"], cfg["thr"])
    ts = ppo.new_training_state(env, nets, cfg["B"], 17)
    ots = _oracle_state(oe, onet, cfg["B"], 17)
    assert tuple(int(x) for x in ots.rng_key) == tuple(ts.rng_key)
    reset_key, _ = hprng.split(ts.rng_key)
    _, env2, tr = rollout.unroll_env(env, ts.env_states, nets, ts.network_states, cfg["T"], reset_key)
    oenv2, oro = oppo.unroll_env(oe, ots.env_state, onet, cfg["T"], np.array(reset_key, np.uint32))
    # reset masks and episode bookkeeping: bit exact
    assert np.array_equal(tr.done.cpu().numpy(), oro.done)
    assert np.array_equal(tr.truncated.cpu().numpy(), oro.truncated)
    assert np.array_equal(env2.step_counter.cpu().numpy(), oenv2.step_counter)
    assert np.array_equal(u32(env2.term_state), oenv2.term_state)
    assert oro.done.sum() > 0 and (cfg["thr"] == 0 or (oro.done & ~oro.truncated).sum() > 0)
    assert oro.truncated.sum() > 0
    # floats: float32 tolerance (errors accumulate over T steps of the env recurrence)
    tol = dict(rtol=2e-4, atol=2e-4)
    assert np.allclose(tr.obs.cpu().numpy(), oro.obs, **tol)
    assert np.allclose(tr.rollout_extras[1]["action"][-1].cpu().numpy(), oro.raw_action, **tol)
    assert np.allclose(tr.network_output.actions.cpu().numpy(), oro.action, **tol)
    assert np.allclose(tr.network_output.loglikelihoods.cpu().numpy(), oro.loglik, rtol=1e-3, atol=1e-3)
    assert np.allclose(tr.network_output.value_estimates.cpu().numpy(), oro.value, **tol)
    assert np.allclose(tr.rewards.cpu().numpy(), oro.reward, **tol)
    assert np.allclose(tr.next_obs.cpu().numpy(), oro.next_obs_last, **tol)
    assert np.allclose(env2.obs.cpu().numpy(), oenv2.obs, **tol)
    assert nets.layers[-1].action.layers[-1].rng.count == onet.rng_count


# ------------------------------------------------------------------------------------------
# K3/K4 one minibatch update against the oracle's loss + analytic gradients + Adam
# ------------------------------------------------------------------------------------------
def _dbg(eng, which, n):
    p = eng.lib.b200ppo_update_debug_ptr(eng.net.plan, eng.T, eng.mb, eng.ws.data_ptr(), which)
    off = (p - eng.ws.data_ptr()) // 4
    return eng.ws[off:off + n].cpu().numpy()


@pytest.mark.parametrize("cfg", [
    dict(O=64, A=8, ah=[64] * 4, ch=[256] * 2, B=256, T=32, M=2, act="relu", clip=None, wd=None),
    dict(O=5, A=1, ah=[64] * 4, ch=[256] * 2, B=96, T=30, M=2, act="tanh", clip=0.5, wd=None),
    dict(O=24, A=5, ah=[48, 40], ch=[72], B=70, T=7, M=2, act="swish", clip=None, wd=1e-3),
    dict(O=300, A=3, ah=[32], ch=[20, 20], B=40, T=5, M=1, act="relu", clip=None, wd=None),
    # actor wider than the critic + a slow (tanh) epilogue: the next chain's weight stages used to land on the
    # previous chain's epilogue buffer (update_tc.cuh tc_chain_setup)
    dict(O=24, A=3, ah=[48], ch=[32], B=96, T=9, M=2, act="tanh", clip=None, wd=None)])
@pytest.mark.parametrize("gemm", [0, 1, 2])
def test_single_update_matches_oracle(dev, cfg, gemm):
    """gemm = 0: fp32 FFMA kernels, 1: tcgen05 3xTF32 (default), 2: tcgen05 plain TF32 (loose)."""
    _lib.load().b200ppo_set_gemm_mode(gemm)
    try:
        _single_update(dev, cfg, 300.0 if gemm == 2 else 1.0)
    finally:
        _lib.load().b200ppo_set_gemm_mode(1)


def _single_update(dev, cfg, loose):
    O, A, B, T, M = cfg["O"], cfg["A"], cfg["B"], cfg["T"], cfg["M"]
    if "obs_sizes" in cfg:          # dict observations -> per-key encoders (Concat) -> trunk
        from nnx_ppo_b200.networks.factories import make_dict_actor_critic
        from oracle import dictnet
        nets = make_dict_actor_critic(cfg["obs_sizes"], A, cfg["enc"], cfg["ah"], cfg["ch"], Rngs(5), activation=cfg["act"])
        onet = dictnet.make_dict_actor_critic(cfg["obs_sizes"], A, cfg["enc"], cfg["ah"], cfg["ch"], seed=5,
                                              activation=cfg["act"])
    elif "trunk" in cfg:            # shared trunk feeding both PPOAdapter ports (tutorial 02_composition)
        nets, onet = _trunk_pair(O, A, cfg["trunk"], cfg["ah"], cfg["ch"], 5, cfg["act"])
    else:
        nets, onet = _pair(O, A, cfg["ah"], cfg["ch"], 5, cfg["act"])
    env = SyntheticEnv(O, A, max_len=12, term_thresh16=2500)
    oe = oenv.SyntheticEnv(O, A, max_len=12, term_thresh16=2500)
    ts = ppo.new_training_state(env, nets, B, 3, learning_rate=1e-3, gradient_clipping=cfg["clip"],
                                weight_decay=cfg["wd"])
    ots = _oracle_state(oe, onet, B, 3)
    net = compile_network(nets)
    # non-trivial normalizer statistics on both sides
    g = np.random.default_rng(2)
    hist = (g.standard_normal((3, 40, O))).astype(np.float32)
    net.normalizer.update_statistics(torch.from_numpy(hist).to(dev))
    onet.update_statistics(hist)
    eng = PPOEngine(net, env, ts.optimizer, B, T, 1, M, 0.95, 0.99, 0.2, True, 1.0, use_graph=False)
    reset_key, new_key = hprng.split(ts.rng_key)
    k = np.array([*reset_key, *new_key], np.uint32).view(np.int32)
    eng.iter_keys.copy_(torch.from_numpy(k.copy()))
    net.sync_counters_to_device()
    eng._enqueue_rollout(ts.env_states)
    lib = eng.lib
    _lib.check(lib.b200ppo_permutation(_lib.current_stream(), eng.iter_keys.data_ptr() + 8, B, 1,
                                       eng.inds.data_ptr(), eng.perm_scratch.data_ptr()))
    # oracle: same rollout, same indices
    _, oro = oppo.unroll_env(oe, ots.env_state, onet, T, np.array(reset_key, np.uint32))
    oinds = oppo.minibatch_indices(np.array(new_key, np.uint32), B, 1, M)
    assert np.array_equal(eng.inds.cpu().numpy().reshape(M, B // M), oinds)
    # make a good fraction of samples clip: perturb the stored old log-probs on both sides
    noise = (0.3 * g.standard_normal((T, B))).astype(np.float32)
    eng.loglik += torch.from_numpy(noise).to(dev)
    # use the GPU rollout as THE rollout on both sides so this test isolates the update kernels
    oro.obs[:] = eng.obs.cpu().numpy(); oro.raw_action[:] = eng.raw_action.cpu().numpy()
    oro.loglik[:] = eng.loglik.cpu().numpy(); oro.reward[:] = eng.reward.cpu().numpy()
    oro.next_obs_last[:] = eng.next_obs_last.cpu().numpy()
    mb = B // M
    base = onet.rng_count
    total, m, grads = oppo.ppo_loss_and_grads(onet, oro, oinds[0], base)
    _lib.check(lib.b200ppo_update(_lib.current_stream(), net.plan, eng.hp, eng.bufs[0], T, B, mb, 2 * T, 0,
                                  _lib.STAGE_FWD | _lib.STAGE_GAE | _lib.STAGE_LOSS | _lib.STAGE_BWD | _lib.STAGE_RED))
    torch.cuda.synchronize()
    R = T * mb
    tol = dict(rtol=1e-4 * loose, atol=1e-5 * loose)
    assert np.allclose(_dbg(eng, 1, R + mb)[:R].reshape(T, mb), m["values"], **tol)
    assert np.allclose(_dbg(eng, 1, R + mb)[R:], m["v_last"], **tol)
    assert np.allclose(_dbg(eng, 0, R).reshape(T, mb), m["adv"], rtol=1e-4 * loose, atol=1e-4 * loose)   # advantages
    sums = eng.adv_sums.cpu().numpy()
    assert abs(sums[0] / R - m["adv_mean"]) < 1e-5 * loose and abs(np.sqrt(sums[1] / R - (sums[0] / R) ** 2) - m["adv_std"]) < 1e-4 * loose
    met = eng.metrics[0].cpu().numpy()
    assert abs(met[0] - m["losses/actor"]) < 1e-5 * loose and abs(met[1] - m["losses/critic"]) < 1e-4 * loose * max(1, abs(m["losses/critic"]))
    assert abs(met[2] - m["losses/regularization"]) < 1e-5 * loose
    # ACTOR_EXTRA / CRITIC_EXTRA statistics of the same launch (ppo.py:514-527)
    assert abs(met[4] - m["losses/clipping_fraction"]) <= (2.0 / R) * loose + 1e-7      # a borderline ratio may flip
    r2 = 1.0 - 2.0 * met[1] / (max(met[6] - met[5] ** 2, 0.0) + 1e-8)
    assert abs(r2 - m["losses/critic_R^2"]) < 2e-3 * loose * max(1.0, abs(m["losses/critic_R^2"]))
    # [7], [8]: moments of the NORMALISED advantages (ppo.py:477-480 reassigns the name before it is logged at
    # :523); [9], [10]: the normalisation constants
    an = (m["adv"] - m["adv_mean"]) / (m["adv_std"] + 1e-8)
    assert abs(met[7] - an.mean()) < 2e-5 * loose and abs(met[8] - (an.astype(np.float64) ** 2).mean()) < 2e-4 * loose
    assert abs(met[9] - m["adv_mean"]) < 1e-5 * loose and abs(met[10] - (m["adv_std"] + 1e-8)) < 1e-4 * loose
    if loose > 1.0:
        # plain TF32 is not fp32 parity (clip decisions can flip): only require a sane gradient
        got_g = net.params_logical(eng.grad)
        assert np.isfinite(got_g).all()
        assert np.linalg.norm(got_g - grads) < 0.2 * np.linalg.norm(grads)
        return
    dy = _dbg(eng, 3, R * 2 * A).reshape(R, 2 * A)
    # a probability ratio within rounding of 1 +- clip_range may fall on the other side of the clip (its
    # sample's gradient is then zero on one side only): allow one such row per 4096 samples
    bad = (np.abs(dy - m["d_y"]) >= 1e-4 * loose * max(np.abs(m["d_y"]).max(), 1e-6) + 1e-9).any(axis=1)
    assert bad.sum() <= R // 4096, (int(bad.sum()), R)
    dv = _dbg(eng, 4, R)
    assert np.abs(dv - m["d_v"]).max() < 1e-4 * loose * np.abs(m["d_v"]).max() + 1e-10
    got_g = net.params_logical(eng.grad)
    gs = np.abs(grads).max()
    assert np.abs(got_g - grads).max() < 2e-4 * loose * gs, (np.abs(got_g - grads).max(), gs)
    # K4: optax update from the SAME gradient on both sides
    p_before = net.params_logical()
    ost = oppo.AdamState(np.zeros_like(p_before), np.zeros_like(p_before), 0)
    p_ref = oppo.adam_update(p_before, got_g, ost, lr=1e-3, gradient_clipping=cfg["clip"], weight_decay=cfg["wd"])
    _lib.check(lib.b200ppo_update(_lib.current_stream(), net.plan, eng.hp, eng.bufs[0], T, B, mb, 2 * T, 0,
                                  _lib.STAGE_ADAM))
    torch.cuda.synchronize()
    p_after = net.params_logical()
    assert np.abs(p_after - p_ref).max() < 2e-7, np.abs(p_after - p_ref).max()
    assert np.abs(net.params_logical(ts.optimizer.mu) - ost.mu).max() < 1e-7 * max(1.0, gs)
    if cfg["clip"] is not None:
        assert abs(eng.metrics[0, 3].item() - np.sqrt((got_g.astype(np.float64) ** 2).sum())) < 1e-4 * max(1, gs)


@pytest.mark.parametrize("cfg", [
    dict(O=64, A=8, ah=[64] * 4, ch=[256] * 2, B=256, T=32, M=2, act="relu"),
    dict(O=24, A=5, ah=[48, 40], ch=[72], B=70, T=7, M=2, act="swish"),
    dict(O=300, A=3, ah=[32], ch=[20, 20], B=40, T=5, M=1, act="tanh")])
def test_update_kernel_structure_switches_agree(dev, cfg):
    """The two structural choices of the update (b200ppo_set_update_paths) against their plain versions on one
    rollout: GAE + loss as ONE launch must be bit-identical to the two launches (same blocks, same partial-sum
    order); the weight-gradient kernel on MN-major TMA operands (hi = truncated tile) must agree with the
    transposing kernel (hi = rounded) to fp32-GEMM accuracy."""
    O, A, B, T, M = cfg["O"], cfg["A"], cfg["B"], cfg["T"], cfg["M"]
    nets, _ = _pair(O, A, cfg["ah"], cfg["ch"], 5, cfg["act"])
    env = SyntheticEnv(O, A, max_len=12, term_thresh16=2500)
    ts = ppo.new_training_state(env, nets, B, 3, learning_rate=1e-3)
    net = compile_network(nets)
    lib = _lib.load()
    mb, R = B // M, T * (B // M)
    out = {}
    prev = lib.b200ppo_set_update_paths(-1, -1)
    try:
        for name, (fuse, mn) in (("new", (1, 1)), ("plain", (0, 0))):
            lib.b200ppo_set_update_paths(fuse, mn)
            eng = PPOEngine(net, env, ts.optimizer, B, T, 1, M, 0.95, 0.99, 0.2, True, 1.0, use_graph=False)
            reset_key, new_key = hprng.split(ts.rng_key)
            eng.iter_keys.copy_(torch.from_numpy(np.array([*reset_key, *new_key], np.uint32).view(np.int32).copy()))
            net.sync_counters_to_device()
            state = type(ts.env_states)(ts.env_states.obs.clone(), ts.env_states.step_counter.clone(),
                                        ts.env_states.term_state.clone())
            eng._enqueue_rollout(state)
            _lib.check(lib.b200ppo_permutation(_lib.current_stream(), eng.iter_keys.data_ptr() + 8, B, 1,
                                               eng.inds.data_ptr(), eng.perm_scratch.data_ptr()))
            _lib.check(lib.b200ppo_update(_lib.current_stream(), net.plan, eng.hp, eng.bufs[0], T, B, mb, 2 * T, 0,
                                          _lib.STAGE_FWD | _lib.STAGE_GAE | _lib.STAGE_LOSS | _lib.STAGE_BWD | _lib.STAGE_RED))
            torch.cuda.synchronize()
            out[name] = dict(adv=_dbg(eng, 0, R).copy(), dy=_dbg(eng, 3, R * 2 * A).copy(), dv=_dbg(eng, 4, R).copy(),
                             met=eng.metrics[0].cpu().numpy().copy(), sums=eng.adv_sums.cpu().numpy().copy(),
                             grad=net.params_logical(eng.grad).copy())
    finally:
        lib.b200ppo_set_update_paths(prev & 1, (prev >> 1) & 1)
    a, b = out["new"], out["plain"]
    for k in ("adv", "dy", "dv", "sums"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["met"][:3], b["met"][:3]) and np.array_equal(a["met"][4:11], b["met"][4:11])
    gs = np.abs(b["grad"]).max()
    assert np.isfinite(a["grad"]).all() and np.abs(a["grad"] - b["grad"]).max() < 2e-5 * gs, (np.abs(a["grad"] - b["grad"]).max(), gs)


# ------------------------------------------------------------------------------------------
# full iterations through the public API (eager first iteration, captured graph afterwards)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [
    dict(O=64, A=8, ah=[64] * 4, ch=[256] * 2, B=256, T=32, E=2, M=4, iters=3),
    dict(O=5, A=1, ah=[64] * 4, ch=[256] * 2, B=128, T=30, E=4, M=4, iters=2)])
def test_ppo_step_matches_oracle_over_iterations(dev, cfg):
    _iterations(dev, cfg)


@pytest.mark.parametrize("gemm", [0, 1])
def test_dict_observation_network_matches_oracle(dev, gemm):
    """BASELINE configs[3] topology at test size: dict observations routed to per-key encoders
    (containers.py Concat), lowered to block-diagonal layers with masked structural zeros."""
    _lib.load().b200ppo_set_gemm_mode(gemm)
    try:
        _iterations(dev, dict(O=40, A=5, B=128, T=12, E=2, M=2, iters=2, ah=[32], ch=[48],
                              obs_sizes={"proprio": 16, "target": 24}, enc={"proprio": [24, 12], "target": [40, 20]}))
    finally:
        _lib.load().b200ppo_set_gemm_mode(1)
    # structural zeros stay exactly zero, and a dict of tensors is accepted at the API boundary
    import torch
    from nnx_ppo_b200.networks.factories import make_dict_actor_critic
    nets = make_dict_actor_critic({"a": 8, "b": 8}, 2, {"a": [8], "b": [8]}, [8], [8], Rngs(3))
    env = SyntheticEnv(16, 2, max_len=16, term_thresh16=700)
    ts = ppo.new_training_state(env, nets, 64, 5)
    for _ in range(2):
        ts, _m = ppo.ppo_step(env, ts, 64, 8, 0.95, 0.99, 0.2, True, False, 2, 2)
    net = compile_network(nets)
    assert net.param_mask is not None
    assert float(net.arena[net.param_mask == 0].abs().max()) == 0.0
    obs = torch.randn(10, 16, device=dev)
    o1 = nets(nets.initialize_state(10), {"a": obs[:, :8], "b": obs[:, 8:]})
    net.sampler.rng.count -= 2
    o2 = nets(nets.initialize_state(10), obs)
    assert torch.equal(o1.output.actions, o2.output.actions)


def test_ppo_step_matches_oracle_ffma_engine(dev):
    """Same check with the fp32 CUDA-core GEMM kernels (B200PPO_GEMM=ffma)."""
    _lib.load().b200ppo_set_gemm_mode(0)
    try:
        _iterations(dev, dict(O=64, A=8, ah=[64] * 4, ch=[256] * 2, B=256, T=32, E=2, M=4, iters=2))
    finally:
        _lib.load().b200ppo_set_gemm_mode(1)


def _iterations(dev, cfg):
    O, A, B, T, E, M = cfg["O"], cfg["A"], cfg["B"], cfg["T"], cfg["E"], cfg["M"]
    if "obs_sizes" in cfg:          # dict observations -> per-key encoders (Concat) -> trunk
        from nnx_ppo_b200.networks.factories import make_dict_actor_critic
        from oracle import dictnet
        nets = make_dict_actor_critic(cfg["obs_sizes"], A, cfg["enc"], cfg["ah"], cfg["ch"], Rngs(0))
        onet = dictnet.make_dict_actor_critic(cfg["obs_sizes"], A, cfg["enc"], cfg["ah"], cfg["ch"], seed=0)
    elif "trunk" in cfg:
        nets, onet = _trunk_pair(O, A, cfg["trunk"], cfg["ah"], cfg["ch"], 0)
    else:
        nets, onet = _pair(O, A, cfg["ah"], cfg["ch"], 0)
    ekw = dict(max_len=cfg.get("max_len", 48), term_thresh16=cfg.get("thresh", 700))
    env = SyntheticEnv(O, A, **ekw)
    oe = oenv.SyntheticEnv(O, A, **ekw)
    ts = ppo.new_training_state(env, nets, B, 17)
    ots = _oracle_state(oe, onet, B, 17)
    net = compile_network(nets)
    for it in range(cfg["iters"]):
        ts, metrics = ppo.ppo_step(env, ts, B, T, 0.95, 0.99, 0.2, True, False, E, M)
        tr = {}
        ots, om = oppo.ppo_step(oe, ots, B, T, n_epochs=E, n_minibatches=M, trace=tr)
        eng = next(iter(net.engines.values()))
        # bit exact: minibatch permutation indices, reset masks, episode bookkeeping, counters
        assert np.array_equal(eng.inds.cpu().numpy().reshape(E * M, B // M), tr["indices"])
        assert np.array_equal(eng.done.cpu().numpy().astype(bool), tr["rollout"].done)
        assert np.array_equal(eng.trunc.cpu().numpy().astype(bool), tr["rollout"].truncated)
        assert np.array_equal(ts.env_states.step_counter.cpu().numpy(), ots.env_state.step_counter)
        assert np.array_equal(u32(ts.env_states.term_state), ots.env_state.term_state)
        assert tuple(ts.rng_key) == tuple(int(x) for x in ots.rng_key)
        assert float(ts.steps_taken) == float(ots.steps_taken) == (it + 1) * T * B
        assert nets.layers[-1].action.layers[-1].rng.count == onet.rng_count
        cnt = u32(net.counters)
        assert cnt[2] == onet.rng_count and cnt[3] == (it + 1) * E * M == ots.opt.count
        assert float(net.normalizer.counter.numpy()[0]) == (it + 1) * T * B            # ppo_test.py:344-349
        # float32 tolerance
        for k in ("losses/actor/mean", "losses/critic/mean", "losses/regularization/mean"):
            assert abs(metrics[k] - om[k]) < 2e-4 * max(1.0, abs(om[k])), (k, metrics[k], om[k])
        p, po = net.params_logical(), onet.flat_params()
        # Adam moves each parameter by at most lr per update; rounding differences in tiny
        # gradients can flip individual steps, so compare against a few-steps budget
        assert np.abs(p - po).max() < cfg.get("pmax", 1e-4 * 4), np.abs(p - po).max()
        assert np.mean(np.abs(p - po)) < cfg.get("pmean", 2e-6), np.mean(np.abs(p - po))
        assert np.allclose(net.normalizer.mean.numpy(), onet.mean, rtol=1e-4, atol=1e-4)
        assert np.allclose(net.normalizer.M2.numpy(), onet.M2, rtol=2e-3)
        assert np.allclose(ts.env_states.obs.cpu().numpy(), ots.env_state.obs, rtol=1e-3, atol=1e-3)
    assert eng.graph is not None or cfg["iters"] < 2    # iterations >= 2 ran from the captured CUDA graph


# ------------------------------------------------------------------------------------------
# shared trunk: Sequential([Normalizer, trunk, PPOAdapter(action=head+sampler, value=head)])
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [
    dict(O=40, A=6, trunk=[64, 48], ah=[32], ch=[64, 16], B=128, T=16, M=2, act="relu", clip=None, wd=None),
    dict(O=17, A=2, trunk=[96], ah=[], ch=[24], B=70, T=7, M=2, act="tanh", clip=0.5, wd=1e-3)])
@pytest.mark.parametrize("gemm", [0, 1])
def test_shared_trunk_single_update_matches_oracle(dev, cfg, gemm):
    """The trunk's gradient is the SUM of the actor path's and the critic path's (the tied copies of the plan):
    same per-stage checks as test_single_update_matches_oracle, oracle/sharednet.py as the reference."""
    _lib.load().b200ppo_set_gemm_mode(gemm)
    try:
        _single_update(dev, cfg, 1.0)
    finally:
        _lib.load().b200ppo_set_gemm_mode(1)


def test_shared_trunk_iterations_match_oracle(dev):
    cfg = dict(O=40, A=6, trunk=[64, 48], ah=[32], ch=[64, 16], B=256, T=16, E=2, M=4, iters=3)
    _iterations(dev, cfg)
    # the two copies of the trunk (actor chain's, critic chain's) stayed bit-identical through 24 updates, in
    # the parameters and in both Adam moments; the user-visible Params are views of the live arena
    from nnx_ppo_b200.networks.plan import compile_network as cn
    # (the engine / compiled plan of the last _iterations call is reachable through its network only; rebuild)
    nets, _ = _trunk_pair(40, 6, [64, 48], [32], [64, 16], 0)
    env = SyntheticEnv(40, 6, max_len=48, term_thresh16=700)
    ts = ppo.new_training_state(env, nets, 256, 17, gradient_clipping=0.3)
    net = cn(nets)
    for _ in range(3):
        ts, m = ppo.ppo_step(env, ts, 256, 16, 0.95, 0.99, 0.2, True, False, 2, 4)
    dup, src = net._tie_pairs
    assert dup.numel() == 40 * 64 + 64 + 64 * 48 + 48
    for arr in (net.arena, ts.optimizer.mu, ts.optimizer.nu):
        assert torch.equal(arr[dup], arr[src])
    trunk0 = nets.layers[1].layers[0]
    assert torch.equal(trunk0.linear.kernel.value.reshape(-1), net.arena[net.plan.actor.w_off[0]:net.plan.actor.w_off[0] + 40 * 64])
    assert not np.array_equal(trunk0.linear.kernel.numpy(), _trunk_pair(40, 6, [64, 48], [32], [64, 16], 0)[0].layers[1].layers[0].linear.kernel.numpy())


def test_adam_refreshes_split_weight_planes(dev, monkeypatch):
    """The Adam kernel re-splits the updated weights into the tensor-core operand planes, so updates
    2..E*M skip the prep launch (B200PPO_STAGE_NO_PREP): must be bit-identical to always prepping."""
    res = []
    for fuse in ("1", "0"):
        monkeypatch.setenv("B200PPO_FUSE_PREP", fuse)
        nets, _ = _pair(40, 6, [64, 48], [96, 32], 0)
        env = SyntheticEnv(40, 6, max_len=48, term_thresh16=700)
        ts = ppo.new_training_state(env, nets, 128, 17)
        net = compile_network(nets)
        for _ in range(3):
            ts, _m = ppo.ppo_step(env, ts, 128, 8, 0.95, 0.99, 0.2, True, False, 2, 4)
        eng = next(iter(net.engines.values()))
        assert eng.fuse_prep == (fuse == "1")
        res.append(net.arena.cpu().numpy().copy())
    assert np.array_equal(res[0], res[1])


@pytest.mark.parametrize("shape", [dict(O=64, A=8, B=4096, T=32, E=4, M=8),      # BASELINE configs[1]
                                   dict(O=5, A=1, B=1024, T=30, E=4, M=4)])      # configs[0] (CartpoleBalance shapes)
def test_full_size_iteration_properties(dev, shape):
    """BASELINE configs at full size: the oracle needs minutes here, so check size-independent
    properties instead — two runs from the same seed are bit-identical (fixed-order reductions
    everywhere), every minibatch is a permutation slice, masks are consistent, counters / step
    bookkeeping are exact, everything stays finite."""
    B, T, E, M = shape["B"], shape["T"], shape["E"], shape["M"]
    runs = []
    for _ in range(2):
        nets, _o = _pair(shape["O"], shape["A"], [64] * 4, [256] * 2, 0)
        env = SyntheticEnv(shape["O"], shape["A"], max_len=64, term_thresh16=512)
        ts = ppo.new_training_state(env, nets, B, 17)
        net = compile_network(nets)
        for _it in range(3):
            ts, m = ppo.ppo_step(env, ts, B, T, 0.95, 0.99, 0.2, True, False, E, M)
        eng = next(iter(net.engines.values()))
        runs.append(dict(p=net.arena.cpu().numpy().copy(), inds=eng.inds.cpu().numpy().copy(),
                         done=eng.done.cpu().numpy().copy(), trunc=eng.trunc.cpu().numpy().copy(),
                         mean=net.normalizer.mean.numpy().copy(), m=dict(m), cnt=u32(net.counters).copy(),
                         steps=float(ts.steps_taken), ncount=float(net.normalizer.counter.numpy()[0])))
    a, b = runs
    assert np.array_equal(a["p"], b["p"]) and np.array_equal(a["mean"], b["mean"])      # deterministic
    assert np.array_equal(a["inds"], b["inds"]) and np.array_equal(a["done"], b["done"])
    for e in range(E):                                                                  # permutations
        assert np.array_equal(np.sort(a["inds"][e]), np.arange(B))
    assert not np.any(a["trunc"].astype(bool) & ~a["done"].astype(bool))                # truncated => done
    assert 0.005 < a["done"].mean() < 0.2
    assert a["steps"] == 3 * T * B == a["ncount"]
    assert a["cnt"][3] == 3 * E * M                                                     # Adam count
    assert a["cnt"][2] - u32(compile_network(_pair(shape["O"], shape["A"], [64] * 4, [256] * 2, 0)[0]).counters)[2] == \
        3 * (2 * T + E * M * 2 * (T + 1))                                               # sampler draws per iteration
    assert np.all(np.isfinite(a["p"])) and all(np.isfinite(v) for v in a["m"].values())
    assert eng.graph is not None


def test_checkpoint_resume_is_bit_identical(dev, tmp_path):
    """checkpointing.py:42-204 API: save after 2 iterations, keep training; restore into fresh
    templates and train the same number of iterations -> identical parameters, statistics, keys."""
    from nnx_ppo_b200.algorithms import checkpointing
    hyper = (64, 8, 0.95, 0.99, 0.2, True, False, 2, 2)

    def fresh():
        nets = make_mlp_actor_critic(12, 3, [16, 16], [16], Rngs(7))
        env = SyntheticEnv(12, 3, max_len=16, term_thresh16=2000)
        return env, nets, ppo.new_training_state(env, nets, 64, 3)

    env, nets, ts = fresh()
    for _ in range(2):
        ts, _m = ppo.ppo_step(env, ts, *hyper)
    ck = checkpointing.make_checkpoint_fn(str(tmp_path), config=TrainConfig())
    ck(ts, 1024)
    for _ in range(2):
        ts, _m = ppo.ppo_step(env, ts, *hyper)
    net = compile_network(nets)
    ref = (net.params_logical().copy(), net.normalizer.mean.numpy().copy(), tuple(ts.rng_key), float(ts.steps_taken))

    env2, nets2, ts2 = fresh()
    out = checkpointing.load_checkpoint(os.path.join(str(tmp_path), "step_0000001024"), nets2, ts2.optimizer)
    assert out["step"] == 1024 and isinstance(out["config"], TrainConfig)
    ts2 = out["training_state"]
    for _ in range(2):
        ts2, _m = ppo.ppo_step(env2, ts2, *hyper)
    net2 = compile_network(nets2)
    assert np.array_equal(net2.params_logical(), ref[0]) and np.array_equal(net2.normalizer.mean.numpy(), ref[1])
    assert tuple(ts2.rng_key) == ref[2] and float(ts2.steps_taken) == ref[3]


def test_logging_levels(dev):
    """metrics.py:17-121 keys for the levels this build produces (values against the rollout buffers)."""
    from nnx_ppo_b200.algorithms.types import LoggingLevel
    nets = make_mlp_actor_critic(12, 3, [16, 16], [16], Rngs(7))
    env = SyntheticEnv(12, 3, max_len=16, term_thresh16=2000)
    ts = ppo.new_training_state(env, nets, 64, 3, gradient_clipping=0.5)
    lvl = (LoggingLevel.LOSSES | LoggingLevel.TRAIN_ROLLOUT_STATS | LoggingLevel.ACTOR_EXTRA | LoggingLevel.WEIGHTS
           | LoggingLevel.GRAD_NORM | LoggingLevel.CRITIC_EXTRA)
    ts, m = ppo.ppo_step(env, ts, 64, 8, 0.95, 0.99, 0.2, True, False, 2, 2, logging_level=lvl)
    eng = next(iter(compile_network(nets).engines.values()))
    assert abs(m["rollout_batch/reward/mean"] - float(eng.reward.mean())) < 1e-6
    assert abs(m["rollout_batch/done_rate"] - float(eng.done.float().mean())) < 1e-7
    assert {"rollout_batch/action/std", "rollout_batch/truncation_rate", "loglikelihood/mean", "weights/std"} <= set(m)
    gn = eng.metrics[:, 3].cpu().numpy()                      # one scalar per update, logged through _log_metric
    assert gn.shape == (4,) and np.all(gn > 0)
    assert abs(m["grad_norm/mean"] - gn.mean()) < 1e-6 and abs(m["grad_norm/std"] - gn.std()) < 1e-6
    assert 0.0 <= m["losses/clipping_fraction/mean"] <= 1.0 and m["losses/critic_R^2/mean"] <= 1.0
    # the logged advantages are the normalised ones (ppo.py:477-480, 523): mean 0, std 1 per update
    assert abs(m["losses/advantages/mean"]) < 1e-4 and abs(m["losses/advantages/std"] - 1.0) < 1e-3
    ts, m = ppo.ppo_step(env, ts, 64, 8, 0.95, 0.99, 0.2, True, False, 2, 2, logging_level=lvl,
                         logging_percentiles=(0, 50, 100))
    assert m["rollout_batch/reward/p0"] <= m["rollout_batch/reward/p50"] <= m["rollout_batch/reward/p100"]
    assert "losses/actor/p50" in m and "weights/p100" in m and "grad_norm/p50" in m
    assert m["losses/advantages/p0"] < -0.5 and abs(m["losses/advantages/p50"]) < 0.5 < m["losses/advantages/p100"]
    # the sampler's metrics (sampling_layers.py:111) under Transition.metrics["net"] (rollout.py:31-34)
    ts, m = ppo.ppo_step(env, ts, 64, 8, 0.95, 0.99, 0.2, True, False, 2, 2,
                         logging_level=LoggingLevel.LOSSES | LoggingLevel.TRAINING_ENV_METRICS)
    base = "net/1/action/3"                                   # Sequential[1] = PPOAdapter, action Sequential[3] = sampler
    assert {f"{base}/mu/mean", f"{base}/mu/std", f"{base}/sigma/mean", f"{base}/sigma/std"} <= set(m)
    assert m[f"{base}/sigma/mean"] > 0.1                      # min_std
    out = nets(nets.initialize_state(64), eng.obs[3])         # the per-step call reports the same tree
    compile_network(nets).sampler.rng.count -= 2
    mu = out.metrics[1]["action"][3]["mu"]
    assert mu.shape == (64, 3) and torch.isfinite(mu).all()


def test_train_ppo_api(dev):
    """ppo_test.py:213-227 / 307-349 style: total steps, counter, finite metrics, log cadence."""
    env = SyntheticEnv(16, 4, max_len=32)
    nets = make_mlp_actor_critic(16, 4, [32, 32], [32, 32], Rngs(1))
    logged = []
    ckpts = []
    cfgt = TrainConfig(ppo=PPOConfig(n_envs=64, rollout_length=10, total_steps=64 * 10 * 5, n_minibatches=4),
                       eval=EvalConfig(enabled=False), checkpoint_every_steps=64 * 10 * 2)
    res = ppo.train_ppo(env, nets, cfgt, seed=3, log_fn=lambda m, s: logged.append((s, dict(m))),
                        checkpoint_fn=lambda st, s: ckpts.append(s))
    assert res.total_steps == 64 * 10 * 5 and res.total_iterations == 5
    assert [s for s, _ in logged] == [640 * i for i in range(1, 6)]
    assert ckpts == [0, 1280, 2560]
    assert all(np.isfinite(v) for _, m in logged for v in m.values())
    assert float(nets.layers[0].counter.numpy()[0]) == 3200
    p = compile_network(nets).params_logical()
    assert np.isfinite(p).all()
    with pytest.raises(ValueError):
        ppo.new_training_state(env, nets, 30, 0) and ppo.ppo_step(env, ppo.new_training_state(env, nets, 30, 0),
                                                                  30, 4, .95, .99, .2, True, False, 1, 4)


def test_eval_rollout_bookkeeping(dev):
    """rollout.py:97-148 on a batched torch env: sticky done, lifespan and reward accumulation."""
    import dataclasses

    @dataclasses.dataclass
    class S:
        obs: torch.Tensor
        reward: torch.Tensor
        done: torch.Tensor
        info: dict
        metrics: dict
        t: torch.Tensor

    class CountEnv:
        observation_size, action_size = 3, 2

        def reset(self, keys):
            B = keys.shape[0]
            life = (keys[:, 0].abs() % 5 + 2).float()
            return S(torch.zeros(B, 3, device=keys.device), torch.zeros(B, device=keys.device),
                     torch.zeros(B, device=keys.device), {"life": life}, {}, torch.zeros(B, device=keys.device))

        def step(self, s, a):
            t = s.t + 1
            return S(s.obs, torch.ones_like(t), (t >= s.info["life"]).float(), s.info, {}, t)

    env = CountEnv()
    nets = make_mlp_actor_critic(3, 2, [8], [8], Rngs(0))
    m = rollout.eval_rollout(env, nets, 32, 12, hprng.key(4), None)
    keys = rollout.split_keys_device(hprng.key(4), 32, dev)
    life = (keys[:, 0].abs() % 5 + 2).float()
    # lifespan counts the steps before the first done; reward also counts the terminal step
    assert abs(m["lifespan_mean"] - float((life - 1).mean())) < 1e-6
    assert abs(m["episode_reward/mean"] - float(life.mean())) < 1e-6


# ------------------------------------------------------------------------------------------
# fused evaluation rollout: one launch per episode batch
# ------------------------------------------------------------------------------------------
class _PerStepEnv:
    """Hides ``fused_rollout`` so eval_rollout takes the per-step path on the same env."""
    fused_rollout = False

    def __init__(self, env):
        self.reset, self.step = env.reset, env.step


@pytest.mark.parametrize("cfg", [dict(O=64, A=8, ah=[64] * 4, ch=[256] * 2, B=200, L=60, max_len=24, thr=1500),
                                  dict(O=5, A=1, ah=[32, 32], ch=[32], B=37, L=7, max_len=16, thr=3000),
                                  dict(O=12, A=3, ah=[48], ch=[16, 16], B=64, L=20, max_len=40, thr=0),
                                  dict(O=900, A=4, ah=[32], ch=[16], B=20, L=6, max_len=8, thr=1500)])
@pytest.mark.parametrize("deterministic", [True, False])
def test_fused_eval_rollout_matches_oracle(dev, cfg, deterministic):
    nets, onet = _pair(cfg["O"], cfg["A"], cfg["ah"], cfg["ch"], 5)
    g = np.random.default_rng(2)
    hist = (0.5 * g.standard_normal((3, 40, cfg["O"]))).astype(np.float32)
    nets.layers[0].update_statistics(torch.from_numpy(hist).to(dev))
    onet.update_statistics(hist)
    env = SyntheticEnv(cfg["O"], cfg["A"], cfg["max_len"], cfg["thr"])
    oe = oenv.SyntheticEnv(cfg["O"], cfg["A"], cfg["max_len"], cfg["thr"])
    key = hprng.key(9)
    net = compile_network(nets)
    sampler = nets.layers[1].action.layers[-1]
    if deterministic:
        nets.eval()
    c0 = sampler.rng.count
    st = env.reset(rollout.split_keys_device(key, cfg["B"], dev))
    cuml, life = rollout._eval_fused(env, net, st, cfg["B"], cfg["L"])
    assert sampler.rng.count == c0 + cfg["L"] * (1 if deterministic else 2)
    onet.rng_count = c0
    ocuml, olife = oppo.eval_rollout(oe, onet, cfg["B"], cfg["L"], np.array(key, np.uint32), deterministic)
    assert onet.rng_count == sampler.rng.count
    assert np.array_equal(life.cpu().numpy(), olife)                       # integer-valued: exact
    assert np.allclose(cuml.cpu().numpy(), ocuml, rtol=2e-4, atol=2e-4)
    assert 0 < olife.min() + 1 and olife.max() <= cfg["L"]
    # the public entry point: fused and per-step paths report the same metrics and RNG counts
    sampler.rng.count = c0
    m_fused = rollout.eval_rollout(env, nets, cfg["B"], cfg["L"], key, (0, 50, 100))
    c_fused = sampler.rng.count
    sampler.rng.count = c0
    m_step = rollout.eval_rollout(_PerStepEnv(env), nets, cfg["B"], cfg["L"], key, (0, 50, 100))
    assert sampler.rng.count == c_fused
    nets.train()
    assert set(m_fused) == set(m_step)
    for k in m_fused:
        assert abs(m_fused[k] - m_step[k]) <= 2e-4 * max(1.0, abs(m_step[k])), k
    assert abs(m_fused["lifespan_mean"] - float(olife.mean())) < 1e-5


def test_predicted_value_metric(dev):
    """losses/predicted_value (metrics.py:62-68) = the critic on the rollout's observations with the
    ROLLOUT-time parameters.  The pass is switched on when CRITIC_EXTRA is first asked for, which
    re-captures the iteration graph."""
    from nnx_ppo_b200.algorithms.types import LoggingLevel
    nets = make_mlp_actor_critic(12, 3, [16, 16], [16], Rngs(7))
    env = SyntheticEnv(12, 3, max_len=16, term_thresh16=2000)
    ts = ppo.new_training_state(env, nets, 64, 3)
    net = compile_network(nets)
    args = (64, 8, 0.95, 0.99, 0.2, True, False, 2, 2)
    for _ in range(2):                                   # eager iteration, then the captured graph
        ts, m = ppo.ppo_step(env, ts, *args)
        assert "losses/predicted_value/mean" not in m
    eng = next(iter(net.engines.values()))
    assert eng.value is None
    lvl = LoggingLevel.LOSSES | LoggingLevel.CRITIC_EXTRA
    for _ in range(2):                                   # re-capture, then replay with the pass inside
        p0, mean0 = net.arena.clone(), net.normalizer.mean._dev.clone()
        ts, m = ppo.ppo_step(env, ts, *args, logging_level=lvl)
        p1, mean1 = net.arena.clone(), net.normalizer.mean._dev.clone()
        assert not torch.equal(p0, p1)
        net.arena.copy_(p0); net.normalizer.mean._dev.copy_(mean0)
        v = rollout.policy_values(net, eng.obs.reshape(-1, 12)).reshape(8, 64)
        net.arena.copy_(p1); net.normalizer.mean._dev.copy_(mean1)
        assert torch.equal(v, eng.value)
        assert abs(m["losses/predicted_value/mean"] - float(v.mean())) < 1e-6
        assert abs(m["losses/predicted_value/std"] - float(v.std(unbiased=False))) < 1e-6


def test_per_step_env_path_matches_fused_path(dev):
    """ppo_step on an env stepped from Python (one policy-step launch per time step, rollout.py:11-45)
    against the fused rollout on the same env definition: same masks / counters, same losses and
    parameters to float32 accuracy; env metrics are logged under env/* (metrics.py:38-40)."""
    from nnx_ppo_b200.algorithms.types import LoggingLevel

    class MetricEnv(_PerStepEnv):
        def __init__(self, env):
            super().__init__(env)
            inner = env.step

            def step(s, a):
                n = inner(s, a)
                n.metrics = {"speed": 2.0 * n.reward, "group": {"alive": 1.0 - n.done}}
                return n
            self.step = step

    kw = dict(max_len=16, term_thresh16=2000)
    env_f, env_p = SyntheticEnv(12, 3, **kw), MetricEnv(SyntheticEnv(12, 3, **kw))
    nets_f = make_mlp_actor_critic(12, 3, [16, 16], [16], Rngs(7))
    nets_p = make_mlp_actor_critic(12, 3, [16, 16], [16], Rngs(7))
    ts_f = ppo.new_training_state(env_f, nets_f, 64, 3, gradient_clipping=0.5)
    ts_p = ppo.new_training_state(env_p, nets_p, 64, 3, gradient_clipping=0.5)
    lvl = LoggingLevel.ALL
    args = (64, 8, 0.95, 0.99, 0.2, True, False, 2, 2)
    for it in range(3):
        ts_f, m_f = ppo.ppo_step(env_f, ts_f, *args, logging_level=lvl)
        ts_p, m_p = ppo.ppo_step(env_p, ts_p, *args, logging_level=lvl)
        assert tuple(ts_f.rng_key) == tuple(ts_p.rng_key) and ts_f.steps_taken == ts_p.steps_taken
        assert torch.equal(ts_f.env_states.step_counter, ts_p.env_states.step_counter)       # bit exact
        assert torch.equal(ts_f.env_states.term_state, ts_p.env_states.term_state)
        assert m_f["rollout_batch/done_rate"] == m_p["rollout_batch/done_rate"]
        assert m_f["rollout_batch/truncation_rate"] == m_p["rollout_batch/truncation_rate"]
        assert set(m_f) <= set(m_p)
        assert set(m_p) - set(m_f) == {"env/speed/mean", "env/speed/std", "env/group/alive/mean",
                                       "env/group/alive/std"}
        for k in m_f:
            a, b = np.asarray(m_f[k], np.float64), np.asarray(m_p[k], np.float64)
            assert np.allclose(a, b, rtol=2e-3, atol=2e-4), (it, k, a, b)
        assert abs(m_p["env/speed/mean"] - 2.0 * m_p["rollout_batch/reward/mean"]) < 1e-6
        assert abs(m_p["env/group/alive/mean"] - (1.0 - m_p["rollout_batch/done_rate"])) < 1e-6
    pf, pp = compile_network(nets_f).params_logical(), compile_network(nets_p).params_logical()
    assert np.abs(pf - pp).mean() < 5e-6 and np.abs(pf - pp).max() < 4e-4
    assert compile_network(nets_f).rng_count == compile_network(nets_p).rng_count


def test_grad_norm_metric_without_clipping(dev):
    """LoggingLevel.GRAD_NORM with gradient_clipping=None (ppo.py:313-315): the norm is reported per
    update and the update itself is bit-identical to a run that does not log it."""
    from nnx_ppo_b200.algorithms.types import LoggingLevel
    env = SyntheticEnv(12, 3, max_len=16, term_thresh16=2000)
    args = (64, 8, 0.95, 0.99, 0.2, True, False, 2, 2)
    runs = []
    for lvl in (LoggingLevel.LOSSES, LoggingLevel.LOSSES | LoggingLevel.GRAD_NORM):
        nets = make_mlp_actor_critic(12, 3, [16, 16], [16], Rngs(7))
        ts = ppo.new_training_state(env, nets, 64, 3)
        for _ in range(3):
            ts, m = ppo.ppo_step(env, ts, *args, logging_level=lvl)
        net = compile_network(nets)
        eng = next(iter(net.engines.values()))
        runs.append((net.arena.clone(), m, float(eng.grad.double().norm())))
    assert torch.equal(runs[0][0], runs[1][0])
    assert "grad_norm/mean" not in runs[0][1]
    assert runs[1][1]["grad_norm/mean"] > 0 and np.isfinite(runs[1][1]["grad_norm/std"])
    gn = eng.metrics[:, 3].cpu().numpy()
    assert gn.shape == (4,) and np.all(gn > 0) and np.all(np.isfinite(gn))
    assert abs(gn[-1] - runs[1][2]) < 1e-5 * max(1.0, runs[1][2])
