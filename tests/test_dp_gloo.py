"""World-size-2 gloo test (CPU) of the data-parallel scheme of DESIGN.md §5.

Each rank computes its shard with the oracle; the collectives go through the same helpers the
engine uses (nnx_ppo_b200/parallel.py).  Rank 0 then checks that the all-reduced gradient equals
the gradient of ONE minibatch made of both ranks' envs, computed independently with torch.autograd
in float64 — i.e. W ranks are equivalent to one big minibatch (global advantage normalisation,
global means)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nnx_ppo_b200 import parallel
    from oracle import env as oenv, nets as onets, ppo as oppo, prng as oprng
    from test_oracle_ppo import _torch_loss

    assert parallel.dist_info() == (world, rank)
    O, A, B, T, mb = 10, 3, 24, 6, 12
    env = oenv.SyntheticEnv(O, A, max_len=8, term_thresh16=3000)
    net = onets.make_mlp_actor_critic(O, A, [16, 16], [24], seed=4, activation="tanh")
    g = np.random.default_rng(0)
    net.update_statistics(g.standard_normal((3, 20, O)).astype(np.float32))   # same stats on all ranks
    key, training_key = parallel.rank_keys(11, rank)
    if rank == 0:
        ref = oppo.new_training_state(env, net, B, 11)          # rank 0 == the single-device keys
        assert tuple(int(x) for x in ref.rng_key) == training_key
    es = env.reset_fast(oprng.split(np.array(key, np.uint32), B))
    reset_key, new_key = oprng.split(np.array(training_key, np.uint32))
    _, ro = oppo.unroll_env(env, es, net, T, reset_key)
    ro.loglik += (0.3 * np.random.default_rng(rank).standard_normal(ro.loglik.shape)).astype(np.float32)
    inds = oppo.minibatch_indices(new_key, B, 1, B // mb)[0]
    base = net.rng_count
    # pass 1: local advantage moment sums -> all-reduce (what upd_gae_kernel + NCCL do)
    _, m, _ = oppo.ppo_loss_and_grads(net, ro, inds, base, want_grads=False)
    adv = m["adv"].astype(np.float64)
    sums = torch.tensor([adv.sum(), (adv * adv).sum()], dtype=torch.float64)
    parallel.all_reduce_sum(sums)
    n_global = parallel.global_sample_count(T * mb, world)
    mean = sums[0].item() / n_global
    std = np.sqrt(max(sums[1].item() / n_global - mean * mean, 0.0))
    # pass 2: local gradient of the globally-normalised loss -> all-reduce(SUM)
    total, m2, grads = oppo.ppo_loss_and_grads(net, ro, inds, base, n_global=n_global,
                                               adv_stats=(mean, std))
    gt = torch.from_numpy(grads.copy())
    parallel.all_reduce_sum(gt)
    lt = torch.tensor([float(total)], dtype=torch.float64)
    parallel.all_reduce_sum(lt)
    # normalizer: per-rank batch statistics gathered in rank order
    flat = ro.obs.reshape(-1, O).astype(np.float64)
    bs = torch.from_numpy(np.concatenate([flat.mean(0), ((flat - flat.mean(0)) ** 2).sum(0)]))
    allbs = torch.zeros(world * 2 * O, dtype=torch.float64)
    parallel.all_gather_into(allbs, bs)
    payload = dict(obs=ro.obs[:, inds], raw=ro.raw_action[:, inds], ll=ro.loglik[:, inds],
                   rew=ro.reward[:, inds], done=ro.done[:, inds], trunc=ro.truncated[:, inds],
                   nol=ro.next_obs_last[inds], eps2=m2["eps2"], allobs=ro.obs)
    gathered = [None] * world
    dist.all_gather_object(gathered, payload)
    if rank == 0:
        cat = lambda k, ax: np.concatenate([p[k] for p in gathered], axis=ax)
        union = oppo.Rollout(cat("obs", 1), cat("raw", 1), np.tanh(cat("raw", 1)), cat("ll", 1),
                             np.zeros_like(cat("ll", 1)), cat("rew", 1), cat("done", 1), cat("trunc", 1),
                             cat("nol", 0))
        ref_total, ref_grads = _torch_loss(net, union, np.arange(world * mb), cat("eps2", 1), act="tanh")
        scale = np.abs(ref_grads).max()
        assert abs(lt.item() - ref_total) < 2e-5, (lt.item(), ref_total)
        assert np.abs(gt.numpy() - ref_grads).max() < 3e-5 * max(scale, 1.0)
        # Chan merge in rank order == moments of the union
        allobs = np.concatenate([p["allobs"].reshape(-1, O) for p in gathered]).astype(np.float64)
        st = allbs.numpy().reshape(world, 2, O)
        n = float(T * B)
        bm, bM2, bn = st[0, 0].copy(), st[0, 1].copy(), n
        for r in range(1, world):
            d = st[r, 0] - bm
            tot = bn + n
            bm = bm + d * n / tot
            bM2 = bM2 + st[r, 1] + d * d * bn * n / tot
            bn = tot
        assert np.allclose(bm, allobs.mean(0), atol=1e-12)
        assert np.allclose(bM2, ((allobs - allobs.mean(0)) ** 2).sum(0), rtol=1e-10)
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_equals_one_big_minibatch(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_rank_keys_rank0_matches_reference_derivation():
    from nnx_ppo_b200 import parallel, prng
    k, tk = parallel.rank_keys(17, 0)
    k0, tk0 = prng.split(prng.key(17))
    assert (k, tk) == (k0, tk0)
    k1, tk1 = parallel.rank_keys(17, 1)
    assert (k1, tk1) == (prng.fold_in(k0, 1), prng.fold_in(tk0, 1)) and tk1 != tk0
