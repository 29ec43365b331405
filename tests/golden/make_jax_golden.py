#!/usr/bin/env python
"""Pin the oracle against the REAL reference: run this once on any machine where ``jax``, ``flax``,
``optax`` and the reference checkout are importable (not this image: SURVEY.md F7), commit the
``jax_golden.npz`` it writes next to it, and ``tests/test_jax_golden.py`` turns every "parity
unpinned" item of DESIGN.md section 4 into a known-answer test of ``oracle/`` (and, through the GPU
parity tests, of the kernels):

    PYTHONPATH=/path/to/nnx-ppo python tests/golden/make_jax_golden.py [--out tests/golden/jax_golden.npz]

What is dumped (all float32 / int32 / uint32 NumPy arrays, a few KB in total):
  prng/*        jax.random.{split, fold_in, bits, normal, randint, permutation} on fixed keys
  rngs/*        flax.nnx.Rngs stream: keys of the first draws, count after make_mlp_actor_critic
  init/*        the parameters make_mlp_actor_critic builds from Rngs(0) (nnx_ppo/networks/factories.py:72-146)
  sampler/*     NormalTanhSampler outputs on fixed inputs (sampling_layers.py:82-147), incl. stream counts
  norm/*        Normalizer after two update_statistics calls (normalizer.py:98-136)
  gae/*         ppo.gae on the inputs of the reference's own test (ppo_test.py:229-264)
  adam/*        three optax steps of chain(clip_by_global_norm, adam) and of adamw on fixed gradients
  lstm/*        one nnx.OptimizedLSTMCell step with known kernels (gate order, bias placement)
  step/*        ONE nnx_ppo.algorithms.ppo.ppo_step on the synthetic env of SURVEY.md section 8(d)
                restated in JAX below: permutation indices, masks, losses, parameters after the step
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(HERE, "jax_golden.npz"))
    args = ap.parse_args()
    os.environ.setdefault("JAX_PLATFORMS", "cpu")
    import jax
    import jax.numpy as jp
    import optax
    from flax import nnx
    from nnx_ppo.algorithms import ppo
    from nnx_ppo.algorithms.types import LoggingLevel
    from nnx_ppo.networks import factories
    from nnx_ppo.networks.normalizer import Normalizer
    from nnx_ppo.networks.sampling_layers import NormalTanhSampler

    out: dict[str, np.ndarray] = {}

    def put(name, x):
        a = np.asarray(x)
        if a.dtype == np.float64:
            a = a.astype(np.float32)
        out[name] = a

    def kd(k):
        return np.asarray(jax.random.key_data(k), np.uint32)

    # ---------------- prng
    k = jax.random.key(17)
    put("prng/key17", kd(k))
    put("prng/split3", kd(jax.random.split(k, 3)))
    put("prng/split_2x3", kd(jax.random.split(k, (2, 3))))
    put("prng/fold_in_5", kd(jax.random.fold_in(k, 5)))
    put("prng/bits_7", jax.random.bits(k, (7,), jp.uint32))
    put("prng/normal_4x3", jax.random.normal(k, (4, 3)))
    put("prng/normal_1000_moments", np.array([float(jax.random.normal(k, (1000,)).mean()),
                                              float(jax.random.normal(k, (1000,)).std())], np.float32))
    put("prng/randint_32", jax.random.randint(k, (32,), 0, 32))
    put("prng/randint_scalar_500", jax.random.randint(k, (), 0, 500))
    for n in (1, 7, 256, 1024, 4096, 5000):
        put(f"prng/permutation_{n}", jax.random.permutation(jax.random.fold_in(k, n), n))

    # ---------------- nnx.Rngs stream + factory init
    r = nnx.Rngs(0)
    put("rngs/first_keys", np.stack([kd(r()) for _ in range(3)]))
    rngs = nnx.Rngs(0)
    nets = factories.make_mlp_actor_critic(6, 2, [8, 8], [8], rngs)
    put("rngs/count_after_factory", np.asarray(rngs.default.count.value if hasattr(rngs, "default") else -1))
    leaves = jax.tree.leaves(nnx.state(nets, nnx.Param))
    put("init/flat_params", np.concatenate([np.asarray(p).ravel() for p in leaves]))
    put("init/leaf_shapes", np.array([list(np.asarray(p).shape) + [0] * (2 - np.asarray(p).ndim) for p in leaves], np.int32))

    # ---------------- sampler
    s_rngs = nnx.Rngs(3)
    sampler = NormalTanhSampler(s_rngs, entropy_weight=1e-2, min_std=1e-1, std_scale=1.0)
    x = jp.asarray(np.linspace(-2.0, 2.0, 5 * 6, dtype=np.float32).reshape(5, 6))
    o1 = sampler((), x)
    put("sampler/in", x)
    put("sampler/raw_action", o1.rollout_extras)
    put("sampler/action", o1.output["action"])
    put("sampler/loglik", o1.output["log_likelihood"])
    put("sampler/reg", o1.regularization_loss)
    o2 = sampler((), x, o1.rollout_extras)                      # replay: same raw action, fresh entropy noise
    put("sampler/replay_loglik", o2.output["log_likelihood"])
    put("sampler/replay_reg", o2.regularization_loss)
    put("sampler/count_after_two_calls", np.asarray(s_rngs.default.count.value))

    # ---------------- Normalizer
    g = np.random.default_rng(0)
    nz = Normalizer(4)
    b1 = (2.0 + 3.0 * g.standard_normal((5, 7, 4))).astype(np.float32)
    b2 = (-1.0 + 0.5 * g.standard_normal((3, 7, 4))).astype(np.float32)
    put("norm/default_out", nz((), jp.asarray(b1[0])).output)
    nz.update_statistics(jp.asarray(b1)); nz.update_statistics(jp.asarray(b2))
    put("norm/b1", b1); put("norm/b2", b2)
    put("norm/mean", nz.mean.value); put("norm/M2", nz.M2.value); put("norm/counter", nz.counter.value)
    put("norm/out", nz((), jp.asarray(b2[0])).output)

    # ---------------- gae on the inputs of the reference's own test (ppo_test.py:229-264; gae_kat.npz holds them)
    kat = np.load(os.path.join(HERE, "gae_kat.npz"))
    rewards, values = kat["rewards_f32"], kat["values_f32"]
    T, B = rewards.shape
    done = np.unpackbits(kat["done_bits"])[:T * B].reshape(T, B).astype(bool)
    trunc = np.unpackbits(kat["trunc_bits"])[:T * B].reshape(T, B).astype(bool)
    adv = ppo.gae(jp.asarray(rewards), jp.asarray(values[:-1]), jp.asarray(values[-1]), jp.asarray(done),
                  jp.asarray(trunc), 0.95, 0.8)
    put("gae/advantages_f32", adv)

    # ---------------- optax
    p0 = (0.1 * g.standard_normal(50)).astype(np.float32)
    grads = [(0.5 * g.standard_normal(50)).astype(np.float32) for _ in range(3)]
    put("adam/p0", p0); put("adam/grads", np.stack(grads))
    for name, tx in (("clip_adam", optax.chain(optax.clip_by_global_norm(0.5), optax.adam(1e-3))),
                     ("adamw", optax.adamw(1e-3, weight_decay=1e-2)), ("adam", optax.adam(1e-3))):
        p, st = jp.asarray(p0), None
        st = tx.init(p)
        for gr in grads:
            upd, st = tx.update(jp.asarray(gr), st, p)
            p = optax.apply_updates(p, upd)
        put(f"adam/{name}_p3", p)

    # ---------------- LSTM cell
    cell = nnx.OptimizedLSTMCell(3, 4, rngs=nnx.Rngs(1))
    cs = nnx.state(cell, nnx.Param)
    flat, _ = jax.tree_util.tree_flatten_with_path(cs)
    for path, leaf in flat:
        put("lstm/param/" + "/".join(str(getattr(q, "key", getattr(q, "name", q))) for q in path), leaf)
    c0 = jp.asarray(g.standard_normal((2, 4)).astype(np.float32))
    h0 = jp.asarray(g.standard_normal((2, 4)).astype(np.float32))
    xin = jp.asarray(g.standard_normal((2, 3)).astype(np.float32))
    (c1, h1), y = cell((c0, h0), xin)
    put("lstm/c0", c0); put("lstm/h0", h0); put("lstm/x", xin); put("lstm/c1", c1); put("lstm/h1", h1)

    # ---------------- one ppo_step on the synthetic env, restated in JAX
    from oracle import env as oenv
    O, A, MAXLEN, THR = 6, 2, 16, 2000
    Wo_np, Wa_np = oenv.make_env_weights(O, A, 0)
    Wo, Wa = jp.asarray(Wo_np), jp.asarray(Wa_np)
    from nnx_ppo.algorithms.types import EnvState as _ES   # protocol only
    import dataclasses as _dc
    from nnx_ppo.jax_dataclass import JaxDataclass

    @_dc.dataclass
    class S(JaxDataclass):
        obs: jax.Array
        reward: jax.Array
        done: jax.Array
        info: dict
        metrics: dict

    class SynthEnv:
        def reset(self, rng):
            k_base, k_cnt = jax.random.split(rng)
            obs = jax.random.normal(k_base, (O,))
            cnt = jax.random.randint(k_cnt, (), 0, MAXLEN // 2)
            kc = jax.random.key_data(k_cnt)
            term = (kc[0] ^ kc[1]).astype(jp.uint32)
            return S(obs, jp.float32(0.0), jp.float32(0.0),
                     {"step_counter": cnt, "term_state": term, "truncated": jp.array(False)}, {})

        def step(self, s, a):
            obs = jp.tanh(s.obs @ Wo + a @ Wa)
            reward = -jp.mean(obs * obs)
            cnt = s.info["step_counter"] + 1
            term = s.info["term_state"] * jp.uint32(1664525) + jp.uint32(1013904223)
            terminated = (term >> 16) < THR
            truncated = cnt >= MAXLEN
            done = terminated | truncated
            return S(obs, reward, done.astype(jp.float32),
                     {"step_counter": cnt, "term_state": term, "truncated": truncated}, {})

    env = SynthEnv()
    nets2 = factories.make_mlp_actor_critic(O, A, [8, 8], [8], nnx.Rngs(0))
    Bn, Tn, En, Mn = 16, 5, 2, 2
    ts = ppo.new_training_state(env, nets2, Bn, 17)
    put("step/rng_key0", kd(ts.rng_key))
    put("step/env_obs0", ts.env_states.obs)
    put("step/env_counter0", ts.env_states.info["step_counter"])
    ts1, metrics = ppo.ppo_step(env, ts, Bn, Tn, 0.95, 0.99, 0.2, True, False, En, Mn, 1.0, LoggingLevel.LOSSES)
    reset_key, new_key = jax.random.split(ts.rng_key)
    put("step/indices", np.concatenate([np.asarray(jax.random.permutation(jax.random.fold_in(new_key, e), Bn)).reshape(Mn, Bn // Mn)
                                         for e in range(En)]))
    for k_, v in metrics.items():
        put("step/metric/" + k_, v)
    put("step/params1", np.concatenate([np.asarray(p).ravel() for p in jax.tree.leaves(nnx.state(ts1.networks, nnx.Param))]))
    put("step/env_obs1", ts1.env_states.obs)
    put("step/env_counter1", ts1.env_states.info["step_counter"])
    put("step/rng_key1", kd(ts1.rng_key))
    put("step/steps_taken", ts1.steps_taken)
    put("step/shape", np.array([O, A, MAXLEN, THR, Bn, Tn, En, Mn], np.int32))

    put("meta/versions", np.array([f"jax {jax.__version__}", f"optax {optax.__version__}"]))
    np.savez_compressed(args.out, **out)
    print(f"wrote {args.out}: {len(out)} arrays, {sum(a.nbytes for a in out.values())} bytes")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
