"""GPU parity of the recurrent-actor step kernels (csrc/recurrent.cu) against oracle/recurrent.py:
the replay forward over T steps from the minibatch's start carry with the rollout's resets, and the
BPTT backward fed with the oracle's d loss / d y.  SURVEY section 8 row a15 (first CUDA path; the
public-API integration of recurrent networks is the next step)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nnx_ppo_b200 import _lib                                    # noqa: E402
from oracle import env as oenv, prng, recurrent as orec          # noqa: E402

pytestmark = pytest.mark.gpu
F = np.float32


def torch_equal_rows(mat, vec):
    import torch
    return bool(torch.equal(mat, vec.to(mat.device)[None, :].expand_as(mat)))


def _plan(net):
    """flat layout = RecurrentActorCritic.param_list(): W1, b1, Wi, Wh, bl, W2, b2, critic..."""
    p = _lib.LstmPlan()
    O, P, H, Y = net.obs_dim, net.pre.W[0].shape[1], net.lstm.hidden, net.post.W[0].shape[1]
    p.obs_dim, p.pre_dim, p.hidden, p.out_dim = O, P, H, Y
    p.act, p.normalize = net.pre.act, 1 if net.normalize else 0
    o = 0
    p.w1_off = o; o += O * P
    p.b1_off = o; o += P
    p.wcat_off = o; o += (P + H) * 4 * H
    p.bl_off = o; o += 4 * H
    if net.init_c is not None:                 # trainable_initial_state: after the LSTM bias in param_list()
        assert o % 4 == 0
        p.init_c_off = o; o += H
        p.init_h_off = o; o += H
    p.w2_off = o; o += H * Y
    p.b2_off = o; o += Y
    p.n_params = net.flat_params().size
    return p, o


@pytest.mark.parametrize("cfg", [dict(O=10, A=3, B=12, T=9, P=7, H=8, act="tanh", mb=[0, 3, 4, 7, 9, 11]),
                                 dict(O=64, A=8, B=40, T=16, P=64, H=256, act="relu", mb=list(range(0, 40, 2))),
                                 dict(O=16, A=4, B=37, T=20, P=32, H=32, act="relu", mb=list(range(37)))])
def test_recurrent_replay_and_bptt_match_oracle(cuda_device, cfg):
    import torch
    dev = cuda_device
    lib = _lib.load()
    O, A, B, T, H = cfg["O"], cfg["A"], cfg["B"], cfg["T"], cfg["H"]
    net = orec.make_recurrent_actor_critic(O, A, [cfg["P"]], H, [], [6], seed=3, activation=cfg["act"])
    e = oenv.SyntheticEnv(O, A, max_len=6, term_thresh16=9000)
    es = e.reset(prng.split(prng.key(5), B))
    es, carry, ro, start = orec.unroll_env(e, es, net, net.initialize_state(B), T, prng.key(11))
    net.update_statistics(ro.obs)                                 # a non-trivial normaliser
    assert ro.done.sum() > 0
    inds = np.asarray(cfg["mb"], np.int32)
    mb = len(inds)
    base = net.rng_count
    total, m, g_ref = orec.ppo_loss_and_grads(net, ro, start, inds, base)

    plan, n_rec = _plan(net)
    t = lambda a, dt=torch.float32: torch.tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    params = t(net.flat_params())
    mean, std = t(net.mean), t(net.norm_std())
    obs, done = t(ro.obs), t(ro.done.astype(np.uint8), torch.uint8)
    ind_d = t(inds, torch.int32)
    c, h = t(start[0][inds]), t(start[1][inds])
    Y = 2 * A
    y = torch.zeros(T, mb, Y, device=dev)
    cf = int(lib.b200ppo_lstm_cache_floats(plan, mb))
    cache = torch.zeros(T, cf, device=dev)
    s = _lib.current_stream()
    for k in range(T):
        _lib.check(lib.b200ppo_lstm_step_fwd(s, plan, params.data_ptr(), mean.data_ptr(), std.data_ptr(),
                                             obs[k].data_ptr(), ind_d.data_ptr(), done[k].data_ptr(), mb,
                                             c.data_ptr(), h.data_ptr(), y[k].data_ptr(), cache[k].data_ptr()), "fwd")
    torch.cuda.synchronize()
    # replay forward: actor outputs of every step (the oracle's loss consumed exactly these)
    A2 = m["loglik"].shape                                       # (T, mb)
    y_ref = None
    # recompute the oracle's y through its own step function
    cc, hh = start[0][inds].copy(), start[1][inds].copy()
    ys = []
    xs = net.normalize_obs(ro.obs[:, inds].reshape(T * mb, -1)).reshape(T, mb, -1)
    for k in range(T):
        cc, hh, yk, _ = orec.actor_step(net, cc, hh, xs[k])
        ys.append(yk)
        keep = (~ro.done[k, inds])[:, None]
        cc, hh = cc * keep, hh * keep
    y_ref = np.stack(ys)
    assert np.abs(y.cpu().numpy() - y_ref).max() < 2e-5 * max(1.0, np.abs(y_ref).max())
    assert np.abs(c.cpu().numpy() - cc).max() < 2e-5 and np.abs(h.cpu().numpy() - hh).max() < 2e-5
    d_last = ro.done[-1, inds]
    assert np.all(c.cpu().numpy()[d_last] == 0) and np.all(h.cpu().numpy()[d_last] == 0)     # reset: exact zeros

    # BPTT with the oracle's d loss / d y
    d_y = t(m["d_y"].reshape(T, mb, Y))
    grad = torch.zeros(plan.n_params, device=dev)
    dc, dh = torch.zeros(mb, H, device=dev), torch.zeros(mb, H, device=dev)
    for k in reversed(range(T)):
        _lib.check(lib.b200ppo_lstm_step_bwd(s, plan, params.data_ptr(), d_y[k].data_ptr(), cache[k].data_ptr(),
                                             ind_d.data_ptr(), done[k].data_ptr(), mb, dc.data_ptr(), dh.data_ptr(),
                                             grad.data_ptr(), 0, 0, 0, 0), "bwd")
    torch.cuda.synchronize()
    g = grad.cpu().numpy()[:n_rec]
    ref = g_ref[:n_rec]
    scale = np.abs(ref).max()
    assert np.abs(g - ref).max() < 3e-4 * scale, (np.abs(g - ref).max(), scale)
    assert np.all(grad.cpu().numpy()[n_rec:] == 0)               # the critic's slots are not touched here

    # deterministic variant: the steps write their GEMM operands, three batched GEMMs reduce them
    P = cfg["P"]
    cat, hn = torch.zeros(T, mb, P + H, device=dev), torch.zeros(T, mb, H, device=dev)
    da, dz = torch.zeros(T, mb, 4 * H, device=dev), torch.zeros(T, mb, P, device=dev)
    scratch = torch.zeros(int(lib.b200ppo_lstm_wgrad_scratch_floats(plan, T * mb)), device=dev)
    outs = []
    for _rep in range(2):
        grad2 = torch.full((plan.n_params,), 7.0, device=dev)        # must be overwritten, not accumulated
        dc.zero_(); dh.zero_()
        for k in reversed(range(T)):
            _lib.check(lib.b200ppo_lstm_step_bwd(s, plan, params.data_ptr(), d_y[k].data_ptr(), cache[k].data_ptr(),
                                                 ind_d.data_ptr(), done[k].data_ptr(), mb, dc.data_ptr(), dh.data_ptr(),
                                                 0, cat[k].data_ptr(), hn[k].data_ptr(), da[k].data_ptr(),
                                                 dz[k].data_ptr()), "bwd(deferred)")
        _lib.check(lib.b200ppo_lstm_weight_grads(s, plan, cache.data_ptr(), cat.data_ptr(), hn.data_ptr(),
                                                 da.data_ptr(), dz.data_ptr(), d_y.data_ptr(), T * mb,
                                                 grad2.data_ptr(), scratch.data_ptr()), "weight_grads")
        torch.cuda.synchronize()
        outs.append(grad2.cpu().numpy().copy())
    assert np.abs(outs[0][:n_rec] - ref).max() < 3e-4 * scale
    assert np.array_equal(outs[0], outs[1])                      # fixed-order sums: bit-reproducible
    assert np.all(outs[0][n_rec:] == 7.0)


@pytest.mark.parametrize("fused_env", ["1", "0"])
@pytest.mark.parametrize("trainable", [False, True])
def test_recurrent_ppo_step_matches_oracle(cuda_device, trainable, fused_env, monkeypatch):
    """Public API (`ppo.ppo_step`) with an LSTM actor against oracle/recurrent.py over iterations:
    bit-exact masks / minibatch indices / counters, float32-tolerance losses and parameters
    (reference: recurrent_test.py:285-330 only checks finiteness and that parameters change).
    fused_env: the synthetic env's step as three kernels (b200ppo_synth_env_step) or through the generic RLEnv
    protocol (env.step / env.reset / tree_where in torch ops)."""
    monkeypatch.setenv("B200PPO_REC_FUSED_ENV", fused_env)
    from nnx_ppo_b200 import Rngs
    from nnx_ppo_b200.algorithms import ppo
    from nnx_ppo_b200.envs import SyntheticEnv
    from nnx_ppo_b200.networks.factories import make_recurrent_actor_critic
    from nnx_ppo_b200.networks.plan import compile_network
    O, A, B, T, E, M, P, H = 16, 4, 64, 12, 2, 2, 32, 32
    nets = make_recurrent_actor_critic(O, A, P, H, [48], Rngs(1), trainable_initial_state=trainable)
    onet = orec.make_recurrent_actor_critic(O, A, [P], H, [], [48], seed=1, trainable_initial_state=trainable)
    if trainable:       # trainable_initial_state (recurrent.py:85-87): start from a non-zero learned carry on both sides
        gi = np.random.default_rng(8)
        v0, v1 = (0.3 * gi.standard_normal(H)).astype(F), (0.3 * gi.standard_normal(H)).astype(F)
        lstm = nets.layers[1].action.layers[1]
        lstm.initial_h.set(v0); lstm.initial_c.set(v1)
        onet.init_c[:], onet.init_h[:] = v0, v1
    env = SyntheticEnv(O, A, max_len=10, term_thresh16=3000)
    oe = oenv.SyntheticEnv(O, A, max_len=10, term_thresh16=3000)
    ts = ppo.new_training_state(env, nets, B, 17)
    ots = orec.new_training_state(oe, onet, B, 17)
    net = compile_network(nets)
    assert net.recurrent
    u32 = lambda t: t.cpu().numpy().view(np.uint32)
    for it in range(3):             # eager iteration, then the captured CUDA graph (capture + replay, replay)
        ts, m = ppo.ppo_step(env, ts, B, T, 0.95, 0.99, 0.2, True, False, E, M)
        tr = {}
        ots, om = orec.ppo_step(oe, ots, B, T, n_epochs=E, n_minibatches=M, trace=tr)
        eng = next(iter(net.engines.values()))
        assert np.array_equal(eng.inds.cpu().numpy().reshape(E * M, B // M), tr["indices"])
        assert np.array_equal(eng.done.cpu().numpy().astype(bool), tr["rollout"].done)
        assert np.array_equal(eng.trunc.cpu().numpy().astype(bool), tr["rollout"].truncated)
        assert tr["rollout"].done.sum() > 0
        assert np.abs(eng.loglik.cpu().numpy() - tr["rollout"].loglik).max() < 2e-4
        assert tuple(ts.rng_key) == tuple(int(x) for x in ots.rng_key)
        assert float(ts.steps_taken) == float(ots.steps_taken) == (it + 1) * T * B
        cnt = u32(net.counters)
        assert cnt[2] == onet.rng_count and cnt[3] == (it + 1) * E * M == ots.opt.count
        for k in ("losses/actor/mean", "losses/critic/mean", "losses/regularization/mean"):
            assert abs(m[k] - om[k]) < 3e-4 * max(1.0, abs(om[k])), (k, m[k], om[k])
        p, po = net.params_logical(), onet.flat_params()
        assert p.shape == po.shape
        assert np.abs(p - po).max() < 1e-4 * 4 and np.mean(np.abs(p - po)) < 3e-6
        c, h = net.get_carry(ts.network_states)
        assert np.allclose(c.cpu().numpy(), ots.carry[0], atol=2e-4) and np.allclose(h.cpu().numpy(), ots.carry[1], atol=2e-4)
        assert np.allclose(net.normalizer.mean.numpy(), onet.mean, rtol=1e-4, atol=1e-4)
    assert eng.r_seq and eng.r_graph is not None and eng.kernel_launches_per_iter > 0    # tensor-core sequence kernels, one graph
    if trainable:
        lstm = nets.layers[1].action.layers[1]
        assert np.abs(lstm.initial_h.numpy() - v0).max() > 1e-5 and np.abs(lstm.initial_c.numpy() - v1).max() > 1e-5   # learned
        assert np.allclose(lstm.initial_h.numpy(), onet.init_c, atol=4e-4) and np.allclose(lstm.initial_c.numpy(), onet.init_h, atol=4e-4)
        # initialize_state / reset_state hand out the learned vectors broadcast over the batch (recurrent_test.py:113-150)
        st = nets.initialize_state(5)
        c0, h0 = net.get_carry(st)
        assert c0.shape == (5, H) and torch_equal_rows(c0, lstm.initial_h.value) and torch_equal_rows(h0, lstm.initial_c.value)


def test_recurrent_ffma_fallback_matches_tensor_core_path(cuda_device, monkeypatch):
    """Sizes the tensor-core kernels do not take (hidden % 16 != 0) run the per-step FFMA kernels; on a size both
    take, the two paths agree to float32 accuracy."""
    from nnx_ppo_b200 import Rngs
    from nnx_ppo_b200.algorithms import ppo
    from nnx_ppo_b200.envs import SyntheticEnv
    from nnx_ppo_b200.networks.factories import make_recurrent_actor_critic
    from nnx_ppo_b200.networks.plan import compile_network
    res = []
    for mode in ("tc", "ffma"):
        monkeypatch.setenv("B200PPO_LSTM", mode)
        nets = make_recurrent_actor_critic(16, 4, 32, 32, [48], Rngs(1))
        env = SyntheticEnv(16, 4, max_len=10, term_thresh16=3000)
        ts = ppo.new_training_state(env, nets, 64, 17)
        for _ in range(3):
            ts, m = ppo.ppo_step(env, ts, 64, 12, 0.95, 0.99, 0.2, True, False, 2, 2)
        net = compile_network(nets)
        eng = next(iter(net.engines.values()))
        assert eng.r_seq == (mode == "tc")
        res.append((net.params_logical().copy(), eng.done.cpu().numpy().copy(), m))
    assert np.array_equal(res[0][1], res[1][1])
    assert np.abs(res[0][0] - res[1][0]).max() < 4e-4 and np.mean(np.abs(res[0][0] - res[1][0])) < 3e-6
    nets = make_recurrent_actor_critic(10, 3, 7, 8, [6], Rngs(1))            # hidden 8: FFMA kernels only
    env = SyntheticEnv(10, 3, max_len=10, term_thresh16=3000)
    ts = ppo.new_training_state(env, nets, 16, 17)
    monkeypatch.delenv("B200PPO_LSTM")
    for _ in range(3):
        ts, m = ppo.ppo_step(env, ts, 16, 6, 0.95, 0.99, 0.2, True, False, 2, 2)
    eng = next(iter(compile_network(nets).engines.values()))
    assert not eng.r_seq and all(np.isfinite(float(v)) for v in m.values())


def test_recurrent_network_call_and_eval(cuda_device):
    """`networks(state, obs)` with an LSTM actor: functional carry (input state untouched), replay of the
    stored raw actions reproduces the log-likelihoods (adapter_test.py:61-75), eval_rollout runs."""
    import torch
    from nnx_ppo_b200 import Rngs, prng as pprng
    from nnx_ppo_b200.algorithms import rollout
    from nnx_ppo_b200.envs import SyntheticEnv
    from nnx_ppo_b200.networks.factories import make_recurrent_actor_critic
    from nnx_ppo_b200.networks.plan import compile_network
    O, A, B, H = 12, 3, 20, 16
    nets = make_recurrent_actor_critic(O, A, 8, H, [10], Rngs(2))
    onet = orec.make_recurrent_actor_critic(O, A, [8], H, [], [10], seed=2)
    net = compile_network(nets)
    obs = torch.randn(B, O, device=cuda_device)
    s0 = nets.initialize_state(B)
    out1 = nets(s0, obs)
    c0, h0 = net.get_carry(s0)
    assert float(c0.abs().max()) == 0.0 and float(h0.abs().max()) == 0.0          # input state not modified
    c1, h1 = net.get_carry(out1.next_state)
    oc, oh = onet.initialize_state(B)
    (oc, oh), oo = orec.policy_forward(onet, (oc, oh), obs.cpu().numpy())
    assert np.allclose(h1.cpu().numpy(), oh, atol=2e-5) and np.allclose(c1.cpu().numpy(), oc, atol=2e-5)
    assert np.allclose(out1.output.actions.cpu().numpy(), oo["action"], atol=2e-5)
    assert np.allclose(out1.output.loglikelihoods.cpu().numpy(), oo["loglik"], atol=2e-4)
    assert np.allclose(out1.output.value_estimates.cpu().numpy(), oo["value"], atol=2e-5)
    assert np.allclose(out1.regularization_loss.cpu().numpy(), oo["reg"], atol=2e-5)
    assert net.rng_count == onet.rng_count
    out2 = nets(s0, obs, out1.rollout_extras)                                      # replay
    assert np.allclose(out2.output.loglikelihoods.cpu().numpy(), out1.output.loglikelihoods.cpu().numpy(), atol=1e-6)
    env = SyntheticEnv(O, A, max_len=10, term_thresh16=3000)
    nets.eval()
    m = rollout.eval_rollout(env, nets, 16, 12, pprng.key(3))
    nets.train()
    assert np.isfinite(m["episode_reward/mean"]) and 0 <= m["lifespan_mean"] <= 12


def test_train_ppo_with_recurrent_actor(cuda_device):
    """The drop-in entry point with an LSTM actor (recurrent_test.py:285-330 through train_ppo):
    iterations run, evaluation (deterministic, carry threaded through eval_rollout) runs, parameters move."""
    from nnx_ppo_b200 import Rngs
    from nnx_ppo_b200.algorithms import ppo
    from nnx_ppo_b200.algorithms.config import EvalConfig, PPOConfig, TrainConfig
    from nnx_ppo_b200.envs import SyntheticEnv
    from nnx_ppo_b200.networks.factories import make_recurrent_actor_critic
    from nnx_ppo_b200.networks.plan import compile_network
    env = SyntheticEnv(12, 3, max_len=16)
    nets = make_recurrent_actor_critic(12, 3, 16, 16, [24], Rngs(4))
    p0 = compile_network(nets).params_logical().copy()
    logged = []
    cfg = TrainConfig(ppo=PPOConfig(n_envs=32, rollout_length=8, total_steps=32 * 8 * 3, n_minibatches=2, n_epochs=2),
                      eval=EvalConfig(enabled=True, every_steps=32 * 8 * 2, n_envs=8, max_episode_length=10))
    res = ppo.train_ppo(env, nets, cfg, seed=5, log_fn=lambda m, s: logged.append((s, dict(m))))
    assert res.total_iterations == 3 and res.total_steps == 32 * 8 * 3
    assert all(np.isfinite(float(v)) for _, m in logged for v in m.values() if np.isscalar(v) or hasattr(v, "__float__"))
    assert any(k.startswith("eval") or "episode_reward" in k for _, m in logged for k in m)
    assert np.abs(compile_network(nets).params_logical() - p0).max() > 0
