"""ppo_step for a recurrent (LSTM) actor — reference ppo.py:254-348 with the carry handling of
rollout.py:11-45 (reset on done) and ppo.py:409-431 (replay from the rollout's start carry).

SURVEY section 8 row a15.  Only ``h_{t-1} Wh`` is sequential, so a minibatch replay is a handful of batched
tensor-core GEMMs plus T step launches (csrc/recurrent_tc.cu: ``b200ppo_lstm_seq_forward`` /
``_backward``); the critic, GAE, the loss head, the gradient reduction and Adam are the MLP path's kernels
driven stage by stage (networks/rplan.py).  The rollout runs the same sequence kernel with T = 1 per env
step.  The whole iteration — rollout, permutation, E*M updates, Normalizer statistics — is a fixed launch
sequence over engine-owned buffers and is captured into ONE CUDA graph on its second run (the reference
jits ``ppo_step``); per-iteration inputs (PRNG keys, hyper-parameters) reach the kernels through the
engine's device block, so the graph is never re-captured.  Networks whose sizes the tensor-core kernels do
not take (hidden % 16, pre_dim % 4) use the per-step FFMA kernels of csrc/recurrent.cu instead, launched
from the host.  Data parallel like the MLP path (envs sharded over ranks): the advantage-moment and gradient
exchanges run inside the GAE / loss / Adam kernels over peer memory (the Adam stage exchanges whatever gradient
it is handed, so the recurrent actor's gradient travels with the critic's); the peer path is required.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .. import _lib, prng
from ..networks.plan import compile_network
from .rollout import policy_values, tree_where
from .types import LoggingLevel


def _engine(net, opt, env, n_envs, T, E, M, lam, gamma, clip, norm_adv, cw):
    from .engine import cached_engine
    from . import ppo as _ppo
    world, group = _ppo._dist_info()
    shape_key = (n_envs, T, E, M, bool(norm_adv), opt.gradient_clipping is not None, opt.wd_value >= 0.0, world)
    eng = cached_engine(net, "recurrent", env, opt, shape_key,
                        lambda: _build_engine(net, opt, n_envs, T, E, M, lam, gamma, clip, norm_adv, cw, world, group))
    eng.set_hparams(lam, gamma, clip, cw)
    return eng


def _ptr(v) -> int:
    return int(C.cast(v, C.c_void_p).value)


def _build_engine(net, opt, n_envs, T, E, M, lam, gamma, clip, norm_adv, cw, world=1, group=None):
    from .engine import PPOEngine
    import torch
    eng = PPOEngine.__new__(PPOEngine)
    fake_env = type("E", (), {"fused_rollout": True})()
    PPOEngine.__init__(eng, net, fake_env, opt, n_envs, T, E, M, lam, gamma, clip, norm_adv, cw,
                       world_size=world, group=group, use_graph=False)
    if world > 1 and not eng.p2p:
        raise NotImplementedError("data-parallel recurrent training needs the peer-memory exchange (B200PPO_P2P=0 is "
                                  "set, or the GPUs have no peer access)")
    lib, lp, mb, dev = eng.lib, net.lplan, eng.mb, net.device
    H, Y, P = lp.hidden, lp.out_dim, lp.pre_dim
    f32 = dict(dtype=torch.float32, device=dev)
    eng.r_seq = bool(lib.b200ppo_lstm_seq_supported(lp)) and os.environ.get("B200PPO_LSTM", "tc") != "ffma"
    if net.init_views is not None and not eng.r_seq:
        raise NotImplementedError("trainable_initial_state is implemented by the sequence kernels only (B200PPO_LSTM=ffma is set)")
    eng.r_c, eng.r_h = torch.zeros(mb, H, **f32), torch.zeros(mb, H, **f32)
    eng.r_y = torch.zeros(n_envs, Y, **f32)
    eng.r_carry = (torch.zeros(n_envs, H, **f32), torch.zeros(n_envs, H, **f32))      # live carry of every env
    eng.r_start = (torch.zeros(n_envs, H, **f32), torch.zeros(n_envs, H, **f32))      # carry the rollout started from
    eng.r_inds64 = torch.zeros(E * n_envs, dtype=torch.int64, device=dev)
    eng.r_keys = torch.zeros(T * n_envs, 2, dtype=torch.int32, device=dev)
    wsp = eng.ws.data_ptr()
    eng.r_y_ptr = _ptr(lib.b200ppo_update_debug_ptr(net.plan, T, mb, wsp, 2))
    eng.r_dy_ptr = _ptr(lib.b200ppo_update_debug_ptr(net.plan, T, mb, wsp, 3))
    eng.r_xhat_ptr = _ptr(lib.b200ppo_update_debug_ptr(net.plan, T, mb, wsp, 5))
    eng.r_grad_ptr = _ptr(lib.b200ppo_update_grad_ptr(net.plan, T, mb, wsp))
    if eng.r_seq:
        n_roll = int(lib.b200ppo_lstm_seq_workspace_floats(lp, 1, n_envs))
        n_upd = int(lib.b200ppo_lstm_seq_workspace_floats(lp, T, mb))
        eng.r_ws = torch.zeros(max(n_roll, n_upd) + 64, **f32)
        eng.r_grad2 = torch.zeros(int(net.n_params), **f32)       # the recurrent actor's gradient (see _enqueue_iteration)
    else:
        eng.r_cache = torch.zeros(T, int(lib.b200ppo_lstm_cache_floats(lp, mb)), **f32)
        eng.r_dc, eng.r_dh = torch.zeros(mb, H, **f32), torch.zeros(mb, H, **f32)
        # GEMM operands of the deterministic weight-gradient pass, step-major [T][mb][width]
        eng.r_cat = torch.zeros(T, mb, P + H, **f32)
        eng.r_hn = torch.zeros(T, mb, H, **f32)
        eng.r_da = torch.zeros(T, mb, 4 * H, **f32)
        eng.r_dz = torch.zeros(T, mb, P, **f32)
        eng.r_scratch = torch.zeros(int(lib.b200ppo_lstm_wgrad_scratch_floats(lp, T * mb)), **f32)
    # measured on B200 (same box, configs[2]): 43.1 / 41.3 ms per iteration with the overlap, 40.9 / 40.9 without - the
    # full-device GEMM kernels of the critic's backward delay the latency-critical step chain more than they hide
    eng.r_overlap = os.environ.get("B200PPO_REC_OVERLAP", "0") == "1"
    eng.r_env = None            # engine-owned env state (the object a captured graph points at)
    eng.r_graph = None
    eng.r_graphable = False
    eng.r_iters = 0
    return eng


def _policy(eng, net, obs, t, s, sample=True):
    """One network call of the rollout on all envs: recurrent actor + sampler (rollout.py:18)."""
    lib, lp, plan = eng.lib, net.lplan, net.plan
    B, A = eng.B, plan.act_dim
    c, h = eng.r_carry
    mean_p, std_p = net.norm_ptrs()
    arena = net.arena.data_ptr()
    if eng.r_seq:
        # the parameters do not change during the rollout: the split weight planes of step 0 serve all T steps
        _lib.check(lib.b200ppo_lstm_seq_forward(s, lp, arena, mean_p, std_p, obs.data_ptr(), 0, 0, B, c.data_ptr(),
                                                h.data_ptr(), 1, B, eng.r_ws.data_ptr(), eng.r_y.data_ptr(), 0 if t == 0 else 2),
                   "lstm_seq_forward(rollout)")
        n = int(lib.b200ppo_lstm_seq_num_launches(lp, 1, B, 0)) - (0 if t == 0 else 1)
    else:
        _lib.check(lib.b200ppo_lstm_step_fwd(s, lp, arena, mean_p, std_p, obs.data_ptr(), 0, 0, B, c.data_ptr(),
                                             h.data_ptr(), eng.r_y.data_ptr(), 0), "lstm_step_fwd(rollout)")
        n = 1
    if not sample:
        return n                      # the fused env step samples from eng.r_y itself (same key, same element index)
    _lib.check(lib.b200ppo_sampler_step(s, eng.r_y.data_ptr(), B, A, 0, plan.min_std, plan.std_scale,
                                        plan.entropy_weight, net.counters.data_ptr(), 2 * t, 0,
                                        eng.raw_action[t].data_ptr(), eng.action[t].data_ptr(),
                                        eng.loglik[t].data_ptr(), 0), "sampler_step")
    return n + 1


def _enqueue_iteration(eng, net, env, env_state):
    """The whole iteration on the current stream.  Returns (final env state, kernel launches of this library)."""
    import torch
    lib, lp, plan, mb = eng.lib, net.lplan, net.plan, eng.mb
    T, B = eng.T, eng.B
    Y = lp.out_dim
    s = _lib.current_stream()
    n = 0
    if net.normalizer is not None:
        net.normalizer.prepare(s); n += 1
    mean_p, std_p = net.norm_ptrs()
    arena = net.arena.data_ptr()
    c, h = eng.r_carry
    eng.r_start[0].copy_(c)                                              # network_state the replay starts from
    eng.r_start[1].copy_(h)
    # rng_keys_for_env_reset = split(reset_key, (T, B)) (rollout.py:57-59), from the device key block
    _lib.check(lib.b200ppo_split_keys_dev(s, eng.iter_keys.data_ptr(), 0, T * B, eng.r_keys.data_ptr()), "split_keys"); n += 1
    keys_all = eng.r_keys.view(T, B, 2)
    # ---------------- rollout (rollout.py:48-73)
    # Device envs with fused step kernels (the synthetic env): sampler, env step, reward / done / truncation, reset
    # select and the transition record are three launches per step (include/b200ppo.h, b200ppo_synth_env_step)
    # instead of ~20 torch ops incl. an env.reset of every env; same arithmetic, same keys (B200PPO_REC_FUSED_ENV=0
    # keeps the generic protocol below, which any RLEnv takes).
    fused_env = (eng.r_seq and getattr(env, "fused_rollout", False) and hasattr(env, "c_struct")
                 and os.environ.get("B200PPO_REC_FUSED_ENV", "1") != "0")
    if fused_env:
        es = env.c_struct(eng.dev)
        O, A = plan.obs_dim, plan.act_dim
        nb = int(lib.b200ppo_synth_env_step_workspace_bytes(O, A, B))
        if getattr(eng, "r_envws", None) is None or eng.r_envws.numel() * 4 < nb:
            eng.r_envws = torch.zeros((nb + 3) // 4, dtype=torch.float32, device=eng.dev)
        wsp = eng.r_envws.data_ptr()
        _lib.check(lib.b200ppo_synth_env_begin(s, es, B, env_state.obs.data_ptr(), eng.obs.data_ptr(), wsp, nb),
                   "synth_env_begin"); n += 2
        for t in range(T):
            n += _policy(eng, net, eng.obs[t], t, s, sample=False)
            _lib.check(lib.b200ppo_synth_env_step(
                s, es, eng.r_y.data_ptr(), Y, plan.min_std, plan.std_scale, net.counters.data_ptr(),
                eng.iter_keys.data_ptr(), t, T, B, env_state.obs.data_ptr(), env_state.step_counter.data_ptr(),
                env_state.term_state.data_ptr(), eng.obs.data_ptr(), eng.raw_action.data_ptr(), eng.action.data_ptr(),
                eng.loglik.data_ptr(), eng.reward.data_ptr(), eng.done.data_ptr(), eng.trunc.data_ptr(),
                eng.next_obs_last.data_ptr(), wsp, nb), "synth_env_step"); n += 3
            done = eng.done[t].bool()
            if net.init_views is None:                                   # reset_state -> zeros (rollout.py:33-40)
                keep = (~done).to(torch.float32)[:, None]
                c.mul_(keep)
                h.mul_(keep)
            else:
                torch.where(done[:, None], net.init_views[0], c, out=c)
                torch.where(done[:, None], net.init_views[1], h, out=h)
    for t in range(0 if fused_env else T):                               # generic RLEnv protocol
        obs = env_state.obs.contiguous()
        n += _policy(eng, net, obs, t, s)
        nxt = env.step(env_state, eng.action[t])
        done = nxt.done.bool()
        tr = nxt.info.get("truncated", torch.zeros_like(done)) if isinstance(nxt.info, dict) else torch.zeros_like(done)
        eng.obs[t].copy_(obs)
        eng.reward[t].copy_(nxt.reward.float())
        eng.done[t].copy_(done.to(torch.uint8))
        eng.trunc[t].copy_(tr.bool().to(torch.uint8))
        if t == T - 1:
            eng.next_obs_last.copy_(nxt.obs)
        if net.init_views is None:                                       # reset_state -> zeros (rollout.py:33-40)
            keep = (~done).to(torch.float32)[:, None]
            c.mul_(keep)
            h.mul_(keep)
        else:                                                            # ... or the learned initial carry
            torch.where(done[:, None], net.init_views[0], c, out=c)
            torch.where(done[:, None], net.init_views[1], h, out=h)
        env_state = tree_where(done, env.reset(keys_all[t].contiguous()), nxt)
    # ---------------- E x M minibatch updates (ppo.py:284-328)
    _lib.check(lib.b200ppo_permutation(s, eng.iter_keys.data_ptr() + 8, B, eng.E, eng.inds.data_ptr(),
                                       eng.perm_scratch.data_ptr()), "permutation"); n += 1
    eng.r_inds64.copy_(eng.inds.view(-1))
    ST = _lib
    for u in range(eng.n_updates):
        off = 2 * T + u * 2 * (T + 1)
        args = (s, plan, eng.hp, eng.bufs[u], T, B, mb, off, u)
        ip = eng.inds.data_ptr() + 4 * u * mb
        _lib.check(lib.b200ppo_update(*args, ST.STAGE_FWD), "update/fwd")            # critic values (+ stand-in actor)
        n += int(lib.b200ppo_update_num_launches(plan, eng.hp, T, mb, ST.STAGE_FWD))
        idx = eng.r_inds64[u * mb:(u + 1) * mb]
        torch.index_select(eng.r_start[0], 0, idx, out=eng.r_c)
        torch.index_select(eng.r_start[1], 0, idx, out=eng.r_h)
        if eng.r_seq:                                                                 # replay scan, ppo.py:409-431
            _lib.check(lib.b200ppo_lstm_seq_forward(s, lp, arena, 0, 0, eng.r_xhat_ptr, eng.done.data_ptr(), ip, B,
                                                    eng.r_c.data_ptr(), eng.r_h.data_ptr(), T, mb, eng.r_ws.data_ptr(),
                                                    eng.r_y_ptr, 1), "lstm_seq_forward(replay)")
            n += int(lib.b200ppo_lstm_seq_num_launches(lp, T, mb, 0))
        else:
            for t in range(T):
                _lib.check(lib.b200ppo_lstm_step_fwd(s, lp, arena, mean_p, std_p, eng.obs[t].data_ptr(), ip,
                                                     eng.done[t].data_ptr(), mb, eng.r_c.data_ptr(), eng.r_h.data_ptr(),
                                                     eng.r_y_ptr + 4 * t * mb * Y, eng.r_cache[t].data_ptr()),
                           "lstm_step_fwd(replay)")
            n += T
        _lib.check(lib.b200ppo_update(*args, ST.STAGE_GAE | ST.STAGE_LOSS), "update/loss")
        n += int(lib.b200ppo_update_num_launches(plan, eng.hp, T, mb, ST.STAGE_GAE | ST.STAGE_LOSS | ST.STAGE_BWD | ST.STAGE_RED))
        if eng.r_seq:                                                                 # BPTT + weight gradients
            # B200PPO_REC_OVERLAP=1 (off by default, see _build_engine): the critic's backward and the recurrent actor's
            # BPTT only share their inputs and can run as two branches of the captured graph.  The reduction writes
            # the whole flat gradient, so BPTT then writes the actor's entries to its own buffer, copied in after the join.
            if eng.r_overlap:
                cur = torch.cuda.current_stream()
                eng._side.wait_stream(cur)
                with torch.cuda.stream(eng._side):
                    _lib.check(lib.b200ppo_update(_lib.current_stream(), *args[1:], ST.STAGE_BWD | ST.STAGE_RED), "update/bwd")
                _lib.check(lib.b200ppo_lstm_seq_backward(s, lp, arena, eng.r_xhat_ptr, eng.r_dy_ptr, eng.done.data_ptr(), ip, B,
                                                         T, mb, eng.r_ws.data_ptr(), eng.r_grad2.data_ptr()), "lstm_seq_backward")
                cur.wait_stream(eng._side)
                eng.grad[:net.n_recurrent].copy_(eng.r_grad2[:net.n_recurrent])
            else:
                _lib.check(lib.b200ppo_update(*args, ST.STAGE_BWD | ST.STAGE_RED), "update/bwd")
                _lib.check(lib.b200ppo_lstm_seq_backward(s, lp, arena, eng.r_xhat_ptr, eng.r_dy_ptr, eng.done.data_ptr(), ip, B,
                                                         T, mb, eng.r_ws.data_ptr(), eng.r_grad_ptr), "lstm_seq_backward")
            n += int(lib.b200ppo_lstm_seq_num_launches(lp, T, mb, 1))
        else:
            _lib.check(lib.b200ppo_update(*args, ST.STAGE_BWD | ST.STAGE_RED), "update/bwd")
            eng.r_dc.zero_()
            eng.r_dh.zero_()
            for t in reversed(range(T)):
                _lib.check(lib.b200ppo_lstm_step_bwd(s, lp, arena, eng.r_dy_ptr + 4 * t * mb * Y,
                                                     eng.r_cache[t].data_ptr(), ip, eng.done[t].data_ptr(), mb,
                                                     eng.r_dc.data_ptr(), eng.r_dh.data_ptr(), 0,
                                                     eng.r_cat[t].data_ptr(), eng.r_hn[t].data_ptr(),
                                                     eng.r_da[t].data_ptr(), eng.r_dz[t].data_ptr()), "lstm_step_bwd")
            # weight gradients of the recurrent actor: three batched GEMMs over all T steps, fixed-order sums
            _lib.check(lib.b200ppo_lstm_weight_grads(s, lp, eng.r_cache.data_ptr(), eng.r_cat.data_ptr(),
                                                     eng.r_hn.data_ptr(), eng.r_da.data_ptr(), eng.r_dz.data_ptr(),
                                                     eng.r_dy_ptr, T * mb, eng.r_grad_ptr, eng.r_scratch.data_ptr()),
                       "lstm_weight_grads")
            n += T + 9
        _lib.check(lib.b200ppo_update(*args, ST.STAGE_ADAM), "update/adam")
        n += int(lib.b200ppo_update_num_launches(plan, eng.hp, T, mb, ST.STAGE_ADAM))
    if eng.world > 1:
        from .. import parallel
        # every column is pre-divided by the GLOBAL sample count except the gradient norm [3], identical on every rank
        if eng.hp.grad_clip > 0.0 and parallel.dist_info()[1] != 0:
            eng.metrics[:, 3].zero_()
        eng._allreduce(eng.metrics)
    # ---------------- Normalizer statistics, counters (ppo.py:329-346)
    if net.normalizer is not None:
        nz = net.normalizer
        _lib.check(lib.b200ppo_norm_batch_stats(s, eng.obs.data_ptr(), T * B, nz.size, eng.batch_stats.data_ptr(),
                                                eng.norm_scratch.data_ptr()), "norm_batch_stats")
        src = eng.batch_stats
        if eng.world > 1:                                  # per-rank moments merged in rank order, as the MLP engine does
            parallel.all_gather_into(eng.batch_stats_all.view(-1), eng.batch_stats, eng.group)
            src = eng.batch_stats_all
        _lib.check(lib.b200ppo_norm_merge(s, src.data_ptr(), eng.world, float(T * B), nz.size,
                                          nz.mean._dev.data_ptr(), nz.M2._dev.data_ptr(),
                                          nz.counter._dev.data_ptr()), "norm_merge")
        n += 3
    _lib.check(lib.b200ppo_iter_finalize(s, net.counters.data_ptr(), eng.rng_per_iter, eng.n_updates,
                                         eng.comm_epoch.data_ptr()), "iter_finalize"); n += 1
    return env_state, n


def _adopt_env_state(eng, env_state):
    """Engine-owned env state: the first state passed in is adopted, a different object later (checkpoint,
    rollback) is copied over the live tensors (same rule as PPOEngine.step)."""
    import dataclasses
    import torch
    if eng.r_env is None:
        eng.r_env = env_state
        return
    if env_state is eng.r_env:
        return
    for f in dataclasses.fields(env_state) if dataclasses.is_dataclass(env_state) else []:
        a, b = getattr(eng.r_env, f.name), getattr(env_state, f.name)
        if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a.shape == b.shape:
            a.copy_(b)


def _write_back_env(eng, final):
    """Copy the rollout's final env state into the engine-owned tensors (inside the captured sequence)."""
    import dataclasses
    import torch
    if final is eng.r_env or not dataclasses.is_dataclass(final):
        eng.r_env = final if not dataclasses.is_dataclass(final) else eng.r_env
        return
    for f in dataclasses.fields(final):
        a, b = getattr(eng.r_env, f.name), getattr(final, f.name)
        if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a.shape == b.shape:
            a.copy_(b.to(a.dtype))
        elif isinstance(a, dict) and isinstance(b, dict):
            for k in a:
                if isinstance(a[k], torch.Tensor) and k in b and isinstance(b[k], torch.Tensor) and a[k].shape == b[k].shape:
                    a[k].copy_(b[k].to(a[k].dtype))


def ppo_step_recurrent(env, training_state, n_envs, rollout_length, gae_lambda, discounting_factor,
                       clip_range, normalize_advantages, n_epochs, n_minibatches, critic_loss_weight,
                       logging_level, logging_percentiles):
    import dataclasses
    import torch
    from . import ppo as _ppo
    net = compile_network(training_state.networks)
    opt = training_state.optimizer
    T, B = rollout_length, n_envs
    eng = _engine(net, opt, env, B, T, n_epochs, n_minibatches, gae_lambda, discounting_factor, clip_range,
                  normalize_advantages, critic_loss_weight)
    if LoggingLevel.GRAD_NORM in logging_level:
        eng.enable_grad_norm()
    reset_key, new_key = prng.split(training_state.rng_key)              # ppo.py:271
    eng._upload_block(reset_key, new_key)
    if eng.r_iters == 0 or net.adam_step != opt.step or eng._rng_mirror != net.rng_count:
        net.adam_step = opt.step
        net.sync_counters_to_device()
    # the carry and the env state live in engine-owned tensors that the launch sequence updates in place
    c_in, h_in = net.get_carry(training_state.network_states)
    if c_in is not eng.r_carry[0]:
        eng.r_carry[0].copy_(c_in)
        eng.r_carry[1].copy_(h_in)
    _adopt_env_state(eng, training_state.env_states)
    # graph capture needs an env whose state is a dataclass of tensors stepped with device-only torch ops / kernels
    graphable = (dataclasses.is_dataclass(eng.r_env) and getattr(env, "graph_capturable", getattr(env, "fused_rollout", False))
                 and os.environ.get("B200PPO_GRAPH", "1") != "0")
    if not graphable or eng.r_iters == 0:
        final, eng.kernel_launches_per_iter = _enqueue_iteration(eng, net, env, eng.r_env)
        if dataclasses.is_dataclass(final) and dataclasses.is_dataclass(eng.r_env):
            _write_back_env(eng, final)
        else:
            eng.r_env = final
    else:
        if eng.r_graph is None:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                final, _ = _enqueue_iteration(eng, net, env, eng.r_env)
                _write_back_env(eng, final)
            eng.r_graph = g
        eng.r_graph.replay()
    eng.r_iters += 1
    eng.iters_run = eng.r_iters
    eng.graph = eng.r_graph
    network_states = net.set_carry(training_state.network_states, eng.r_carry)
    net.advance_rng(eng.rng_per_iter)
    eng._rng_mirror = net.rng_count
    opt.step += eng.n_updates
    net.adam_step = opt.step
    if LoggingLevel.CRITIC_EXTRA in logging_level:
        # rollout-time value estimates (logging only; the MLP critic does not see the carry)
        eng.value = policy_values(net, eng.obs.reshape(T * B, -1)).reshape(T, B)
    per_update = eng.metrics.cpu().numpy()
    total_steps = np.float32(training_state.steps_taken + np.float32(T * B))
    metrics = _ppo._iteration_metrics(per_update, eng, logging_level, logging_percentiles)
    metrics["total_steps"] = total_steps
    return training_state.replace(network_states=network_states, env_states=eng.r_env, rng_key=new_key,
                                  steps_taken=total_steps), metrics
