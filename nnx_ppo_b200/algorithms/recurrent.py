"""ppo_step for a recurrent (LSTM) actor — reference ppo.py:254-348 with the carry handling of
rollout.py:11-45 (reset on done) and ppo.py:409-431 (replay from the rollout's start carry).

First CUDA path of SURVEY section 8 row a15: the time loops live on the host (one launch of the
recurrent step kernel per time step, forward and backward), the critic / GAE / loss head /
gradient reduction / Adam are the MLP path's kernels driven stage by stage (networks/rplan.py).
Single GPU, no CUDA graph yet.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib, prng
from ..networks.plan import compile_network
from .rollout import policy_values, split_keys_device, tree_where
from .types import LoggingLevel


def _engine(net, opt, n_envs, T, E, M, lam, gamma, clip, norm_adv, cw):
    from .engine import PPOEngine
    import torch
    from .engine import cached_engine
    shape_key = (n_envs, T, E, M, bool(norm_adv), opt.gradient_clipping is not None, opt.wd_value >= 0.0)
    eng = cached_engine(net, "recurrent", None, opt, shape_key,
                        lambda: _build_engine(net, opt, n_envs, T, E, M, lam, gamma, clip, norm_adv, cw))
    eng.set_hparams(lam, gamma, clip, cw)
    return eng


def _build_engine(net, opt, n_envs, T, E, M, lam, gamma, clip, norm_adv, cw):
    from .engine import PPOEngine
    import torch
    eng = PPOEngine.__new__(PPOEngine)
    fake_env = type("E", (), {"fused_rollout": True})()
    PPOEngine.__init__(eng, net, fake_env, opt, n_envs, T, E, M, lam, gamma, clip, norm_adv, cw,
                       world_size=1, group=None, use_graph=False)
    lib, lp, mb, dev = eng.lib, net.lplan, eng.mb, net.device
    H, Y = lp.hidden, lp.out_dim
    f32 = dict(dtype=torch.float32, device=dev)
    eng.r_cache = torch.zeros(T, int(lib.b200ppo_lstm_cache_floats(lp, mb)), **f32)
    eng.r_c, eng.r_h = torch.zeros(mb, H, **f32), torch.zeros(mb, H, **f32)
    eng.r_dc, eng.r_dh = torch.zeros(mb, H, **f32), torch.zeros(mb, H, **f32)
    eng.r_y = torch.zeros(n_envs, Y, **f32)
    # GEMM operands of the deterministic weight-gradient pass, step-major [T][mb][width]
    P = lp.pre_dim
    eng.r_cat = torch.zeros(T, mb, P + H, **f32)
    eng.r_hn = torch.zeros(T, mb, H, **f32)
    eng.r_da = torch.zeros(T, mb, 4 * H, **f32)
    eng.r_dz = torch.zeros(T, mb, P, **f32)
    eng.r_scratch = torch.zeros(int(lib.b200ppo_lstm_wgrad_scratch_floats(lp, T * mb)), **f32)
    wsp = eng.ws.data_ptr()
    eng.r_y_ptr = int(C.cast(lib.b200ppo_update_debug_ptr(net.plan, T, mb, wsp, 2), C.c_void_p).value)
    eng.r_dy_ptr = int(C.cast(lib.b200ppo_update_debug_ptr(net.plan, T, mb, wsp, 3), C.c_void_p).value)
    eng.r_grad_ptr = int(C.cast(lib.b200ppo_update_grad_ptr(net.plan, T, mb, wsp), C.c_void_p).value)
    return eng


def ppo_step_recurrent(env, training_state, n_envs, rollout_length, gae_lambda, discounting_factor,
                       clip_range, normalize_advantages, n_epochs, n_minibatches, critic_loss_weight,
                       logging_level, logging_percentiles):
    import torch
    from . import ppo as _ppo
    net = compile_network(training_state.networks)
    if _ppo._dist_info()[0] != 1:
        raise NotImplementedError("the recurrent path is single-GPU for now")
    opt = training_state.optimizer
    T, B = rollout_length, n_envs
    eng = _engine(net, opt, B, T, n_epochs, n_minibatches, gae_lambda, discounting_factor, clip_range,
                  normalize_advantages, critic_loss_weight)
    if LoggingLevel.GRAD_NORM in logging_level:
        eng.enable_grad_norm()
    lib, lp, plan, mb, dev = eng.lib, net.lplan, net.plan, eng.mb, net.device
    A, Y, H = plan.act_dim, lp.out_dim, lp.hidden
    s = _lib.current_stream()
    reset_key, new_key = prng.split(training_state.rng_key)              # ppo.py:271
    net.adam_step = opt.step
    net.sync_counters_to_device()
    if net.normalizer is not None:
        net.normalizer.prepare()
    mean_p, std_p = net.norm_ptrs()
    arena = net.arena.data_ptr()

    # ---------------- rollout (rollout.py:48-73), one recurrent step launch per time step
    c, h = net.get_carry(training_state.network_states)
    c, h = c.contiguous(), h.contiguous()
    start_c, start_h = c.clone(), h.clone()                              # network_state the replay starts from
    env_state = training_state.env_states
    keys_all = split_keys_device(reset_key, T * B, dev).reshape(T, B, 2)
    for t in range(T):
        obs = env_state.obs.contiguous()
        _lib.check(lib.b200ppo_lstm_step_fwd(s, lp, arena, mean_p, std_p, obs.data_ptr(), 0, 0, B, c.data_ptr(),
                                             h.data_ptr(), eng.r_y.data_ptr(), 0), "lstm_step_fwd(rollout)")
        _lib.check(lib.b200ppo_sampler_step(s, eng.r_y.data_ptr(), B, A, 0, plan.min_std, plan.std_scale,
                                            plan.entropy_weight, net.counters.data_ptr(), 2 * t, 0,
                                            eng.raw_action[t].data_ptr(), eng.action[t].data_ptr(),
                                            eng.loglik[t].data_ptr(), 0), "sampler_step")
        nxt = env.step(env_state, eng.action[t])
        done = nxt.done.bool()
        tr = nxt.info.get("truncated", torch.zeros_like(done)) if isinstance(nxt.info, dict) else torch.zeros_like(done)
        eng.obs[t].copy_(obs)
        eng.reward[t].copy_(nxt.reward.float())
        eng.done[t].copy_(done.to(torch.uint8))
        eng.trunc[t].copy_(tr.bool().to(torch.uint8))
        if t == T - 1:
            eng.next_obs_last.copy_(nxt.obs)
        keep = (~done).to(torch.float32)[:, None]                        # reset_state -> zeros (rollout.py:33-40)
        c.mul_(keep)
        h.mul_(keep)
        env_state = tree_where(done, env.reset(keys_all[t].contiguous()), nxt)
    network_states = net.set_carry(training_state.network_states, (c, h))
    if LoggingLevel.CRITIC_EXTRA in logging_level:
        # rollout-time value estimates (logging only; the MLP critic does not see the carry)
        eng.value = policy_values(net, eng.obs.reshape(T * B, -1)).reshape(T, B)

    # ---------------- E x M minibatch updates (ppo.py:284-328)
    eng._upload_block(reset_key, new_key)
    _lib.check(lib.b200ppo_permutation(s, eng.iter_keys.data_ptr() + 8, B, eng.E, eng.inds.data_ptr(),
                                       eng.perm_scratch.data_ptr()), "permutation")
    inds_flat = eng.inds.view(-1)
    ST = _lib
    for u in range(eng.n_updates):
        off = 2 * T + u * 2 * (T + 1)
        args = (s, plan, eng.hp, eng.bufs[u], T, B, mb, off, u)
        ip = eng.inds.data_ptr() + 4 * u * mb
        _lib.check(lib.b200ppo_update(*args, ST.STAGE_FWD), "update/fwd")            # critic values (+ stand-in actor)
        idx = inds_flat[u * mb:(u + 1) * mb].long()
        eng.r_c.copy_(start_c[idx])
        eng.r_h.copy_(start_h[idx])
        for t in range(T):                                                            # replay scan, ppo.py:409-431
            _lib.check(lib.b200ppo_lstm_step_fwd(s, lp, arena, mean_p, std_p, eng.obs[t].data_ptr(), ip,
                                                 eng.done[t].data_ptr(), mb, eng.r_c.data_ptr(), eng.r_h.data_ptr(),
                                                 eng.r_y_ptr + 4 * t * mb * Y, eng.r_cache[t].data_ptr()),
                       "lstm_step_fwd(replay)")
        _lib.check(lib.b200ppo_update(*args, ST.STAGE_GAE | ST.STAGE_LOSS), "update/loss")
        _lib.check(lib.b200ppo_update(*args, ST.STAGE_BWD | ST.STAGE_RED), "update/bwd")
        eng.r_dc.zero_()
        eng.r_dh.zero_()
        for t in reversed(range(T)):                                                  # BPTT
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
        _lib.check(lib.b200ppo_update(*args, ST.STAGE_ADAM), "update/adam")

    # ---------------- Normalizer statistics, counters (ppo.py:329-346)
    if net.normalizer is not None:
        nz = net.normalizer
        _lib.check(lib.b200ppo_norm_batch_stats(s, eng.obs.data_ptr(), T * B, nz.size, eng.batch_stats.data_ptr(),
                                                eng.norm_scratch.data_ptr()), "norm_batch_stats")
        _lib.check(lib.b200ppo_norm_merge(s, eng.batch_stats.data_ptr(), 1, float(T * B), nz.size,
                                          nz.mean._dev.data_ptr(), nz.M2._dev.data_ptr(),
                                          nz.counter._dev.data_ptr()), "norm_merge")
    adv = 2 * T + eng.n_updates * 2 * (T + 1)
    _lib.check(lib.b200ppo_iter_finalize(s, net.counters.data_ptr(), adv, eng.n_updates, eng.comm_epoch.data_ptr()),
               "iter_finalize")
    net.advance_rng(adv)
    opt.step += eng.n_updates
    net.adam_step = opt.step
    per_update = eng.metrics.cpu().numpy()
    total_steps = np.float32(training_state.steps_taken + np.float32(T * B))
    metrics = _ppo._iteration_metrics(per_update, eng, logging_level, logging_percentiles)
    metrics["total_steps"] = total_steps
    return training_state.replace(network_states=network_states, env_states=env_state, rng_key=new_key,
                                  steps_taken=total_steps), metrics
