"""Rollout functions — mirrors nnx_ppo/algorithms/rollout.py (``unroll_env`` :48-73,
``eval_rollout`` :97-148, ``tree_where`` :270-279).

Two env kinds are supported:
* device envs with a fused rollout kernel (``env.fused_rollout``; envs/synthetic.py): the whole
  T-step rollout is ONE persistent kernel (csrc/rollout.cu);
* batched torch envs (``reset(keys[B,2]) -> state``, ``step(state, action[B,A]) -> state`` on CUDA
  tensors): one fused policy-step kernel (K1) per time step, env arithmetic in user code.
"""
from __future__ import annotations

from typing import Any, Optional

import numpy as np

from .. import _lib, prng
from ..networks.plan import call_network, compile_network
from ..networks.types import PPONetworkOutput
from ..networks.utils import tree_leaves
from .types import LoggingLevel, Transition


def tree_where(cond, on_true: Any, on_false: Any) -> Any:
    """rollout.py:270-279 on torch tensors / dicts / lists / dataclass-like env states."""
    import torch
    if isinstance(on_true, torch.Tensor):
        if on_true.dim() == 0 or on_true.shape[0] != cond.shape[0]:
            return on_true
        c = cond.reshape(cond.shape + (1,) * (on_true.dim() - cond.dim()))
        return torch.where(c, on_true, on_false)
    if isinstance(on_true, dict):
        return {k: tree_where(cond, on_true[k], on_false[k]) for k in on_true}
    if isinstance(on_true, (list, tuple)):
        return type(on_true)(tree_where(cond, a, b) for a, b in zip(on_true, on_false))
    if hasattr(on_true, "__dataclass_fields__"):
        import dataclasses
        return dataclasses.replace(on_true, **{f: tree_where(cond, getattr(on_true, f), getattr(on_false, f))
                                               for f in on_true.__dataclass_fields__})
    return on_true


def _keys_tensor(keys: list, device):
    import torch
    a = np.array(keys, np.uint32).reshape(-1, 2).view(np.int32)
    return torch.from_numpy(a.copy()).to(device)


def split_keys_device(key, n: int, device):
    """jax.random.split(key, n) materialised on the device as an int32 [n, 2] tensor."""
    import torch
    lib = _lib.load()
    out = torch.empty(n, 2, dtype=torch.int32, device=device)
    _lib.check(lib.b200ppo_synth_init_keys(_lib.current_stream(), key[0], key[1], n, _lib.ptr(out)), "split")
    return out


def reset_envs(env, key, n_envs: int, device):
    return env.reset(split_keys_device(key, n_envs, device))


def unroll_env(env, env_state, networks, network_state, unroll_length: int, rng_key_for_env_reset):
    """rollout.py:48-73 -> (final_network_state, final_env_state, Transition[T, B, ...])."""
    import torch
    net = compile_network(networks)
    lib = _lib.load()
    T = unroll_length
    O, A = net.plan.obs_dim, net.plan.act_dim
    dev = net.device
    if getattr(env, "fused_rollout", False):
        B = env_state.obs.shape[0]
        f32 = dict(dtype=torch.float32, device=dev)
        obs, raw, act = torch.empty(T, B, O, **f32), torch.empty(T, B, A, **f32), torch.empty(T, B, A, **f32)
        ll, rew = torch.empty(T, B, **f32), torch.empty(T, B, **f32)
        done = torch.empty(T, B, dtype=torch.uint8, device=dev)
        trunc = torch.empty(T, B, dtype=torch.uint8, device=dev)
        nol = torch.empty(B, O, **f32)
        keys = _keys_tensor([rng_key_for_env_reset, (0, 0)], dev).reshape(-1)
        s = _lib.current_stream()
        if net.normalizer is not None:
            net.normalizer.prepare(s)
        net.sync_counters_to_device()
        mean_p, std_p = net.norm_ptrs()
        new_env = type(env_state)(env_state.obs.clone(), env_state.step_counter.clone(),
                                  env_state.term_state.clone(), env_state.reward, env_state.done,
                                  dict(env_state.info), dict(env_state.metrics))
        ws_p, ws_n = net.rollout_workspace(B)
        _lib.check(lib.b200ppo_rollout_synth_ws(
            s, net.plan, env.c_struct(dev), net.arena.data_ptr(), mean_p, std_p, net.counters.data_ptr(),
            keys.data_ptr(), T, B, new_env.obs.data_ptr(), new_env.step_counter.data_ptr(),
            new_env.term_state.data_ptr(), obs.data_ptr(), raw.data_ptr(), act.data_ptr(), ll.data_ptr(),
            rew.data_ptr(), done.data_ptr(), trunc.data_ptr(), nol.data_ptr(), ws_p, ws_n), "rollout_synth")
        net.advance_rng(2 * T)
        net.sync_counters_to_device()
        # value estimates of the rollout (metrics only, ppo.py never trains on them): K1 in
        # value-only use over the T*B recorded observations
        values = policy_values(net, obs.reshape(T * B, O)).reshape(T, B)
        extras = _extras(net, obs, raw)
        tr = Transition(obs=obs, network_output=PPONetworkOutput(act, ll, values), rewards=rew,
                        done=done.bool(), truncated=trunc.bool(), next_obs=nol, metrics={},
                        rollout_extras=extras)
        return network_state, new_env, tr
    return _unroll_generic(env, env_state, networks, network_state, T, rng_key_for_env_reset)


def _stack_tree(steps: list, dev):
    """Stack a list of per-step metric pytrees (nested dicts of [B] tensors / scalars) along time."""
    import torch
    first = steps[0]
    if isinstance(first, dict):
        return {k: _stack_tree([m[k] for m in steps], dev) for k in first if all(k in m for m in steps)}
    return torch.stack([torch.as_tensor(m, device=dev) for m in steps])


def _extras(net, obs, raw):
    na, nc = len(net.actor_layers), len(net.critic_layers)
    ad = {"action": [None] * na + [raw], "value": [None] * nc}
    return net.wrap(obs, ad, None) if not net.recurrent else ([obs, ad] if net.normalizer is not None else ad)


def policy_values(net, obs_flat):
    """Critic values for [N, O] observations via the policy-step kernel in replay mode (no RNG is
    consumed: the stored-action path ignores the sample key and reg_loss is not requested)."""
    import torch
    lib = _lib.load()
    N = obs_flat.shape[0]
    A = net.plan.act_dim
    dev = obs_flat.device
    dummy = torch.zeros(N, A, device=dev)
    raw, act = torch.empty(N, A, device=dev), torch.empty(N, A, device=dev)
    ll, val = torch.empty(N, device=dev), torch.empty(N, device=dev)
    mean_p, std_p = net.norm_ptrs()
    _lib.check(lib.b200ppo_policy_step(_lib.current_stream(), net.plan, net.arena.data_ptr(), mean_p, std_p,
                                       obs_flat.contiguous().data_ptr(), N, 1, net.counters.data_ptr(), 0,
                                       dummy.data_ptr(), raw.data_ptr(), act.data_ptr(), ll.data_ptr(),
                                       val.data_ptr(), 0, 0), "policy_step(values)")
    return val


def _unroll_generic(env, env_state, networks, network_state, T: int, reset_key):
    """single_transition (rollout.py:11-45) per step for batched torch envs."""
    import torch
    net = compile_network(networks)
    dev = net.device
    B = tree_leaves(env_state.obs)[0].shape[0]
    # split(reset_key, (T, B)): element t*B + b
    keys_all = split_keys_device(reset_key, T * B, dev).reshape(T, B, 2)
    rec = {k: [] for k in ("obs", "raw", "act", "ll", "val", "rew", "done", "trunc")}
    env_metrics: list = []
    net_metrics: list = []
    next_obs = None
    for t in range(T):
        out = call_network(networks, network_state, env_state.obs)
        po = out.output
        nxt = env.step(env_state, po.actions)
        done = nxt.done.bool()
        if net.recurrent:                                                 # rollout.py:33-40: reset the carry on done
            network_state = net.set_carry(out.next_state, net.reset_carry(net.get_carry(out.next_state), done))
        tr = nxt.info.get("truncated", torch.zeros_like(done)) if isinstance(nxt.info, dict) else torch.zeros_like(done)
        rec["obs"].append(net.flat_obs(env_state.obs))               # what the kernels saw (adapters applied)
        extras = net.adapter_extras(out.rollout_extras)
        rec["raw"].append(extras["action"][-1]); rec["act"].append(po.actions)
        rec["ll"].append(po.loglikelihoods); rec["val"].append(po.value_estimates)
        rec["rew"].append(nxt.reward.float()); rec["done"].append(done); rec["trunc"].append(tr.bool())
        env_metrics.append(nxt.metrics if isinstance(getattr(nxt, "metrics", None), dict) else {})
        net_metrics.append(out.metrics if isinstance(out.metrics, dict) else {})
        next_obs = net.flat_obs(nxt.obs) if t == T - 1 else None
        reset_states = env.reset(keys_all[t].contiguous())
        env_state = tree_where(done, reset_states, nxt)
    st = {k: torch.stack(v) for k, v in rec.items()}
    # Transition.metrics = {"env": ..., "net": ...} (rollout.py:31-34), stacked over time
    stacked_env = _stack_tree(env_metrics, dev) if env_metrics else {}
    stacked_net = _stack_tree(net_metrics, dev) if net_metrics else {}
    tr = Transition(obs=st["obs"], network_output=PPONetworkOutput(st["act"], st["ll"], st["val"]),
                    rewards=st["rew"], done=st["done"], truncated=st["trunc"], next_obs=next_obs,
                    metrics={"env": stacked_env, "net": stacked_net},
                    rollout_extras=_extras(net, st["obs"], st["raw"]))
    return network_state, env_state, tr


def ppo_step_generic(env, training_state, n_envs, rollout_length, gae_lambda, discounting_factor,
                     clip_range, normalize_advantages, n_epochs, n_minibatches, critic_loss_weight,
                     logging_level, logging_percentiles):
    """ppo_step for batched torch envs: per-step K1 launches, then the same update sequence."""
    import torch
    from . import ppo as _ppo
    from .engine import PPOEngine
    net = compile_network(training_state.networks)
    opt = training_state.optimizer
    world, group = _ppo._dist_info()

    def build():
        eng = PPOEngine.__new__(PPOEngine)
        _FakeEnv = type("E", (), {"fused_rollout": True})
        PPOEngine.__init__(eng, net, _FakeEnv(), opt, n_envs, rollout_length, n_epochs, n_minibatches,
                           gae_lambda, discounting_factor, clip_range, normalize_advantages,
                           critic_loss_weight, world_size=world, group=group, use_graph=False)
        return eng

    from .engine import cached_engine
    shape_key = (n_envs, rollout_length, n_epochs, n_minibatches, bool(normalize_advantages),
                 opt.gradient_clipping is not None, opt.wd_value >= 0.0, world)
    eng = cached_engine(net, "generic", env, opt, shape_key, build)
    eng.set_hparams(gae_lambda, discounting_factor, clip_range, critic_loss_weight)
    if _ppo.LoggingLevel.CRITIC_EXTRA in logging_level and logging_percentiles:
        eng.enable_adv_log()
    if _ppo.LoggingLevel.GRAD_NORM in logging_level:
        eng.enable_grad_norm()
    reset_key, new_key = prng.split(training_state.rng_key)
    _, next_env_state, tr = _unroll_generic(env, training_state.env_states, training_state.networks,
                                            training_state.network_states, rollout_length, reset_key)
    eng.obs.copy_(tr.obs); eng.raw_action.copy_(net.adapter_extras(tr.rollout_extras)["action"][-1])
    eng.loglik.copy_(tr.network_output.loglikelihoods); eng.reward.copy_(tr.rewards)
    eng.action.copy_(tr.network_output.actions)
    eng.value = tr.network_output.value_estimates                         # logging only
    eng.env_metrics = tr.metrics
    eng.done.copy_(tr.done.to(torch.uint8)); eng.trunc.copy_(tr.truncated.to(torch.uint8))
    eng.next_obs_last.copy_(tr.next_obs)
    eng._upload_block(reset_key, new_key)
    net.adam_step = opt.step
    net.sync_counters_to_device()
    if net.normalizer is not None:
        net.normalizer.prepare()
    adv = eng.n_updates * 2 * (rollout_length + 1)
    eng._enqueue_updates(0, adv)
    net.advance_rng(adv)
    opt.step += eng.n_updates
    net.adam_step = opt.step
    per_update = eng.metrics.cpu().numpy()
    total_steps = np.float32(training_state.steps_taken + np.float32(rollout_length * n_envs))
    metrics = _ppo._iteration_metrics(per_update, eng, logging_level, logging_percentiles)
    metrics["total_steps"] = total_steps
    return training_state.replace(env_states=next_env_state, rng_key=new_key, steps_taken=total_steps), metrics


def eval_rollout(env, networks, n_envs: int, max_episode_length: int, key,
                 logging_percentiles: Optional[tuple[int, ...]] = None) -> dict[str, Any]:
    """rollout.py:97-148: sticky done, reward accumulated only while alive, lifespan counter."""
    import torch
    net = compile_network(networks)
    dev = net.device
    env_state = env.reset(split_keys_device(key, n_envs, dev))
    if getattr(env, "fused_rollout", False) and not net.recurrent:
        cuml, lifespan = _eval_fused(env, net, env_state, n_envs, max_episode_length)
        return _eval_metrics(cuml, lifespan, logging_percentiles)
    net_state = networks.initialize_state(n_envs)
    cuml = torch.zeros(n_envs, device=dev)
    lifespan = torch.zeros(n_envs, device=dev)
    prev_done = env_state.done.bool() if env_state.done is not None else torch.zeros(n_envs, dtype=torch.bool, device=dev)
    for _ in range(max_episode_length):
        out = call_network(networks, net_state, env_state.obs)
        if net.recurrent:
            net_state = out.next_state                                    # rollout.py:111-113 carries the state
        nxt = env.step(env_state, out.output.actions)
        done = torch.logical_or(nxt.done.bool(), prev_done)               # rollout.py:115-117
        cuml = cuml + torch.where(prev_done, torch.zeros_like(cuml), nxt.reward.float())
        lifespan = lifespan + torch.where(done, 0.0, 1.0)
        prev_done = done
        env_state = nxt
    return _eval_metrics(cuml, lifespan, logging_percentiles)


def _eval_fused(env, net, env_state, n_envs: int, L: int):
    """The whole evaluation episode in ONE launch of the persistent eval kernel (csrc/rollout.cu):
    policy step + env step + the sticky-done bookkeeping for all L steps, env tile resident in
    shared memory.  Consumes the sampler counts the per-step path would (sampling_layers.py:93-96)."""
    import torch
    lib = _lib.load()
    dev = net.device
    s = _lib.current_stream()
    if net.normalizer is not None:
        net.normalizer.prepare(s)
    net.sync_counters_to_device()
    mean_p, std_p = net.norm_ptrs()
    deterministic = bool(net.sampler.deterministic)
    cuml = torch.empty(n_envs, dtype=torch.float32, device=dev)
    lifespan = torch.empty(n_envs, dtype=torch.float32, device=dev)
    _lib.check(lib.b200ppo_eval_synth(
        s, net.plan, env.c_struct(dev), net.arena.data_ptr(), mean_p, std_p, net.counters.data_ptr(),
        2 if deterministic else 0, L, n_envs, env_state.obs.data_ptr(), env_state.step_counter.data_ptr(),
        env_state.term_state.data_ptr(), cuml.data_ptr(), lifespan.data_ptr()), "eval_synth")
    net.advance_rng(L * (1 if deterministic else 2))
    net.sync_counters_to_device()
    return cuml, lifespan


def _eval_metrics(cuml, lifespan, logging_percentiles) -> dict[str, Any]:
    """rollout.py:141-148 / _add_reward_metrics :76-94."""
    import torch
    dev = cuml.device
    metrics = {"lifespan_mean": float(lifespan.mean()), "lifespan_std": float(lifespan.std(unbiased=False))}
    if logging_percentiles is not None:
        q = torch.tensor([p / 100.0 for p in logging_percentiles], device=dev)
        for pl, p in zip(logging_percentiles, torch.quantile(cuml, q).tolist()):
            metrics[f"episode_reward/p{int(pl)}"] = p
        for pl, p in zip(logging_percentiles, torch.quantile(lifespan, q).tolist()):
            metrics[f"lifespan/p{int(pl)}"] = p
    else:
        metrics["episode_reward/mean"] = float(cuml.mean())
        metrics["episode_reward/std"] = float(cuml.std(unbiased=False))
    return metrics
