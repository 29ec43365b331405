"""PPO driver — same public API as nnx_ppo/algorithms/ppo.py (``train_ppo`` :41, ``ppo_step``
:254, ``gae`` :351, ``new_training_state`` :534, ``default_config`` :29, ``_should_run`` :34),
with the jitted device program replaced by the CUDA engine (algorithms/engine.py).
"""
from __future__ import annotations

import dataclasses
import time
from collections.abc import Callable
from typing import Any, Optional

import os

import numpy as np

from .. import _lib, parallel, prng
from ..networks.plan import compile_network
from ..networks.types import StatefulModule
from . import rollout
from .config import EvalConfig, PPOConfig, TrainConfig, TrainResult, VideoConfig, VideoData  # noqa: F401
from .engine import AdamOptimizer, PPOEngine, cached_engine
from .types import LoggingLevel, RLEnv, TrainingState


def default_config() -> TrainConfig:
    return TrainConfig()


def _should_run(steps: int, last_step: int, every_steps: int) -> bool:
    if every_steps <= 0:
        return False
    return (steps // every_steps) > (last_step // every_steps)


def _dist_info():
    """(world_size, process_group) when torch.distributed is initialised, else (1, None)."""
    return parallel.dist_info()[0], None


def new_training_state(env: RLEnv, networks: StatefulModule, n_envs: int, seed: int,
                       learning_rate: float = 1e-4, gradient_clipping: Optional[float] = None,
                       weight_decay: Optional[float] = None) -> TrainingState:
    """ppo.py:534-572.  Under torch.distributed every rank is an independent replica of this
    function on its own env shard: rank r folds r into the seed-derived keys (rank 0 is
    bit-identical to a single-process run)."""
    net = compile_network(networks)
    key, training_key = parallel.rank_keys(seed, parallel.dist_info()[1])   # ppo.py:544-545
    if getattr(env, "fused_rollout", False):
        env_states = env.reset_from_split(key, n_envs, net.device)       # ppo.py:548-549
    else:
        env_states = rollout.reset_envs(env, key, n_envs, net.device)
    network_states = networks.initialize_state(n_envs)                   # ppo.py:552
    optimizer = AdamOptimizer(net, learning_rate, gradient_clipping, weight_decay)
    return TrainingState(networks, network_states, env_states, optimizer, training_key,
                         np.float32(0.0))


def _engine_for(env, training_state: TrainingState, n_envs, rollout_length, gae_lambda,
                discounting_factor, clip_range, normalize_advantages, n_epochs, n_minibatches,
                critic_loss_weight) -> PPOEngine:
    """The engine for these SHAPES (cached per network, LRU); gamma / lambda / clip range / critic
    weight / learning rate are run-time data of the captured iteration, refreshed here on every call
    (the reference traces gae_lambda / discounting_factor, ppo.py:105)."""
    net = compile_network(training_state.networks)
    opt: AdamOptimizer = training_state.optimizer
    world, group = _dist_info()
    shape_key = (n_envs, rollout_length, n_epochs, n_minibatches, bool(normalize_advantages),
                 opt.gradient_clipping is not None, opt.wd_value >= 0.0, world)
    eng = cached_engine(net, "fused", env, opt, shape_key, lambda: PPOEngine(
        net, env, opt, n_envs, rollout_length, n_epochs, n_minibatches, gae_lambda, discounting_factor,
        clip_range, normalize_advantages, critic_loss_weight, world_size=world, group=group))
    eng.set_hparams(gae_lambda, discounting_factor, clip_range, critic_loss_weight)
    return eng


def ppo_step(env: RLEnv, training_state: TrainingState, n_envs: int, rollout_length: int,
             gae_lambda, discounting_factor, clip_range, normalize_advantages: bool,
             combine_advantages: bool, n_epochs: int, n_minibatches: int,
             critic_loss_weight=1.0, logging_level: LoggingLevel = LoggingLevel.LOSSES,
             logging_percentiles: Optional[tuple[int, ...]] = None
             ) -> tuple[TrainingState, dict[str, Any]]:
    """One PPO iteration (ppo.py:254-348).  The input state is consumed: env state, parameters,
    optimizer moments and Normalizer statistics are updated in place on the device."""
    if combine_advantages:
        raise NotImplementedError("combine_advantages needs dict rewards; single scalar reward only")
    if compile_network(training_state.networks).recurrent:
        from . import recurrent as _rec
        return _rec.ppo_step_recurrent(env, training_state, n_envs, rollout_length, gae_lambda,
                                       discounting_factor, clip_range, normalize_advantages, n_epochs,
                                       n_minibatches, critic_loss_weight, logging_level, logging_percentiles)
    if not getattr(env, "fused_rollout", False):
        return rollout.ppo_step_generic(env, training_state, n_envs, rollout_length, gae_lambda,
                                        discounting_factor, clip_range, normalize_advantages,
                                        n_epochs, n_minibatches, critic_loss_weight, logging_level,
                                        logging_percentiles)
    eng = _engine_for(env, training_state, n_envs, rollout_length, gae_lambda, discounting_factor,
                      clip_range, normalize_advantages, n_epochs, n_minibatches, critic_loss_weight)
    if LoggingLevel.CRITIC_EXTRA in logging_level:
        eng.enable_values()
        if logging_percentiles:
            eng.enable_adv_log()
    if LoggingLevel.TRAINING_ENV_METRICS in logging_level:
        eng.enable_values(net_metrics=True)
    if LoggingLevel.GRAD_NORM in logging_level:
        eng.enable_grad_norm()
    reset_key, new_key = prng.split(training_state.rng_key)              # ppo.py:271
    total_steps = np.float32(training_state.steps_taken + np.float32(rollout_length * n_envs))
    if not (logging_level & _EAGER_LEVELS) and _LAZY_METRICS:
        # only the per-update loss rows are logged: return without waiting for the device.  The dict
        # materialises (event wait + mean / std over the updates) when it is first read - the reference's
        # jitted ppo_step returns asynchronously dispatched arrays in the same way.
        pending = eng.step(training_state.env_states, reset_key, new_key, fetch_metrics="lazy")

        def build():
            m = _iteration_metrics(pending.wait(), eng, logging_level, logging_percentiles)
            m["total_steps"] = total_steps                               # ppo.py:333
            return m
        metrics = LazyMetrics(build)
    else:
        per_update = eng.step(training_state.env_states, reset_key, new_key, fetch_metrics=True)
        metrics = _iteration_metrics(per_update, eng, logging_level, logging_percentiles)
        metrics["total_steps"] = total_steps                             # ppo.py:333
    # the engine advances ITS env-state tensors in place: hand those back, so that the caller's next
    # TrainingState points at the live state whatever object came in (ppo.py:341-346)
    new_state = training_state.replace(env_states=eng.env_state, rng_key=new_key, steps_taken=total_steps)
    return new_state, metrics


# levels whose metrics read the rollout buffers / parameters (which the next iteration overwrites): computed eagerly
_EAGER_LEVELS = (LoggingLevel.CRITIC_EXTRA | LoggingLevel.ACTOR_EXTRA | LoggingLevel.TRAIN_ROLLOUT_STATS
                 | LoggingLevel.TRAINING_ENV_METRICS | LoggingLevel.WEIGHTS)
_LAZY_METRICS = os.environ.get("B200PPO_LAZY_METRICS", "1") != "0"


class LazyMetrics(dict):
    """The metrics dict of one iteration, filled in on first access (any read forces it)."""

    def __init__(self, build):
        super().__init__()
        self._build = build

    def _force(self):
        b = self._build
        if b is not None:
            self._build = None
            dict.update(self, b())
        return self

    def __getitem__(self, k):
        return dict.__getitem__(self._force(), k)

    def __iter__(self):
        return dict.__iter__(self._force())

    def __len__(self):
        return dict.__len__(self._force())

    def __contains__(self, k):
        return dict.__contains__(self._force(), k)

    def __repr__(self):
        return dict.__repr__(self._force())

    def __eq__(self, other):
        return dict.__eq__(self._force(), other)

    __hash__ = None

    def __setitem__(self, k, v):
        dict.__setitem__(self._force(), k, v)

    def __delitem__(self, k):
        dict.__delitem__(self._force(), k)

    def get(self, k, default=None):
        return dict.get(self._force(), k, default)

    def keys(self):
        return dict.keys(self._force())

    def values(self):
        return dict.values(self._force())

    def items(self):
        return dict.items(self._force())

    def copy(self):
        return dict(self._force())

    def update(self, *a, **kw):
        dict.update(self._force(), *a, **kw)

    def pop(self, *a):
        return dict.pop(self._force(), *a)

    def setdefault(self, k, default=None):
        return dict.setdefault(self._force(), k, default)

    def __or__(self, other):
        return dict(self._force()) | other

    def __ror__(self, other):
        return other | dict(self._force())

    def __reduce__(self):
        return (dict, (dict(self._force()),))


def _iteration_metrics(per_update: np.ndarray, eng, logging_level, percentiles) -> dict[str, Any]:
    """compute_metrics + the grad-norm / weight extras of ppo_step (ppo.py:313-315,329-335) from the
    per-update metric rows and the engine's rollout buffers; shared by the fused, per-step and
    recurrent paths."""
    metrics = _loss_metrics(per_update, logging_level, percentiles, getattr(eng, "adv_log", None))
    if LoggingLevel.GRAD_NORM in logging_level and eng.hp.grad_clip > 0.0:
        # loss_metrics["grad_norm"] is one scalar per update and goes through _log_metric like every
        # other loss metric (ppo.py:313-315, metrics.py:36-37): grad_norm/mean, /std or /pN
        _log_np(metrics, "grad_norm", per_update[:, 3], percentiles)
    _extra_metrics(metrics, eng.net, eng, logging_level, percentiles)
    return metrics


def _log_np(m: dict, name: str, x: np.ndarray, percentiles) -> None:
    """metrics.py:72-100 on a host array."""
    if percentiles:
        for pl, p in zip(percentiles, np.percentile(x, percentiles)):
            m[f"{name}/p{int(pl)}"] = np.float32(p)
    else:
        m[f"{name}/mean"] = x.mean(dtype=np.float32)
        m[f"{name}/std"] = x.std(dtype=np.float32)


def _log_metric(m: dict, name: str, x, percentiles) -> None:
    """metrics.py:72-100 on a CUDA tensor: bool -> fraction true; else mean / std or percentiles.
    (Logging only: a handful of reductions over buffers the kernels already wrote.)"""
    import torch
    if x.dtype in (torch.bool, torch.uint8):
        m[name] = np.float32(x.float().mean().item())
    elif not percentiles:
        x = x.float()
        m[f"{name}/mean"] = np.float32(x.mean().item())
        m[f"{name}/std"] = np.float32(x.std(unbiased=False).item())
    else:
        q = torch.tensor([p / 100.0 for p in percentiles], device=x.device)
        xs = x.float().reshape(-1)
        if xs.numel() > 1_000_000:                       # torch.quantile's input limit: subsample evenly
            xs = xs[:: xs.numel() // 1_000_000 + 1]
        for pl, v in zip(percentiles, torch.quantile(xs, q).tolist()):
            m[f"{name}/p{int(pl)}"] = np.float32(v)


def _log_tree(m: dict, name: str, x, percentiles) -> None:
    """metrics.py:87-90: nested mappings are logged leaf by leaf under ``name/key``."""
    if isinstance(x, dict):
        for k, v in x.items():
            _log_tree(m, f"{name}/{k}", v, percentiles)
    else:
        _log_metric(m, name, x, percentiles)


def _extra_metrics(m: dict, net, eng, logging_level, percentiles) -> None:
    """The parts of metrics.compute_metrics / log_weight_stats (metrics.py:17-121) that only need the
    rollout buffers and the parameter arena.  ROLLOUT_OBS logs nothing in the reference either
    (metrics.py:55-56)."""
    if LoggingLevel.TRAINING_ENV_METRICS in logging_level:                # metrics.py:38-40
        for k, v in getattr(eng, "env_metrics", {}).items():
            _log_tree(m, k, v, percentiles)
        ms = getattr(eng, "mu_sigma", None)                               # Transition.metrics["net"], rollout.py:31-34
        if ms is not None:
            A = ms.shape[-1] // 2
            base = net.sampler_metric_path()
            _log_metric(m, f"{base}/mu", ms[..., :A], percentiles)
            _log_metric(m, f"{base}/sigma", ms[..., A:], percentiles)
    if LoggingLevel.TRAIN_ROLLOUT_STATS in logging_level:
        _log_metric(m, "rollout_batch/reward", eng.reward, percentiles)
        _log_metric(m, "rollout_batch/action", eng.action, percentiles)
        _log_metric(m, "rollout_batch/done_rate", eng.done, percentiles)
        _log_metric(m, "rollout_batch/truncation_rate", eng.trunc, percentiles)
    if LoggingLevel.ACTOR_EXTRA in logging_level:
        _log_metric(m, "loglikelihood", eng.loglik, percentiles)
    if LoggingLevel.CRITIC_EXTRA in logging_level and getattr(eng, "value", None) is not None:
        _log_metric(m, "losses/predicted_value", eng.value, percentiles)  # metrics.py:62-68
    if LoggingLevel.WEIGHTS in logging_level:
        # log_weight_stats (metrics.py:103-121): every nnx.Param leaf, i.e. the arena without its
        # alignment padding and without the structural zeros of block-diagonal encoder layers
        if hasattr(net, "logical_index_dev"):
            w = net.arena[net.logical_index_dev()]
        else:
            w = net.arena if getattr(net, "param_mask", None) is None else net.arena[net.param_mask == 1]
        _log_metric(m, "weights", w, percentiles)


def _loss_metrics(per_update: np.ndarray, logging_level, percentiles, adv_log=None) -> dict[str, Any]:
    """metrics.py:17-100 for the ``losses/*`` keys (mean / std or percentiles over the updates)."""
    m: dict[str, Any] = {}
    if LoggingLevel.LOSSES in logging_level:
        for i, name in enumerate(("losses/actor", "losses/critic", "losses/regularization")):
            _log_np(m, name, per_update[:, i], percentiles)
    if per_update.shape[1] > 8:
        if LoggingLevel.ACTOR_EXTRA in logging_level:                      # ppo.py:514-520
            _log_np(m, "losses/clipping_fraction", per_update[:, 4], percentiles)
        if LoggingLevel.CRITIC_EXTRA in logging_level:                     # ppo.py:522-527
            var_t = np.maximum(per_update[:, 6].astype(np.float64) - per_update[:, 5].astype(np.float64) ** 2, 0.0)
            _log_np(m, "losses/critic_R^2", (1.0 - 2.0 * per_update[:, 1] / (var_t + 1e-8)).astype(np.float32), percentiles)
            # losses/advantages is the [updates, T, mb] array of the advantages the surrogate used, i.e. the
            # NORMALISED ones when normalize_advantages (ppo.py:477-480 reassigns the name before :523).
            # Columns 7 / 8 hold E[a] / E[a^2] of exactly that tensor per update.
            if percentiles and adv_log is not None:
                import torch
                mean = torch.from_numpy(per_update[:, 9].copy()).to(adv_log.device)[:, None]
                den = torch.from_numpy(per_update[:, 10].copy()).to(adv_log.device)[:, None]
                _log_metric(m, "losses/advantages", (adv_log - mean) / den, percentiles)
            elif not percentiles:
                mean = per_update[:, 7].astype(np.float64).mean()
                m["losses/advantages/mean"] = np.float32(mean)
                m["losses/advantages/std"] = np.float32(np.sqrt(max(per_update[:, 8].astype(np.float64).mean() - mean * mean, 0.0)))
    return m


def gae(rewards, values_excl_last, last_value, done, truncation, lambda_, gamma):
    """ppo.py:351-394 on CUDA tensors: [T, B] float32 rewards / values, [B] last_value,
    [T, B] bool done / truncation -> [T, B] advantages (K2, csrc/misc.cu)."""
    import torch
    lib = _lib.load()
    T, B = rewards.shape
    r = rewards.contiguous().float()
    v = values_excl_last.contiguous().float()
    lv = last_value.contiguous().float()
    d = done.to(torch.uint8).contiguous()
    tr = truncation.to(torch.uint8).contiguous()
    _lib.require_cuda(r, v, lv, d, tr)
    out = torch.empty_like(r)
    _lib.check(lib.b200ppo_gae(_lib.current_stream(), _lib.ptr(r), _lib.ptr(v), _lib.ptr(lv), _lib.ptr(d),
                               _lib.ptr(tr), T, B, float(lambda_), float(gamma), _lib.ptr(out)), "gae")
    return out


def ppo_loss(networks: StatefulModule, network_state: Any, rollout_data, clip_range, normalize_advantages: bool,
             combine_advantages: bool, discounting_factor, gae_lambda, critic_loss_weight,
             logging_level: LoggingLevel, *, return_grads: bool = False):
    """ppo.py:397-531 as a callable: ``(total_loss, loss_metrics)`` of one minibatch ``Transition``
    ([T, mb, ...] leaves, e.g. ``jax.tree.map(lambda x: x[:, inds], rollout_data)`` of ppo.py:297) under
    the CURRENT parameters — stages FWD | GAE | LOSS of the update kernels (K3) on the rows in their
    given order.  Like the reference call it advances the sampler stream by 2 * (T + 1) counts (one
    network call per replayed step plus the bootstrap call, two draws each).  The reference
    differentiates this function with ``nnx.grad``; here the analytic backward is part of the same
    kernels: ``return_grads=True`` also runs BWD | RED and returns the flat gradient in the parameter
    order of ``CompiledNet.params_logical`` as a third element."""
    import torch
    if combine_advantages:
        raise NotImplementedError("combine_advantages needs dict rewards; single scalar reward only")
    net = compile_network(networks)
    if net.recurrent:
        raise NotImplementedError("ppo_loss as a standalone callable supports the MLP plans; recurrent networks "
                                  "are differentiated inside ppo_step")
    lib = _lib.load()
    T, mb = rollout_data.rewards.shape

    def build():
        eng = PPOEngine.__new__(PPOEngine)
        fake_env = type("E", (), {"fused_rollout": True})()
        PPOEngine.__init__(eng, net, fake_env, AdamOptimizer(net), mb, T, 1, 1, gae_lambda, discounting_factor,
                           clip_range, normalize_advantages, critic_loss_weight, world_size=1, group=None,
                           use_graph=False)
        eng.inds.copy_(torch.arange(mb, dtype=torch.int32, device=net.device).reshape(1, mb))
        return eng

    eng = cached_engine(net, "loss", None, None, (T, mb, bool(normalize_advantages)), build)
    eng.set_hparams(gae_lambda, discounting_factor, clip_range, critic_loss_weight)
    eng._upload_block((0, 0), (0, 0))
    obs = net.flat_obs(rollout_data.obs)
    eng.obs.copy_(obs.reshape(T, mb, -1))
    eng.raw_action.copy_(net.adapter_extras(rollout_data.rollout_extras)["action"][-1].reshape(T, mb, -1))
    eng.loglik.copy_(rollout_data.network_output.loglikelihoods)
    eng.reward.copy_(rollout_data.rewards)
    eng.done.copy_(rollout_data.done.to(torch.uint8))
    eng.trunc.copy_(rollout_data.truncated.to(torch.uint8))
    nxt = net.flat_obs(rollout_data.next_obs)
    eng.next_obs_last.copy_(nxt[-1] if nxt.dim() == 3 else nxt)           # ppo.py:433: only next_obs[-1] is read
    s = _lib.current_stream()
    if net.normalizer is not None:
        net.normalizer.prepare(s)
    net.sync_counters_to_device()
    stages = _lib.STAGE_FWD | _lib.STAGE_GAE | _lib.STAGE_LOSS
    if return_grads:
        stages |= _lib.STAGE_BWD | _lib.STAGE_RED
    _lib.check(lib.b200ppo_update(s, net.plan, eng.hp, eng.bufs[0], T, mb, mb, 0, 0, stages), "ppo_loss")
    net.advance_rng(2 * (T + 1))
    net.sync_counters_to_device()
    row = eng.metrics[0].cpu().numpy()
    total = np.float32(row[0] + np.float32(critic_loss_weight) * row[1] + row[2])   # ppo.py:529
    loss_metrics: dict[str, Any] = {}
    if LoggingLevel.LOSSES in logging_level:                                        # ppo.py:510-513
        loss_metrics["losses/actor"] = np.float32(row[0])
        loss_metrics["losses/critic"] = np.float32(row[1])
        loss_metrics["losses/regularization"] = np.float32(row[2])
    if LoggingLevel.ACTOR_EXTRA in logging_level:                                   # ppo.py:514-520
        loss_metrics["losses/clipping_fraction"] = np.float32(row[4])
    if LoggingLevel.CRITIC_EXTRA in logging_level:                                  # ppo.py:522-527
        wsp = eng.ws.data_ptr()
        o = (int(lib.b200ppo_update_debug_ptr(net.plan, T, mb, wsp, 0)) - wsp) // 4
        adv = eng.ws[o:o + T * mb].reshape(T, mb)
        loss_metrics["losses/advantages"] = ((adv - float(row[9])) / float(row[10])).clone()
        var_t = max(float(row[6]) - float(row[5]) ** 2, 0.0)
        loss_metrics["losses/critic_R^2"] = np.float32(1.0 - 2.0 * float(row[1]) / (var_t + 1e-8))
    if return_grads:
        return total, loss_metrics, net.params_logical(eng.grad)
    return total, loss_metrics


def train_ppo(env: RLEnv, networks: StatefulModule, config: Optional[TrainConfig] = None, *,
              total_steps: Optional[int] = None, seed: Optional[int] = None,
              log_fn: Optional[Callable[[dict[str, Any], int], None]] = None,
              video_fn: Optional[Callable[[VideoData], None]] = None,
              checkpoint_fn: Optional[Callable[[TrainingState, int], None]] = None,
              eval_env: Optional[RLEnv] = None,
              initial_state: Optional[TrainingState] = None) -> TrainResult:
    """Train a PPO agent — host loop of ppo.py:41-251 (same cadence rules for eval / checkpoint /
    logging; video rendering is out of scope and ignored)."""
    if config is None:
        config = default_config()
    if total_steps is not None:
        config = dataclasses.replace(config, ppo=dataclasses.replace(config.ppo, total_steps=total_steps))
    if seed is not None:
        config = dataclasses.replace(config, seed=seed)
    if eval_env is None:
        eval_env = env
    if initial_state is None:
        training_state = new_training_state(env, networks, config.ppo.n_envs, config.seed,
                                            config.ppo.learning_rate, config.ppo.gradient_clipping,
                                            config.ppo.weight_decay)
    else:
        training_state = initial_state

    eval_history: list[dict[str, Any]] = []
    last_eval_step = -config.eval.every_steps
    last_checkpoint_step = -config.checkpoint_every_steps
    metrics: dict[str, Any] = {}
    n_iterations = 0
    measure_throughput = LoggingLevel.THROUGHPUT in config.ppo.logging_level

    def run_eval(steps: int) -> dict[str, Any]:
        networks.eval()
        t0 = time.perf_counter() if measure_throughput else None
        em = rollout.eval_rollout(eval_env, networks, config.eval.n_envs, config.eval.max_episode_length,
                                  prng.key(config.seed), config.eval.logging_percentiles)
        if measure_throughput:
            em["throughput/eval_sps"] = (config.eval.n_envs * config.eval.max_episode_length
                                         / (time.perf_counter() - t0))
        networks.train()
        return dict(em)

    steps = int(training_state.steps_taken)
    if config.eval.enabled:
        em = run_eval(steps)
        metrics.update(em)
        eval_history.append({"step": steps, **em})
        last_eval_step = steps
    if checkpoint_fn is not None and _should_run(steps, last_checkpoint_step, config.checkpoint_every_steps):
        checkpoint_fn(training_state, steps)
        last_checkpoint_step = steps
    if log_fn is not None and metrics:
        log_fn(metrics, steps)

    while int(training_state.steps_taken) < config.ppo.total_steps:
        t0 = time.perf_counter() if measure_throughput else None
        training_state, metrics = ppo_step(
            env, training_state, config.ppo.n_envs, config.ppo.rollout_length, config.ppo.gae_lambda,
            config.ppo.discounting_factor, config.ppo.clip_range, config.ppo.normalize_advantages,
            config.ppo.combine_advantages, config.ppo.n_epochs, config.ppo.n_minibatches,
            config.ppo.critic_loss_weight, config.ppo.logging_level, config.ppo.logging_percentiles)
        n_iterations += 1
        steps = int(training_state.steps_taken)
        if measure_throughput:
            len(metrics)                                 # wait for the iteration (ppo_step may return before it ends)
            metrics["throughput/train_sps"] = (config.ppo.n_envs * config.ppo.rollout_length
                                               / (time.perf_counter() - t0))
        if config.eval.enabled and _should_run(steps, last_eval_step, config.eval.every_steps):
            em = run_eval(steps)
            metrics.update(em)
            eval_history.append({"step": steps, **em})
            last_eval_step = steps
        if checkpoint_fn is not None and _should_run(steps, last_checkpoint_step, config.checkpoint_every_steps):
            checkpoint_fn(training_state, steps)
            last_checkpoint_step = steps
        if log_fn is not None:
            log_fn(metrics, steps)

    return TrainResult(training_state=training_state, final_metrics=metrics, eval_history=eval_history,
                       total_steps=int(training_state.steps_taken), total_iterations=n_iterations)
