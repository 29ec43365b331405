"""Policy distillation — same public API as nnx_ppo/algorithms/distillation.py (``train_distillation`` :422,
``distillation_loss`` :160, ``distillation_step`` :235, ``new_distillation_state`` :363, ``default_distillation_config`` :63) on the
kernels of the PPO path.

One iteration is the PPO launch sequence with two changes (include/b200ppo.h, B200PPO_STAGE_NLL):

* after the student's fused rollout the TEACHER is evaluated in deterministic mode on the rollout's
  observations (one ``b200ppo_policy_step`` launch over all T * B rows): its raw action is its mean, which is
  what the reference stores as ``teacher_rollout_extras`` (distillation.py:81-104);
* the updates replay the student on those targets: the loss head is the mean negative log-likelihood of the
  teacher's mean plus the student's entropy regulariser (distillation.py:160-232) instead of the clipped
  surrogate; no GAE stage, no gradient into the value head, 2 * T sampler counts per update (no bootstrap call).
"""
from __future__ import annotations

import dataclasses
from collections.abc import Callable
from typing import Any, Optional

import numpy as np

from .. import _lib, parallel, prng
from ..networks.plan import compile_network
from ..networks.types import StatefulModule
from . import rollout
from .config import (DistillationConfig, DistillationTrainConfig, DistillationTrainResult, EvalConfig,  # noqa: F401
                     VideoConfig, VideoData)
from .engine import AdamOptimizer, PPOEngine, cached_engine
from .ppo import LazyMetrics, _EAGER_LEVELS, _LAZY_METRICS, _extra_metrics, _log_np, _should_run
from .types import DistillationState, LoggingLevel, RLEnv


def default_distillation_config() -> DistillationTrainConfig:
    return DistillationTrainConfig()


class DistillationEngine(PPOEngine):
    """PPOEngine with the teacher pass after the rollout and the NLL loss head in the updates."""

    def __init__(self, net, teacher_net, env, opt, n_envs, rollout_length, n_epochs, n_minibatches,
                 world_size=1, group=None, use_graph=None):
        import torch
        if teacher_net.recurrent or net.recurrent:
            raise NotImplementedError("distillation runs on the MLP plans")
        if teacher_net.plan.obs_dim != net.plan.obs_dim or teacher_net.plan.act_dim != net.plan.act_dim:
            raise ValueError("teacher and student must share the observation and action shapes "
                             "(isomorphic rollout_extras trees, distillation.py:23-26)")
        super().__init__(net, env, opt, n_envs, rollout_length, n_epochs, n_minibatches, 0.0, 0.0, 0.0, False, 0.0,
                         world_size=world_size, group=group, use_graph=use_graph)
        self.teacher = teacher_net
        T, B, A = self.T, self.B, net.plan.act_dim
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.teacher_mu = torch.zeros(T, B, A, **f32)
        self._t_scratch = (torch.empty(T * B, A, **f32), torch.empty(T * B, **f32), torch.empty(T * B, **f32))
        for b in self.bufs:
            b.raw_action = self.teacher_mu.data_ptr()          # the replay's rollout_extras are the teacher's
        self.nll = True
        self.rng_per_update = 2 * T
        self.rng_per_iter = 2 * T + self.n_updates * self.rng_per_update

    def _enqueue_rollout(self, env_state) -> int:
        n = super()._enqueue_rollout(env_state)
        t = self.teacher
        s = _lib.current_stream()
        if t.normalizer is not None:
            t.normalizer.prepare(s); n += 1
        mean_p, std_p = t.norm_ptrs()
        act, ll, val = self._t_scratch
        # deterministic mode: raw_action out = mu (sampling_layers.py:93-96)
        _lib.check(self.lib.b200ppo_policy_step(
            s, t.plan, t.arena.data_ptr(), mean_p, std_p, self.obs.data_ptr(), self.T * self.B, 2,
            t.counters.data_ptr(), 0, 0, self.teacher_mu.data_ptr(), act.data_ptr(), ll.data_ptr(),
            val.data_ptr(), 0, 0), "policy_step(teacher)")
        return n + 1


def new_distillation_state(env: RLEnv, teacher: StatefulModule, student: StatefulModule, n_envs: int, seed: int,
                           learning_rate: float = 1e-4, gradient_clipping: Optional[float] = None,
                           weight_decay: Optional[float] = None) -> DistillationState:
    """distillation.py:363-419."""
    net = compile_network(student)
    compile_network(teacher)
    key, training_key = parallel.rank_keys(seed, parallel.dist_info()[1])   # :387-388
    if not getattr(env, "fused_rollout", False):
        raise NotImplementedError("distillation needs a device env with a fused rollout kernel")
    env_states = env.reset_from_split(key, n_envs, net.device)             # :390-391
    optimizer = AdamOptimizer(net, learning_rate, gradient_clipping, weight_decay)
    return DistillationState(student, student.initialize_state(n_envs), teacher.initialize_state(n_envs),
                             env_states, optimizer, training_key, np.float32(0.0))


def _engine_for(env, teacher, ds: DistillationState, n_envs, rollout_length, n_epochs, n_minibatches):
    net, tnet = compile_network(ds.student), compile_network(teacher)
    opt: AdamOptimizer = ds.optimizer
    world = parallel.dist_info()[0]
    shape_key = (id(tnet), n_envs, rollout_length, n_epochs, n_minibatches, opt.gradient_clipping is not None,
                 opt.wd_value >= 0.0, world)
    eng = cached_engine(net, "distill", env, opt, shape_key, lambda: DistillationEngine(
        net, tnet, env, opt, n_envs, rollout_length, n_epochs, n_minibatches, world_size=world, group=None))
    if eng.teacher is not tnet:                                             # id() recycled after a collection
        eng.close()
        raise RuntimeError("stale distillation engine: teacher network changed identity")
    eng.set_hparams()
    return eng


def _distillation_metrics(per_update: np.ndarray, eng, logging_level, percentiles) -> dict[str, Any]:
    m: dict[str, Any] = {}
    if LoggingLevel.LOSSES in logging_level:                               # distillation.py:226-229, :330-332
        _log_np(m, "losses/distillation_nll", per_update[:, 0], percentiles)
        _log_np(m, "losses/regularization", per_update[:, 2], percentiles)
    lvl = logging_level & (LoggingLevel.TRAIN_ROLLOUT_STATS | LoggingLevel.TRAINING_ENV_METRICS)   # :334-349
    _extra_metrics(m, eng.net, eng, lvl, percentiles)
    return m


def distillation_loss(student: StatefulModule, student_state: Any, rollout_data, logging_level: LoggingLevel, *,
                      return_grads: bool = False):
    """distillation.py:160-232 as a callable: ``(total_loss, loss_metrics)`` of one minibatch ``DistillationTransition``
    ([T, mb, ...] leaves, the teacher's rollout_extras holding its raw-space mean at the sampler position) under the
    student's CURRENT parameters - stages FWD | LOSS (NLL head) of the update kernels on the rows in their given order.
    Like the reference call it advances the student's sampler stream by 2 * T counts.  The reference differentiates
    this function with ``nnx.grad``; here the analytic backward is part of the same kernels: ``return_grads=True`` also
    runs BWD | RED and returns the flat gradient (``CompiledNet.params_logical`` order) as a third element."""
    import torch
    net = compile_network(student)
    if net.recurrent:
        raise NotImplementedError("distillation runs on the MLP plans")
    lib = _lib.load()
    T, mb = rollout_data.rewards.shape

    def build():
        eng = PPOEngine.__new__(PPOEngine)
        fake_env = type("E", (), {"fused_rollout": True})()
        PPOEngine.__init__(eng, net, fake_env, AdamOptimizer(net), mb, T, 1, 1, 0.0, 0.0, 0.0, False, 0.0, world_size=1,
                           group=None, use_graph=False)
        eng.inds.copy_(torch.arange(mb, dtype=torch.int32, device=net.device).reshape(1, mb))
        return eng

    eng = cached_engine(net, "distill_loss", None, None, (T, mb), build)
    eng._upload_block((0, 0), (0, 0))
    eng.obs.copy_(net.flat_obs(rollout_data.obs).reshape(T, mb, -1))
    eng.raw_action.copy_(net.adapter_extras(rollout_data.teacher_rollout_extras)["action"][-1].reshape(T, mb, -1))
    eng.done.copy_(rollout_data.done.to(torch.uint8))
    s = _lib.current_stream()
    if net.normalizer is not None:
        net.normalizer.prepare(s)
    net.sync_counters_to_device()
    stages = _lib.STAGE_FWD | _lib.STAGE_LOSS | _lib.STAGE_NLL
    if return_grads:
        stages |= _lib.STAGE_BWD | _lib.STAGE_RED
    _lib.check(lib.b200ppo_update(s, net.plan, eng.hp, eng.bufs[0], T, mb, mb, 0, 0, stages), "distillation_loss")
    net.advance_rng(2 * T)
    net.sync_counters_to_device()
    row = eng.metrics[0].cpu().numpy()
    total = np.float32(row[0] + row[2])                                      # :222
    loss_metrics: dict[str, Any] = {}
    if LoggingLevel.LOSSES in logging_level:                                # :225-228
        loss_metrics["losses/distillation_nll"] = np.float32(row[0])
        loss_metrics["losses/regularization"] = np.float32(row[2])
    if return_grads:
        return total, loss_metrics, net.params_logical(eng.grad)
    return total, loss_metrics


def distillation_step(env: RLEnv, teacher: StatefulModule, distillation_state: DistillationState, n_envs: int,
                      rollout_length: int, n_epochs: int, n_minibatches: int,
                      logging_level: LoggingLevel = LoggingLevel.LOSSES,
                      logging_percentiles: Optional[tuple[int, ...]] = None
                      ) -> tuple[DistillationState, dict[str, Any]]:
    """distillation.py:235-360.  The teacher must be in eval mode (train_distillation puts it there, :469): the
    teacher pass always evaluates its mean."""
    ds = distillation_state
    eng = _engine_for(env, teacher, ds, n_envs, rollout_length, n_epochs, n_minibatches)
    reset_key, new_key = prng.split(ds.rng_key)                            # :264
    total_steps = np.float32(ds.steps_taken + np.float32(rollout_length * n_envs))
    if not (logging_level & _EAGER_LEVELS) and _LAZY_METRICS:
        pending = eng.step(ds.env_states, reset_key, new_key, fetch_metrics="lazy")

        def build():
            m = _distillation_metrics(pending.wait(), eng, logging_level, logging_percentiles)
            m["total_steps"] = total_steps
            return m
        metrics = LazyMetrics(build)
    else:
        per_update = eng.step(ds.env_states, reset_key, new_key, fetch_metrics=True)
        metrics = _distillation_metrics(per_update, eng, logging_level, logging_percentiles)
        metrics["total_steps"] = total_steps
    # the teacher's sampler still draws its entropy noise once per rollout step in eval mode (sampling_layers.py:143)
    compile_network(teacher).advance_rng(rollout_length)
    return ds.replace(env_states=eng.env_state, rng_key=new_key, steps_taken=total_steps), metrics


def train_distillation(env: RLEnv, teacher: StatefulModule, student: StatefulModule,
                       config: Optional[DistillationTrainConfig] = None, *, total_steps: Optional[int] = None,
                       seed: Optional[int] = None,
                       log_fn: Optional[Callable[[dict[str, Any], int], None]] = None,
                       video_fn: Optional[Callable[[VideoData], None]] = None,
                       checkpoint_fn: Optional[Callable[[DistillationState, int], None]] = None,
                       eval_env: Optional[RLEnv] = None,
                       initial_state: Optional[DistillationState] = None) -> DistillationTrainResult:
    """Host loop of distillation.py:422-592 (same eval / checkpoint / logging cadence; video rendering is out of
    scope and ignored)."""
    if config is None:
        config = default_distillation_config()
    if total_steps is not None:
        config = dataclasses.replace(config, distillation=dataclasses.replace(config.distillation, total_steps=total_steps))
    if seed is not None:
        config = dataclasses.replace(config, seed=seed)
    if eval_env is None:
        eval_env = env
    teacher.eval()                                                         # :469
    dc = config.distillation
    if initial_state is None:
        ds = new_distillation_state(env, teacher, student, dc.n_envs, config.seed, dc.learning_rate,
                                    dc.gradient_clipping, dc.weight_decay)
    else:
        ds = initial_state
    eval_history: list[dict[str, Any]] = []
    last_eval_step = -config.eval.every_steps
    last_checkpoint_step = -config.checkpoint_every_steps
    metrics: dict[str, Any] = {}
    n_iterations = 0

    def run_eval() -> dict[str, Any]:
        student.eval()
        em = rollout.eval_rollout(eval_env, student, config.eval.n_envs, config.eval.max_episode_length,
                                  prng.key(config.seed), config.eval.logging_percentiles)
        student.train()
        return dict(em)

    steps = int(ds.steps_taken)
    if config.eval.enabled:
        em = run_eval()
        metrics.update(em)
        eval_history.append({"step": steps, **em})
        last_eval_step = steps
    if checkpoint_fn is not None and _should_run(steps, last_checkpoint_step, config.checkpoint_every_steps):
        checkpoint_fn(ds, steps)
        last_checkpoint_step = steps
    if log_fn is not None and metrics:
        log_fn(metrics, steps)
    while int(ds.steps_taken) < dc.total_steps:
        ds, metrics = distillation_step(env, teacher, ds, dc.n_envs, dc.rollout_length, dc.n_epochs, dc.n_minibatches,
                                        dc.logging_level, dc.logging_percentiles)
        n_iterations += 1
        steps = int(ds.steps_taken)
        if config.eval.enabled and _should_run(steps, last_eval_step, config.eval.every_steps):
            em = run_eval()
            metrics.update(em)
            eval_history.append({"step": steps, **em})
            last_eval_step = steps
        if checkpoint_fn is not None and _should_run(steps, last_checkpoint_step, config.checkpoint_every_steps):
            checkpoint_fn(ds, steps)
            last_checkpoint_step = steps
        if log_fn is not None:
            log_fn(metrics, steps)
    return DistillationTrainResult(training_state=ds, final_metrics=metrics, eval_history=eval_history,
                                   total_steps=int(ds.steps_taken), total_iterations=n_iterations)
