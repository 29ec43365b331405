"""Runtime types for the PPO algorithm — mirrors nnx_ppo/algorithms/types.py:15-150."""
from __future__ import annotations

import dataclasses
import enum
from typing import Any, Protocol, runtime_checkable


@runtime_checkable
class EnvState(Protocol):
    @property
    def obs(self) -> Any: ...
    @property
    def done(self) -> Any: ...
    @property
    def reward(self) -> Any: ...
    @property
    def info(self) -> dict: ...
    @property
    def metrics(self) -> dict: ...


@runtime_checkable
class RLEnv(Protocol):
    def reset(self, rng: Any) -> Any: ...
    def step(self, state: Any, action: Any) -> Any: ...


@dataclasses.dataclass(frozen=True)
class TrainingState:
    """Reference: algorithms/types.py:48-56.  ``rng_key`` is the raw (hi, lo) uint32 pair and
    ``steps_taken`` a float32 scalar, as in the reference (ppo.py:571)."""
    networks: Any
    network_states: Any
    env_states: Any
    optimizer: Any
    rng_key: Any
    steps_taken: Any

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)


@dataclasses.dataclass(frozen=True)
class Transition:
    """Reference: algorithms/types.py:59-80.  Leaves are time-major [T, B, ...] CUDA tensors."""
    obs: Any
    network_output: Any
    rewards: Any
    done: Any
    truncated: Any
    next_obs: Any
    metrics: dict
    rollout_extras: Any = None


@dataclasses.dataclass(frozen=True)
class DistillationTransition:
    """Reference: algorithms/types.py:85-106.  The teacher's rollout_extras at the sampler position (its action
    mean in raw space, the teacher running in eval mode) is the distillation target."""
    obs: Any
    student_output: Any
    rewards: Any
    done: Any
    truncated: Any
    next_obs: Any
    metrics: dict
    student_rollout_extras: Any = None
    teacher_rollout_extras: Any = None


@dataclasses.dataclass(frozen=True)
class DistillationState:
    """Reference: algorithms/types.py:111-125.  The teacher is an argument of the step (like the env), only its
    per-env carry is tracked here."""
    student: Any
    student_states: Any
    teacher_states: Any
    env_states: Any
    optimizer: Any
    rng_key: Any
    steps_taken: Any

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)


class LoggingLevel(enum.Flag):
    LOSSES = enum.auto()
    CRITIC_EXTRA = enum.auto()
    ACTOR_EXTRA = enum.auto()
    TRAIN_ROLLOUT_STATS = enum.auto()
    ROLLOUT_OBS = enum.auto()
    TRAINING_ENV_METRICS = enum.auto()
    GRAD_NORM = enum.auto()
    WEIGHTS = enum.auto()
    THROUGHPUT = enum.auto()
    BASIC = LOSSES
    ALL = (LOSSES | ACTOR_EXTRA | CRITIC_EXTRA | TRAIN_ROLLOUT_STATS | TRAINING_ENV_METRICS
           | GRAD_NORM | WEIGHTS | ROLLOUT_OBS | THROUGHPUT)
    NONE = 0
