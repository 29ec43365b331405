"""Configuration dataclasses of the drop-in API.

Field names, order and defaults are the reference's (nnx_ppo/algorithms/config.py:11-116): user code
constructs these positionally / by keyword, so they are part of the boundary.  The comments say where
each value ends up in the B200 build.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np

from .types import DistillationState, LoggingLevel, TrainingState


@dataclass
class PPOConfig:
    n_envs: int = 256                       # envs per rank; rows of every [T, B, ...] rollout buffer
    rollout_length: int = 20                # T: steps per fused rollout launch
    total_steps: int = 512_000              # train_ppo stops once steps_taken >= total_steps
    gae_lambda: float = 0.95                # b200ppo_hparams.lambda_
    discounting_factor: float = 0.99        # b200ppo_hparams.gamma
    clip_range: float = 0.2                 # b200ppo_hparams.clip_range (loss kernel)
    learning_rate: float = 1e-4             # Adam kernel
    normalize_advantages: bool = True       # global (all ranks) mean / std of the minibatch advantages
    combine_advantages: bool = False        # dict rewards only: not supported (single scalar reward)
    n_epochs: int = 4                       # E permutations of the envs per iteration
    n_minibatches: int = 4                  # M: n_envs must be divisible by it; E*M updates per iteration
    critic_loss_weight: float = 1.0         # b200ppo_hparams.critic_loss_weight
    gradient_clipping: Optional[float] = None   # optax.clip_by_global_norm before Adam (grad-norm kernel)
    weight_decay: Optional[float] = None        # None: adam; True: adamw(1e-4); float: adamw(value)
    logging_level: LoggingLevel = LoggingLevel.LOSSES
    logging_percentiles: Optional[tuple[int, ...]] = None   # None: mean / std per metric


@dataclass
class EvalConfig:
    enabled: bool = True
    every_steps: int = 50_000               # cadence in env steps (also evaluated at step 0)
    n_envs: int = 64
    max_episode_length: int = 1000          # eval_rollout length (sticky done)
    logging_level: LoggingLevel = LoggingLevel.BASIC
    logging_percentiles: Optional[tuple[int, ...]] = (0, 25, 50, 75, 100)


def _default_render_kwargs() -> dict[str, Any]:
    return {"height": 480, "width": 640}


@dataclass
class VideoConfig:
    enabled: bool = False                   # needs env.render; the synthetic env has none
    every_steps: int = 200_000
    episode_length: int = 1000
    render_kwargs: dict[str, Any] = field(default_factory=_default_render_kwargs)


@dataclass
class TrainConfig:
    ppo: PPOConfig = field(default_factory=PPOConfig)
    eval: EvalConfig = field(default_factory=EvalConfig)
    video: VideoConfig = field(default_factory=VideoConfig)
    seed: int = 17                          # key(seed) -> (env reset key, training key); rank r folds r in
    checkpoint_every_steps: int = 500_000   # checkpoint_fn cadence (also at step 0)


@dataclass
class DistillationConfig:
    """config.py:72-84 of the reference."""
    n_envs: int = 256
    rollout_length: int = 20
    total_steps: int = 512_000
    learning_rate: float = 1e-4
    n_epochs: int = 4
    n_minibatches: int = 4
    gradient_clipping: Optional[float] = None
    weight_decay: Optional[float] = None
    logging_level: LoggingLevel = LoggingLevel.LOSSES
    logging_percentiles: Optional[tuple[int, ...]] = None


@dataclass
class DistillationTrainConfig:
    """config.py:88-95 of the reference."""
    distillation: DistillationConfig = field(default_factory=DistillationConfig)
    eval: EvalConfig = field(default_factory=EvalConfig)
    video: VideoConfig = field(default_factory=VideoConfig)
    seed: int = 17
    checkpoint_every_steps: int = 500_000


@dataclass
class VideoData:
    frames: np.ndarray                      # (T, H, W, C) uint8
    step: int
    episode_reward: float
    episode_length: int


@dataclass
class TrainResult:
    training_state: TrainingState
    final_metrics: dict[str, Any]
    eval_history: list[dict[str, Any]]
    total_steps: int
    total_iterations: int


@dataclass
class DistillationTrainResult:
    training_state: DistillationState
    final_metrics: dict[str, Any]
    eval_history: list[dict[str, Any]]
    total_steps: int
    total_iterations: int
