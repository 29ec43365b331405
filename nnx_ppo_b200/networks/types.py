"""Module protocol of the network tree — mirrors nnx_ppo/networks/types.py:14-113."""
from __future__ import annotations

import abc
import dataclasses
from typing import Any

ModuleState = Any


@dataclasses.dataclass(frozen=True)
class PPONetworkOutput:
    """PPO-specific forward output (reference: networks/types.py:14-26)."""
    actions: Any
    loglikelihoods: Any
    value_estimates: Any


@dataclasses.dataclass(frozen=True)
class StatefulModuleOutput:
    """Reference: networks/types.py:29-36."""
    next_state: ModuleState
    output: Any
    regularization_loss: Any
    metrics: dict
    rollout_extras: Any = None


class StatefulModule(abc.ABC):
    """Interface between network modules and the RL algorithm (reference: networks/types.py:39-113).

    ``__call__(module_state, obs, rollout_extras=None)``: ``rollout_extras is None`` means ROLLOUT /
    INFERENCE (sample fresh, emit snapshots), a value means LOSS_REPLAY (consume snapshots).
    In this build the arithmetic of a whole actor-critic tree runs in fused CUDA kernels; a tree
    is first compiled to a flat "plan" (networks/plan.py) and only plan-compilable trees can be
    called or trained.  Anything else raises NotImplementedError — there is no eager fallback.
    """

    training: bool = True

    def __call__(self, module_state: ModuleState, obs: Any, rollout_extras: Any = None):
        raise NotImplementedError(
            f"{type(self).__name__} cannot be evaluated on its own in the B200 build: call the "
            "enclosing actor-critic network (Sequential([Normalizer?, PPOAdapter(...)]))")

    def initialize_state(self, batch_size: int) -> ModuleState:
        return ()

    def reset_state(self, prev_state: ModuleState) -> ModuleState:
        return prev_state

    def update_statistics(self, rollout_extras: Any) -> None:
        del rollout_extras
        return None

    # nnx.Module.train()/eval(): flips `deterministic` on samplers (ppo.py:122,139)
    def _children(self):
        return []

    def iter_modules(self):
        yield self
        for c in self._children():
            yield from c.iter_modules()

    def eval(self):
        for m in self.iter_modules():
            m.training = False
            if hasattr(m, "deterministic"):
                m.deterministic = True

    def train(self):
        for m in self.iter_modules():
            m.training = True
            if hasattr(m, "deterministic"):
                m.deterministic = False
