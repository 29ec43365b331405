"""RewardScalingWrapper — the env wrapper of nnx_ppo/wrappers/reward_scaling_wrapper.py:8-28 for
batched torch envs (states are dataclasses of [B, ...] tensors; see algorithms/rollout.py).

Pure host-side plumbing: it multiplies the reward leaf of whatever state the wrapped env returns,
in ``reset`` as well as in ``step``, exactly as the reference does."""
from __future__ import annotations

import dataclasses
from typing import Any


def _scale(reward: Any, scale: float) -> Any:
    if isinstance(reward, dict):                       # dict rewards keep their structure
        return {k: _scale(v, scale) for k, v in reward.items()}
    return scale * reward


class RewardScalingWrapper:
    def __init__(self, env, reward_scale: float) -> None:
        self.env = env
        self.reward_scale = reward_scale

    def reset(self, rng):
        state = self.env.reset(rng)
        return dataclasses.replace(state, reward=_scale(state.reward, self.reward_scale))

    def step(self, state, action):
        nxt = self.env.step(state, action)
        return dataclasses.replace(nxt, reward=_scale(nxt.reward, self.reward_scale))

    @property
    def observation_size(self):
        return self.env.observation_size

    @property
    def action_size(self):
        return self.env.action_size
