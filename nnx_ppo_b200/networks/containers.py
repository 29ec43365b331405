"""Containers — mirrors nnx_ppo/networks/containers.py:14-52 (Sequential)."""
from __future__ import annotations

from typing import Any, Sequence

from .types import ModuleState, StatefulModule


class Sequential(StatefulModule):
    def __init__(self, layers: Sequence[StatefulModule]):
        self.layers = list(layers)

    def _children(self):
        return self.layers

    def __call__(self, network_state, obs: Any, rollout_extras: Any = None):
        from .plan import call_network
        return call_network(self, network_state, obs, rollout_extras)

    def initialize_state(self, batch_size: int) -> list[ModuleState]:
        return [layer.initialize_state(batch_size) for layer in self.layers]

    def reset_state(self, prev_state):
        return [layer.reset_state(s) for layer, s in zip(self.layers, prev_state)]

    def update_statistics(self, rollout_extras: Any) -> None:
        for layer, layer_extras in zip(self.layers, rollout_extras):
            layer.update_statistics(layer_extras)

    def __getitem__(self, ind: int) -> StatefulModule:
        return self.layers[ind]
