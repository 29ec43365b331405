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


class Concat(StatefulModule):
    """Per-key dispatch + concat (containers.py:55-110): every named sub-module sees the same-named
    entry of a dict input, the outputs are concatenated along the last axis, in insertion order.

    In the B200 plan a Concat of Dense stacks at the head of the actor / critic becomes ONE
    block-diagonal Dense layer per depth level over the concatenated observation vector (the
    off-diagonal blocks are structural zeros: masked out of the gradient norm and of Adam), so the
    per-key encoders run inside the same fused kernels as a plain MLP."""

    def __init__(self, modules=None, /, **kwargs):
        if modules is not None and kwargs:
            raise ValueError("Concat: pass either a positional dict or keyword arguments, not both")
        components = modules if modules is not None else kwargs
        if not components:
            raise ValueError("Concat requires at least one component")
        self.components = dict(components)

    def _children(self):
        return list(self.components.values())

    def __call__(self, state, x, rollout_extras=None):
        raise NotImplementedError("Concat is evaluated as part of the enclosing actor-critic network")

    def initialize_state(self, batch_size: int):
        return {k: c.initialize_state(batch_size) for k, c in self.components.items()}

    def reset_state(self, prev_state):
        return {k: c.reset_state(prev_state[k]) for k, c in self.components.items()}

    def update_statistics(self, rollout_extras) -> None:
        for key, component in self.components.items():
            component.update_statistics(rollout_extras[key] if rollout_extras is not None else None)
