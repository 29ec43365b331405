"""Containers — mirrors nnx_ppo/networks/containers.py: Sequential (:14-52), Concat (:55-112), Parallel (:115-176),
Splitter (:179-218)."""
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


class Parallel(StatefulModule):
    """Several named sub-modules on the SAME input, outputs as a dict keyed by name (containers.py:115-176).

    Parameter-free routing: callable on its own when its components are (``Splitter``, ``Filter`` ...).  As the
    tail of a shared trunk - ``Sequential([trunk..., Parallel(action_params=actor_head, value=critic_head)])`` -
    the plan compiler (networks/plan.py) lowers the branches like the ports of a ``PPOAdapter`` placed after the
    trunk."""

    def __init__(self, modules=None, /, **kwargs):
        if modules is not None and kwargs:
            raise ValueError("Parallel: pass either a positional dict or keyword arguments, not both")
        components = modules if modules is not None else kwargs
        if not components:
            raise ValueError("Parallel requires at least one sub-module")
        self.components = dict(components)

    def _children(self):
        return list(self.components.values())

    def __call__(self, state, x, rollout_extras=None):
        from .types import StatefulModuleOutput
        new_state, new_extras, outputs, metrics = {}, {}, {}, {}
        reg = 0.0
        for key, component in self.components.items():
            child_extras = None if rollout_extras is None else rollout_extras[key]
            out = component(state[key], x, child_extras)
            new_state[key], new_extras[key], outputs[key], metrics[key] = (out.next_state, out.rollout_extras,
                                                                            out.output, out.metrics)
            reg = reg + out.regularization_loss
        return StatefulModuleOutput(new_state, outputs, reg, metrics, new_extras)

    def initialize_state(self, batch_size: int):
        return {k: c.initialize_state(batch_size) for k, c in self.components.items()}

    def reset_state(self, prev_state):
        return {k: c.reset_state(prev_state[k]) for k, c in self.components.items()}

    def update_statistics(self, rollout_extras) -> None:
        for key, component in self.components.items():
            component.update_statistics(rollout_extras[key] if rollout_extras is not None else None)


class Splitter(StatefulModule):
    """A tensor split into named slices along the last axis, in keyword order (containers.py:179-218); excess
    input features are ignored.  Parameter-free: plain slicing of whatever tensor comes in."""

    def __init__(self, **sizes: int):
        if not sizes:
            raise ValueError("Splitter requires at least one named slice")
        for k, v in sizes.items():
            if v <= 0:
                raise ValueError(f"slice size for {k!r} must be positive, got {v}")
        self._sizes = dict(sizes)

    def __call__(self, state, x, rollout_extras=None):
        from .types import StatefulModuleOutput
        outputs, offset = {}, 0
        for key, size in self._sizes.items():
            outputs[key] = x[..., offset:offset + size]
            offset += size
        return StatefulModuleOutput((), outputs, 0.0, {}, None)
