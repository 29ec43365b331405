"""Observation adapters — the stateless pytree projections of nnx_ppo/networks/utils.py:
``Flattener`` (:65-117) and ``Filter`` (:119-165).

They hold no parameters and do no arithmetic: reshapes, concatenations and dict look-ups on whatever
tensors the env hands over.  That is host plumbing, so unlike the Dense / LSTM / sampler layers they
can be called on their own, and a network may put them in front of its Normalizer / PPOAdapter: the
plan compiler (networks/plan.py) recognises leading adapters and applies them to the observation
before the fused kernels see it.
"""
from __future__ import annotations

from typing import Any, Callable, Union

from .types import StatefulModule, StatefulModuleOutput

FilterSpec = Union[str, tuple, Callable[[Any], Any]]


def tree_leaves(x: Any) -> list:
    """Leaves in ``jax.tree.flatten`` order: dict values by sorted key, sequences in order, None skipped."""
    if x is None:
        return []
    if isinstance(x, dict):
        return [leaf for k in sorted(x) for leaf in tree_leaves(x[k])]
    if isinstance(x, (list, tuple)):
        return [leaf for v in x for leaf in tree_leaves(v)]
    return [x]


def _flatten_at_depth(x: Any, preserve_levels: int) -> Any:
    import torch
    if preserve_levels == 0:
        return torch.cat([a.reshape(a.shape[0], -1) for a in tree_leaves(x)], dim=-1)
    if isinstance(x, dict):
        return {k: _flatten_at_depth(v, preserve_levels - 1) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_flatten_at_depth(v, preserve_levels - 1) for v in x)
    raise TypeError("Flattener(preserve_levels > 0) requires dict/list/tuple at each preserved level; "
                    f"encountered a leaf of type {type(x).__name__} with {preserve_levels} levels still to preserve.")


class Flattener(StatefulModule):
    """Every leaf reshaped to (B, -1) and concatenated along the last axis; with ``preserve_levels=N``
    the top N levels of dict / list / tuple structure are kept and only the sub-trees are flattened."""

    def __init__(self, preserve_levels: int = 0):
        if preserve_levels < 0:
            raise ValueError(f"preserve_levels must be >= 0, got {preserve_levels}")
        self.preserve_levels = preserve_levels

    def __call__(self, state, x: Any, rollout_extras: Any = None) -> StatefulModuleOutput:
        return StatefulModuleOutput((), _flatten_at_depth(x, self.preserve_levels), 0.0, {}, None)


class Filter(StatefulModule):
    """``{output_key: extraction}``: a string takes ``x[k]``, a tuple walks a nested path, a callable
    is applied to the whole input.  Everything not named is dropped."""

    def __init__(self, spec: dict):
        if not isinstance(spec, dict):
            raise TypeError(f"Filter spec must be a dict; got {type(spec).__name__}")
        for out_key, sub in spec.items():
            if not isinstance(sub, (str, tuple)) and not callable(sub):
                raise TypeError(f"Filter spec for {out_key!r} must be str, tuple, or callable; got {type(sub).__name__}")
        self._spec = dict(spec)

    def __call__(self, state, x: Any, rollout_extras: Any = None) -> StatefulModuleOutput:
        output = {}
        for out_key, sub in self._spec.items():
            if isinstance(sub, str):
                output[out_key] = x[sub]
            elif isinstance(sub, tuple):
                v = x
                for p in sub:
                    v = v[p]
                output[out_key] = v
            else:
                output[out_key] = sub(x)
        return StatefulModuleOutput((), output, 0.0, {}, None)
