"""Action samplers — mirrors nnx_ppo/networks/sampling_layers.py:45-147 (arithmetic in K1/K3)."""
from __future__ import annotations

import abc

from .. import prng
from .types import StatefulModule


class ActionSampler(StatefulModule, abc.ABC):
    deterministic: bool = False


class NormalTanhSampler(ActionSampler):
    """Normal distribution followed by tanh (sampling_layers.py:66-147).  Draws two keys from the
    shared ``Rngs`` stream per forward call: one for the sample (skipped when deterministic), one
    for the Monte-Carlo entropy estimate."""

    def __init__(self, rng: prng.Rngs, entropy_weight: float, min_std: float = 1e-3,
                 std_scale: float = 1.0):
        self.rng = rng
        self.min_std = min_std
        self.std_scale = std_scale
        self.deterministic = False
        self.entropy_weight = entropy_weight
