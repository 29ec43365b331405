"""LSTM layer — mirrors nnx_ppo/networks/recurrent.py:16-161 (flax ``OptimizedLSTMCell`` underneath).

Parameters: ``kernel_i`` [in, 4H] (the four input kernels ii|if|ig|io, no bias), ``kernel_h`` [H, 4H]
(hi|hf|hg|ho) and their ``bias`` [4H].  The carry is the reference's 2-tuple of [B, H] arrays, zeros
at start and after a reset - or, with ``trainable_initial_state`` (recurrent.py:85-87), the learned
``initial_h`` / ``initial_c`` vectors broadcast over the batch into carry slot 0 / slot 1 (the reference
hands the tuple to flax untouched, whose cell reads slot 0 as c: recurrent.py:109, 135-141).
Initialisation: uniform
variance scaling for both kernels (flax's lecun-normal / orthogonal defaults cannot be reproduced
bit-for-bit without jax).  The arithmetic runs in csrc/recurrent.cu.
"""
from __future__ import annotations

import numpy as np

from .. import prng
from .feedforward import Param
from .types import StatefulModule


class LSTM(StatefulModule):
    def __init__(self, in_features: int, hidden_features: int, rngs: prng.Rngs, *,
                 trainable_initial_state: bool = False, **unsupported):
        bad = {k: v for k, v in unsupported.items() if v is not None and k not in ("use_optimized",)}
        if bad:
            raise NotImplementedError(f"unsupported LSTM options: {sorted(bad)}")
        self.in_features = in_features
        self.hidden_features = hidden_features
        H = hidden_features
        self.kernel_i = Param(prng.variance_scaling_uniform(rngs.params(), in_features, 4 * H, 1.0))
        self.kernel_h = Param(prng.variance_scaling_uniform(rngs.params(), H, 4 * H, 1.0))
        self.bias = Param(np.zeros(4 * H, np.float32))
        self.trainable_initial_state = bool(trainable_initial_state)
        if self.trainable_initial_state:                                  # recurrent.py:85-87
            self.initial_h = Param(np.zeros(H, np.float32))
            self.initial_c = Param(np.zeros(H, np.float32))

    def _initial(self, like0, like1):
        import torch
        if not self.trainable_initial_state:
            return (torch.zeros_like(like0), torch.zeros_like(like1))
        as_dev = lambda p, like: (p.value if isinstance(p.value, torch.Tensor) else torch.from_numpy(p.numpy())).to(like.device)
        return (as_dev(self.initial_h, like0).expand_as(like0).clone(), as_dev(self.initial_c, like1).expand_as(like1).clone())

    def initialize_state(self, batch_size: int):
        import torch
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        z = lambda: torch.zeros(batch_size, self.hidden_features, dtype=torch.float32, device=dev)
        return self._initial(z(), z())

    def reset_state(self, prev_state):
        return self._initial(prev_state[0], prev_state[1])
