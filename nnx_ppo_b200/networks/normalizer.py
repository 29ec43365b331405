"""Online (Welford) input normalizer — mirrors nnx_ppo/networks/normalizer.py:35-136.

The forward ``(x - mean) / std`` is fused into the policy / update kernels (K1, K3); the batched
Welford merge of ``update_statistics`` is K5 (csrc/misc.cu).  Flat observation shapes only; the
dict-observation variant (normalizer.py:52-59) is a SURVEY §8 "next" row.
"""
from __future__ import annotations

from typing import Any

import numpy as np

from .feedforward import Param
from .types import StatefulModule, StatefulModuleOutput


class Normalizer(StatefulModule):
    def __init__(self, shape):
        if isinstance(shape, (tuple, list)):
            if len(shape) != 1:
                raise NotImplementedError("Normalizer: only flat [obs_size] shapes are supported")
            shape = int(shape[0])
        if not isinstance(shape, (int, np.integer)):
            raise NotImplementedError("Normalizer: dict / pytree observation shapes are not supported yet")
        self.size = int(shape)
        self.mean = Param(np.zeros(self.size, np.float32))
        self.M2 = Param(np.zeros(self.size, np.float32))
        self.counter = Param(np.zeros(1, np.float32))
        self.epsilon = 1e-6
        self._std = None  # device scratch, filled by norm_prepare

    def _bind(self, device) -> None:
        import torch
        if self.mean._dev is None:
            for p in (self.mean, self.M2, self.counter):
                p._dev = torch.from_numpy(p._host.copy()).to(device)
            self._std = torch.empty(self.size, dtype=torch.float32, device=device)
            self._batch_stats = torch.zeros(2 * self.size, dtype=torch.float32, device=device)

    def prepare(self, stream=None) -> None:
        """std = counter > 0 ? sqrt(max(M2 / counter, eps)) : 10   (normalizer.py:72-77,92-96)."""
        from .. import _lib
        lib = _lib.load()
        _lib.check(lib.b200ppo_norm_prepare(stream if stream is not None else _lib.current_stream(),
                                            _lib.ptr(self.M2._dev), _lib.ptr(self.counter._dev),
                                            self.size, _lib.ptr(self._std)), "norm_prepare")

    def __call__(self, state, x, rollout_extras: Any = None) -> StatefulModuleOutput:
        """Standalone forward through the fused policy kernel is not exposed; the normalised
        values are produced by the NORM-only CUDA path below for parity tests."""
        raise NotImplementedError("call the enclosing actor-critic network")

    def update_statistics(self, rollout_extras: Any) -> None:
        """Fold a ``[T, B, O]`` (or ``[N, O]``) history of raw inputs into the running statistics
        with one batched Welford merge on the GPU (normalizer.py:98-136)."""
        import torch
        from .. import _lib
        x = rollout_extras
        if not isinstance(x, torch.Tensor):
            raise TypeError("update_statistics expects a CUDA float32 tensor [T, B, O]")
        self._bind(x.device)
        x = x.reshape(-1, self.size)
        _lib.require_cuda(x)
        lib = _lib.load()
        scratch = torch.zeros(int(lib.b200ppo_norm_scratch_bytes(self.size)) // 4,
                              dtype=torch.float32, device=x.device)
        s = _lib.current_stream()
        _lib.check(lib.b200ppo_norm_batch_stats(s, _lib.ptr(x), x.shape[0], self.size,
                                                _lib.ptr(self._batch_stats), _lib.ptr(scratch)),
                   "norm_batch_stats")
        _lib.check(lib.b200ppo_norm_merge(s, _lib.ptr(self._batch_stats), 1, float(x.shape[0]),
                                          self.size, _lib.ptr(self.mean._dev), _lib.ptr(self.M2._dev),
                                          _lib.ptr(self.counter._dev)), "norm_merge")
