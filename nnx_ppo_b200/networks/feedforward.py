"""Dense layer — mirrors nnx_ppo/networks/feedforward.py:13-51."""
from __future__ import annotations

from typing import Any, Callable, Optional

import numpy as np

from .. import prng
from .types import StatefulModule


def relu(x):  # activation tokens: the CUDA plan only needs to know WHICH activation is used
    raise RuntimeError("activation tokens are not callable on the host")


def swish(x):
    raise RuntimeError("activation tokens are not callable on the host")


def tanh(x):
    raise RuntimeError("activation tokens are not callable on the host")


silu = swish
_ACT_NAMES = {relu: "relu", swish: "swish", tanh: "tanh", None: "none",
              "relu": "relu", "swish": "swish", "silu": "swish", "tanh": "tanh", "none": "none"}


def activation_name(act) -> str:
    if act in _ACT_NAMES:
        return _ACT_NAMES[act]
    name = getattr(act, "__name__", None)
    if name in ("relu", "swish", "silu", "tanh"):
        return "swish" if name == "silu" else name
    raise NotImplementedError(f"unsupported activation {act!r}: the B200 plan supports relu / swish / tanh")


class Param:
    """A parameter tensor: host float32 array until the network is compiled, then a torch view
    into the flat device parameter arena (so training mutates the user's network in place)."""

    def __init__(self, value: np.ndarray):
        self._host = np.ascontiguousarray(value, np.float32)
        self._dev = None
        self._after_set = None          # compiled shared-trunk parameters: re-sync the tied second copy

    @property
    def shape(self):
        return self._host.shape

    @property
    def value(self):
        return self._dev if self._dev is not None else self._host

    def numpy(self) -> np.ndarray:
        if self._dev is not None:
            return self._dev.detach().cpu().numpy().copy()
        return self._host.copy()

    def set(self, value) -> None:
        value = np.ascontiguousarray(value, np.float32).reshape(self._host.shape)
        if self._dev is not None:
            import torch
            self._dev.copy_(torch.from_numpy(value))
            if self._after_set is not None:
                self._after_set()
        self._host = value.copy()

    def __getitem__(self, idx):
        return self.value[idx]


class Linear:
    """``flax.nnx.Linear`` stand-in: kernel [in, out], bias [out]; draws two keys from ``rngs``
    (kernel, then bias) exactly like the reference's layers do."""

    def __init__(self, in_features: int, out_features: int, rngs: prng.Rngs, kernel_init=None):
        kkey = rngs.params()
        if kernel_init is None:
            w = prng.lecun_normal_like_default(kkey, in_features, out_features)
        else:
            w = kernel_init(kkey, (in_features, out_features))
        rngs.params()  # bias key (zeros initializer)
        self.kernel = Param(w)
        self.bias = Param(np.zeros(out_features, np.float32))


class Dense(StatefulModule):
    def __init__(self, in_features: int, out_features: int, rngs: prng.Rngs,
                 activation: Optional[Callable] = None, **linear_kwargs: Any):
        self.in_features = in_features
        self.out_features = out_features
        self.linear = Linear(in_features, out_features, rngs, **linear_kwargs)
        self.activation = activation
        self.activation_name = activation_name(activation)
