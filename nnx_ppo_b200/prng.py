"""Host-side key arithmetic (threefry2x32, JAX ``key/split/fold_in`` layouts, ``nnx.Rngs`` stream).

Only the scalar key flow of the training loop runs here (a handful of threefry blocks per
iteration — ppo.py:271,544-548) plus parameter initialisation; all bulk random numbers (sampler
noise, env resets, permutations) are generated on the device by csrc/common.cuh.
"""
from __future__ import annotations

import math

import numpy as np

_M = 0xFFFFFFFF
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def threefry2x32(k0: int, k1: int, x0: int, x1: int) -> tuple[int, int]:
    ks = (k0 & _M, k1 & _M, (k0 ^ k1 ^ 0x1BD11BDA) & _M)
    x0 = (x0 + ks[0]) & _M
    x1 = (x1 + ks[1]) & _M
    for i in range(5):
        for r in _ROT[i % 2]:
            x0 = (x0 + x1) & _M
            x1 = ((x1 << r) | (x1 >> (32 - r))) & _M
            x1 ^= x0
        x0 = (x0 + ks[(i + 1) % 3]) & _M
        x1 = (x1 + ks[(i + 2) % 3] + i + 1) & _M
    return x0, x1


Key = tuple  # (uint32, uint32)


def key(seed: int) -> Key:
    """``jax.random.key(seed)`` raw data."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return (seed >> 32, seed & _M)


def split(k: Key, n: int = 2) -> list[Key]:
    """``jax.random.split(key, n)`` (threefry_partitionable layout)."""
    return [threefry2x32(k[0], k[1], (j >> 32) & _M, j & _M) for j in range(n)]


def fold_in(k: Key, data: int) -> Key:
    return threefry2x32(k[0], k[1], 0, int(data) & _M)


def _threefry_np(k: Key, n: int):
    """Vectorised threefry over counters 0..n-1 (parameter initialisation only)."""
    U = np.uint32
    with np.errstate(over="ignore"):
        j = np.arange(n, dtype=np.uint64)
        x0 = (j >> np.uint64(32)).astype(U)
        x1 = (j & np.uint64(_M)).astype(U)
        ks = (U(k[0]), U(k[1]), U(k[0] ^ k[1] ^ 0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = (x1 << U(r)) | (x1 >> U(32 - r))
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + U(i + 1)
    return x0, x1


def uniform(k: Key, shape, minval=0.0, maxval=1.0) -> np.ndarray:
    """``jax.random.uniform(key, shape, float32, minval, maxval)``."""
    n = int(np.prod(shape))
    o0, o1 = _threefry_np(k, n)
    bits = o0 ^ o1
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo, hi = np.float32(minval), np.float32(maxval)
    return np.maximum(lo, f * (hi - lo) + lo).reshape(shape)


class Rngs:
    """Mirror of ``flax.nnx.Rngs(seed)``: every stream call returns ``fold_in(key, count)`` and
    bumps the count.  Named streams given as keyword seeds are accepted; modules here only use
    the default stream (as the reference's Linear / NormalTanhSampler do via ``rngs()`` /
    ``rngs.params()``)."""

    def __init__(self, default: int = 0, **named: int):
        self.key = key(default)
        self.count = 0
        self.named = dict(named)

    def __call__(self) -> Key:
        k = fold_in(self.key, self.count)
        self.count = (self.count + 1) & _M
        return k

    params = __call__


def variance_scaling_uniform(k: Key, fan_in: int, fan_out: int, scale: float = 1.0) -> np.ndarray:
    """``nnx.initializers.variance_scaling(scale, "fan_in", "uniform")`` on a [in, out] kernel."""
    u = uniform(k, (fan_in, fan_out), -1.0, 1.0)
    return (u * np.float32(math.sqrt(3.0 * scale / fan_in))).astype(np.float32)


def lecun_normal_like_default(k: Key, fan_in: int, fan_out: int) -> np.ndarray:
    """Default ``nnx.Linear`` kernel init is lecun_normal (truncated normal); the hot path's
    factory always passes variance_scaling-uniform, so this helper only provides a same-scale
    uniform stand-in for hand-built Dense layers (documented deviation)."""
    return variance_scaling_uniform(k, fan_in, fan_out, 1.0)
