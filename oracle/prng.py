"""NumPy restatement of the JAX PRNG pieces the PPO hot path uses (ORACLE — test infrastructure).

Third-party arithmetic restated here (absent from /root/reference, unpinned in its
pyproject.toml:25-27): jax's threefry2x32 PRNG with ``jax_threefry_partitionable=True`` (the
default since JAX 0.5), ``jax.random.{key,split,fold_in,bits,uniform,normal,randint,permutation}``
and flax.nnx's ``Rngs`` stream.  Reference call sites: ``ppo.py:271,288-289,544-548``,
``rollout.py:57-59``, ``sampling_layers.py:96,144``, ``episode_wrapper.py:25-29``.

Known answers (see tests/test_oracle_prng.py): Random123 threefry2x32 vectors and the JAX
``split(key(0))`` / ``fold_in(key(0), 1)`` values listed in SURVEY.md App. B.
"""

from __future__ import annotations

import math

import numpy as np

U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds (Salmon et al. 2011; jax/_src/prng.py ``threefry2x32``)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, U32)
        k1 = np.asarray(k1, U32)
        x0 = np.asarray(x0, U32).copy()
        x1 = np.asarray(x1, U32).copy()
        ks = (k0, k1, k0 ^ k1 ^ U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + U32(i + 1)
    return x0, x1


def key(seed: int) -> np.ndarray:
    """``jax.random.key(seed)`` → raw key data ``(hi32, lo32)``."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=U32)


def _counts(n: int):
    j = np.arange(n, dtype=np.uint64)
    return (j >> np.uint64(32)).astype(U32), (j & np.uint64(0xFFFFFFFF)).astype(U32)


def split(k: np.ndarray, shape=2) -> np.ndarray:
    """``jax.random.split`` (partitionable layout): element j = threefry(key, (hi(j), lo(j)))."""
    if isinstance(shape, int):
        shape = (shape,)
    n = int(np.prod(shape))
    hi, lo = _counts(n)
    o0, o1 = threefry2x32(k[0], k[1], hi, lo)
    return np.stack([o0, o1], axis=-1).reshape(tuple(shape) + (2,))


def fold_in(k: np.ndarray, data: int) -> np.ndarray:
    """``jax.random.fold_in``: threefry(key, (0, data))."""
    o0, o1 = threefry2x32(k[0], k[1], U32(0), U32(int(data) & 0xFFFFFFFF))
    return np.array([o0, o1], dtype=U32)


def random_bits(k: np.ndarray, shape) -> np.ndarray:
    """``jax.random.bits(key, shape, uint32)`` (partitionable): out0 ^ out1 per element."""
    if isinstance(shape, int):
        shape = (shape,)
    n = int(np.prod(shape)) if len(shape) else 1
    hi, lo = _counts(n)
    o0, o1 = threefry2x32(k[0], k[1], hi, lo)
    return (o0 ^ o1).reshape(shape)


def bits_to_uniform(bits: np.ndarray, minval=0.0, maxval=1.0) -> np.ndarray:
    """float32 uniform from 32 random bits exactly as ``jax.random.uniform`` does."""
    minval = np.float32(minval)
    maxval = np.float32(maxval)
    fb = (bits >> U32(9)) | U32(0x3F800000)
    f = fb.view(np.float32) - np.float32(1.0)
    return np.maximum(minval, f * (maxval - minval) + minval)


def uniform(k, shape, minval=0.0, maxval=1.0):
    return bits_to_uniform(random_bits(k, shape), minval, maxval)


def erfinv_f32(x: np.ndarray) -> np.ndarray:
    """XLA's float32 ``erf_inv`` (Giles 2010 single-precision polynomial on w = -log1p(-x²))."""
    x = np.asarray(x, np.float32)
    f = np.float32
    with np.errstate(divide="ignore", invalid="ignore"):
        w = -np.log1p(-(x * x)).astype(np.float32)
        wl = w - f(2.5)
        p = f(2.81022636e-08)
        for c in (3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
                  -0.00125372503, -0.00417768164, 0.246640727, 1.50140941):
            p = f(c) + p * wl
        ws = np.sqrt(w).astype(np.float32) - f(3.0)
        q = f(-0.000200214257)
        for c in (0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
                  -0.0076224613, 0.00943887047, 1.00167406, 2.83297682):
            q = f(c) + q * ws
        out = np.where(w < f(5.0), p, q) * x
    return np.where(np.abs(x) == f(1.0), np.copysign(np.float32(np.inf), x), out).astype(np.float32)


_NORMAL_LO = np.nextafter(np.float32(-1.0), np.float32(0.0))
_SQRT2 = np.float32(np.sqrt(2))


def bits_to_normal(bits: np.ndarray) -> np.ndarray:
    u = bits_to_uniform(bits, _NORMAL_LO, 1.0)
    return (_SQRT2 * erfinv_f32(u)).astype(np.float32)


def normal(k, shape) -> np.ndarray:
    """``jax.random.normal(key, shape, float32)``: sqrt(2)·erfinv(uniform(nextafter(-1,0), 1))."""
    return bits_to_normal(random_bits(k, shape))


def randint(k, shape, minval: int, maxval: int) -> np.ndarray:
    """``jax.random.randint`` for int32 (jax/_src/random.py ``_randint``)."""
    k1, k2 = split(k)
    hi_bits = random_bits(k1, shape).astype(np.uint64)
    lo_bits = random_bits(k2, shape).astype(np.uint64)
    span = max(int(maxval) - int(minval), 1)
    mult = (2 ** 16) % span
    mult = (mult * mult) % span
    with np.errstate(over="ignore"):
        off = ((hi_bits % np.uint64(span)).astype(U32) * U32(mult)
               + (lo_bits % np.uint64(span)).astype(U32))
    off = off % U32(span)
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int32)


def permutation_rounds(n: int) -> int:
    """Number of sort rounds of ``jax.random.permutation`` (jax/_src/random.py ``_shuffle``)."""
    return int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))


def permutation(k, n: int) -> np.ndarray:
    """``jax.random.permutation(key, n)``: repeated stable sort by fresh 32-bit keys."""
    x = np.arange(n, dtype=np.int32)
    k = np.asarray(k, U32)
    for _ in range(permutation_rounds(n)):
        k, sub = split(k)
        sort_keys = random_bits(sub, (n,))
        order = np.argsort(sort_keys, kind="stable")
        x = x[order]
    return x


class Rngs:
    """flax.nnx ``Rngs(seed)`` default stream: ``rngs()`` = fold_in(stream_key, count); count += 1."""

    def __init__(self, seed: int = 0, count: int = 0, **_named):
        self.key = key(seed)
        self.count = int(count)

    def __call__(self) -> np.ndarray:
        k = fold_in(self.key, self.count)
        self.count = (self.count + 1) & 0xFFFFFFFF
        return k

    params = __call__


def variance_scaling_uniform(k, fan_in: int, fan_out: int, scale: float = 1.0) -> np.ndarray:
    """``nnx.initializers.variance_scaling(scale, "fan_in", "uniform")`` for a [in, out] kernel."""
    u = uniform(k, (fan_in, fan_out), -1.0, 1.0)
    return (u * np.float32(math.sqrt(3.0 * scale / fan_in))).astype(np.float32)
