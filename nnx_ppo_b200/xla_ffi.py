"""XLA-FFI binding of libb200ppo.so for callers that live inside a JAX program (the reference's own
jitted ``ppo_step``): builds ``csrc/xla/b200ppo_xla.cc`` against ``jax.ffi.include_dir()``, registers the
handlers as CUDA custom-call targets and exposes jit-compatible wrappers with the reference's
signatures.  JAX is an OPTIONAL dependency of this module only — nothing else in the package imports it,
and without JAX every function here raises ``B200PPOError`` (there is no fallback)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import _lib, build as _build

XLA_SRC = os.path.join(_build.CSRC, "xla", "b200ppo_xla.cc")
XLA_LIB = os.path.join(_build.LIB_DIR, "libb200ppo_xla.so")
TARGETS = {"b200ppo_gae": "B200ppoGae", "b200ppo_permutation": "B200ppoPermutation", "b200ppo_update": "B200ppoUpdate"}
_registered = False


def _jax():
    try:
        import jax
        return jax
    except Exception as e:                                    # noqa: BLE001
        raise _lib.B200PPOError(f"the XLA-FFI binding needs jax ({e}); use the ctypes binding (nnx_ppo_b200._lib)")


def build(force: bool = False) -> str:
    """Compile the adapter next to libb200ppo.so (g++, no nvcc: it only forwards pointers)."""
    jax = _jax()
    _build.build()
    if not force and os.path.exists(XLA_LIB) and os.path.getmtime(XLA_LIB) >= os.path.getmtime(XLA_SRC):
        return XLA_LIB
    inc = os.path.join(os.path.dirname(_build.HERE), "include")
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(_build._nvcc())), "include")
    cmd = ["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-I", jax.ffi.include_dir(), "-I", inc, "-I", cuda_inc,
           XLA_SRC, "-L", _build.LIB_DIR, "-lb200ppo", "-Wl,-rpath,$ORIGIN", "-o", XLA_LIB]
    subprocess.check_call(cmd)
    return XLA_LIB


def register() -> None:
    """jax.ffi.register_ffi_target for every handler (idempotent)."""
    global _registered
    if _registered:
        return
    jax = _jax()
    lib = ctypes.CDLL(build())
    for target, symbol in TARGETS.items():
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(lib, symbol)), platform="CUDA")
    _registered = True


def gae(rewards, values_excl_last, last_value, done, truncation, lambda_, gamma):
    """Drop-in body for the reference's ``ppo.gae`` (ppo.py:351-394) inside jit."""
    jax = _jax()
    import jax.numpy as jp
    register()
    call = jax.ffi.ffi_call("b200ppo_gae", jax.ShapeDtypeStruct(rewards.shape, jp.float32))
    adv = call(rewards.astype(jp.float32), values_excl_last.astype(jp.float32), last_value.astype(jp.float32),
               done.astype(jp.bool_), truncation.astype(jp.bool_), lambda_=np.float32(lambda_), gamma=np.float32(gamma))
    return jax.lax.stop_gradient(adv)


def permutation_indices(new_key, n_envs: int, n_epochs: int):
    """``vmap(lambda e: jax.random.permutation(fold_in(new_key, e), n_envs))(arange(n_epochs))`` of
    ppo.py:287-294, bit-exact, as int32 [n_epochs, n_envs]."""
    jax = _jax()
    import jax.numpy as jp
    register()
    scratch = int(_lib.load().b200ppo_permutation_scratch_bytes(n_envs, n_epochs)) + 256
    call = jax.ffi.ffi_call("b200ppo_permutation", (jax.ShapeDtypeStruct((n_epochs, n_envs), jp.int32),
                                                    jax.ShapeDtypeStruct((scratch,), jp.uint8)))
    inds, _ = call(jax.random.key_data(new_key).astype(jp.uint32))
    return inds


def update(plan: _lib.Plan, hp: _lib.HParams, rollout: dict, inds, norm_mean, norm_std, counters, params, adam_mu,
           adam_nu, workspace, rng_count_offset: int, update_index: int, stages: int = _lib.STAGE_ALL):
    """One ``update_step`` (ppo.py:296-317): returns (params, adam_mu, adam_nu, workspace, metrics) with the
    first four aliased to their inputs (donate them).  ``rollout`` holds the time-major record
    (obs, raw_action, loglik_old, reward, done, truncated, next_obs_last)."""
    jax = _jax()
    import jax.numpy as jp
    register()
    outs = (jax.ShapeDtypeStruct(params.shape, jp.float32), jax.ShapeDtypeStruct(adam_mu.shape, jp.float32),
            jax.ShapeDtypeStruct(adam_nu.shape, jp.float32), jax.ShapeDtypeStruct(workspace.shape, jp.float32),
            jax.ShapeDtypeStruct((_lib.METRICS_STRIDE,), jp.float32))
    call = jax.ffi.ffi_call("b200ppo_update", outs, input_output_aliases={11: 0, 12: 1, 13: 2, 14: 3})
    return call(rollout["obs"], rollout["raw_action"], rollout["loglik_old"], rollout["reward"],
                rollout["done"].astype(jp.bool_), rollout["truncated"].astype(jp.bool_), rollout["next_obs_last"],
                inds.astype(jp.int32), norm_mean, norm_std, counters.astype(jp.uint32), params, adam_mu, adam_nu, workspace,
                plan=np.frombuffer(bytes(plan), np.uint8), hparams=np.frombuffer(bytes(hp), np.uint8),
                rng_count_offset=np.int32(rng_count_offset), update_index=np.int32(update_index), stages=np.int32(stages))
