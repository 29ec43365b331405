// Shared device helpers for the B200 PPO kernels: threefry2x32 + JAX random transforms,
// activation / sampler math, small reductions.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200ppo.h"

#define B200PPO_LAUNCH_CHECK()                                   \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return static_cast<int>(e__);        \
  } while (0)

namespace b200ppo {

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL) between the kernels of one update.  A kernel launched with the
// programmatic-stream-serialization attribute may start while its predecessor in the stream is still
// draining: its CTAs become resident, run their prologue (shared-memory / TMEM / mbarrier set-up) and
// block in griddepcontrol.wait until the predecessor grid has completed and flushed.  Every kernel of
// the chain executes launch_dependents first (so ITS successor can be scheduled as soon as all its own
// CTAs are resident) and wait before its first dependent global access; completion stays transitive
// because every grid waits.  Without the attribute both instructions are no-ops.  The attribute is OFF
// unless B200PPO_PDL=1 / b200ppo_set_pdl(1): inside the captured iteration graph the programmatic edges
// measured slower than plain kernel-to-kernel edges on B200 (profiles/r2_notes.md).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();   // misc.cu (reads B200PPO_PDL once)
int pdl_mode();       // 0 off, 1 every launch, 2 only the small kernels (GAE / loss / Adam), 3 those + the kernel after one

// launch class: 0 = large kernel following a large kernel, 1 = small latency-bound kernel, 2 = large kernel
// following a small one
template <class... KArgs, class... Args>
inline cudaError_t launch_kc(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                             Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  const int m = pdl_mode();
  cfg.numAttrs = (m == 1 || (m == 2 && cls == 1) || (m == 3 && cls >= 1)) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                            Args... args) {
  return launch_kc(0, kernel, grid, block, smem, s, args...);
}

// ------------------------------------------------------------------------------------------
// threefry2x32 (20 rounds) — jax/_src/prng.py; reference call sites ppo.py:271,288-289,
// rollout.py:57-59, sampling_layers.py:96,144.  Bit-exact integer work.
// ------------------------------------------------------------------------------------------
struct Key {
  uint32_t a, b;
};

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) {
  return (x << r) | (x >> (32 - r));
}

__host__ __device__ __forceinline__ Key threefry2x32(Key k, uint32_t x0, uint32_t x1) {
  const uint32_t ks0 = k.a, ks1 = k.b, ks2 = k.a ^ k.b ^ 0x1BD11BDAu;
  x0 += ks0;
  x1 += ks1;
#define TF_ROUND(r) \
  x0 += x1;         \
  x1 = rotl32(x1, r); \
  x1 ^= x0;
  TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
  x0 += ks1; x1 += ks2 + 1u;
  TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24)
  x0 += ks2; x1 += ks0 + 2u;
  TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
  x0 += ks0; x1 += ks1 + 3u;
  TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24)
  x0 += ks1; x1 += ks2 + 4u;
  TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
  x0 += ks2; x1 += ks0 + 5u;
#undef TF_ROUND
  return Key{x0, x1};
}

// jax.random.fold_in(key, d) and element j of jax.random.split(key, n) (partitionable layout).
__host__ __device__ __forceinline__ Key fold_in(Key k, uint32_t d) { return threefry2x32(k, 0u, d); }
__host__ __device__ __forceinline__ Key split_at(Key k, uint32_t j) { return threefry2x32(k, 0u, j); }
// element j of jax.random.bits(key, shape, uint32) (partitionable layout, j < 2^32).
__host__ __device__ __forceinline__ uint32_t random_bits_at(Key k, uint32_t j) {
  Key o = threefry2x32(k, 0u, j);
  return o.a ^ o.b;
}

// XLA float32 erf_inv (Giles' polynomial on w = -log1p(-x^2)).
__device__ __forceinline__ float erfinv_giles(float x) {
  float w = -log1pf(-__fmul_rn(x, x));
  float p;
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = __fadd_rn(3.43273939e-07f, __fmul_rn(p, w));
    p = __fadd_rn(-3.5233877e-06f, __fmul_rn(p, w));
    p = __fadd_rn(-4.39150654e-06f, __fmul_rn(p, w));
    p = __fadd_rn(0.00021858087f, __fmul_rn(p, w));
    p = __fadd_rn(-0.00125372503f, __fmul_rn(p, w));
    p = __fadd_rn(-0.00417768164f, __fmul_rn(p, w));
    p = __fadd_rn(0.246640727f, __fmul_rn(p, w));
    p = __fadd_rn(1.50140941f, __fmul_rn(p, w));
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = __fadd_rn(0.000100950558f, __fmul_rn(p, w));
    p = __fadd_rn(0.00134934322f, __fmul_rn(p, w));
    p = __fadd_rn(-0.00367342844f, __fmul_rn(p, w));
    p = __fadd_rn(0.00573950773f, __fmul_rn(p, w));
    p = __fadd_rn(-0.0076224613f, __fmul_rn(p, w));
    p = __fadd_rn(0.00943887047f, __fmul_rn(p, w));
    p = __fadd_rn(1.00167406f, __fmul_rn(p, w));
    p = __fadd_rn(2.83297682f, __fmul_rn(p, w));
  }
  return __fmul_rn(p, x);
}

// jax.random.normal from 32 random bits: sqrt(2) * erfinv(uniform(nextafter(-1, 0), 1)).
__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
  const float lo = -0.99999994f;  // nextafter(-1, 0) in float32
  float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  float u = __fadd_rn(__fmul_rn(f, __fsub_rn(1.0f, lo)), lo);
  u = fmaxf(lo, u);
  return __fmul_rn(1.41421356f, erfinv_giles(u));
}

// jax.random.randint(key, (), 0, span) for int32 (jax/_src/random.py _randint).
__device__ __forceinline__ int32_t randint_scalar(Key k, uint32_t span) {
  Key k1 = split_at(k, 0u), k2 = split_at(k, 1u);
  uint32_t hi = random_bits_at(k1, 0u), lo = random_bits_at(k2, 0u);
  if (span == 0u) span = 1u;
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  uint32_t off = (hi % span) * mult + (lo % span);
  return static_cast<int32_t>(off % span);
}

// ------------------------------------------------------------------------------------------
// activation / sampler math (feedforward.py:48-50; sampling_layers.py:88-147)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_f(float x) {
  float e = expf(-fabsf(x));
  return x >= 0.0f ? 1.0f / (1.0f + e) : e / (1.0f + e);
}
__device__ __forceinline__ float softplus_f(float x) {
  return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x)));
}
// relu / identity stay inline; the transcendental activations are out-of-line so that heavily
// unrolled GEMM prologues / epilogues do not carry 16 copies of tanhf / expf (instruction cache)
static __device__ __noinline__ float act_fwd_slow(float z, int act) {
  return act == B200PPO_ACT_TANH ? tanhf(z) : z * sigmoid_f(z);
}
static __device__ __noinline__ float act_grad_slow(float z, int act) {
  if (act == B200PPO_ACT_TANH) { const float h = tanhf(z); return 1.0f - h * h; }
  const float s = sigmoid_f(z);
  return s * (1.0f + z * (1.0f - s));
}
__device__ __forceinline__ float act_fwd(float z, int act) {
  if (act == B200PPO_ACT_RELU) return fmaxf(z, 0.0f);
  if (act == B200PPO_ACT_NONE) return z;
  return act_fwd_slow(z, act);
}
__device__ __forceinline__ float act_grad(float z, int act) {
  if (act == B200PPO_ACT_RELU) return z > 0.0f ? 1.0f : 0.0f;
  if (act == B200PPO_ACT_NONE) return 1.0f;
  return act_grad_slow(z, act);
}

#define B200PPO_LOG2 0.69314718f
#define B200PPO_HALF_LOG_2PI 0.91893853f

// log|d tanh(z)/dz| in the Brax-stable form (sampling_layers.py:131).
__device__ __forceinline__ float log_det_jac(float z) {
  return 2.0f * (B200PPO_LOG2 - z - softplus_f(-2.0f * z));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct SamplerOut {
  float raw, action, llterm, regterm;
};

// NormalTanhSampler for one (row, action-dim) element — sampling_layers.py:88-147.
__device__ __forceinline__ SamplerOut sampler_elem(float mu, float rho, float min_std, float std_scale,
                                                   float entropy_weight, int mode, float raw_in,
                                                   Key k_sample, Key k_ent, uint32_t j, bool want_reg) {
  SamplerOut o;
  const float sigma = (softplus_f(rho) + min_std) * std_scale;
  float z;
  if (mode & 1) {
    z = raw_in;                                   // LOSS_REPLAY: stored raw action
  } else if (mode & 2) {
    z = mu;                                       // deterministic
  } else {
    const float eps = bits_to_normal(random_bits_at(k_sample, j));
    z = __fadd_rn(mu, __fmul_rn(sigma, eps));
  }
  o.raw = z;
  o.action = tanhf(z);
  const float q = (z - mu) / sigma;
  o.llterm = -0.5f * q * q - (B200PPO_HALF_LOG_2PI + logf(sigma)) - log_det_jac(z);
  o.regterm = 0.0f;
  if (want_reg) {
    const float eps2 = bits_to_normal(random_bits_at(k_ent, j));
    const float zp = __fadd_rn(mu, __fmul_rn(sigma, eps2));
    o.regterm = -entropy_weight * (0.5f + B200PPO_HALF_LOG_2PI + logf(sigma) + log_det_jac(zp));
  }
  return o;
}

// ------------------------------------------------------------------------------------------
// reset of the synthetic env for one env (see oracle/env.py for the definition)
// ------------------------------------------------------------------------------------------
struct ResetScalars {
  Key k_base;
  int32_t counter;
  uint32_t term;
};
__device__ __forceinline__ ResetScalars synth_reset_scalars(Key key, int max_len) {
  ResetScalars r;
  r.k_base = split_at(key, 0u);
  const Key k_cnt = split_at(key, 1u);
  r.counter = randint_scalar(k_cnt, static_cast<uint32_t>(max_len / 2));
  r.term = k_cnt.a ^ k_cnt.b;
  return r;
}

// ------------------------------------------------------------------------------------------
// batched ("wide") rollout: the per-step launch sequence for networks / envs whose weights do not fit an SM's
// shared memory (csrc/recurrent_tc.cu, rollout_wide).  Shared between the two translation units.
// ------------------------------------------------------------------------------------------
struct RolloutWideArgs {
  const b200ppo_plan* plan;
  const float* Wenv;          // [(O + A)][O]
  int max_len, term_thresh16;
  const float* params; const float* mean; const float* std;
  const uint32_t* rng_state; const uint32_t* iter_keys;
  int T, B;
  float* env_obs; int32_t* env_counter; uint32_t* env_term;
  float* obs; float* raw_action; float* action; float* loglik; float* reward;
  uint8_t* done; uint8_t* trunc; float* next_obs_last;
  float* ws;
};
int64_t rollout_wide_ws_floats(const b200ppo_plan* plan, int B);
int rollout_wide_num_launches(const b200ppo_plan* plan, int T);
int rollout_wide(cudaStream_t s, const RolloutWideArgs& a);

static inline int cdiv(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace b200ppo
