// K1: policy step and the persistent fused rollout on the synthetic env.  sm_100a.
//
// Reference computation replaced: rollout.single_transition / unroll_env (rollout.py:11-73) with
// the network of make_mlp_actor_critic (factories.py:72-146): Normalizer.__call__
// (normalizer.py:63-96) -> actor Dense stack (feedforward.py:48-50) -> NormalTanhSampler
// (sampling_layers.py:82-147) and critic Dense stack -> value squeeze (adapter.py:98).
//
// Design: every env is an independent unit, so one CTA owns a tile of TE envs for ALL T steps
// (no grid-wide synchronisation, one launch per rollout).  Actor + env weights are staged once
// in shared memory and re-used for T steps; activations ping-pong between two smem tiles.
// Two Dense engines: rollout_synth_mma_kernel (default) runs a layer as warp-level tensor-core tiles
// (mma.sync m16n8k8, 3xTF32, weights in fragment order); rollout_synth_kernel / eval_synth_kernel /
// policy_step_kernel compute a 2-env x 4-column FFMA register tile per thread, and layers with fewer
// tiles than threads split their k range over the idle threads (dense_tile).  Networks whose weights do
// not fit shared memory go to the batched per-step path (rollout_wide, csrc/recurrent_tc.cu) when the
// caller passes a workspace (b200ppo_rollout_synth_ws).  The evaluation rollout (rollout.py:97-148) is
// the same loop without the transition record and with sticky done flags.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

using namespace b200ppo;

namespace {

constexpr int TE = 16;          // envs per CTA (16: two CTAs share an SM and hide each other's barrier / latency bubbles)
constexpr int NT = 256;         // threads per CTA
constexpr int SMEM_LIMIT = 227 * 1024;

// out[e][n] = act( sum_k in[e][k] * W[k][n] + bias[n] )  for e < TE, n < N.
// `in`/`out`/`scratch` are shared-memory tiles; W/bias live in shared memory (staged) or in global memory.
//
// Each work item is an RP-row x 4-column register tile.  When a layer has fewer items than the CTA has
// threads (a 64-wide layer on 16 envs is 128 items), the k range is split over the otherwise idle
// threads and the partial tiles are summed through `scratch` — the layer's latency is the k loop, so
// this is what shortens a rollout step.  Activation rows are read 4 k at a time (tile strides are
// multiples of 4 floats, see tile_ld) so a 4-k step is 2 + 4 vector loads for 32 FMAs.
template <int RP, int NC = 0>   // NC: compile-time layer width (the weight rows become immediate offsets), 0 = runtime
__device__ __forceinline__ void dense_accumulate(const float* __restrict__ x0, int ldin,
                                                 const float* __restrict__ W, int Nrt, int n0, int k0, int k1,
                                                 bool vec, float (&acc)[RP][4]) {
  const int N = NC ? NC : Nrt;
  int k = k0;
  if (vec) {
#pragma unroll 2
    for (; k + 4 <= k1; k += 4) {
      float4 xv[RP];
#pragma unroll
      for (int r = 0; r < RP; ++r) xv[r] = *reinterpret_cast<const float4*>(x0 + r * ldin + k);
      const float* wp = W + k * N + n0;
      const float4 w0 = *reinterpret_cast<const float4*>(wp);
      const float4 w1 = *reinterpret_cast<const float4*>(wp + N);
      const float4 w2 = *reinterpret_cast<const float4*>(wp + 2 * N);
      const float4 w3 = *reinterpret_cast<const float4*>(wp + 3 * N);
#pragma unroll
      for (int r = 0; r < RP; ++r) {
        acc[r][0] = fmaf(xv[r].x, w0.x, acc[r][0]); acc[r][1] = fmaf(xv[r].x, w0.y, acc[r][1]);
        acc[r][2] = fmaf(xv[r].x, w0.z, acc[r][2]); acc[r][3] = fmaf(xv[r].x, w0.w, acc[r][3]);
        acc[r][0] = fmaf(xv[r].y, w1.x, acc[r][0]); acc[r][1] = fmaf(xv[r].y, w1.y, acc[r][1]);
        acc[r][2] = fmaf(xv[r].y, w1.z, acc[r][2]); acc[r][3] = fmaf(xv[r].y, w1.w, acc[r][3]);
        acc[r][0] = fmaf(xv[r].z, w2.x, acc[r][0]); acc[r][1] = fmaf(xv[r].z, w2.y, acc[r][1]);
        acc[r][2] = fmaf(xv[r].z, w2.z, acc[r][2]); acc[r][3] = fmaf(xv[r].z, w2.w, acc[r][3]);
        acc[r][0] = fmaf(xv[r].w, w3.x, acc[r][0]); acc[r][1] = fmaf(xv[r].w, w3.y, acc[r][1]);
        acc[r][2] = fmaf(xv[r].w, w3.z, acc[r][2]); acc[r][3] = fmaf(xv[r].w, w3.w, acc[r][3]);
      }
    }
    for (; k < k1; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(W + k * N + n0);
#pragma unroll
      for (int r = 0; r < RP; ++r) {
        const float a0 = x0[r * ldin + k];
        acc[r][0] = fmaf(a0, w.x, acc[r][0]); acc[r][1] = fmaf(a0, w.y, acc[r][1]);
        acc[r][2] = fmaf(a0, w.z, acc[r][2]); acc[r][3] = fmaf(a0, w.w, acc[r][3]);
      }
    }
  } else {
    for (; k < k1; ++k) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w = (n0 + j < N) ? W[k * N + n0 + j] : 0.0f;
#pragma unroll
        for (int r = 0; r < RP; ++r) acc[r][j] = fmaf(x0[r * ldin + k], w, acc[r][j]);
      }
    }
  }
}

template <int RP>
__device__ __forceinline__ void dense_finish(float (&acc)[RP][4], const float* __restrict__ bias, int N, int n0,
                                             int e0, float* __restrict__ out, int ldout, int act, bool vec) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float b = (bias != nullptr && n0 + j < N) ? bias[n0 + j] : 0.0f;
#pragma unroll
    for (int r = 0; r < RP; ++r) acc[r][j] = act_fwd(acc[r][j] + b, act);
  }
  if (vec) {
#pragma unroll
    for (int r = 0; r < RP; ++r)
      *reinterpret_cast<float4*>(out + (e0 + r) * ldout + n0) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n0 + j < N) {
#pragma unroll
        for (int r = 0; r < RP; ++r) out[(e0 + r) * ldout + n0 + j] = acc[r][j];
      }
    }
  }
}

constexpr int RP = 2;   // rows per item: 1 row per thread was measured slower (twice the weight loads per FMA)

// Every thread of the CTA must call this (the split path has a barrier inside); the caller
// synchronises before `out` is read.  `log2S` > 0 (host-computed, ksplit_log2) selects the k-split
// path: N / 4 is a power of two there, so a thread's (part, row group, column group) are shifts.
__device__ __forceinline__ void dense_tile(const float* __restrict__ in, int ldin, int K,
                                           const float* __restrict__ W, const float* __restrict__ bias,
                                           int N, float* __restrict__ out, int ldout, int act,
                                           float* __restrict__ scratch, int log2S) {
  if (log2S == 0) {
    const int ng = (N + 3) >> 2;
    const int items = (TE / RP) * ng;
    const bool vec = (N & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) & 15) == 0) && ((ldin | ldout) & 3) == 0;
    for (int item = threadIdx.x; item < items; item += NT) {
      const int eg = item / ng, nq = item - eg * ng;
      float acc[RP][4] = {};
      dense_accumulate<RP>(in + eg * RP * ldin, ldin, W, N, nq * 4, 0, K, vec, acc);
      dense_finish<RP>(acc, bias, N, nq * 4, eg * RP, out, ldout, act, vec);
    }
    return;
  }
  static_assert(TE / RP == 8, "row-group decomposition below assumes 8 row groups");
  const int ng = N >> 2;
  const int q = static_cast<int>(threadIdx.x) >> (__ffs(ng) - 1);
  const int nq = threadIdx.x & (ng - 1), eg = q & 7, part = q >> 3;
  const int e0 = eg * RP, n0 = nq * 4;
  const int S = 1 << log2S;
  const int kc = (((K + S - 1) >> log2S) + 3) & ~3;    // k per part, a multiple of 4 (vector loads of `in`)
  float acc[RP][4] = {};
  if (part < S) {
    const int k0 = part * kc, k1 = min(K, k0 + kc);
    if (N == 64) dense_accumulate<RP, 64>(in + e0 * ldin, ldin, W, N, n0, k0, k1, true, acc);
    else dense_accumulate<RP>(in + e0 * ldin, ldin, W, N, n0, k0, k1, true, acc);
    if (part > 0) {
#pragma unroll
      for (int r = 0; r < RP; ++r)
        *reinterpret_cast<float4*>(scratch + ((part - 1) * TE + e0 + r) * N + n0) =
            make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
  }
  __syncthreads();
  if (part == 0) {
    for (int p = 1; p < S; ++p) {                        // fixed order: the sum is deterministic
#pragma unroll
      for (int r = 0; r < RP; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(scratch + ((p - 1) * TE + e0 + r) * N + n0);
        acc[r][0] += v.x; acc[r][1] += v.y; acc[r][2] += v.z; acc[r][3] += v.w;
      }
    }
    dense_finish<RP>(acc, bias, N, n0, e0, out, ldout, act, true);
  }
}

// Runs one Dense chain on a smem tile; returns the buffer holding the last layer's output.
// split[l]: log2 of layer l's k-split factor (0 when there is no scratch tile).
__device__ __forceinline__ float* run_chain(const b200ppo_chain& ch, const float* __restrict__ P,
                                            float* bufA, float* bufB, int ld, float* scratch,
                                            const int8_t* __restrict__ split) {
  float* cur = bufA;
  float* nxt = bufB;
  for (int l = 0; l < ch.n_layers; ++l) {
    const int act = (l + 1 < ch.n_layers) ? ch.act : B200PPO_ACT_NONE;
    dense_tile(cur, ld, ch.dims[l], P + ch.w_off[l], P + ch.b_off[l], ch.dims[l + 1], nxt, ld, act, scratch,
               scratch != nullptr ? split[l] : 0);
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
  return cur;
}

__global__ void synth_init_keys_kernel(Key k, int B, uint32_t* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    const Key o = split_at(k, static_cast<uint32_t>(i));
    keys[2 * i] = o.a;
    keys[2 * i + 1] = o.b;
  }
}

__global__ void split_keys_dev_kernel(const uint32_t* __restrict__ key, int64_t first, int count,
                                      uint32_t* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    const Key o = split_at(Key{key[0], key[1]}, static_cast<uint32_t>(first + i));
    keys[2 * i] = o.a;
    keys[2 * i + 1] = o.b;
  }
}

// element `index` of jax.random.split(key_i) for every row key (vmapped env.reset key flows, wrappers)
__global__ void split_rows_kernel(const uint32_t* __restrict__ keys, int rows, uint32_t index, uint32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) {
    const Key o = split_at(Key{keys[2 * i], keys[2 * i + 1]}, index);
    out[2 * i] = o.a;
    out[2 * i + 1] = o.b;
  }
}

__global__ void synth_reset_kernel(const uint32_t* __restrict__ keys, int B, int O, int max_len,
                                   float* __restrict__ obs, int32_t* __restrict__ counter,
                                   uint32_t* __restrict__ term) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<int64_t>(B) * O) return;
  const int e = static_cast<int>(idx / O), o = static_cast<int>(idx - static_cast<int64_t>(e) * O);
  const Key key{keys[2 * e], keys[2 * e + 1]};
  const Key kb = split_at(key, 0u);
  obs[idx] = bits_to_normal(random_bits_at(kb, static_cast<uint32_t>(o)));
  if (o == 0) {
    const ResetScalars r = synth_reset_scalars(key, max_len);
    counter[e] = r.counter;
    term[e] = r.term;
  }
}

// ------------------------------------------------------------------------------------------
// persistent fused rollout
// ------------------------------------------------------------------------------------------
struct RolloutArgs {
  b200ppo_plan plan;
  int O, A, max_len, term_thresh16;
  const float* Wenv;          // [(O + A)][O]: rows 0..O-1 = Wo, rows O.. = Wa  (global)
  const float* params;
  const float* mean;
  const float* std;
  const uint32_t* rng_state;
  const uint32_t* iter_keys;
  int T, B;
  float* env_obs; int32_t* env_counter; uint32_t* env_term;
  float* obs; float* raw_action; float* action; float* loglik; float* reward;
  uint8_t* done; uint8_t* trunc; float* next_obs_last;
  int8_t split[B200PPO_MAX_LAYERS + 1];   // log2 k-split per actor layer; [MAX_LAYERS] = the env step
  int use_scratch;            // the scratch tile of the k-split layers fits in shared memory
  int stage_actor;            // actor parameters staged in smem
  int stage_env;              // env weights staged in smem
  int actor_span;             // floats of the arena covered by the actor chain (starts at 0)
  int ld;                     // row stride of the activation tiles
};

// SA / SE: actor parameters / env weights staged in shared memory.  Compile-time so that the weight
// loads of the k loops are LDS (or LDG) instead of generic loads with 64-bit address arithmetic.
template <bool SA, bool SE>
__global__ void __launch_bounds__(NT, 2) rollout_synth_kernel(const RolloutArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int O = a.O, A = a.A, ld = a.ld;
  const int env0 = blockIdx.x * TE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sp = smem;
  float* bufA = sp; sp += TE * ld;
  float* bufB = sp; sp += TE * ld;
  float* bufC = nullptr;                // partial tiles of the k-split layers (dropped for very wide tiles)
  if (a.use_scratch) { bufC = sp; sp += TE * ld; }
  float* bufX = sp; sp += TE * ld;      // env-step input tile, resident across steps: [obs | action] per env row
  float* raw_s = sp; sp += TE * A;
  float* llt_s = sp; sp += TE * A;
  float* mean_s = sp; sp += O;
  float* std_s = sp; sp += O;
  int32_t* cnt_s = reinterpret_cast<int32_t*>(sp); sp += TE;
  uint32_t* term_s = reinterpret_cast<uint32_t*>(sp); sp += TE;
  uint32_t* kb_s = reinterpret_cast<uint32_t*>(sp); sp += 2 * TE;
  int32_t* done_s = reinterpret_cast<int32_t*>(sp); sp += TE;
  sp = smem + ((sp - smem + 3) & ~3);
  float* ps = sp; if (SA) sp += (a.actor_span + 3) & ~3;
  float* ws = sp; if (SE) sp += (O + A) * O;
  if (SA) for (int i = threadIdx.x; i < a.actor_span; i += NT) ps[i] = a.params[i];
  if (SE) for (int i = threadIdx.x; i < (O + A) * O; i += NT) ws[i] = a.Wenv[i];
  const float* P = SA ? ps : a.params;
  const float* Wenv = SE ? ws : a.Wenv;
  for (int i = threadIdx.x; i < O; i += NT) {
    mean_s[i] = a.plan.normalize ? a.mean[i] : 0.0f;
    std_s[i] = a.plan.normalize ? a.std[i] : 1.0f;
  }
  // tile loops: one warp per env row, lanes over the row (no integer division by a runtime width)
  for (int e = warp; e < TE; e += NT / 32)
    for (int o = lane; o < O; o += 32)
      bufX[e * ld + o] = (env0 + e < a.B) ? a.env_obs[static_cast<size_t>(env0 + e) * O + o] : 0.0f;
  if (threadIdx.x < TE) {
    const bool ok = env0 + threadIdx.x < a.B;
    cnt_s[threadIdx.x] = ok ? a.env_counter[env0 + threadIdx.x] : 0;
    term_s[threadIdx.x] = ok ? a.env_term[env0 + threadIdx.x] : 0u;
  }
  const Key stream_key{a.rng_state[0], a.rng_state[1]};
  const uint32_t count0 = a.rng_state[2];
  const Key reset_key{a.iter_keys[0], a.iter_keys[1]};
  __syncthreads();

  for (int t = 0; t < a.T; ++t) {
    // (1) record the raw observation, normalise into bufA  (rollout.py:23; normalizer.py:78-80)
    const size_t row0 = static_cast<size_t>(t) * a.B + env0;
    for (int e = warp; e < TE; e += NT / 32)
      for (int o = lane; o < O; o += 32) {
        const float x = bufX[e * ld + o];
        if (env0 + e < a.B) a.obs[(row0 + e) * O + o] = x;
        bufA[e * ld + o] = a.plan.normalize ? __fdiv_rn(x - mean_s[o], std_s[o]) : x;
      }
    __syncthreads();
    // (2) actor MLP
    float* y = run_chain(a.plan.actor, P, bufA, bufB, ld, bufC, a.split);
    // (3) sampler: one fresh normal draw per (env, action dim); count = count0 + 2t
    //     (the entropy draw at count0 + 2t + 1 does not influence the rollout and is skipped).
    //     The action goes straight into the env-step input tile.
    const Key k_sample = fold_in(stream_key, count0 + 2u * static_cast<uint32_t>(t));
    for (int i = threadIdx.x; i < TE * A; i += NT) {
      const int e = i / A, d = i - e * A;
      const uint32_t j = static_cast<uint32_t>(env0 + e) * static_cast<uint32_t>(A) + d;
      const SamplerOut s = sampler_elem(y[e * ld + d], y[e * ld + A + d], a.plan.min_std,
                                        a.plan.std_scale, a.plan.entropy_weight, 0, 0.0f, k_sample,
                                        k_sample, j, false);
      raw_s[i] = s.raw;
      bufX[e * ld + O + d] = s.action;
      llt_s[i] = s.llterm;
      if (env0 + e < a.B) {
        a.raw_action[row0 * A + i] = s.raw;
        a.action[row0 * A + i] = s.action;
      }
    }
    __syncthreads();
    // (4) episode bookkeeping (integer-exact) and reset scalars: independent of the env arithmetic, so the
    //     16 threads that own an env do it while the others already start the env-step GEMM
    if (threadIdx.x < TE) {
      const int e = threadIdx.x;
      const int ge = env0 + e;
      float ll = 0.0f;
      for (int d = 0; d < A; ++d) ll += llt_s[e * A + d];
      const int32_t c = cnt_s[e] + 1;
      const uint32_t ts = term_s[e] * 1664525u + 1013904223u;
      const bool terminated = (ts >> 16) < static_cast<uint32_t>(a.term_thresh16);
      const bool truncated = c >= a.max_len;
      const bool dn = terminated || truncated;
      if (ge < a.B) {
        a.loglik[row0 + e] = ll;
        a.done[row0 + e] = dn ? 1 : 0;
        a.trunc[row0 + e] = truncated ? 1 : 0;
      }
      done_s[e] = dn ? 1 : 0;
      if (dn) {
        // rng_keys_for_env_reset[t][env] = split(reset_key, (T, B))[t, env]  (rollout.py:57-59)
        const Key k = split_at(reset_key, static_cast<uint32_t>(t) * static_cast<uint32_t>(a.B) +
                                              static_cast<uint32_t>(ge));
        const ResetScalars r = synth_reset_scalars(k, a.max_len);
        cnt_s[e] = r.counter;
        term_s[e] = r.term;
        kb_s[2 * e] = r.k_base.a;
        kb_s[2 * e + 1] = r.k_base.b;
      } else {
        cnt_s[e] = c;
        term_s[e] = ts;
      }
    }
    // (5) env step: obs' = tanh([obs, action] @ [Wo; Wa]) into the (dead) actor-output tile
    dense_tile(bufX, ld, O + A, Wenv, nullptr, O, y, ld, B200PPO_ACT_TANH, bufC, a.split[B200PPO_MAX_LAYERS]);
    __syncthreads();
    // (6) one warp per env: reward = -mean(obs'^2); next_obs[-1] is the pre-reset observation
    //     (rollout.py:30); then tree_where(done, reset, next) back into the input tile
    for (int e = warp; e < TE; e += NT / 32) {
      const bool dn = done_s[e] != 0;
      const Key kb{kb_s[2 * e], kb_s[2 * e + 1]};
      float sq = 0.0f;
      for (int o = lane; o < O; o += 32) {
        const float v = y[e * ld + o];
        sq = fmaf(v, v, sq);
        if (t == a.T - 1 && env0 + e < a.B) a.next_obs_last[static_cast<size_t>(env0 + e) * O + o] = v;
        bufX[e * ld + o] = dn ? bits_to_normal(random_bits_at(kb, static_cast<uint32_t>(o))) : v;
      }
      sq = warp_sum(sq);
      if (lane == 0 && env0 + e < a.B) a.reward[row0 + e] = -(sq / static_cast<float>(O));
    }
    __syncthreads();
  }
  for (int e = warp; e < TE; e += NT / 32)
    for (int o = lane; o < O; o += 32)
      if (env0 + e < a.B) a.env_obs[static_cast<size_t>(env0 + e) * O + o] = bufX[e * ld + o];
  if (threadIdx.x < TE && env0 + threadIdx.x < a.B) {
    a.env_counter[env0 + threadIdx.x] = cnt_s[threadIdx.x];
    a.env_term[env0 + threadIdx.x] = term_s[threadIdx.x];
  }
}

// ------------------------------------------------------------------------------------------
// The same rollout with the Dense layers on the tensor cores: warp-level mma.sync m16n8k8 (TF32 operands, fp32
// accumulate), error-compensated with the 3-product split (x = hi + lo: lo*hi + hi*lo + hi*hi), so the result
// keeps fp32-level accuracy.  A tile of TE = 16 envs is exactly one m16 row tile; a layer's 8-column n-tiles
// go round-robin over the 8 warps.  Why not tcgen05 here: a CTA-wide tcgen05 tile is 128 envs, i.e. 32 CTAs
// for 4096 envs, and every layer would pay the commit -> mbarrier -> TMEM-load hand-off (~2.5 K cycles in the
// update kernels) on a chain of 6 dependent layers per step; the warp-level MMA has a ~30-cycle dependent
// latency and keeps 256 CTAs (2 per SM) busy.  What the tensor path buys over the FFMA tiles above is operand
// traffic: the FFMA version needs 6 LDS.128 per 32 FMAs (the LSU pipe, not the FMA pipe, was its bound), here
// a k-step of 1024 MACs takes 4 LDS.32 + 1 LDS.64.
// Weights are staged once per CTA in FRAGMENT order: for n-tile j and k-step s the 32 lanes' (b0, b1) pairs
// are 256 contiguous bytes (lane = 4 g + t holds W[8s + t][8j + g], W[8s + t + 4][8j + g]); rows / columns
// beyond the layer's K / N are zero, so ragged widths need no predicates in the loop.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// floats of one layer's fragment-ordered weights
__host__ __device__ inline int frag_floats(int K, int N) { return ((K + 7) & ~7) * ((N + 7) & ~7); }

// stage W [K][N] (row-major, global) into fragment order
__device__ __forceinline__ void stage_frags(const float* __restrict__ W, int K, int N, float2* __restrict__ dst) {
  const int KS = (K + 7) >> 3, NJ = (N + 7) >> 3;
  for (int idx = threadIdx.x; idx < NJ * KS * 32; idx += NT) {
    const int lane = idx & 31, q = idx >> 5, s = q % KS, j = q / KS;
    const int k = 8 * s + (lane & 3), n = 8 * j + (lane >> 2);
    float2 v = make_float2(0.0f, 0.0f);
    if (n < N) {
      if (k < K) v.x = W[static_cast<size_t>(k) * N + n];
      if (k + 4 < K) v.y = W[static_cast<size_t>(k + 4) * N + n];
    }
    dst[idx] = v;
  }
}

// out[e][n] = act(sum_k in[e][k] W[k][n] + bias[n]) for the 16 rows of the tile; `in` columns >= K are finite
// (zero-initialised padding), `out` columns up to the padded width are written (padding -> act(0) = 0).
// Two accumulator sets (even / odd k-steps) of three products each: six independent MMA chains per warp.
__device__ __forceinline__ void mma_layer(const float* __restrict__ in, int ld, int K, const float2* __restrict__ wf,
                                          const float* __restrict__ bias /*padded, smem; nullptr: none*/, int N,
                                          float* __restrict__ out, int act) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int KS = (K + 7) >> 3, NJ = (N + 7) >> 3;
  const float* r0 = in + g * ld + t;
  const float* r1 = in + (g + 8) * ld + t;
  for (int j = warp; j < NJ; j += NT / 32) {
    float acc[2][3][4];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[p][q][i] = 0.0f;
    const float2* wj = wf + static_cast<size_t>(j) * KS * 32 + lane;
#pragma unroll 2
    for (int s = 0; s < KS; ++s) {
      const float x[4] = {r0[8 * s], r1[8 * s], r0[8 * s + 4], r1[8 * s + 4]};
      const float2 w = wj[s * 32];
      // split by truncation: the MMA ignores the low 13 mantissa bits of a tf32 operand, so hi is the masked
      // value (2 instructions per element: LOP3 + FADD; cvt.rna.tf32 is 4 + the FADD) and lo = x - hi is exact
      // in fp32, itself truncated by the MMA: |error| <= 2^-20 |x|
      uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ah[i] = __float_as_uint(x[i]) & 0xFFFFE000u;
        al[i] = __float_as_uint(x[i] - __uint_as_float(ah[i]));
      }
      bh[0] = __float_as_uint(w.x) & 0xFFFFE000u; bl[0] = __float_as_uint(w.x - __uint_as_float(bh[0]));
      bh[1] = __float_as_uint(w.y) & 0xFFFFE000u; bl[1] = __float_as_uint(w.y - __uint_as_float(bh[1]));
      const int p = s & 1;
      mma_tf32_16x8x8(acc[p][0], al, bh);
      mma_tf32_16x8x8(acc[p][1], ah, bl);
      mma_tf32_16x8x8(acc[p][2], ah, bh);
    }
    const int col = 8 * j + 2 * t;
    float z[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)     // small terms first
      z[i] = ((acc[0][0][i] + acc[1][0][i]) + (acc[0][1][i] + acc[1][1][i])) + (acc[0][2][i] + acc[1][2][i]);
    if (bias != nullptr) {
      const float b0 = bias[col], b1 = bias[col + 1];
      z[0] += b0; z[1] += b1; z[2] += b0; z[3] += b1;
    }
    *reinterpret_cast<float2*>(out + g * ld + col) = make_float2(act_fwd(z[0], act), act_fwd(z[1], act));
    *reinterpret_cast<float2*>(out + (g + 8) * ld + col) = make_float2(act_fwd(z[2], act), act_fwd(z[3], act));
  }
}

static_assert(TE == 16, "the tensor-core rollout maps a CTA's env tile onto one m16 MMA row tile");

__global__ void __launch_bounds__(NT, 2) rollout_synth_mma_kernel(const RolloutArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int O = a.O, A = a.A, ld = a.ld;
  const int env0 = blockIdx.x * TE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const b200ppo_chain& ch = a.plan.actor;
  const int L = ch.n_layers;
  float* sp = smem;
  float* bufA = sp; sp += TE * ld;
  float* bufB = sp; sp += TE * ld;
  float* bufX = sp; sp += TE * ld;      // env-step input tile, resident across steps: [obs | action] per env row
  float* raw_s = sp; sp += TE * A;
  float* llt_s = sp; sp += TE * A;
  float* mean_s = sp; sp += O;
  float* std_s = sp; sp += O;
  int32_t* cnt_s = reinterpret_cast<int32_t*>(sp); sp += TE;
  uint32_t* term_s = reinterpret_cast<uint32_t*>(sp); sp += TE;
  uint32_t* kb_s = reinterpret_cast<uint32_t*>(sp); sp += 2 * TE;
  int32_t* done_s = reinterpret_cast<int32_t*>(sp); sp += TE;
  sp = smem + ((sp - smem + 3) & ~3);
  __shared__ int woff_s[B200PPO_MAX_LAYERS + 1], boff_s[B200PPO_MAX_LAYERS];
  // fragment-ordered weights and padded biases of every actor layer, then the env's [Wo; Wa]
  float* wbase = sp;
  {
    int off = 0;
    for (int l = 0; l < L; ++l) {
      if (threadIdx.x == 0) woff_s[l] = off;
      off += frag_floats(ch.dims[l], ch.dims[l + 1]);
    }
    if (threadIdx.x == 0) woff_s[L] = off;
    off += frag_floats(O + A, O);
    for (int l = 0; l < L; ++l) {
      if (threadIdx.x == 0) boff_s[l] = off;
      off += (ch.dims[l + 1] + 7) & ~7;
    }
  }
  for (int i = threadIdx.x; i < 3 * TE * ld; i += NT) smem[i] = 0.0f;      // padding columns stay zero
  __syncthreads();
  for (int l = 0; l < L; ++l) {
    stage_frags(a.params + ch.w_off[l], ch.dims[l], ch.dims[l + 1], reinterpret_cast<float2*>(wbase + woff_s[l]));
    const int N = ch.dims[l + 1], Np = (N + 7) & ~7;
    for (int i = threadIdx.x; i < Np; i += NT) wbase[boff_s[l] + i] = i < N ? a.params[ch.b_off[l] + i] : 0.0f;
  }
  stage_frags(a.Wenv, O + A, O, reinterpret_cast<float2*>(wbase + woff_s[L]));
  for (int i = threadIdx.x; i < O; i += NT) {
    mean_s[i] = a.plan.normalize ? a.mean[i] : 0.0f;
    std_s[i] = a.plan.normalize ? a.std[i] : 1.0f;
  }
  for (int e = warp; e < TE; e += NT / 32)
    for (int o = lane; o < O; o += 32)
      bufX[e * ld + o] = (env0 + e < a.B) ? a.env_obs[static_cast<size_t>(env0 + e) * O + o] : 0.0f;
  if (threadIdx.x < TE) {
    const bool ok = env0 + threadIdx.x < a.B;
    cnt_s[threadIdx.x] = ok ? a.env_counter[env0 + threadIdx.x] : 0;
    term_s[threadIdx.x] = ok ? a.env_term[env0 + threadIdx.x] : 0u;
  }
  const Key stream_key{a.rng_state[0], a.rng_state[1]};
  const uint32_t count0 = a.rng_state[2];
  const Key reset_key{a.iter_keys[0], a.iter_keys[1]};
  __syncthreads();

  for (int t = 0; t < a.T; ++t) {
    // (1) record the raw observation, normalise into bufA  (rollout.py:23; normalizer.py:78-80)
    const size_t row0 = static_cast<size_t>(t) * a.B + env0;
    for (int e = warp; e < TE; e += NT / 32)
      for (int o = lane; o < O; o += 32) {
        const float x = bufX[e * ld + o];
        if (env0 + e < a.B) a.obs[(row0 + e) * O + o] = x;
        bufA[e * ld + o] = a.plan.normalize ? __fdiv_rn(x - mean_s[o], std_s[o]) : x;
      }
    __syncthreads();
    // (2) actor MLP: activations ping-pong between bufA and bufB
    float* cur = bufA;
    float* nxt = bufB;
    for (int l = 0; l < L; ++l) {
      mma_layer(cur, ld, ch.dims[l], reinterpret_cast<const float2*>(wbase + woff_s[l]), wbase + boff_s[l],
                ch.dims[l + 1], nxt, l + 1 < L ? ch.act : B200PPO_ACT_NONE);
      __syncthreads();
      float* tmp = cur; cur = nxt; nxt = tmp;
    }
    float* y = cur;
    // (3) sampler (count = count0 + 2t; the entropy draw does not influence the rollout): action -> env input tile
    const Key k_sample = fold_in(stream_key, count0 + 2u * static_cast<uint32_t>(t));
    for (int i = threadIdx.x; i < TE * A; i += NT) {
      const int e = i / A, d = i - e * A;
      const uint32_t j = static_cast<uint32_t>(env0 + e) * static_cast<uint32_t>(A) + d;
      const SamplerOut s = sampler_elem(y[e * ld + d], y[e * ld + A + d], a.plan.min_std,
                                        a.plan.std_scale, a.plan.entropy_weight, 0, 0.0f, k_sample,
                                        k_sample, j, false);
      raw_s[i] = s.raw;
      bufX[e * ld + O + d] = s.action;
      llt_s[i] = s.llterm;
      if (env0 + e < a.B) {
        a.raw_action[row0 * A + i] = s.raw;
        a.action[row0 * A + i] = s.action;
      }
    }
    __syncthreads();
    // (4) episode bookkeeping (integer-exact) and reset scalars, beside the env-step GEMM
    if (threadIdx.x < TE) {
      const int e = threadIdx.x;
      const int ge = env0 + e;
      float ll = 0.0f;
      for (int d = 0; d < A; ++d) ll += llt_s[e * A + d];
      const int32_t c = cnt_s[e] + 1;
      const uint32_t ts = term_s[e] * 1664525u + 1013904223u;
      const bool terminated = (ts >> 16) < static_cast<uint32_t>(a.term_thresh16);
      const bool truncated = c >= a.max_len;
      const bool dn = terminated || truncated;
      if (ge < a.B) {
        a.loglik[row0 + e] = ll;
        a.done[row0 + e] = dn ? 1 : 0;
        a.trunc[row0 + e] = truncated ? 1 : 0;
      }
      done_s[e] = dn ? 1 : 0;
      if (dn) {
        const Key k = split_at(reset_key, static_cast<uint32_t>(t) * static_cast<uint32_t>(a.B) +
                                              static_cast<uint32_t>(ge));
        const ResetScalars r = synth_reset_scalars(k, a.max_len);
        cnt_s[e] = r.counter;
        term_s[e] = r.term;
        kb_s[2 * e] = r.k_base.a;
        kb_s[2 * e + 1] = r.k_base.b;
      } else {
        cnt_s[e] = c;
        term_s[e] = ts;
      }
    }
    // (5) env step: obs' = tanh([obs, action] @ [Wo; Wa]) into the tile the actor no longer needs
    float* yo = nxt;
    mma_layer(bufX, ld, O + A, reinterpret_cast<const float2*>(wbase + woff_s[L]), nullptr, O, yo, B200PPO_ACT_TANH);
    __syncthreads();
    // (6) reward = -mean(obs'^2); next_obs[-1] is the pre-reset observation; tree_where(done, reset, next)
    for (int e = warp; e < TE; e += NT / 32) {
      const bool dn = done_s[e] != 0;
      const Key kb{kb_s[2 * e], kb_s[2 * e + 1]};
      float sq = 0.0f;
      for (int o = lane; o < O; o += 32) {
        const float v = yo[e * ld + o];
        sq = fmaf(v, v, sq);
        if (t == a.T - 1 && env0 + e < a.B) a.next_obs_last[static_cast<size_t>(env0 + e) * O + o] = v;
        bufX[e * ld + o] = dn ? bits_to_normal(random_bits_at(kb, static_cast<uint32_t>(o))) : v;
      }
      sq = warp_sum(sq);
      if (lane == 0 && env0 + e < a.B) a.reward[row0 + e] = -(sq / static_cast<float>(O));
    }
    __syncthreads();
  }
  for (int e = warp; e < TE; e += NT / 32)
    for (int o = lane; o < O; o += 32)
      if (env0 + e < a.B) a.env_obs[static_cast<size_t>(env0 + e) * O + o] = bufX[e * ld + o];
  if (threadIdx.x < TE && env0 + threadIdx.x < a.B) {
    a.env_counter[env0 + threadIdx.x] = cnt_s[threadIdx.x];
    a.env_term[env0 + threadIdx.x] = term_s[threadIdx.x];
  }
}

// ------------------------------------------------------------------------------------------
// Second version of the tensor-core rollout: ONE 512-thread CTA per SM holding two independent 8-warp env groups
// (16 envs each, own activation tiles, own named barrier) that share one copy of the fragment-ordered weights.  The room
// this frees (no second copy of the 88 KB of weights) holds every activation tile PRE-SPLIT (a hi tile and a lo tile,
// written once by the producing epilogue), so a k-step's A fragments are two `ldmatrix.x4` instead of four LDS.32 plus
// eight split instructions that all eight warps of a group repeated: 13 instead of 23 issued instructions per k-step
// (profiles/r2b_notes.md).  The two groups hide each other's barrier and latency bubbles the way the two CTAs per SM of
// the first version did.
// ------------------------------------------------------------------------------------------
constexpr int NT2 = 512;

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
// value of a pre-split tile element / store of one (hi = the tf32-truncated value, lo = the exact remainder: hi + lo == v)
__device__ __forceinline__ float tile_get(const float* t, int tlo, int i) { return t[i] + t[i + tlo]; }
__device__ __forceinline__ void tile_put(float* t, int tlo, int i, float v) {
  const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  t[i] = hi;
  t[i + tlo] = v - hi;
}

__device__ __forceinline__ void mma_layer2(const float* __restrict__ in, int tlo, int ld, int K, const float2* __restrict__ wf,
                                           const float* __restrict__ bias, int N, float* __restrict__ out, int act, int lwarp,
                                           int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int KS = (K + 7) >> 3, NJ = (N + 7) >> 3;
  // ldmatrix row addresses of the m16 x k8 fp32 fragment: matrices 0..3 = (rows 0-7 | 8-15) x (k 0-3 | 4-7)
  const int frow = (lane & 7) + ((lane >> 3) & 1) * 8, fcol = (lane >> 4) * 4;
  const uint32_t a_hi = static_cast<uint32_t>(__cvta_generic_to_shared(in + frow * ld + fcol));
  const uint32_t a_lo = a_hi + static_cast<uint32_t>(tlo) * 4u;
  for (int j = lwarp; j < NJ; j += 8) {
    float acc[2][3][4];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[p][q][i] = 0.0f;
    const float2* wj = wf + static_cast<size_t>(j) * KS * 32 + lane;
#pragma unroll 2
    for (int s = 0; s < KS; ++s) {
      uint32_t ah[4], al[4], bh[2], bl[2];
      ldmatrix_x4(ah, a_hi + 32u * s);
      ldmatrix_x4(al, a_lo + 32u * s);
      const float2 w = wj[s * 32];
      bh[0] = __float_as_uint(w.x) & 0xFFFFE000u; bl[0] = __float_as_uint(w.x - __uint_as_float(bh[0]));
      bh[1] = __float_as_uint(w.y) & 0xFFFFE000u; bl[1] = __float_as_uint(w.y - __uint_as_float(bh[1]));
      const int p = s & 1;
      mma_tf32_16x8x8(acc[p][0], al, bh);
      mma_tf32_16x8x8(acc[p][1], ah, bl);
      mma_tf32_16x8x8(acc[p][2], ah, bh);
    }
    const int col = 8 * j + 2 * t;
    float z[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)     // small terms first
      z[i] = ((acc[0][0][i] + acc[1][0][i]) + (acc[0][1][i] + acc[1][1][i])) + (acc[0][2][i] + acc[1][2][i]);
    if (bias != nullptr) {
      const float b0 = bias[col], b1 = bias[col + 1];
      z[0] += b0; z[1] += b1; z[2] += b0; z[3] += b1;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) z[i] = act_fwd(z[i], act);
    float h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __uint_as_float(__float_as_uint(z[i]) & 0xFFFFE000u);
    float* o0 = out + g * ld + col;
    float* o1 = out + (g + 8) * ld + col;
    *reinterpret_cast<float2*>(o0) = make_float2(h[0], h[1]);
    *reinterpret_cast<float2*>(o1) = make_float2(h[2], h[3]);
    *reinterpret_cast<float2*>(o0 + tlo) = make_float2(z[0] - h[0], z[1] - h[1]);
    *reinterpret_cast<float2*>(o1 + tlo) = make_float2(z[2] - h[2], z[3] - h[3]);
  }
}

__global__ void __launch_bounds__(NT2, 1) rollout_synth_mma2_kernel(const RolloutArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int O = a.O, A = a.A, ld = a.ld;
  const int gi = threadIdx.x >> 8, lt = threadIdx.x & 255;        // env group, thread within the group
  const int env0 = (blockIdx.x * 2 + gi) * TE;
  const int warp = lt >> 5, lane = lt & 31;
  const b200ppo_chain& ch = a.plan.actor;
  const int L = ch.n_layers;
  const int tlo = TE * ld;                                          // lo tile follows its hi tile
  const int per_group = 6 * TE * ld + 2 * TE * A + 5 * TE;
  float* sp = smem + gi * ((per_group + 3) & ~3);
  float* bufA = sp; sp += 2 * TE * ld;
  float* bufB = sp; sp += 2 * TE * ld;
  float* bufX = sp; sp += 2 * TE * ld;  // env-step input tile, resident across steps: [obs | action] per env row
  float* raw_s = sp; sp += TE * A;
  float* llt_s = sp; sp += TE * A;
  int32_t* cnt_s = reinterpret_cast<int32_t*>(sp); sp += TE;
  uint32_t* term_s = reinterpret_cast<uint32_t*>(sp); sp += TE;
  uint32_t* kb_s = reinterpret_cast<uint32_t*>(sp); sp += 2 * TE;
  int32_t* done_s = reinterpret_cast<int32_t*>(sp); sp += TE;
  float* shared0 = smem + 2 * ((per_group + 3) & ~3);
  float* mean_s = shared0;
  float* std_s = mean_s + O;
  float* wbase = shared0 + ((2 * O + 3) & ~3);
  __shared__ int woff_s[B200PPO_MAX_LAYERS + 1], boff_s[B200PPO_MAX_LAYERS];
  {
    int off = 0;
    for (int l = 0; l < L; ++l) {
      if (threadIdx.x == 0) woff_s[l] = off;
      off += frag_floats(ch.dims[l], ch.dims[l + 1]);
    }
    if (threadIdx.x == 0) woff_s[L] = off;
    off += frag_floats(O + A, O);
    for (int l = 0; l < L; ++l) {
      if (threadIdx.x == 0) boff_s[l] = off;
      off += (ch.dims[l + 1] + 7) & ~7;
    }
  }
  for (int i = lt; i < 6 * TE * ld; i += 256) bufA[i] = 0.0f;        // all six tiles of the group; padding stays zero
  __syncthreads();
  for (int l = 0; l < L; ++l) {
    const int K = ch.dims[l], N = ch.dims[l + 1], Np = (N + 7) & ~7;
    const int KS = (K + 7) >> 3, NJ = (N + 7) >> 3;
    float2* dst = reinterpret_cast<float2*>(wbase + woff_s[l]);
    const float* W = a.params + ch.w_off[l];
    for (int idx = threadIdx.x; idx < NJ * KS * 32; idx += NT2) {
      const int ln = idx & 31, q = idx >> 5, s = q % KS, j = q / KS;
      const int k = 8 * s + (ln & 3), n = 8 * j + (ln >> 2);
      float2 v = make_float2(0.0f, 0.0f);
      if (n < N) {
        if (k < K) v.x = W[static_cast<size_t>(k) * N + n];
        if (k + 4 < K) v.y = W[static_cast<size_t>(k + 4) * N + n];
      }
      dst[idx] = v;
    }
    for (int i = threadIdx.x; i < Np; i += NT2) wbase[boff_s[l] + i] = i < N ? a.params[ch.b_off[l] + i] : 0.0f;
  }
  {
    const int K = O + A, N = O, KS = (K + 7) >> 3, NJ = (N + 7) >> 3;
    float2* dst = reinterpret_cast<float2*>(wbase + woff_s[L]);
    for (int idx = threadIdx.x; idx < NJ * KS * 32; idx += NT2) {
      const int ln = idx & 31, q = idx >> 5, s = q % KS, j = q / KS;
      const int k = 8 * s + (ln & 3), n = 8 * j + (ln >> 2);
      float2 v = make_float2(0.0f, 0.0f);
      if (n < N) {
        if (k < K) v.x = a.Wenv[static_cast<size_t>(k) * N + n];
        if (k + 4 < K) v.y = a.Wenv[static_cast<size_t>(k + 4) * N + n];
      }
      dst[idx] = v;
    }
  }
  for (int i = threadIdx.x; i < O; i += NT2) {
    mean_s[i] = a.plan.normalize ? a.mean[i] : 0.0f;
    std_s[i] = a.plan.normalize ? a.std[i] : 1.0f;
  }
  for (int e = warp; e < TE; e += 8)
    for (int o = lane; o < O; o += 32)
      tile_put(bufX, tlo, e * ld + o, (env0 + e < a.B) ? a.env_obs[static_cast<size_t>(env0 + e) * O + o] : 0.0f);
  if (lt < TE) {
    const bool ok = env0 + lt < a.B;
    cnt_s[lt] = ok ? a.env_counter[env0 + lt] : 0;
    term_s[lt] = ok ? a.env_term[env0 + lt] : 0u;
  }
  const Key stream_key{a.rng_state[0], a.rng_state[1]};
  const uint32_t count0 = a.rng_state[2];
  const Key reset_key{a.iter_keys[0], a.iter_keys[1]};
  __syncthreads();
  // from here on the two groups never meet again: barrier 1 + gi, 256 threads
#define GROUP_SYNC() asm volatile("bar.sync %0, 256;" ::"r"(1 + gi) : "memory")

  for (int t = 0; t < a.T; ++t) {
    // (1) record the raw observation, normalise into bufA  (rollout.py:23; normalizer.py:78-80)
    const size_t row0 = static_cast<size_t>(t) * a.B + env0;
    for (int e = warp; e < TE; e += 8)
      for (int o = lane; o < O; o += 32) {
        const float x = tile_get(bufX, tlo, e * ld + o);
        if (env0 + e < a.B) a.obs[(row0 + e) * O + o] = x;
        tile_put(bufA, tlo, e * ld + o, a.plan.normalize ? __fdiv_rn(x - mean_s[o], std_s[o]) : x);
      }
    GROUP_SYNC();
    // (2) actor MLP: activations ping-pong between bufA and bufB
    float* cur = bufA;
    float* nxt = bufB;
    for (int l = 0; l < L; ++l) {
      mma_layer2(cur, tlo, ld, ch.dims[l], reinterpret_cast<const float2*>(wbase + woff_s[l]), wbase + boff_s[l],
                 ch.dims[l + 1], nxt, l + 1 < L ? ch.act : B200PPO_ACT_NONE, warp, lane);
      GROUP_SYNC();
      float* tmp = cur; cur = nxt; nxt = tmp;
    }
    const float* y = cur;
    // (3) sampler (count = count0 + 2t; the entropy draw does not influence the rollout): action -> env input tile
    const Key k_sample = fold_in(stream_key, count0 + 2u * static_cast<uint32_t>(t));
    for (int i = lt; i < TE * A; i += 256) {
      const int e = i / A, d = i - e * A;
      const uint32_t j = static_cast<uint32_t>(env0 + e) * static_cast<uint32_t>(A) + d;
      const SamplerOut s = sampler_elem(tile_get(y, tlo, e * ld + d), tile_get(y, tlo, e * ld + A + d), a.plan.min_std,
                                        a.plan.std_scale, a.plan.entropy_weight, 0, 0.0f, k_sample, k_sample, j, false);
      raw_s[i] = s.raw;
      tile_put(bufX, tlo, e * ld + O + d, s.action);
      llt_s[i] = s.llterm;
      if (env0 + e < a.B) {
        a.raw_action[row0 * A + i] = s.raw;
        a.action[row0 * A + i] = s.action;
      }
    }
    GROUP_SYNC();
    // (4) episode bookkeeping (integer-exact) and reset scalars, beside the env-step GEMM
    if (lt < TE) {
      const int e = lt;
      const int ge = env0 + e;
      float ll = 0.0f;
      for (int d = 0; d < A; ++d) ll += llt_s[e * A + d];
      const int32_t c = cnt_s[e] + 1;
      const uint32_t ts = term_s[e] * 1664525u + 1013904223u;
      const bool terminated = (ts >> 16) < static_cast<uint32_t>(a.term_thresh16);
      const bool truncated = c >= a.max_len;
      const bool dn = terminated || truncated;
      if (ge < a.B) {
        a.loglik[row0 + e] = ll;
        a.done[row0 + e] = dn ? 1 : 0;
        a.trunc[row0 + e] = truncated ? 1 : 0;
      }
      done_s[e] = dn ? 1 : 0;
      if (dn) {
        const Key k = split_at(reset_key, static_cast<uint32_t>(t) * static_cast<uint32_t>(a.B) + static_cast<uint32_t>(ge));
        const ResetScalars r = synth_reset_scalars(k, a.max_len);
        cnt_s[e] = r.counter;
        term_s[e] = r.term;
        kb_s[2 * e] = r.k_base.a;
        kb_s[2 * e + 1] = r.k_base.b;
      } else {
        cnt_s[e] = c;
        term_s[e] = ts;
      }
    }
    // (5) env step: obs' = tanh([obs, action] @ [Wo; Wa]) into the tile the actor no longer needs
    float* yo = nxt;
    mma_layer2(bufX, tlo, ld, O + A, reinterpret_cast<const float2*>(wbase + woff_s[L]), nullptr, O, yo, B200PPO_ACT_TANH, warp,
               lane);
    GROUP_SYNC();
    // (6) reward = -mean(obs'^2); next_obs[-1] is the pre-reset observation; tree_where(done, reset, next)
    for (int e = warp; e < TE; e += 8) {
      const bool dn = done_s[e] != 0;
      const Key kb{kb_s[2 * e], kb_s[2 * e + 1]};
      float sq = 0.0f;
      for (int o = lane; o < O; o += 32) {
        const float v = tile_get(yo, tlo, e * ld + o);
        sq = fmaf(v, v, sq);
        if (t == a.T - 1 && env0 + e < a.B) a.next_obs_last[static_cast<size_t>(env0 + e) * O + o] = v;
        tile_put(bufX, tlo, e * ld + o, dn ? bits_to_normal(random_bits_at(kb, static_cast<uint32_t>(o))) : v);
      }
      sq = warp_sum(sq);
      if (lane == 0 && env0 + e < a.B) a.reward[row0 + e] = -(sq / static_cast<float>(O));
    }
    GROUP_SYNC();
  }
#undef GROUP_SYNC
  for (int e = warp; e < TE; e += 8)
    for (int o = lane; o < O; o += 32)
      if (env0 + e < a.B) a.env_obs[static_cast<size_t>(env0 + e) * O + o] = tile_get(bufX, tlo, e * ld + o);
  if (lt < TE && env0 + lt < a.B) {
    a.env_counter[env0 + lt] = cnt_s[lt];
    a.env_term[env0 + lt] = term_s[lt];
  }
}

// ------------------------------------------------------------------------------------------
// persistent fused evaluation rollout (rollout.py:97-148): no transition record, no env reset;
// done is sticky, the reward is accumulated while the env was alive BEFORE the step, lifespan
// counts the steps that did not end in done.  Once every env of the tile is done nothing
// observable changes any more, so the CTA leaves the time loop early.
// ------------------------------------------------------------------------------------------
struct EvalArgs {
  b200ppo_plan plan;
  int O, A, max_len, term_thresh16;
  const float* Wenv;
  const float* params;
  const float* mean;
  const float* std;
  const uint32_t* rng_state;
  int L, B, mode;
  const float* env_obs; const int32_t* env_counter; const uint32_t* env_term;
  float* episode_reward; float* lifespan;
  int8_t split[B200PPO_MAX_LAYERS + 1];
  int use_scratch;
  int stage_actor, stage_env, actor_span, ld;
};

template <bool SA, bool SE>
__global__ void __launch_bounds__(NT, 2) eval_synth_kernel(const EvalArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int O = a.O, A = a.A, ld = a.ld;
  const int env0 = blockIdx.x * TE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sp = smem;
  float* bufA = sp; sp += TE * ld;
  float* bufB = sp; sp += TE * ld;
  float* bufC = nullptr;
  if (a.use_scratch) { bufC = sp; sp += TE * ld; }
  float* obs_s = sp; sp += TE * O;
  float* act_s = sp; sp += TE * A;
  float* mean_s = sp; sp += O;
  float* std_s = sp; sp += O;
  float* rew_s = sp; sp += TE;
  sp = smem + ((sp - smem + 3) & ~3);
  float* ps = sp; if (SA) sp += (a.actor_span + 3) & ~3;
  float* ws = sp; if (SE) sp += (O + A) * O;
  if (SA) for (int i = threadIdx.x; i < a.actor_span; i += NT) ps[i] = a.params[i];
  if (SE) for (int i = threadIdx.x; i < (O + A) * O; i += NT) ws[i] = a.Wenv[i];
  const float* P = SA ? ps : a.params;
  const float* Wenv = SE ? ws : a.Wenv;
  for (int i = threadIdx.x; i < O; i += NT) {
    mean_s[i] = a.plan.normalize ? a.mean[i] : 0.0f;
    std_s[i] = a.plan.normalize ? a.std[i] : 1.0f;
  }
  for (int e = warp; e < TE; e += NT / 32)
    for (int o = lane; o < O; o += 32)
      obs_s[e * O + o] = (env0 + e < a.B) ? a.env_obs[static_cast<size_t>(env0 + e) * O + o] : 0.0f;
  // per-env scalars live in the registers of thread e < TE
  const bool mine = threadIdx.x < TE && env0 + threadIdx.x < a.B;
  int32_t cnt = mine ? a.env_counter[env0 + threadIdx.x] : 0;
  uint32_t term = mine ? a.env_term[env0 + threadIdx.x] : 0u;
  bool prev_done = !mine;            // env.reset leaves done = 0; padding rows never hold the CTA back
  float cuml = 0.0f, life = 0.0f;
  const Key stream_key{a.rng_state[0], a.rng_state[1]};
  const uint32_t count0 = a.rng_state[2];
  const bool deterministic = (a.mode & 2) != 0;
  __syncthreads();

  for (int t = 0; t < a.L; ++t) {
    for (int e = warp; e < TE; e += NT / 32)
      for (int o = lane; o < O; o += 32) {
        const float x = obs_s[e * O + o];
        bufA[e * ld + o] = a.plan.normalize ? __fdiv_rn(x - mean_s[o], std_s[o]) : x;
      }
    __syncthreads();
    const float* y = run_chain(a.plan.actor, P, bufA, bufB, ld, bufC, a.split);
    // sampling_layers.py:93-96: one count per call when deterministic (the entropy draw), two otherwise
    const Key k_sample = deterministic ? stream_key : fold_in(stream_key, count0 + 2u * static_cast<uint32_t>(t));
    for (int i = threadIdx.x; i < TE * A; i += NT) {
      const int e = i / A, d = i - e * A;
      const uint32_t j = static_cast<uint32_t>(env0 + e) * static_cast<uint32_t>(A) + d;
      act_s[i] = sampler_elem(y[e * ld + d], y[e * ld + A + d], a.plan.min_std, a.plan.std_scale,
                              a.plan.entropy_weight, a.mode & 2, 0.0f, k_sample, k_sample, j, false).action;
    }
    __syncthreads();
    float* xin = (y == bufA) ? bufB : bufA;
    float* xout = (y == bufA) ? bufA : bufB;
    for (int e = warp; e < TE; e += NT / 32)
      for (int c = lane; c < O + A; c += 32)
        xin[e * ld + c] = c < O ? obs_s[e * O + c] : act_s[e * A + (c - O)];
    __syncthreads();
    dense_tile(xin, ld, O + A, Wenv, nullptr, O, xout, ld, B200PPO_ACT_TANH, bufC, a.split[B200PPO_MAX_LAYERS]);
    __syncthreads();
    for (int e = warp; e < TE; e += NT / 32) {      // reward and the next observation of one env per warp
      float s = 0.0f;
      for (int o = lane; o < O; o += 32) {
        const float v = xout[e * ld + o];
        s = fmaf(v, v, s);
        obs_s[e * O + o] = v;
      }
      s = warp_sum(s);
      if (lane == 0) rew_s[e] = -(s / static_cast<float>(O));
    }
    __syncthreads();
    bool all_done = true;
    if (threadIdx.x < TE) {
      cnt += 1;
      term = term * 1664525u + 1013904223u;
      const bool dn = (term >> 16) < static_cast<uint32_t>(a.term_thresh16) || cnt >= a.max_len;
      const bool done = dn || prev_done;                       // rollout.py:115-117
      if (!prev_done) cuml = __fadd_rn(cuml, rew_s[threadIdx.x]);   // rollout.py:119-123
      if (!done) life += 1.0f;                                 // rollout.py:124
      prev_done = done;
      all_done = done;
    }
    if (__syncthreads_and(all_done)) break;
  }
  if (mine) {
    a.episode_reward[env0 + threadIdx.x] = cuml;
    a.lifespan[env0 + threadIdx.x] = life;
  }
}

// ------------------------------------------------------------------------------------------
// generic single policy step on B rows (used for host/torch envs, replay checks and eval)
// ------------------------------------------------------------------------------------------
struct PolicyArgs {
  b200ppo_plan plan;
  const float* params; const float* mean; const float* std; const float* obs;
  int B, mode, ld, use_scratch;
  int8_t split_a[B200PPO_MAX_LAYERS], split_c[B200PPO_MAX_LAYERS];
  const uint32_t* rng_state; uint32_t count_offset;
  const float* raw_in;
  float* raw; float* action; float* loglik; float* value; float* reg;
  float* musig;              // nullable: [B][2A] = [mu | sigma], the sampler's `metrics` (sampling_layers.py:111)
};

__global__ void __launch_bounds__(NT, 1) policy_step_kernel(const PolicyArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int O = a.plan.obs_dim, A = a.plan.act_dim, ld = a.ld;
  const int row0 = blockIdx.x * TE;
  float* bufA = smem;
  float* bufB = bufA + TE * ld;
  float* bufC = a.use_scratch ? bufB + TE * ld : nullptr;   // partial tiles of the k-split layers
  float* x_s = bufB + (a.use_scratch ? 2 : 1) * TE * ld;    // normalised obs, kept for the critic
  float* llt_s = x_s + TE * ld;
  float* reg_s = llt_s + TE * A;
  for (int i = threadIdx.x; i < TE * O; i += NT) {
    const int e = i / O, o = i - e * O;
    float x = (row0 + e < a.B) ? a.obs[static_cast<size_t>(row0) * O + i] : 0.0f;
    if (a.plan.normalize) x = __fdiv_rn(x - a.mean[o], a.std[o]);
    x_s[e * ld + o] = x;
    bufA[e * ld + o] = x;
  }
  __syncthreads();
  const float* y = run_chain(a.plan.actor, a.params, bufA, bufB, ld, bufC, a.split_a);
  const Key stream_key{a.rng_state[0], a.rng_state[1]};
  uint32_t c = a.rng_state[2] + a.count_offset;
  const bool deterministic = (a.mode & 2) != 0;
  // sampling_layers.py:93-96: the sample key is only drawn when not deterministic
  Key k_sample = stream_key;
  if (!deterministic) { k_sample = fold_in(stream_key, c); c += 1u; }
  const Key k_ent = fold_in(stream_key, c);
  const bool want_reg = a.reg != nullptr;
  for (int i = threadIdx.x; i < TE * A; i += NT) {
    const int e = i / A, d = i - e * A;
    const int row = row0 + e;
    if (row < a.B) {
      const uint32_t j = static_cast<uint32_t>(row) * static_cast<uint32_t>(A) + d;
      const float rin = (a.mode & 1) ? a.raw_in[static_cast<size_t>(row) * A + d] : 0.0f;
      const SamplerOut s = sampler_elem(y[e * ld + d], y[e * ld + A + d], a.plan.min_std,
                                        a.plan.std_scale, a.plan.entropy_weight, a.mode, rin, k_sample,
                                        k_ent, j, want_reg);
      a.raw[static_cast<size_t>(row) * A + d] = s.raw;
      a.action[static_cast<size_t>(row) * A + d] = s.action;
      if (a.musig != nullptr) {
        a.musig[static_cast<size_t>(row) * 2 * A + d] = y[e * ld + d];
        a.musig[static_cast<size_t>(row) * 2 * A + A + d] = (softplus_f(y[e * ld + A + d]) + a.plan.min_std) * a.plan.std_scale;
      }
      llt_s[i] = s.llterm;
      reg_s[i] = s.regterm;
    } else {
      llt_s[i] = 0.0f;
      reg_s[i] = 0.0f;
    }
  }
  __syncthreads();
  if (threadIdx.x < TE && row0 + threadIdx.x < a.B) {
    float ll = 0.0f, rg = 0.0f;
    for (int d = 0; d < A; ++d) { ll += llt_s[threadIdx.x * A + d]; rg += reg_s[threadIdx.x * A + d]; }
    a.loglik[row0 + threadIdx.x] = ll;
    if (want_reg) a.reg[row0 + threadIdx.x] = rg;
  }
  __syncthreads();
  // critic on the same normalised input (adapter.py:87-88)
  float* ca = (y == bufA) ? bufB : bufA;
  for (int i = threadIdx.x; i < TE * O; i += NT) {
    const int e = i / O, o = i - e * O;
    ca[e * ld + o] = x_s[e * ld + o];
  }
  __syncthreads();
  float* cb = (ca == bufA) ? bufB : bufA;
  const float* v = run_chain(a.plan.critic, a.params, ca, cb, ld, bufC, a.split_c);
  if (threadIdx.x < TE && row0 + threadIdx.x < a.B) a.value[row0 + threadIdx.x] = v[threadIdx.x * ld];
}

// Row stride of the activation tiles: a multiple of 4 floats (float4 loads / stores of rows) that is
// 4 mod 8, so the row pairs read by the two half-warps of an item group are 8 or 24 banks apart.
int tile_ld(int md) {
  int ld = (md + 3) & ~3;
  if ((ld & 7) == 0) ld += 4;
  return ld;
}

// log2 of the k-split factor of a K x N layer on a TE-row tile (see dense_tile): a power of two that
// fills the CTA, leaves >= 4 k per part and fits the TE * ld scratch tile.  0 = no split.
int ksplit_log2(int K, int N, int ld) {
  if (N <= 0 || (N & 3)) return 0;
  const int ng = N >> 2;
  if (ng & (ng - 1)) return 0;
  const int items = (TE / 2) * ng;
  int S = 1, l = 0;
  while (S < 8 && 2 * S * items <= NT && K >= 8 * S && (2 * S - 1) * TE * N <= TE * ld) { S *= 2; ++l; }
  return l;
}
void fill_split(const b200ppo_chain& c, int ld, int8_t* split) {
  for (int l = 0; l < B200PPO_MAX_LAYERS; ++l)
    split[l] = l < c.n_layers ? static_cast<int8_t>(ksplit_log2(c.dims[l], c.dims[l + 1], ld)) : 0;
}

int max_dim(const b200ppo_chain& c) {
  int m = 0;
  for (int l = 0; l <= c.n_layers; ++l) m = c.dims[l] > m ? c.dims[l] : m;
  return m;
}

// Dense engine of the rollout: 1 = tensor cores (default; include/b200ppo.h b200ppo_set_rollout_mode), 0 = fp32 FFMA tiles,
// 2 = the batched per-step path whenever a workspace is passed.  B200PPO_ROLLOUT=ffma|mma|wide
int g_rollout_mode = -1;
int rollout_mode() {
  if (g_rollout_mode < 0) {
    const char* e = std::getenv("B200PPO_ROLLOUT");
    g_rollout_mode = (e && !std::strcmp(e, "ffma")) ? 0 : ((e && !std::strcmp(e, "wide")) ? 2 : 1);
  }
  return g_rollout_mode;
}

// Which tensor-core rollout kernel: 2 = one CTA per SM with two 16-env groups sharing the weights, 1 = one 16-env tile
// per CTA (two CTAs per SM).  B200PPO_ROLLOUT_MMA=1|2 forces one; by default the second version runs when there are
// more 16-env tiles than SMs (configs[1]: 256 tiles, 0.393 -> 0.362 ms) and the first when every tile can have an SM
// of its own (configs[0] shapes, 64 tiles: 0.245 vs 0.278 ms).
int g_rollout_mma_force = 0;        // b200ppo_set_rollout_mode(3 | 4)
int rollout_mma_version(int n_envs) {
  static const int v = [] { const char* e = std::getenv("B200PPO_ROLLOUT_MMA"); return e ? std::atoi(e) : 0; }();
  if (g_rollout_mma_force) return g_rollout_mma_force;
  if (v == 1 || v == 2) return v;
  return cdiv(n_envs, TE) > b200ppo_num_sms() ? 2 : 1;
}

int check_plan(const b200ppo_plan* p) {
  if (!p) return B200PPO_EINVAL;
  if (p->obs_dim <= 0 || p->act_dim <= 0) return B200PPO_EINVAL;
  const b200ppo_chain* cs[2] = {&p->actor, &p->critic};
  for (const b200ppo_chain* c : cs) {
    if (c->n_layers < 1 || c->n_layers > B200PPO_MAX_LAYERS) return B200PPO_EINVAL;
    if (c->dims[0] != p->obs_dim) return B200PPO_EINVAL;
    for (int l = 0; l < c->n_layers; ++l) {
      if (c->dims[l + 1] <= 0) return B200PPO_EINVAL;
      if ((c->w_off[l] & 3) || (c->b_off[l] & 3)) return B200PPO_EALIGN;
      if (c->w_off[l] < 0 || c->w_off[l] + static_cast<int64_t>(c->dims[l]) * c->dims[l + 1] > p->n_params)
        return B200PPO_EINVAL;
      if (c->b_off[l] < 0 || c->b_off[l] + c->dims[l + 1] > p->n_params) return B200PPO_EINVAL;
    }
  }
  if (p->actor.dims[p->actor.n_layers] != 2 * p->act_dim) return B200PPO_EINVAL;
  if (p->critic.dims[p->critic.n_layers] != 1) return B200PPO_EINVAL;
  return 0;
}

}  // namespace

extern "C" int b200ppo_synth_init_keys(void* stream, uint32_t k0, uint32_t k1, int32_t B, uint32_t* keys_out) {
  if (B <= 0 || !keys_out) return B200PPO_EINVAL;
  synth_init_keys_kernel<<<cdiv(B, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(Key{k0, k1}, B, keys_out);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200ppo_split_keys_dev(void* stream, const uint32_t* key, int64_t first, int32_t count, uint32_t* keys_out) {
  if (count < 0 || first < 0 || !key || (count > 0 && !keys_out)) return B200PPO_EINVAL;
  if (count == 0) return 0;
  split_keys_dev_kernel<<<cdiv(count, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(key, first, count, keys_out);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200ppo_split_rows(void* stream, const uint32_t* keys, int32_t rows, uint32_t index, uint32_t* keys_out) {
  if (rows < 0 || (rows > 0 && (!keys || !keys_out))) return B200PPO_EINVAL;
  if (rows == 0) return 0;
  split_rows_kernel<<<cdiv(rows, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(keys, rows, index, keys_out);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200ppo_synth_reset(void* stream, const b200ppo_synth_env* env, const uint32_t* keys,
                                   int32_t B, float* obs, int32_t* step_counter, uint32_t* term_state) {
  if (!env || B <= 0 || !keys || !obs || !step_counter || !term_state || env->obs_dim <= 0) return B200PPO_EINVAL;
  const int64_t n = static_cast<int64_t>(B) * env->obs_dim;
  synth_reset_kernel<<<cdiv(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      keys, B, env->obs_dim, env->max_len, obs, step_counter, term_state);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

namespace {
// weights of the whole actor chain + the env matrix resident in one SM's shared memory?  (fused kernels: yes ->
// every weight is read once per CTA; no -> every CTA re-streams them from L2 on every step)
bool fused_weights_resident(const b200ppo_plan* plan) {
  const int O = plan->obs_dim, A = plan->act_dim;
  int64_t w = static_cast<int64_t>(O + A) * O;
  for (int l = 0; l < plan->actor.n_layers; ++l) w += static_cast<int64_t>(plan->actor.dims[l]) * plan->actor.dims[l + 1];
  return 4 * w <= SMEM_LIMIT - 48 * 1024;
}
// batched (per-step tcgen05 GEMM) path: forced (mode 2) or chosen when the fused kernels would stream weights
bool wide_preferred(const b200ppo_plan* plan, int B) {
  if (rollout_mode() == 2) return true;
  return rollout_mode() == 1 && !fused_weights_resident(plan) && B >= 1024;
}

int rollout_synth_impl(void* stream, const b200ppo_plan* plan, const b200ppo_synth_env* env,
                                     const float* params, const float* norm_mean, const float* norm_std,
                                     const uint32_t* rng_state, const uint32_t* iter_keys, int32_t T,
                                     int32_t B, float* env_obs, int32_t* env_counter, uint32_t* env_term,
                                     float* obs, float* raw_action, float* action, float* loglik,
                                     float* reward, uint8_t* done, uint8_t* truncated,
                                     float* next_obs_last, void* ws, int64_t ws_bytes) {
  int rc = check_plan(plan);
  if (rc) return rc;
  if (!env || !params || !rng_state || !iter_keys || !env_obs || !env_counter || !env_term || !obs ||
      !raw_action || !action || !loglik || !reward || !done || !truncated || !next_obs_last)
    return B200PPO_EINVAL;
  if (plan->normalize && (!norm_mean || !norm_std)) return B200PPO_EINVAL;
  if (T <= 0 || B <= 0) return B200PPO_EINVAL;
  if (env->obs_dim != plan->obs_dim || env->act_dim != plan->act_dim || !env->Wo || !env->Wa) return B200PPO_EINVAL;
  const int O = plan->obs_dim, A = plan->act_dim;
  // Wo [O][O] and Wa [A][O] must be one contiguous [(O + A)][O] block (Wa right after Wo).
  if (env->Wa != env->Wo + static_cast<size_t>(O) * O) return B200PPO_EINVAL;
  RolloutArgs a;
  a.plan = *plan;
  a.O = O; a.A = A; a.max_len = env->max_len; a.term_thresh16 = env->term_thresh16;
  a.Wenv = env->Wo; a.params = params; a.mean = norm_mean; a.std = norm_std;
  a.rng_state = rng_state; a.iter_keys = iter_keys; a.T = T; a.B = B;
  a.env_obs = env_obs; a.env_counter = env_counter; a.env_term = env_term;
  a.obs = obs; a.raw_action = raw_action; a.action = action; a.loglik = loglik; a.reward = reward;
  a.done = done; a.trunc = truncated; a.next_obs_last = next_obs_last;
  if (ws != nullptr && wide_preferred(plan, B) && ws_bytes >= 4 * rollout_wide_ws_floats(plan, B)) {
    RolloutWideArgs w;
    w.plan = plan; w.Wenv = env->Wo; w.max_len = env->max_len; w.term_thresh16 = env->term_thresh16;
    w.params = params; w.mean = norm_mean; w.std = norm_std; w.rng_state = rng_state; w.iter_keys = iter_keys;
    w.T = T; w.B = B; w.env_obs = env_obs; w.env_counter = env_counter; w.env_term = env_term;
    w.obs = obs; w.raw_action = raw_action; w.action = action; w.loglik = loglik; w.reward = reward;
    w.done = done; w.trunc = truncated; w.next_obs_last = next_obs_last; w.ws = static_cast<float*>(ws);
    return rollout_wide(static_cast<cudaStream_t>(stream), w);
  }
  int md = max_dim(plan->actor);
  if (O + A > md) md = O + A;
  if (rollout_mode() >= 1) {
    // tensor-core tiles (mma.sync 3xTF32): fragment-ordered weights of every layer resident in shared memory
    a.ld = ((md + 7) & ~7) + 4;          // = 4 (mod 8): the A-fragment loads (lane = 4 g + t -> g * ld + t) hit 32 distinct banks
    int64_t fl = 3ll * TE * a.ld + 2ll * TE * A + 2ll * O + 5ll * TE + 8;
    for (int l = 0; l < plan->actor.n_layers; ++l)
      fl += frag_floats(plan->actor.dims[l], plan->actor.dims[l + 1]) + ((plan->actor.dims[l + 1] + 7) & ~7);
    fl += frag_floats(O + A, O);
    // second version first: one CTA per SM with two env groups sharing the weights, activation tiles pre-split
    {
      const int64_t per_group = (6ll * TE * a.ld + 2ll * TE * A + 5ll * TE + 3) & ~3ll;
      int64_t f2 = 2 * per_group + ((2ll * O + 3) & ~3ll);
      for (int l = 0; l < plan->actor.n_layers; ++l)
        f2 += frag_floats(plan->actor.dims[l], plan->actor.dims[l + 1]) + ((plan->actor.dims[l + 1] + 7) & ~7);
      f2 += frag_floats(O + A, O);
      if (4 * f2 <= SMEM_LIMIT - 2048 && rollout_mma_version(B) == 2) {
        cudaError_t e = cudaFuncSetAttribute(rollout_synth_mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(4 * f2));
        if (e != cudaSuccess) return static_cast<int>(e);
        rollout_synth_mma2_kernel<<<cdiv(B, 2 * TE), NT2, 4 * f2, static_cast<cudaStream_t>(stream)>>>(a);
        B200PPO_LAUNCH_CHECK();
        return 0;
      }
    }
    if (4 * fl <= SMEM_LIMIT - 2048) {     // the kernel also has ~1 KB of static shared memory
      cudaError_t e = cudaFuncSetAttribute(rollout_synth_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(4 * fl));
      if (e != cudaSuccess) return static_cast<int>(e);
      rollout_synth_mma_kernel<<<cdiv(B, TE), NT, 4 * fl, static_cast<cudaStream_t>(stream)>>>(a);
      B200PPO_LAUNCH_CHECK();
      return 0;
    }
  }
  a.ld = tile_ld(md);
  fill_split(plan->actor, a.ld, a.split);
  a.split[B200PPO_MAX_LAYERS] = static_cast<int8_t>(ksplit_log2(O + A, O, a.ld));
  // the split path loads weight rows as float4: fall back to the unsplit path for unaligned buffers
  if (reinterpret_cast<uintptr_t>(params) & 15) fill_split(b200ppo_chain{}, a.ld, a.split);
  if (reinterpret_cast<uintptr_t>(env->Wo) & 15) a.split[B200PPO_MAX_LAYERS] = 0;
  // actor parameters occupy the arena prefix [0, actor_span)
  int64_t span = 0;
  for (int l = 0; l < plan->actor.n_layers; ++l) {
    int64_t we = plan->actor.w_off[l] + static_cast<int64_t>(plan->actor.dims[l]) * plan->actor.dims[l + 1];
    int64_t be = plan->actor.b_off[l] + plan->actor.dims[l + 1];
    span = we > span ? we : span;
    span = be > span ? be : span;
  }
  int64_t base_bytes = 4ll * (4ll * TE * a.ld + 2ll * TE * A + 2ll * O + 5ll * TE + 8);
  a.use_scratch = 1;
  if (base_bytes > SMEM_LIMIT) {          // very wide tiles: run the layers unsplit, without the scratch tile
    a.use_scratch = 0;
    base_bytes -= 4ll * TE * a.ld;
    for (int8_t& v : a.split) v = 0;
  }
  if (base_bytes > SMEM_LIMIT) return B200PPO_ELIMIT;
  int64_t bytes = base_bytes;
  a.actor_span = static_cast<int>(span);
  a.stage_actor = 0;
  a.stage_env = 0;
  if (bytes + 4 * ((span + 3) & ~3ll) <= SMEM_LIMIT) { a.stage_actor = 1; bytes += 4 * ((span + 3) & ~3ll); }
  const int64_t envw = 4ll * (O + A) * O;
  if (bytes + envw <= SMEM_LIMIT) { a.stage_env = 1; bytes += envw; }
  auto launch = [&](auto kernel) -> int {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return static_cast<int>(e);
    kernel<<<cdiv(B, TE), NT, bytes, static_cast<cudaStream_t>(stream)>>>(a);
    return 0;
  };
  rc = a.stage_actor ? (a.stage_env ? launch(rollout_synth_kernel<true, true>) : launch(rollout_synth_kernel<true, false>))
                     : (a.stage_env ? launch(rollout_synth_kernel<false, true>) : launch(rollout_synth_kernel<false, false>));
  if (rc) return rc;
  B200PPO_LAUNCH_CHECK();
  return 0;
}
}  // namespace

extern "C" int b200ppo_rollout_synth(void* stream, const b200ppo_plan* plan, const b200ppo_synth_env* env,
                                     const float* params, const float* norm_mean, const float* norm_std,
                                     const uint32_t* rng_state, const uint32_t* iter_keys, int32_t T,
                                     int32_t B, float* env_obs, int32_t* env_counter, uint32_t* env_term,
                                     float* obs, float* raw_action, float* action, float* loglik,
                                     float* reward, uint8_t* done, uint8_t* truncated,
                                     float* next_obs_last) {
  return rollout_synth_impl(stream, plan, env, params, norm_mean, norm_std, rng_state, iter_keys, T, B, env_obs,
                            env_counter, env_term, obs, raw_action, action, loglik, reward, done, truncated,
                            next_obs_last, nullptr, 0);
}

extern "C" int64_t b200ppo_rollout_synth_workspace_bytes(const b200ppo_plan* plan, int32_t B) {
  if (check_plan(plan) || B <= 0) return -1;
  return wide_preferred(plan, B) ? 4 * rollout_wide_ws_floats(plan, B) : 0;
}

extern "C" int b200ppo_rollout_synth_num_launches(const b200ppo_plan* plan, int32_t T, int32_t B, int32_t with_ws) {
  if (check_plan(plan) || B <= 0 || T <= 0) return -1;
  return (with_ws && wide_preferred(plan, B)) ? rollout_wide_num_launches(plan, T) : 1;
}

extern "C" int b200ppo_rollout_synth_ws(void* stream, const b200ppo_plan* plan, const b200ppo_synth_env* env,
                                        const float* params, const float* norm_mean, const float* norm_std,
                                        const uint32_t* rng_state, const uint32_t* iter_keys, int32_t T,
                                        int32_t B, float* env_obs, int32_t* env_counter, uint32_t* env_term,
                                        float* obs, float* raw_action, float* action, float* loglik,
                                        float* reward, uint8_t* done, uint8_t* truncated,
                                        float* next_obs_last, void* ws, int64_t ws_bytes) {
  return rollout_synth_impl(stream, plan, env, params, norm_mean, norm_std, rng_state, iter_keys, T, B, env_obs,
                            env_counter, env_term, obs, raw_action, action, loglik, reward, done, truncated,
                            next_obs_last, ws, ws_bytes);
}

extern "C" int b200ppo_set_rollout_mode(int mode) {
  const int prev = rollout_mode();
  if (mode >= 0 && mode <= 2) { g_rollout_mode = mode; g_rollout_mma_force = 0; }
  if (mode == 3 || mode == 4) { g_rollout_mode = 1; g_rollout_mma_force = mode - 2; }   // tensor cores, kernel version 1 / 2
  return prev;
}

extern "C" int b200ppo_eval_synth(void* stream, const b200ppo_plan* plan, const b200ppo_synth_env* env,
                                  const float* params, const float* norm_mean, const float* norm_std,
                                  const uint32_t* rng_state, int32_t mode, int32_t L, int32_t B,
                                  const float* env_obs, const int32_t* env_counter, const uint32_t* env_term,
                                  float* episode_reward, float* lifespan) {
  int rc = check_plan(plan);
  if (rc) return rc;
  if (!env || !params || !rng_state || !env_obs || !env_counter || !env_term || !episode_reward || !lifespan)
    return B200PPO_EINVAL;
  if (plan->normalize && (!norm_mean || !norm_std)) return B200PPO_EINVAL;
  if (L <= 0 || B <= 0 || (mode & ~2)) return B200PPO_EINVAL;
  if (env->obs_dim != plan->obs_dim || env->act_dim != plan->act_dim || !env->Wo || !env->Wa) return B200PPO_EINVAL;
  const int O = plan->obs_dim, A = plan->act_dim;
  if (env->Wa != env->Wo + static_cast<size_t>(O) * O) return B200PPO_EINVAL;
  EvalArgs a;
  a.plan = *plan;
  a.O = O; a.A = A; a.max_len = env->max_len; a.term_thresh16 = env->term_thresh16;
  a.Wenv = env->Wo; a.params = params; a.mean = norm_mean; a.std = norm_std;
  a.rng_state = rng_state; a.L = L; a.B = B; a.mode = mode;
  a.env_obs = env_obs; a.env_counter = env_counter; a.env_term = env_term;
  a.episode_reward = episode_reward; a.lifespan = lifespan;
  int md = max_dim(plan->actor);
  if (O + A > md) md = O + A;
  a.ld = tile_ld(md);
  fill_split(plan->actor, a.ld, a.split);
  a.split[B200PPO_MAX_LAYERS] = static_cast<int8_t>(ksplit_log2(O + A, O, a.ld));
  // the split path loads weight rows as float4: fall back to the unsplit path for unaligned buffers
  if (reinterpret_cast<uintptr_t>(params) & 15) fill_split(b200ppo_chain{}, a.ld, a.split);
  if (reinterpret_cast<uintptr_t>(env->Wo) & 15) a.split[B200PPO_MAX_LAYERS] = 0;
  int64_t span = 0;
  for (int l = 0; l < plan->actor.n_layers; ++l) {
    int64_t we = plan->actor.w_off[l] + static_cast<int64_t>(plan->actor.dims[l]) * plan->actor.dims[l + 1];
    int64_t be = plan->actor.b_off[l] + plan->actor.dims[l + 1];
    span = we > span ? we : span;
    span = be > span ? be : span;
  }
  int64_t bytes = 4ll * (3ll * TE * a.ld + static_cast<int64_t>(TE) * O + static_cast<int64_t>(TE) * A + 2ll * O + TE + 8);
  a.use_scratch = 1;
  if (bytes > SMEM_LIMIT) {
    a.use_scratch = 0;
    bytes -= 4ll * TE * a.ld;
    for (int8_t& v : a.split) v = 0;
  }
  if (bytes > SMEM_LIMIT) return B200PPO_ELIMIT;
  a.actor_span = static_cast<int>(span);
  a.stage_actor = 0;
  a.stage_env = 0;
  if (bytes + 4 * ((span + 3) & ~3ll) <= SMEM_LIMIT) { a.stage_actor = 1; bytes += 4 * ((span + 3) & ~3ll); }
  const int64_t envw = 4ll * (O + A) * O;
  if (bytes + envw <= SMEM_LIMIT) { a.stage_env = 1; bytes += envw; }
  auto launch = [&](auto kernel) -> int {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return static_cast<int>(e);
    kernel<<<cdiv(B, TE), NT, bytes, static_cast<cudaStream_t>(stream)>>>(a);
    return 0;
  };
  rc = a.stage_actor ? (a.stage_env ? launch(eval_synth_kernel<true, true>) : launch(eval_synth_kernel<true, false>))
                     : (a.stage_env ? launch(eval_synth_kernel<false, true>) : launch(eval_synth_kernel<false, false>));
  if (rc) return rc;
  B200PPO_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t b200ppo_policy_workspace_bytes(const b200ppo_plan* plan, int32_t B) {
  if (check_plan(plan) || B < 0) return B200PPO_EINVAL;
  return 8ll * plan->act_dim * B;   // the optional [B][2A] mu / sigma output; everything else lives in shared memory
}

extern "C" int b200ppo_policy_step(void* stream, const b200ppo_plan* plan, const float* params,
                                   const float* norm_mean, const float* norm_std, const float* obs,
                                   int32_t B, int32_t mode, const uint32_t* rng_state,
                                   uint32_t count_offset, const float* raw_action_in, float* raw_action,
                                   float* action, float* loglik, float* value, float* reg_loss, void* ws) {
  int rc = check_plan(plan);
  if (rc) return rc;
  if (B < 0) return B200PPO_EINVAL;
  if (B == 0) return 0;
  if (!params || !obs || !rng_state || !raw_action || !action || !loglik || !value) return B200PPO_EINVAL;
  if (plan->normalize && (!norm_mean || !norm_std)) return B200PPO_EINVAL;
  if ((mode & 1) && !raw_action_in) return B200PPO_EINVAL;
  PolicyArgs a;
  a.plan = *plan; a.params = params; a.mean = norm_mean; a.std = norm_std; a.obs = obs;
  a.B = B; a.mode = mode; a.rng_state = rng_state; a.count_offset = count_offset; a.raw_in = raw_action_in;
  a.raw = raw_action; a.action = action; a.loglik = loglik; a.value = value; a.reg = reg_loss;
  a.musig = static_cast<float*>(ws);
  int md = max_dim(plan->actor);
  const int mc = max_dim(plan->critic);
  md = mc > md ? mc : md;
  a.ld = tile_ld(md);
  fill_split(plan->actor, a.ld, a.split_a);
  fill_split(plan->critic, a.ld, a.split_c);
  if (reinterpret_cast<uintptr_t>(params) & 15) {
    fill_split(b200ppo_chain{}, a.ld, a.split_a);
    fill_split(b200ppo_chain{}, a.ld, a.split_c);
  }
  int64_t bytes = 4ll * (4ll * TE * a.ld + 2ll * TE * plan->act_dim);
  a.use_scratch = 1;
  if (bytes > SMEM_LIMIT) {               // very wide layers: no room for the k-split scratch tile (not needed there)
    a.use_scratch = 0;
    bytes -= 4ll * TE * a.ld;
  }
  if (bytes > SMEM_LIMIT) return B200PPO_ELIMIT;
  cudaError_t e = cudaFuncSetAttribute(policy_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  if (e != cudaSuccess) return static_cast<int>(e);
  policy_step_kernel<<<cdiv(B, TE), NT, bytes, static_cast<cudaStream_t>(stream)>>>(a);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// NormalTanhSampler alone, on actor outputs computed elsewhere (the recurrent actor's step kernel):
// same stream / count conventions as policy_step_kernel (sampling_layers.py:88-147).
// ------------------------------------------------------------------------------------------
namespace {
struct SamplerArgs {
  const float* y; int B, A, mode;
  float min_std, std_scale, entropy_weight;
  const uint32_t* rng_state; uint32_t count_offset;
  const float* raw_in;
  float* raw; float* action; float* loglik; float* reg;
};
__global__ void __launch_bounds__(256) sampler_step_kernel(const SamplerArgs a) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= a.B) return;
  const Key stream_key{a.rng_state[0], a.rng_state[1]};
  uint32_t c = a.rng_state[2] + a.count_offset;
  // sampling_layers.py:93-96: the sample key is only drawn when not deterministic
  Key k_sample = stream_key;
  if (!(a.mode & 2)) { k_sample = fold_in(stream_key, c); c += 1u; }
  const Key k_ent = fold_in(stream_key, c);
  const bool want_reg = a.reg != nullptr;
  float ll = 0.0f, rg = 0.0f;
  for (int d = 0; d < a.A; ++d) {
    const uint32_t j = static_cast<uint32_t>(row) * static_cast<uint32_t>(a.A) + d;
    const float rin = (a.mode & 1) ? a.raw_in[static_cast<size_t>(row) * a.A + d] : 0.0f;
    const SamplerOut s = sampler_elem(a.y[static_cast<size_t>(row) * 2 * a.A + d],
                                      a.y[static_cast<size_t>(row) * 2 * a.A + a.A + d], a.min_std, a.std_scale,
                                      a.entropy_weight, a.mode, rin, k_sample, k_ent, j, want_reg);
    a.raw[static_cast<size_t>(row) * a.A + d] = s.raw;
    a.action[static_cast<size_t>(row) * a.A + d] = s.action;
    ll += s.llterm;
    rg += s.regterm;
  }
  a.loglik[row] = ll;
  if (want_reg) a.reg[row] = rg;
}
}  // namespace

extern "C" int b200ppo_sampler_step(void* stream, const float* y, int32_t B, int32_t A, int32_t mode,
                                    float min_std, float std_scale, float entropy_weight,
                                    const uint32_t* rng_state, uint32_t count_offset,
                                    const float* raw_action_in, float* raw_action, float* action,
                                    float* loglik, float* reg_loss) {
  if (B < 0 || A <= 0) return B200PPO_EINVAL;
  if (B == 0) return 0;
  if (!y || !rng_state || !raw_action || !action || !loglik) return B200PPO_EINVAL;
  if ((mode & 1) && !raw_action_in) return B200PPO_EINVAL;
  SamplerArgs a;
  a.y = y; a.B = B; a.A = A; a.mode = mode; a.min_std = min_std; a.std_scale = std_scale;
  a.entropy_weight = entropy_weight;
  a.rng_state = rng_state; a.count_offset = count_offset; a.raw_in = raw_action_in;
  a.raw = raw_action; a.action = action; a.loglik = loglik; a.reg = reg_loss;
  sampler_step_kernel<<<cdiv(B, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  B200PPO_LAUNCH_CHECK();
  return 0;
}
