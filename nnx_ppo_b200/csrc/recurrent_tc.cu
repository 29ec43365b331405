// Recurrent actor on the tensor cores (reference networks/recurrent.py:89-161
// over flax's OptimizedLSTMCell, replayed over T in ppo.py:411-431).  sm_100a, tcgen05 kind::tf32 with the
// error-compensated 3xTF32 split of tc.cuh (fp32 parity), fp32 accumulators in TMEM.
//
// What is sequential in   obs -> Normalizer -> Dense(act) -> LSTM -> Dense   is only  h_{t-1} Wh  (and its
// transpose in BPTT); everything else is a batch over all T * rows samples.  So a replay is
//   forward : z1 = x W1 + b1, u = act(z1)                      one GEMM over T*rows rows
//             gx = u Wi + bl                                   one GEMM (N = 4H, tiled by 256)
//             for t: a_t = gx_t + h_{t-1} Wh -> gates -> (c, h)  T step launches: 128-row x 16-unit tiles; the
//                                                              gate math, the reset-on-done select and the
//                                                              activation cache live in the GEMM epilogue
//             y = h' W2 + b2                                   one GEMM
//   backward: dh_post = dY W2^T                                one GEMM
//             for t: da_t (element-wise; sums the split-K partials of step t+1), dh_rec = da_t Wh^T
//                                                              T x (element-wise launch + split-K GEMM launch)
//             du = da Wi^T, dz1 = du * act'(z1)                one GEMM
//             d[Wi; Wh], dbl, dW1, db1, dW2, db2               "TN" GEMMs over all T*rows rows, row-split
//                                                              partials reduced in fixed order (deterministic)
// The same forward entry point serves the rollout (rows = n_envs, T = 1 per call, no cache).
//
// The kernels are lock-step pipelines: all 256 threads stage a k-slab of both operands (global -> registers
// one stage ahead -> hi / lo split -> shared memory), one elected lane issues the MMAs, tcgen05.commit hands
// the buffer back.  Operand layout in shared memory: tc.cuh (K-major, no swizzle, padded planes).
#include "common.cuh"
#include "tc.cuh"

#include <cstdio>
#include <cstdlib>

using namespace b200ppo;

namespace {

constexpr int RM = 128;        // rows per CTA (UMMA M)
constexpr int RK = 32;         // k per stage of the NN / NT kernel (4 MMAs of K = 8)
constexpr int RT = 256;        // staging / epilogue threads per CTA
constexpr int RTI = RT + 32;   // + one warp that only issues MMAs (an issuing warp that also stages serialises the
                               //   ~90-cycle-per-MMA issue with its share of the staging: profiles/r2_recurrent_notes.md)
constexpr int RN_MAX = 256;    // widest N tile
constexpr int UT = 16;         // hidden units per CTA of the recurrent step (x 4 gates = 64 accumulator columns)

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------
// C[M x N] = opA(A)[M x K] * Bop[N x K]^T          (one 128-row x N-column tile per CTA)
// ------------------------------------------------------------------------------------------
struct GemmArgs {
  // A: row-major, row r at A + r * lda; reduction columns [a_k0, a_k0 + K); rows >= M read as zero
  const float* A; int lda; int a_k0; int M; int K;
  int act_a;                                   // activation applied to A on load (B200PPO_ACT_*)
  const float* a_mean; const float* a_std;     // nullable: (x - mean[k]) / std[k] on load (Normalizer.__call__)
  // B, two layouts:
  //   b_nt = 0: row-major [K][ldb];  Bop(n, k) = B[(b_k0 + k) * ldb + col(n)],
  //             col(n) = b_col0 + (n >> b_unit_log2) * b_gate_stride + (n & (2^b_unit_log2 - 1))   (gate-interleaved tiles)
  //   b_nt = 1: row-major [N][ldb];  Bop(n, k) = B[(b_col0 + n) * ldb + b_k0 + k]  (reduction index contiguous)
  const float* B; int ldb; int b_nt; int b_col0; int b_unit_log2; int b_gate_stride; int b_k0;
  // optional (rg_gemm_kernel<0>): B pre-split by rg_prep_b_kernel into the shared-memory stage layout, per column
  // tile and 32-k stage one contiguous block [hi planes | lo planes] -> ONE cp.async.bulk per stage, no thread work
  const float* Bplanes;
  int N; int n_log2;                           // MMA N of one tile: a power of two in [16, 256]
  int n_real;                                  // valid columns over all tiles (columns >= n_real are padding)
  int tile_stride;                             // columns per blockIdx.y tile (added to b_col0 / the C column)
  // epilogue: C[r * ldc + n] = acc + bias[n]   (n = global column); optional C2 = act(C); optional
  //           C = acc * act'(Zmul[r * ldz + n])
  float* C; int ldc;
  int c_planes; int c_cols4;                   // c_planes: C is stored as planes [row / 128][c_cols4][128][4] (see recurrent_tc_step.cuh)
  const float* bias;
  float* C2; int ldc2; int act_c2;
  const float* Zmul; int ldz; int act_z;
  // split-K: blockIdx.z = slice; reduction offsets advance by k_slice, C by c_slice_stride floats
  int k_slice; long long c_slice_stride;
};

struct StageRegs {
  float4 a[4];
  float4 b[RN_MAX / 32];
};


__host__ __device__ inline uint32_t stage_bytes_nn(int n) {
  return 2u * (RK / 4) * tc::plane_bytes(RM) + 2u * (RK / 4) * tc::plane_bytes(n);
}

__device__ __forceinline__ float4 ld4_guard(const float* src, int k, int K, bool aligned) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (aligned && k + 3 < K) {
    v = *reinterpret_cast<const float4*>(src);
  } else {
    v.x = src[0];
    if (k + 1 < K) v.y = src[1];
    if (k + 2 < K) v.z = src[2];
    if (k + 3 < K) v.w = src[3];
  }
  return v;
}

// global -> registers for stage s (every load is issued before any is used)
template <int NB>
__device__ __forceinline__ void gemm_load_stage(const GemmArgs& g, int row0, int kbase_a, int kbase_b, int col0, int s,
                                                int nst, StageRegs& x) {
  const int tid = threadIdx.x;
  const int k0 = s * RK;
  const bool a_al = ((g.lda | kbase_a) & 3) == 0 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int idx = tid + j * RT;
    const int q = idx & 7, r = idx >> 3;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int k = k0 + 4 * q;
    if (s < nst && row0 + r < g.M && k < g.K) {
      v = ld4_guard(g.A + static_cast<size_t>(row0 + r) * g.lda + kbase_a + k, k, g.K, a_al);
      if (g.a_mean != nullptr) {
        v.x = __fdiv_rn(v.x - g.a_mean[k], g.a_std[k]);
        if (k + 1 < g.K) v.y = __fdiv_rn(v.y - g.a_mean[k + 1], g.a_std[k + 1]);
        if (k + 2 < g.K) v.z = __fdiv_rn(v.z - g.a_mean[k + 2], g.a_std[k + 2]);
        if (k + 3 < g.K) v.w = __fdiv_rn(v.w - g.a_mean[k + 3], g.a_std[k + 3]);
      }
    }
    x.a[j] = v;
  }
  const bool b_al = ((g.ldb | kbase_b) & 3) == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15) == 0;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int idx = tid + j * RT;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g.b_nt) {
      const int q = idx & 7, n = idx >> 3;
      const int k = k0 + 4 * q;
      if (s < nst && n < g.N && col0 + n < g.n_real && k < g.K)
        v = ld4_guard(g.B + static_cast<size_t>(col0 + n) * g.ldb + kbase_b + k, k, g.K, b_al);
    } else {
      const int n = idx & (g.N - 1), q = idx >> g.n_log2;   // consecutive lanes -> consecutive columns (coalesced)
      const int k = k0 + 4 * q;
      const int col = col0 + (n >> g.b_unit_log2) * g.b_gate_stride + (n & ((1 << g.b_unit_log2) - 1));
      if (s < nst && q < RK / 4 && (g.b_gate_stride != 0 || col < g.n_real) && k < g.K) {
        const float* src = g.B + static_cast<size_t>(kbase_b + k) * g.ldb + col;
        v.x = src[0];
        if (k + 1 < g.K) v.y = src[g.ldb];
        if (k + 2 < g.K) v.z = src[2 * static_cast<size_t>(g.ldb)];
        if (k + 3 < g.K) v.w = src[3 * static_cast<size_t>(g.ldb)];
      }
    }
    x.b[j] = v;
  }
}

// registers -> hi / lo operand planes of the stage buffer `st`
template <int NB>
__device__ __forceinline__ void gemm_store_stage(const GemmArgs& g, uint8_t* st, const StageRegs& x) {
  const int tid = threadIdx.x;
  const uint32_t pa = tc::plane_bytes(RM), pb = tc::plane_bytes(g.N);
  uint8_t* a_hi = st;
  uint8_t* a_lo = a_hi + (RK / 4) * pa;
  uint8_t* b_hi = a_lo + (RK / 4) * pa;
  uint8_t* b_lo = b_hi + (RK / 4) * pb;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int idx = tid + j * RT;
    const int q = idx & 7, r = idx >> 3;
    float4 v = x.a[j];
    if (g.act_a == B200PPO_ACT_RELU) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    } else if (g.act_a != B200PPO_ACT_NONE) {
      v.x = act_fwd(v.x, g.act_a); v.y = act_fwd(v.y, g.act_a); v.z = act_fwd(v.z, g.act_a); v.w = act_fwd(v.w, g.act_a);
    }
    float4 hi, lo;
    tc::split4_fast(v, hi, lo);
    *reinterpret_cast<float4*>(a_hi + q * pa + r * 16) = hi;
    *reinterpret_cast<float4*>(a_lo + q * pa + r * 16) = lo;
  }
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int idx = tid + j * RT;
    int n, q;
    if (g.b_nt) { q = idx & 7; n = idx >> 3; }
    else { n = idx & (g.N - 1); q = idx >> g.n_log2; }
    if (n < g.N && q < RK / 4) {
      float4 hi, lo;
      tc::split4_fast(x.b[j], hi, lo);
      *reinterpret_cast<float4*>(b_hi + q * pb + n * 16) = hi;
      *reinterpret_cast<float4*>(b_lo + q * pb + n * 16) = lo;
    }
  }
}

struct Pipe {
  uint8_t* smem;
  uint64_t* bar_empty;   // [2]: tcgen05.commit of the MMAs that read the buffer
  uint64_t* bar_full;    // [2]: the staging warps finished writing the buffer
  uint64_t* bar_done;
  uint32_t tmem_base;
  uint32_t tmem_cols;
};

__device__ __forceinline__ uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while (c < static_cast<uint32_t>(n)) c <<= 1;
  return c;
}

// bars: [0..ns) empty, [ns..2 ns) full (`full_count` arrivals each), [2 ns] done; ns = ring depth (2, or 3 for the bulk-fed GEMM)
__device__ __forceinline__ void pipe_init(Pipe& p, uint8_t* smem, uint64_t* bars, uint32_t* tmem_slot, int ncols, int full_count,
                                          int ns = 2) {
  const int warp = threadIdx.x >> 5;
  p.tmem_cols = pow2_cols(ncols);
  if (warp == 0) tc::tmem_alloc(tmem_slot, p.tmem_cols);
  if (threadIdx.x == 32) {
    for (int i = 0; i < ns; ++i) {
      tc::mbar_init(&bars[i], 1);
      tc::mbar_init(&bars[ns + i], full_count);
    }
    tc::mbar_init(&bars[2 * ns], 1);
    tc::mbar_init_fence();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  p.smem = smem;
  p.bar_empty = bars;
  p.bar_full = bars + ns;
  p.bar_done = bars + 2 * ns;
  p.tmem_base = *tmem_slot;
}

__device__ __forceinline__ void pipe_fini(Pipe& p) {
  tc::tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc(p.tmem_base, p.tmem_cols);
}

// issue the MMAs of one staged slab (KS k-steps of 8) and commit; called by the whole (converged) issuer warp
template <int KS>
__device__ __forceinline__ void issue_slab(const Pipe& p, uint8_t* st, uint32_t pa, uint32_t pb, uint32_t idesc, bool first,
                                           bool last, int buf) {
  {
    tc::tc_fence_after();
    if (tc::elect_one()) {
      uint8_t* a_hi = st;
      uint8_t* a_lo = a_hi + (KS * 2) * pa;
      uint8_t* b_hi = a_lo + (KS * 2) * pa;
      uint8_t* b_lo = b_hi + (KS * 2) * pb;
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        const uint64_t ah = tc::make_desc(tc::smem_u32(a_hi + 2 * j * pa), pa, 128);
        const uint64_t al = tc::make_desc(tc::smem_u32(a_lo + 2 * j * pa), pa, 128);
        const uint64_t bh = tc::make_desc(tc::smem_u32(b_hi + 2 * j * pb), pb, 128);
        const uint64_t bl = tc::make_desc(tc::smem_u32(b_lo + 2 * j * pb), pb, 128);
        const uint32_t acc0 = (!first || j > 0) ? 1u : 0u;
        tc::mma_tf32(p.tmem_base, al, bh, idesc, acc0);       // small terms first, then the dominant product
        tc::mma_tf32(p.tmem_base, ah, bl, idesc, 1u);
        tc::mma_tf32(p.tmem_base, ah, bh, idesc, 1u);
      }
      tc::commit(&p.bar_empty[buf]);
      if (last) tc::commit(p.bar_done);
    }
    __syncwarp();
  }
}

// operand ring depth: 3 slots for the bulk-fed kernel when three stages fit shared memory (column tiles <= 128), else 2
template <int NB>
__host__ __device__ inline int gemm_ring_depth(int N) { return (NB == 0 && N <= 128) ? 3 : 2; }

// B operand of stage s from the pre-split planes: one bulk copy, completion counted on the stage's full barrier
// (the issuing thread's expect_tx arrival is the extra arrival pipe_init is told about)
__device__ __forceinline__ void gemm_bulk_b(const GemmArgs& g, const Pipe& p, int buf, int s, int nst, uint32_t pa,
                                            uint32_t pb, uint32_t sbytes) {
  const uint32_t bytes = 2u * (RK / 4) * pb;
  uint8_t* dst = p.smem + buf * sbytes + 2u * (RK / 4) * pa;
  const float* src = g.Bplanes + (static_cast<size_t>(blockIdx.y) * nst + s) * (bytes / 4);
  tc::mbar_arrive_expect_tx(&p.bar_full[buf], bytes);
  tc::bulk_g2s(dst, src, bytes, &p.bar_full[buf]);
}

// accumulate into TMEM columns [0, N).  Warps 0-7 stage (two statically named register sets ping-pong so that the
// loads of stage s + 1 are in flight while stage s is split and stored; a rotating `cur = next` copy would make
// every stage wait for a memory round trip: profiles/r1_tc_notes.md, finding 2) and hand each buffer to the issuer
// warp through an mbarrier (one arrive per staging warp); tcgen05.commit hands it back.
template <int NB>
__device__ __forceinline__ void gemm_mainloop(const GemmArgs& g, const Pipe& p, int row0, int kbase_a, int kbase_b,
                                              int col0) {
  const uint32_t pa = tc::plane_bytes(RM), pb = tc::plane_bytes(g.N);
  const uint32_t sbytes = stage_bytes_nn(g.N);
  const int nst = (g.K + RK - 1) / RK;
  if (threadIdx.x >= RT) {                                   // issuer warp
    const uint32_t idesc = tc::make_idesc_tf32(RM, g.N);
    const int ns = gemm_ring_depth<NB>(g.N);
    for (int s = 0; s < nst; ++s) {
      const int buf = s % ns;
      tc::mbar_wait(&p.bar_full[buf], static_cast<uint32_t>(s / ns) & 1u);
      issue_slab<RK / 8>(p, p.smem + buf * sbytes, pa, pb, idesc, s == 0, s == nst - 1, buf);
    }
    return;
  }
  const bool lane0 = (threadIdx.x & 31) == 0;
  if (NB == 0) {
    // bulk-fed B operand: the copy of stage s can only be issued once its ring slot is free, and its latency
    // (~1 K cycles for 33 KB) is exposed with two slots (the slot of stage s frees when the MMAs of stage s - 2 retire,
    // one MMA time before stage s is needed).  Three slots (tiles of <= 128 columns: 3 x 66 KB) give it two MMA times.
    const int ns = gemm_ring_depth<NB>(g.N);
    // lean A path: the generic loaders spend ~280 dependent instructions per warp and stage on addresses, bounds and the
    // activation switch (ncu: 8.7 cycles per issued instruction with 9 warps per SM = 2.4 K cycles per stage, the whole
    // stage time).  Per-thread row pointers and shared-memory offsets are hoisted; a stage that lies fully inside K takes
    // four predicated 16-byte loads.
    const int tid = threadIdx.x;
    const bool a_al = ((g.lda | kbase_a) & 3) == 0 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0;
    const float* ap[4];
    bool av[4];
    uint32_t so[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = tid + j * RT;
      const int q = idx & 7, r = idx >> 3;
      av[j] = row0 + r < g.M;
      ap[j] = g.A + static_cast<size_t>(av[j] ? row0 + r : 0) * g.lda + kbase_a + 4 * q;
      so[j] = q * pa + r * 16;
    }
    const int act_a = g.act_a;
    auto ld = [&](int s, float4 (&x)[4]) {
      const int k0 = s * RK;
      if (s < nst && a_al && k0 + RK <= g.K) {
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = av[j] ? __ldg(reinterpret_cast<const float4*>(ap[j] + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + 4 * ((tid + j * RT) & 7);
          x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (s < nst && av[j] && k < g.K) x[j] = ld4_guard(ap[j] + k0, k, g.K, a_al);
        }
      }
    };
    auto put = [&](int s, const float4 (&x)[4]) {
      const int buf = s % ns;
      if (s >= ns) tc::mbar_wait(&p.bar_empty[buf], static_cast<uint32_t>(s / ns - 1) & 1u);
      if (tid == 0) gemm_bulk_b(g, p, buf, s, nst, pa, pb, sbytes);
      uint8_t* a_hi = p.smem + buf * sbytes;
      uint8_t* a_lo = a_hi + (RK / 4) * pa;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 v = x[j];
        if (act_a == B200PPO_ACT_RELU) {
          v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        } else if (act_a != B200PPO_ACT_NONE) {
          v.x = act_fwd(v.x, act_a); v.y = act_fwd(v.y, act_a); v.z = act_fwd(v.z, act_a); v.w = act_fwd(v.w, act_a);
        }
        float4 hi, lo;
        tc::split4_fast(v, hi, lo);
        *reinterpret_cast<float4*>(a_hi + so[j]) = hi;
        *reinterpret_cast<float4*>(a_lo + so[j]) = lo;
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane0) tc::mbar_arrive(&p.bar_full[buf]);
    };
    float4 y0[4], y1[4];
    ld(0, y0);
    for (int s = 0; s < nst; s += 2) {
      ld(s + 1, y1);
      put(s, y0);
      if (s + 1 < nst) {
        ld(s + 2, y0);
        put(s + 1, y1);
      }
    }
    return;
  }
  uint32_t phase0 = 0u, phase1 = 0u;
  StageRegs x0, x1;
  gemm_load_stage<NB>(g, row0, kbase_a, kbase_b, col0, 0, nst, x0);
  for (int s = 0; s < nst; s += 2) {
    gemm_load_stage<NB>(g, row0, kbase_a, kbase_b, col0, s + 1, nst, x1);
    if (s >= 2) { tc::mbar_wait(&p.bar_empty[0], phase0); phase0 ^= 1u; }
    gemm_store_stage<NB>(g, p.smem, x0);
    tc::fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core
    __syncwarp();
    if (lane0) tc::mbar_arrive(&p.bar_full[0]);
    if (s + 1 < nst) {
      gemm_load_stage<NB>(g, row0, kbase_a, kbase_b, col0, s + 2, nst, x0);
      if (s >= 2) { tc::mbar_wait(&p.bar_empty[1], phase1); phase1 ^= 1u; }
      gemm_store_stage<NB>(g, p.smem + sbytes, x1);
      tc::fence_proxy_async();
      __syncwarp();
      if (lane0) tc::mbar_arrive(&p.bar_full[1]);
    }
  }
}

__device__ __forceinline__ void wait_acc(const Pipe& p) {
  tc::mbar_wait(p.bar_done, 0u);
  tc::tc_fence_after();
}

template <int NB>
__global__ void __launch_bounds__(RTI, 1) rg_gemm_kernel(const GemmArgs g) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[7];
  __shared__ uint32_t tmem_slot;
  Pipe p;
  pipe_init(p, smem, bars, &tmem_slot, g.N, RT / 32 + (NB == 0 ? 1 : 0), gemm_ring_depth<NB>(g.N));
  const int row0 = blockIdx.x * RM;
  const int tile = blockIdx.y, slice = blockIdx.z;
  const int col0 = g.b_col0 + tile * g.tile_stride;
  const int ccol0 = tile * g.tile_stride;
  const int kb_a = g.a_k0 + slice * g.k_slice, kb_b = g.b_k0 + slice * g.k_slice;
  gemm_mainloop<NB>(g, p, row0, kb_a, kb_b, col0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = warp & 3, cg = warp >> 2;
  const int row = row0 + sub * 32 + lane;
  float* Cb = g.C + static_cast<long long>(slice) * g.c_slice_stride;
  const bool vec_c = (g.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(Cb) & 15) == 0;
  const bool vec_c2 = g.C2 != nullptr && (g.ldc2 & 3) == 0 && (reinterpret_cast<uintptr_t>(g.C2) & 15) == 0;
  if (threadIdx.x < RT) wait_acc(p);
  for (int c = cg * 16; c < g.N && threadIdx.x < RT; c += 32) {
    float v[16];
    tc::tmem_ld16(p.tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + static_cast<uint32_t>(c), v);
    if (row < g.M) {
      const int n0 = ccol0 + c;
      // every global operand of the 16 columns is fetched before the first store (a load -> store -> load chain
      // per column cost one memory round trip each: 27 us for a 256-column tile, profiles/r2_recurrent_notes.md)
      float bz[16], zz[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        bz[i] = (g.bias != nullptr && n0 + i < g.n_real) ? g.bias[n0 + i] : 0.0f;
        zz[i] = (g.Zmul != nullptr && n0 + i < g.n_real) ? g.Zmul[static_cast<size_t>(row) * g.ldz + n0 + i] : 0.0f;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] += bz[i];
        if (g.Zmul != nullptr) v[i] *= act_grad(zz[i], g.act_z);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = n0 + 4 * i;
        if (n >= g.n_real) break;
        const float4 val = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        if (g.C == nullptr) {
          // only the activated copy is wanted (the env step's tanh): no pre-activation store
        } else if (g.c_planes) {
          *reinterpret_cast<float4*>(Cb + (static_cast<size_t>(row >> 7) * g.c_cols4 + (n >> 2)) * (RM * 4) + (row & (RM - 1)) * 4) = val;
        } else if (vec_c && n + 3 < g.n_real) {
          *reinterpret_cast<float4*>(Cb + static_cast<size_t>(row) * g.ldc + n) = val;
        } else {
          const float vv[4] = {val.x, val.y, val.z, val.w};
          for (int j = 0; j < 4; ++j)
            if (n + j < g.n_real) Cb[static_cast<size_t>(row) * g.ldc + n + j] = vv[j];
        }
        if (g.C2 != nullptr) {
          const float vv[4] = {val.x, val.y, val.z, val.w};
          if (vec_c2 && n + 3 < g.n_real) {
            *reinterpret_cast<float4*>(g.C2 + static_cast<size_t>(row) * g.ldc2 + n) =
                make_float4(act_fwd(vv[0], g.act_c2), act_fwd(vv[1], g.act_c2), act_fwd(vv[2], g.act_c2), act_fwd(vv[3], g.act_c2));
          } else {
            for (int j = 0; j < 4; ++j)
              if (n + j < g.n_real) g.C2[static_cast<size_t>(row) * g.ldc2 + n + j] = act_fwd(vv[j], g.act_c2);
          }
        }
      }
    }
  }
  pipe_fini(p);
}

#include "recurrent_tc_step.cuh"

// ------------------------------------------------------------------------------------------
// "TN" GEMM for the weight gradients:  part[s][m][n] = sum_{r in split s} A[r][a_col0 + m] * D[r][d_col0 + n]
// Both operands have the reduction index (rows) as the SLOW index in memory, so a slab of TK rows is copied
// raw into shared memory (coalesced) and transposed shared -> shared into the K-major operand planes
// (conflict-free LDS.32 down the rows, one STS.128 per 4-row chunk).  The column sums of D (bias gradients)
// fall out of the same pass.
// ------------------------------------------------------------------------------------------
constexpr int TK = 16;                     // rows per stage (2 MMAs of K = 8)
constexpr int TN_APITCH = RM + 4;          // floats

struct TnArgs {
  const float* A; int lda; int a_col0; int Kdim;     // m tiles of 128 over Kdim columns: blockIdx.x
  int act_a;
  const float* D; int ldd; int N; int n_log2; int n_real; int tile_stride;   // N = MMA N (power of two, 16..256); blockIdx.y
  int rows; int rows_per_split;                      // blockIdx.z
  float* part; long long part_stride; int ldp;       // part[s * part_stride + m * ldp + n]
  float* bpart; long long bpart_stride;              // nullable: bpart[s * bpart_stride + n] = sum_r D[r][n]
};

__host__ __device__ inline uint32_t tn_planes_bytes(int n) {
  return 2u * (TK / 4) * tc::plane_bytes(RM) + 2u * (TK / 4) * tc::plane_bytes(n);
}
__host__ __device__ inline uint32_t tn_raw_bytes(int n) { return static_cast<uint32_t>(TK) * (TN_APITCH + n + 4) * 4u; }

struct TnRegs {
  float4 a[2];
  float4 d[4];
};

__device__ __forceinline__ void tn_load(const TnArgs& t, int m0, int n0, int r_begin, int r_end, int s, TnRegs& x) {
  const int tid = threadIdx.x;
  const int rs = r_begin + s * TK;
  const bool a_al = ((t.lda | (t.a_col0 + m0)) & 3) == 0 && (reinterpret_cast<uintptr_t>(t.A) & 15) == 0;
  const bool d_al = ((t.ldd | n0) & 3) == 0 && (reinterpret_cast<uintptr_t>(t.D) & 15) == 0;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int idx = tid + j * RT;
    const int c4 = idx & 31, r = idx >> 5;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int m = 4 * c4;
    if (rs + r < r_end && m0 + m < t.Kdim)
      v = ld4_guard(t.A + static_cast<size_t>(rs + r) * t.lda + t.a_col0 + m0 + m, m0 + m, t.Kdim, a_al);
    x.a[j] = v;
  }
  const int n4 = t.N >> 2;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int idx = tid + j * RT;
    const int c4 = idx & (n4 - 1), r = idx >> (t.n_log2 - 2);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n = 4 * c4;
    if (r < TK && rs + r < r_end && n0 + n < t.n_real)
      v = ld4_guard(t.D + static_cast<size_t>(rs + r) * t.ldd + n0 + n, n0 + n, t.n_real, d_al);
    x.d[j] = v;
  }
}

__device__ __forceinline__ void tn_store_raw(const TnArgs& t, float* ra, float* rd, const TnRegs& x) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int idx = tid + j * RT;
    const int c4 = idx & 31, r = idx >> 5;
    float4 v = x.a[j];
    if (t.act_a == B200PPO_ACT_RELU) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    } else if (t.act_a != B200PPO_ACT_NONE) {
      v.x = act_fwd(v.x, t.act_a); v.y = act_fwd(v.y, t.act_a); v.z = act_fwd(v.z, t.act_a); v.w = act_fwd(v.w, t.act_a);
    }
    *reinterpret_cast<float4*>(ra + r * TN_APITCH + 4 * c4) = v;
  }
  const int n4 = t.N >> 2, dp = t.N + 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int idx = tid + j * RT;
    const int c4 = idx & (n4 - 1), r = idx >> (t.n_log2 - 2);
    if (r < TK) *reinterpret_cast<float4*>(rd + r * dp + 4 * c4) = x.d[j];
  }
}

// raw slab -> operand planes of buffer `st`; accumulates this thread's share of the column sums of D
__device__ __forceinline__ void tn_transpose(const TnArgs& t, const float* ra, const float* rd, uint8_t* st, float (&bsum)[4]) {
  const int tid = threadIdx.x;
  const uint32_t pa = tc::plane_bytes(RM), pb = tc::plane_bytes(t.N);
  uint8_t* a_hi = st;
  uint8_t* a_lo = a_hi + (TK / 4) * pa;
  uint8_t* b_hi = a_lo + (TK / 4) * pa;
  uint8_t* b_lo = b_hi + (TK / 4) * pb;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int idx = tid + j * RT;
    const int m = idx & (RM - 1), q = idx >> 7;        // 4 planes x 128 columns
    const float* src = ra + (4 * q) * TN_APITCH + m;
    const float4 v = make_float4(src[0], src[TN_APITCH], src[2 * TN_APITCH], src[3 * TN_APITCH]);
    float4 hi, lo;
    tc::split4_fast(v, hi, lo);
    *reinterpret_cast<float4*>(a_hi + q * pa + m * 16) = hi;
    *reinterpret_cast<float4*>(a_lo + q * pa + m * 16) = lo;
  }
  const int dp = t.N + 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int idx = tid + j * RT;
    const int n = idx & (t.N - 1), q = idx >> t.n_log2;   // N is a power of two
    if (q < TK / 4) {
      const float* src = rd + (4 * q) * dp + n;
      const float4 v = make_float4(src[0], src[dp], src[2 * dp], src[3 * dp]);
      bsum[j] += (v.x + v.y) + (v.z + v.w);
      float4 hi, lo;
      tc::split4_fast(v, hi, lo);
      *reinterpret_cast<float4*>(b_hi + q * pb + n * 16) = hi;
      *reinterpret_cast<float4*>(b_lo + q * pb + n * 16) = lo;
    }
  }
}

__global__ void __launch_bounds__(RTI, 1) rg_tn_kernel(const TnArgs t) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[5];
  __shared__ uint32_t tmem_slot;
  __shared__ float bred[4 * RT];
  Pipe p;
  pipe_init(p, smem, bars, &tmem_slot, t.N, 1);
  const int m0 = blockIdx.x * RM, n0 = blockIdx.y * t.tile_stride, sp = blockIdx.z;
  const int r_begin = sp * t.rows_per_split;
  int r_end = r_begin + t.rows_per_split;
  if (r_end > t.rows) r_end = t.rows;
  const int nst = r_end > r_begin ? (r_end - r_begin + TK - 1) / TK : 0;
  const uint32_t pbytes = tn_planes_bytes(t.N);
  float* ra = reinterpret_cast<float*>(smem + 2 * pbytes);
  float* rd = ra + TK * TN_APITCH;
  const uint32_t pa = tc::plane_bytes(RM), pb = tc::plane_bytes(t.N);
  const uint32_t idesc = tc::make_idesc_tf32(RM, t.N);
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool stager = threadIdx.x < RT;
  if (!stager) {                                             // issuer warp
    for (int s = 0; s < nst; ++s) {
      const int buf = s & 1;
      tc::mbar_wait(&p.bar_full[buf], static_cast<uint32_t>(s >> 1) & 1u);
      issue_slab<TK / 8>(p, p.smem + buf * pbytes, pa, pb, idesc, s == 0, s == nst - 1, buf);
    }
  } else {
    uint32_t phase0 = 0u, phase1 = 0u;
    TnRegs x0, x1;
    auto stager_sync = []() { asm volatile("bar.sync 1, %0;" ::"n"(RT) : "memory"); };
    if (nst > 0) tn_load(t, m0, n0, r_begin, r_end, 0, x0);
    for (int s = 0; s < nst; s += 2) {
      if (s + 1 < nst) tn_load(t, m0, n0, r_begin, r_end, s + 1, x1);
      tn_store_raw(t, ra, rd, x0);
      stager_sync();
      if (s >= 2) { tc::mbar_wait(&p.bar_empty[0], phase0); phase0 ^= 1u; }
      tn_transpose(t, ra, rd, p.smem, bsum);
      tc::fence_proxy_async();
      stager_sync();                       // planes complete; everybody is done reading the raw slab
      if (threadIdx.x == 0) tc::mbar_arrive(&p.bar_full[0]);
      if (s + 1 < nst) {
        if (s + 2 < nst) tn_load(t, m0, n0, r_begin, r_end, s + 2, x0);
        tn_store_raw(t, ra, rd, x1);
        stager_sync();
        if (s >= 2) { tc::mbar_wait(&p.bar_empty[1], phase1); phase1 ^= 1u; }
        tn_transpose(t, ra, rd, p.smem + pbytes, bsum);
        tc::fence_proxy_async();
        stager_sync();
        if (threadIdx.x == 0) tc::mbar_arrive(&p.bar_full[1]);
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = warp & 3, cg = warp >> 2;
  const int m = m0 + sub * 32 + lane;
  float* po = t.part + static_cast<long long>(sp) * t.part_stride;
  if (nst > 0 && stager) wait_acc(p);
  for (int c = cg * 16; c < t.N && stager; c += 32) {
    float v[16];
    if (nst > 0) {
      tc::tmem_ld16(p.tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + static_cast<uint32_t>(c), v);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.0f;
    }
    if (m < t.Kdim) {
      float* dst = po + static_cast<size_t>(m) * t.ldp + n0 + c;
      if ((t.ldp & 3) == 0 && n0 + c + 16 <= t.n_real && (reinterpret_cast<uintptr_t>(po) & 15) == 0) {
        // 64 contiguous bytes per lane: whole sectors (scalar stores at a row stride wrote 8x the bytes)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (n0 + c + i < t.n_real) dst[i] = v[i];
      }
    }
  }
  if (t.bpart != nullptr && blockIdx.x == 0) {          // column sums: thread (n, q) partials -> fixed-order sum over q
    if (stager) {
#pragma unroll
      for (int j = 0; j < 4; ++j) bred[threadIdx.x + j * RT] = bsum[j];
    }
    __syncthreads();
    for (int n = threadIdx.x; n < t.N && stager; n += RT) {
      float s = 0.0f;
      for (int q = 0; q < TK / 4; ++q) {
        const int idx = q * t.N + n;                     // chunk id of (n, q): idx = tid + j * RT
        s += bred[idx];
      }
      if (n0 + n < t.n_real) t.bpart[static_cast<long long>(sp) * t.bpart_stride + n0 + n] = s;
    }
  }
  pipe_fini(p);
}

// sum of S partial matrices in fixed order
__global__ void __launch_bounds__(256) rg_reduce_kernel(const float* __restrict__ part, long long n, int S,
                                                        long long stride, float* __restrict__ out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float s = 0.0f;
  int sp = 0;
  for (; sp + 8 <= S; sp += 8) {                       // eight loads in flight, summed in index order
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = part[(sp + j) * stride + i];
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
  }
  for (; sp < S; ++sp) s += part[sp * stride + i];
  out[i] = s;
}

// several fixed-order partial reductions in one launch (the three weight-gradient GEMMs of a sequence backward
// and their bias sums): block b belongs to the segment whose block range contains it
struct ReduceSegs {
  int n_seg;
  int blk0[7];                                  // first block of segment i; blk0[n_seg] = total
  const float* part[6]; long long n[6]; int S[6]; long long stride[6]; float* out[6];
};
__global__ void __launch_bounds__(256) rg_reduce_multi_kernel(const ReduceSegs g) {
  int q = 0;
  while (q + 1 < g.n_seg && static_cast<int>(blockIdx.x) >= g.blk0[q + 1]) ++q;
  const long long i = (static_cast<long long>(blockIdx.x) - g.blk0[q]) * blockDim.x + threadIdx.x;
  if (i >= g.n[q]) return;
  const float* part = g.part[q];
  const long long stride = g.stride[q];
  const int S = g.S[q];
  float s = 0.0f;
  int sp = 0;
  for (; sp + 8 <= S; sp += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = part[(sp + j) * stride + i];
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
  }
  for (; sp < S; ++sp) s += part[sp * stride + i];
  g.out[q][i] = s;
}

// carry in / out of the sequence workspace: dst[r][0..H) = src[r][0..H) with row strides
__global__ void __launch_bounds__(256) rg_copy_rows_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst,
                                                           int ldd, int rows, int H) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * H) return;
  const int r = idx / H, k = idx - r * H;
  dst[static_cast<size_t>(r) * ldd + k] = src[static_cast<size_t>(r) * lds + k];
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct SeqLayout {
  // row-major matrices over the R = T * rows row space
  size_t z1, cat, hn, da, dz1;
  // plane layouts (recurrent_tc_step.cuh)
  size_t gxp, dhpp;                       // over the R row space: [R / 128][cols / 4][PLC]
  size_t hp;                              // A planes of h: [2 buffers][2 halves][tiles][H / 4][PLA]
  size_t cp, dcp;                         // state planes [tiles][H / 4][PLC]
  size_t cache[6];                        // gi gf gg go tc cin: [T][tiles][H / 4][PLC]
  size_t dap;                             // A planes of da: [2 halves][tiles][4 gates][H / 4][PLA]
  size_t dhr;                             // split-K partials [4 * KS][tiles][H / 4][PLC]
  int KS;                                 // k parts per gate block in the backward split-K GEMM
  size_t whf, whb;                        // weight planes
  size_t flags;                           // int32 [tiles][T + 1] arrival counters of the persistent forward kernel
  size_t mdc, mdh, ipart;                 // learned initial carry: masked per-step gradient planes, per-step sums
  size_t part, bpart, part1, bpart1, part2, bpart2, total;   // partial sums of the three weight-gradient GEMMs
  int tiles, NBT, nbt;
  size_t hp_half, hp_buf, st_tile, cache_step, dap_half;
  int S_cat, S_w1, S_w2;
};

inline size_t al64(size_t x) { return (x + 63) & ~static_cast<size_t>(63); }

int split_for(int tiles, int rows) {
  // one wave of CTAs: twice as many splits of half the rows cost the same tensor time but double the partial-sum
  // traffic (S x Kdim x N floats written and read back) and the per-CTA set-up
  int S = 148 / (tiles > 0 ? tiles : 1);
  const int smax = (rows + 4 * TK - 1) / (4 * TK);
  if (S > smax) S = smax;
  if (S > 64) S = 64;
  if (S < 1) S = 1;
  return S;
}

int mma_n(int n);

SeqLayout seq_layout(const b200ppo_lstm_plan& p, int T, int rows) {
  SeqLayout L;
  const size_t R = static_cast<size_t>(T) * rows;
  const size_t P = p.pre_dim, H = p.hidden, Y = p.out_dim;
  const size_t planes = H / 4;
  L.tiles = cdiv(rows, RM);
  const size_t tiles = L.tiles, Rtiles = cdiv(static_cast<int64_t>(R), RM);
  L.NBT = mma_n(static_cast<int>(H < 64 ? H : 64));
  L.nbt = cdiv(static_cast<int64_t>(H), L.NBT);
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o = al64(o + n); return r; };
  L.z1 = take(R * P);
  L.cat = take((R + rows) * (P + H));
  L.hn = take(R * H);
  L.da = take(R * 4 * H);
  L.dz1 = take(R * P);
  L.gxp = take(Rtiles * H * PLC);                     // 4H / 4 = H planes per row tile
  L.dhpp = take(Rtiles * planes * PLC);
  L.hp_half = tiles * planes * PLA;
  L.hp_buf = 2 * L.hp_half;
  L.hp = take(2 * L.hp_buf);
  L.st_tile = planes * PLC;
  L.cp = take(tiles * L.st_tile);
  L.dcp = take(tiles * L.st_tile);
  L.cache_step = tiles * L.st_tile;
  for (int i = 0; i < 6; ++i) L.cache[i] = take(static_cast<size_t>(T) * L.cache_step);
  L.dap_half = tiles * 4 * planes * PLA;
  L.dap = take(2 * L.dap_half);
  // a minibatch of few row tiles leaves SMs idle: cut each gate block's K in two as well (needs 32-k slabs)
  L.KS = (H % 64 == 0 && tiles * L.nbt * 8 <= 2 * 148) ? 2 : 1;
  L.dhr = take(4 * L.KS * tiles * L.st_tile);
  L.whf = take((H / UT) * 2 * planes * plb(4 * UT));
  L.whb = take(static_cast<size_t>(L.nbt) * 4 * 2 * planes * plb(L.NBT));
  L.flags = take(tiles * (static_cast<size_t>(T) + 1));
  const size_t tr = p.init_c_off > 0 ? 1 : 0;
  L.mdc = take(tr * T * L.cache_step);
  L.mdh = take(tr * T * L.cache_step);
  L.ipart = take(tr * T * 2 * H);
  const int Rr = static_cast<int>(R);
  L.S_cat = split_for(cdiv(P + H, RM) * cdiv(4 * H, 256), Rr);
  L.S_w1 = split_for(cdiv(p.obs_dim, RM), Rr);
  L.S_w2 = split_for(cdiv(H, RM), Rr);
  L.part = take(static_cast<size_t>(L.S_cat) * (P + H) * 4 * H);
  L.bpart = take(64 * 4 * H);
  L.part1 = take(static_cast<size_t>(L.S_w1) * p.obs_dim * P);
  L.bpart1 = take(64 * P);
  L.part2 = take(static_cast<size_t>(L.S_w2) * H * Y);
  L.bpart2 = take(64 * Y);
  L.total = o;
  return L;
}

constexpr int PERSIST_SMEM_MAX = 220 * 1024;   // dynamic shared memory of the persistent kernel (227 KB minus static, with margin)
int g_seq_persist = -1;
// one persistent launch for the forward recurrence: T > 1 and every CTA resident at once (one per SM).
// B200PPO_LSTM_PERSIST=1 / b200ppo_lstm_set_persistent(1) selects it; default: one launch per step.
bool seq_persistent(const b200ppo_lstm_plan& p, int T, int rows) {
  if (g_seq_persist < 0) {
    const char* e = std::getenv("B200PPO_LSTM_PERSIST");
    // Off unless asked for: measured on B200 (configs[2], profiles/r2_recurrent_notes.md) the persistent launch
    // is parity-identical but not faster inside the captured iteration (46.4 vs 45.3 ms): a step is bound by the
    // carry hand-off + 264 KB operand stream + gate epilogue, not by the kernel boundary.
    g_seq_persist = (e && e[0] == '1') ? 1 : 0;
  }
  return g_seq_persist != 0 && T > 1 && cdiv(rows, RM) * (p.hidden / UT) <= b200ppo_num_sms();
}

// B200PPO_LSTM_PAIR=0 disables the cluster-pair forward step (A/B measurements)
bool seq_pair_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = std::getenv("B200PPO_LSTM_PAIR");
    mode = (e && e[0] == '0') ? 0 : 1;
  }
  return mode != 0;
}

int check_plan_tc(const b200ppo_lstm_plan* p) {
  if (!p || p->obs_dim <= 0 || p->pre_dim <= 0 || p->hidden <= 0 || p->out_dim <= 0 || p->n_params <= 0) return B200PPO_EINVAL;
  if (p->act < 0 || p->act > 3) return B200PPO_EINVAL;
  // 16-unit tiles, and the carry inside [u | h] rows must stay 16-byte aligned; the FFMA step kernels
  // (b200ppo_lstm_step_fwd / _bwd) have no such limits
  if (p->hidden % UT || p->pre_dim % 4) return B200PPO_ELIMIT;
  if (p->pre_dim > 4096 || p->out_dim > 4096 || p->obs_dim > 65536) return B200PPO_ELIMIT;
  // learned initial carry: both or neither, inside the arena, 16-byte aligned
  if ((p->init_c_off > 0) != (p->init_h_off > 0)) return B200PPO_EINVAL;
  if (p->init_c_off > 0) {
    if ((p->init_c_off & 3) || (p->init_h_off & 3)) return B200PPO_EALIGN;
    if (p->init_c_off + p->hidden > p->n_params || p->init_h_off + p->hidden > p->n_params) return B200PPO_EINVAL;
  }
  return 0;
}

int mma_n(int n) {          // MMA N for n valid columns: a power of two in [16, 256]
  int v = 16;
  while (v < n && v < 256) v <<= 1;
  return v;
}

bool g_attr_tc = false;
int set_attrs_tc() {
  if (g_attr_tc) return 0;
  const int big = 2 * static_cast<int>(stage_bytes_nn(256));
  cudaError_t e;
  e = cudaFuncSetAttribute(rg_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           big > 3 * static_cast<int>(stage_bytes_nn(128)) ? big : 3 * static_cast<int>(stage_bytes_nn(128)));
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(rg_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(rg_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(rg_gemm_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(lstm_step_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(SP_NS * sp_slot_bytes(4 * UT)));
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(lstm_step_fwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(FWD3_XB_BYTES + SP_NS * sp_slot_bytes(4 * UT)));
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(lstm_seq_fwd_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PERSIST_SMEM_MAX);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(lstm_step_bwd_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(SP_NS * sp_slot_bytes(64)));
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(rg_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(2 * tn_planes_bytes(256) + tn_raw_bytes(256)));
  if (e != cudaSuccess) return static_cast<int>(e);
  g_attr_tc = true;
  return 0;
}

GemmArgs gemm_defaults() {
  GemmArgs g;
  g.A = nullptr; g.lda = 0; g.a_k0 = 0; g.M = 0; g.K = 0; g.act_a = B200PPO_ACT_NONE; g.a_mean = nullptr; g.a_std = nullptr;
  g.B = nullptr; g.ldb = 0; g.b_nt = 0; g.b_col0 = 0; g.b_unit_log2 = 30; g.b_gate_stride = 0; g.b_k0 = 0; g.Bplanes = nullptr;
  g.N = 16; g.n_log2 = 4; g.n_real = 0; g.tile_stride = 0;
  g.C = nullptr; g.ldc = 0; g.c_planes = 0; g.c_cols4 = 0; g.bias = nullptr; g.C2 = nullptr; g.ldc2 = 0; g.act_c2 = B200PPO_ACT_NONE;
  g.Zmul = nullptr; g.ldz = 0; g.act_z = B200PPO_ACT_NONE;
  g.k_slice = 0; g.c_slice_stride = 0;
  return g;
}

int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// launch the generic GEMM over n_real columns (tiles of at most n_tile columns) and `slices` k slices
int launch_gemm(cudaStream_t s, GemmArgs g, int n_real, int slices, int n_tile = 256) {
  g.n_real = n_real;
  g.N = mma_n(n_real < n_tile ? n_real : n_tile);
  g.n_log2 = ilog2(g.N);
  g.tile_stride = g.N;
  const int tiles = cdiv(n_real, g.N);
  const dim3 grid(cdiv(g.M, RM), tiles, slices);
  const size_t smem = 2 * static_cast<size_t>(stage_bytes_nn(g.N));
  if (g.Bplanes != nullptr && g.a_mean != nullptr) return B200PPO_EINVAL;     // the bulk-fed kernel's A path does not normalise
  if (g.Bplanes != nullptr) rg_gemm_kernel<0><<<grid, RTI, gemm_ring_depth<0>(g.N) * static_cast<size_t>(stage_bytes_nn(g.N)), s>>>(g);
  else if (g.N <= 32) rg_gemm_kernel<1><<<grid, RTI, smem, s>>>(g);
  else if (g.N <= 64) rg_gemm_kernel<2><<<grid, RTI, smem, s>>>(g);
  else rg_gemm_kernel<8><<<grid, RTI, smem, s>>>(g);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

// weight gradient out[Kdim][n_real] (+ bias gradient) = A^T D over `rows` rows, S row splits, fixed-order reduction
// (segs == nullptr: reduced right away; else the reductions are appended to *segs for one rg_reduce_multi_kernel)
int launch_tn(cudaStream_t s, const float* A, int lda, int a_col0, int Kdim, int act_a, const float* D, int ldd, int n_real,
              int rows, int S, float* part, float* bpart, float* gw, float* gb, ReduceSegs* segs = nullptr) {
  TnArgs t;
  t.A = A; t.lda = lda; t.a_col0 = a_col0; t.Kdim = Kdim; t.act_a = act_a;
  t.D = D; t.ldd = ldd; t.n_real = n_real; t.N = mma_n(n_real); t.n_log2 = ilog2(t.N); t.tile_stride = t.N;
  t.rows = rows;
  t.rows_per_split = cdiv(cdiv(rows, S), TK) * TK;
  S = cdiv(rows, t.rows_per_split);
  t.part = part; t.part_stride = static_cast<long long>(Kdim) * n_real; t.ldp = n_real;
  t.bpart = gb != nullptr ? bpart : nullptr; t.bpart_stride = n_real;
  const dim3 grid(cdiv(Kdim, RM), cdiv(n_real, t.N), S);
  const size_t smem = 2 * static_cast<size_t>(tn_planes_bytes(t.N)) + tn_raw_bytes(t.N);
  rg_tn_kernel<<<grid, RTI, smem, s>>>(t);
  B200PPO_LAUNCH_CHECK();
  const long long n = static_cast<long long>(Kdim) * n_real;
  if (segs != nullptr) {
    auto add = [&](const float* p, long long cnt, float* out) {
      const int q = segs->n_seg++;
      segs->part[q] = p; segs->n[q] = cnt; segs->S[q] = S; segs->stride[q] = cnt; segs->out[q] = out;
      segs->blk0[q + 1] = segs->blk0[q] + static_cast<int>(cdiv(cnt, 256));
    };
    add(part, n, gw);
    if (gb != nullptr) add(bpart, n_real, gb);
    return 0;
  }
  rg_reduce_kernel<<<cdiv(n, 256), 256, 0, s>>>(part, n, S, n, gw);
  B200PPO_LAUNCH_CHECK();
  if (gb != nullptr) {
    rg_reduce_kernel<<<cdiv(n_real, 256), 256, 0, s>>>(bpart, n_real, S, n_real, gb);
    B200PPO_LAUNCH_CHECK();
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// Batched ("wide") rollout on the synthetic env: rollout.single_transition / unroll_env (rollout.py:11-73) as a
// per-step launch sequence for networks whose weights do not fit an SM's shared memory (BASELINE configs[3]:
// 768-wide block-diagonal encoder layers and a 789 x 768 env matrix).  The fused kernel (csrc/rollout.cu) keeps a
// 16-env tile per CTA and re-streams every weight from L2 per tile and step (2.7 GB per step at 8192 envs:
// 1.26 ms per step); here every Dense layer of a step is ONE tcgen05 tile GEMM over all B envs (128-env row
// tiles: the weights are read 64 times per step instead of 512), the element-wise parts are three small kernels:
//   init (once)   X[:, :O] <- env obs; obs[0] <- env obs
//   per step      L layer GEMMs (normalise / activation on load, bias in the epilogue)
//                 sampler: y -> raw action, action (record + the env input tile X[:, O:]), log-likelihood
//                 env GEMM: obs' = tanh([obs | action] [Wo; Wa])
//                 bookkeeping: counters, done / truncation, reward, next_obs[-1], reset select -> X, obs[t + 1]
// Same arithmetic per element as the fused kernels (sampler_elem / synth_reset_scalars in common.cuh), integer
// bookkeeping bit-identical.
// ------------------------------------------------------------------------------------------
// floats of the pre-split B planes of a [K][n_real] weight matrix cut into column tiles of mma_n(min(n_real, n_tile))
inline int planes_ntile(int n_real, int n_tile) { return mma_n(n_real < n_tile ? n_real : n_tile); }
inline size_t planes_floats(int K, int n_real, int n_tile) {
  const int N = planes_ntile(n_real, n_tile);
  return static_cast<size_t>(cdiv(n_real, N)) * cdiv(K, RK) * 2 * (RK / 4) * (N + 1) * 4;
}
// W [K][ldb] row-major -> per (column tile, 32-k stage) the stage's B area exactly as gemm_store_stage lays it out:
// hi planes q = 0..7 ([N + 1] float4 each, entry n = the 4 consecutive k of column n), then the lo planes
__global__ void __launch_bounds__(256) rg_prep_b_kernel(const float* __restrict__ W, int ldb, int K, int n_real, int N,
                                                        int nst, int tiles, float4* __restrict__ out) {
  const int prow = N + 1;
  const size_t total = static_cast<size_t>(tiles) * nst * (RK / 4) * prow;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(idx % prow);
    size_t r = idx / prow;
    const int q = static_cast<int>(r % (RK / 4)); r /= (RK / 4);
    const int s = static_cast<int>(r % nst), tile = static_cast<int>(r / nst);
    const int k = s * RK + 4 * q, col = tile * N + n;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (n < N && col < n_real)
      for (int j = 0; j < 4; ++j)
        if (k + j < K) x[j] = W[static_cast<size_t>(k + j) * ldb + col];
    float4 hi, lo;
    tc::split4_fast(make_float4(x[0], x[1], x[2], x[3]), hi, lo);
    const size_t base = (static_cast<size_t>(tile) * nst + s) * 2 * (RK / 4) * prow;
    out[base + static_cast<size_t>(q) * prow + n] = hi;
    out[base + static_cast<size_t>(RK / 4 + q) * prow + n] = lo;
  }
}

struct WideLayout {
  int ldx, ldz, ldo;
  int n_tile[B200PPO_MAX_LAYERS + 1];                 // column-tile width request of every layer GEMM; [L] = the env GEMM
  size_t planes[B200PPO_MAX_LAYERS + 1];              // pre-split weight planes (rg_prep_b_kernel)
  size_t x, xn, za, zb, zenv, ynext, total;
};
WideLayout wide_layout(const b200ppo_plan& p, int B) {
  WideLayout L;
  const int O = p.obs_dim, A = p.act_dim;
  int mx = 0;
  for (int l = 0; l < p.actor.n_layers; ++l) mx = p.actor.dims[l + 1] > mx ? p.actor.dims[l + 1] : mx;
  L.ldx = (O + A + 3) & ~3;
  L.ldz = (mx + 3) & ~3;
  L.ldo = (O + 3) & ~3;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o = al64(o + n); return r; };
  L.x = take(static_cast<size_t>(B) * L.ldx);
  L.xn = take(static_cast<size_t>(B) * L.ldo);       // normalised observations: the first layer's A operand
  L.za = take(static_cast<size_t>(B) * L.ldz);
  L.zb = take(static_cast<size_t>(B) * L.ldz);
  L.zenv = take(static_cast<size_t>(B) * L.ldo);
  L.ynext = take(static_cast<size_t>(B) * L.ldo);
  const int row_tiles = cdiv(B, RM);
  // 128-column tiles: three operand stages fit shared memory (B200PPO_WIDE_NTILE=256: 256-column tiles with a two-slot
  // ring on layers with enough row tiles to fill the device)
  static const int want256 = [] { const char* e = std::getenv("B200PPO_WIDE_NTILE"); return e && std::atoi(e) == 256 ? 1 : 0; }();
  auto n_tile_for = [&](int N) { return (want256 && row_tiles * cdiv(N, 256) >= b200ppo_num_sms()) ? 256 : 128; };
  for (int l = 0; l < p.actor.n_layers; ++l) {
    L.n_tile[l] = n_tile_for(p.actor.dims[l + 1]);
    L.planes[l] = take(planes_floats(p.actor.dims[l], p.actor.dims[l + 1], L.n_tile[l]));
  }
  L.n_tile[p.actor.n_layers] = n_tile_for(O);
  L.planes[p.actor.n_layers] = take(planes_floats(O + A, O, L.n_tile[p.actor.n_layers]));
  L.total = o;
  return L;
}

__global__ void __launch_bounds__(256) wide_init_kernel(const float* __restrict__ env_obs, int B, int O, float* __restrict__ X,
                                                        int ldx, float* __restrict__ obs0, float* __restrict__ Xn, int ldo,
                                                        const float* __restrict__ mean, const float* __restrict__ stdv) {
  const size_t n = static_cast<size_t>(B) * O;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t e = i / O, o = i - e * O;
    const float v = env_obs[i];
    X[e * ldx + o] = v;
    obs0[i] = v;
    if (Xn != nullptr) Xn[e * ldo + o] = mean != nullptr ? __fdiv_rn(v - mean[o], stdv[o]) : v;
  }
}

// one thread per (env, action dim): count = count0 + 2 t (the entropy draw does not influence the rollout)
__global__ void __launch_bounds__(256) wide_sampler_kernel(const float* __restrict__ y, int ldy, int B, int A, float min_std,
                                                           float std_scale, const uint32_t* __restrict__ rng_state, uint32_t t,
                                                           float* __restrict__ raw, float* __restrict__ action,
                                                           float* __restrict__ X, int ldx, int O, float* __restrict__ llterm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * A) return;
  const int e = i / A, d = i - e * A;
  const Key stream_key{rng_state[0], rng_state[1]};
  const Key k_sample = fold_in(stream_key, rng_state[2] + 2u * t);
  const SamplerOut s = sampler_elem(y[static_cast<size_t>(e) * ldy + d], y[static_cast<size_t>(e) * ldy + A + d], min_std, std_scale,
                                    0.0f, 0, 0.0f, k_sample, k_sample, static_cast<uint32_t>(i), false);
  raw[i] = s.raw;
  action[i] = s.action;
  X[static_cast<size_t>(e) * ldx + O + d] = s.action;
  llterm[i] = s.llterm;
}

struct WideBookArgs {
  int B, O, A, T, t, max_len, term_thresh16, ldx, ldo;
  const uint32_t* iter_keys;
  const float* ynext; const float* llterm;
  float* X; float* obs_next;            // obs[t + 1] (nullptr at the last step)
  float* Xn; const float* mean; const float* stdv;   // normalised next observation (normalizer.py:78-80; mean == nullptr: none)
  float* loglik; float* reward; uint8_t* done; uint8_t* trunc;     // row t of the record
  float* next_obs_last; float* env_obs;                            // written at the last step only
  int32_t* env_counter; uint32_t* env_term;                        // advanced in place every step
};
// one warp per env (rollout.cu steps (4) and (6))
__global__ void __launch_bounds__(256) wide_book_kernel(const WideBookArgs a) {
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= a.B) return;
  int dn = 0;
  uint32_t kba = 0u, kbb = 0u;
  if (lane == 0) {
    float ll = 0.0f;
    for (int d = 0; d < a.A; ++d) ll += a.llterm[static_cast<size_t>(e) * a.A + d];
    const int32_t c = a.env_counter[e] + 1;
    const uint32_t ts = a.env_term[e] * 1664525u + 1013904223u;
    const bool terminated = (ts >> 16) < static_cast<uint32_t>(a.term_thresh16);
    const bool truncated = c >= a.max_len;
    dn = (terminated || truncated) ? 1 : 0;
    a.loglik[e] = ll;
    a.done[e] = dn ? 1 : 0;
    a.trunc[e] = truncated ? 1 : 0;
    if (dn) {
      const Key reset_key{a.iter_keys[0], a.iter_keys[1]};
      const Key k = split_at(reset_key, static_cast<uint32_t>(a.t) * static_cast<uint32_t>(a.B) + static_cast<uint32_t>(e));
      const ResetScalars r = synth_reset_scalars(k, a.max_len);
      a.env_counter[e] = r.counter;
      a.env_term[e] = r.term;
      kba = r.k_base.a; kbb = r.k_base.b;
    } else {
      a.env_counter[e] = c;
      a.env_term[e] = ts;
    }
  }
  dn = __shfl_sync(0xffffffffu, dn, 0);
  kba = __shfl_sync(0xffffffffu, kba, 0);
  kbb = __shfl_sync(0xffffffffu, kbb, 0);
  const Key kb{kba, kbb};
  const bool last = a.t == a.T - 1;
  float sq = 0.0f;
  auto put = [&](int o, float v) {
    if (last) a.next_obs_last[static_cast<size_t>(e) * a.O + o] = v;
    const float nv = dn ? bits_to_normal(random_bits_at(kb, static_cast<uint32_t>(o))) : v;
    a.X[static_cast<size_t>(e) * a.ldx + o] = nv;
    if (a.Xn != nullptr) a.Xn[static_cast<size_t>(e) * a.ldo + o] = a.mean != nullptr ? __fdiv_rn(nv - a.mean[o], a.stdv[o]) : nv;
    if (a.obs_next != nullptr) a.obs_next[static_cast<size_t>(e) * a.O + o] = nv;
    if (last) a.env_obs[static_cast<size_t>(e) * a.O + o] = nv;
  };
  if ((a.O & 3) == 0) {
    // four columns per lane and trip (lane l owns columns 128 i + 4 l .. + 3), 16-byte loads and stores: a fixed
    // summation order of its own, the reward differs from the fused kernels' in the last bits only
    const bool norm = a.mean != nullptr;
    for (int o = 4 * lane; o < a.O; o += 128) {
      const float4 v = *reinterpret_cast<const float4*>(a.ynext + static_cast<size_t>(e) * a.ldo + o);
      sq = fmaf(v.x, v.x, sq); sq = fmaf(v.y, v.y, sq); sq = fmaf(v.z, v.z, sq); sq = fmaf(v.w, v.w, sq);
      float4 nv = v;
      if (dn) {
        nv.x = bits_to_normal(random_bits_at(kb, static_cast<uint32_t>(o)));
        nv.y = bits_to_normal(random_bits_at(kb, static_cast<uint32_t>(o + 1)));
        nv.z = bits_to_normal(random_bits_at(kb, static_cast<uint32_t>(o + 2)));
        nv.w = bits_to_normal(random_bits_at(kb, static_cast<uint32_t>(o + 3)));
      }
      float4 xn = nv;
      if (norm) {
        const float4 m = *reinterpret_cast<const float4*>(a.mean + o), sd = *reinterpret_cast<const float4*>(a.stdv + o);
        xn.x = __fdiv_rn(nv.x - m.x, sd.x); xn.y = __fdiv_rn(nv.y - m.y, sd.y);
        xn.z = __fdiv_rn(nv.z - m.z, sd.z); xn.w = __fdiv_rn(nv.w - m.w, sd.w);
      }
      *reinterpret_cast<float4*>(a.X + static_cast<size_t>(e) * a.ldx + o) = nv;
      if (a.Xn != nullptr) *reinterpret_cast<float4*>(a.Xn + static_cast<size_t>(e) * a.ldo + o) = xn;
      if (a.obs_next != nullptr) *reinterpret_cast<float4*>(a.obs_next + static_cast<size_t>(e) * a.O + o) = nv;
      if (last) {
        *reinterpret_cast<float4*>(a.next_obs_last + static_cast<size_t>(e) * a.O + o) = v;
        *reinterpret_cast<float4*>(a.env_obs + static_cast<size_t>(e) * a.O + o) = nv;
      }
    }
  } else {
    for (int o = lane; o < a.O; o += 32) {
      const float v = a.ynext[static_cast<size_t>(e) * a.ldo + o];
      sq = fmaf(v, v, sq);
      put(o, v);
    }
  }
  sq = warp_sum(sq);
  if (lane == 0) a.reward[e] = -(sq / static_cast<float>(a.O));
}

}  // namespace

namespace b200ppo {
int64_t rollout_wide_ws_floats(const b200ppo_plan* plan, int B) {
  return static_cast<int64_t>(wide_layout(*plan, B).total + static_cast<size_t>(B) * plan->act_dim + 64);
}
int rollout_wide_num_launches(const b200ppo_plan* plan, int T) {
  return plan->actor.n_layers + 2 + T * (plan->actor.n_layers + 3);     // weight planes + init, then the per-step sequence
}

int rollout_wide(cudaStream_t s, const RolloutWideArgs& a) {
  int rc = set_attrs_tc();
  if (rc) return rc;
  const b200ppo_plan& p = *a.plan;
  const int O = p.obs_dim, A = p.act_dim, B = a.B, T = a.T, L = p.actor.n_layers;
  const WideLayout W = wide_layout(p, B);
  float* X = a.ws + W.x;
  float* Xn = a.ws + W.xn;
  const float* nmean = p.normalize ? a.mean : nullptr;
  const float* nstd = p.normalize ? a.std : nullptr;
  float* zbuf[2] = {a.ws + W.za, a.ws + W.zb};
  float* zenv = a.ws + W.zenv;
  float* ynext = a.ws + W.ynext;
  float* llterm = a.ws + W.total;
  // the weights are constant over the rollout: split them once into the GEMMs' shared-memory stage layout
  auto prep = [&](const float* Wm, int K, int N, int li) -> int {
    const int Nt = planes_ntile(N, W.n_tile[li]);
    const int tiles = cdiv(N, Nt), nst = cdiv(K, RK);
    const size_t total = static_cast<size_t>(tiles) * nst * (RK / 4) * (Nt + 1);
    rg_prep_b_kernel<<<cdiv(static_cast<int64_t>(total), 256), 256, 0, s>>>(Wm, N, K, N, Nt, nst, tiles,
                                                                          reinterpret_cast<float4*>(a.ws + W.planes[li]));
    B200PPO_LAUNCH_CHECK();
    return 0;
  };
  for (int l = 0; l < L; ++l) {
    rc = prep(a.params + p.actor.w_off[l], p.actor.dims[l], p.actor.dims[l + 1], l);
    if (rc) return rc;
  }
  rc = prep(a.Wenv, O + A, O, L);
  if (rc) return rc;
  wide_init_kernel<<<cdiv(static_cast<int64_t>(B) * O, 256 * 8), 256, 0, s>>>(a.env_obs, B, O, X, W.ldx, a.obs, Xn, W.ldo, nmean, nstd);
  B200PPO_LAUNCH_CHECK();
  for (int t = 0; t < T; ++t) {
    const size_t row0 = static_cast<size_t>(t) * B;
    const float* in = Xn;
    int ldin = W.ldo;
    for (int l = 0; l < L; ++l) {
      const int K = p.actor.dims[l], N = p.actor.dims[l + 1];
      GemmArgs g = gemm_defaults();
      g.A = in; g.lda = ldin; g.M = B; g.K = K;
      if (l > 0) g.act_a = p.actor.act;
      g.B = a.params + p.actor.w_off[l]; g.ldb = N;
      g.Bplanes = a.ws + W.planes[l];
      g.bias = a.params + p.actor.b_off[l];
      g.C = zbuf[l & 1]; g.ldc = W.ldz;
      rc = launch_gemm(s, g, N, 1, W.n_tile[l]);
      if (rc) return rc;
      in = zbuf[l & 1];
      ldin = W.ldz;
    }
    wide_sampler_kernel<<<cdiv(static_cast<int64_t>(B) * A, 256), 256, 0, s>>>(
        in, ldin, B, A, p.min_std, p.std_scale, a.rng_state, static_cast<uint32_t>(t), a.raw_action + row0 * A,
        a.action + row0 * A, X, W.ldx, O, llterm);
    B200PPO_LAUNCH_CHECK();
    {
      GemmArgs g = gemm_defaults();
      g.A = X; g.lda = W.ldx; g.M = B; g.K = O + A;
      g.B = a.Wenv; g.ldb = O;
      g.Bplanes = a.ws + W.planes[L];
      g.C = nullptr; g.ldc = W.ldo;
      g.C2 = ynext; g.ldc2 = W.ldo; g.act_c2 = B200PPO_ACT_TANH;
      rc = launch_gemm(s, g, O, 1, W.n_tile[L]);
      if (rc) return rc;
    }
    WideBookArgs b;
    b.B = B; b.O = O; b.A = A; b.T = T; b.t = t; b.max_len = a.max_len; b.term_thresh16 = a.term_thresh16;
    b.ldx = W.ldx; b.ldo = W.ldo; b.iter_keys = a.iter_keys; b.ynext = ynext; b.llterm = llterm; b.X = X; b.Xn = Xn; b.mean = nmean; b.stdv = nstd;
    b.obs_next = t + 1 < T ? a.obs + (row0 + B) * O : nullptr;
    b.loglik = a.loglik + row0; b.reward = a.reward + row0; b.done = a.done + row0; b.trunc = a.trunc + row0;
    b.next_obs_last = a.next_obs_last; b.env_obs = a.env_obs; b.env_counter = a.env_counter; b.env_term = a.env_term;
    wide_book_kernel<<<cdiv(B, 8), 256, 0, s>>>(b);
    B200PPO_LAUNCH_CHECK();
  }
  return 0;
}
}  // namespace b200ppo

// ------------------------------------------------------------------------------------------
// The synthetic env's side of one rollout step for policies evaluated elsewhere (the recurrent actor): sampler on the
// actor outputs, env GEMM, bookkeeping - the last three launches of the batched rollout above, same kernels.
// ------------------------------------------------------------------------------------------
namespace {
struct EnvStepLayout {
  int ldx, ldo, n_tile;
  size_t x, zenv, ynext, llterm, planes, total;
};
EnvStepLayout env_step_layout(int O, int A, int B) {
  EnvStepLayout L;
  L.ldx = (O + A + 3) & ~3;
  L.ldo = (O + 3) & ~3;
  L.n_tile = 128;                       // three-slot operand ring (see gemm_ring_depth)
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o = al64(o + n); return r; };
  L.x = take(static_cast<size_t>(B) * L.ldx);
  L.zenv = take(static_cast<size_t>(B) * L.ldo);
  L.ynext = take(static_cast<size_t>(B) * L.ldo);
  L.llterm = take(static_cast<size_t>(B) * A);
  L.planes = take(planes_floats(O + A, O, L.n_tile));
  L.total = o;
  return L;
}
int check_env_step(const b200ppo_synth_env* env, int B, const void* ws, int64_t ws_bytes) {
  if (!env || !env->Wo || !env->Wa || env->obs_dim <= 0 || env->act_dim <= 0 || B <= 0 || !ws) return B200PPO_EINVAL;
  if (env->Wa != env->Wo + static_cast<size_t>(env->obs_dim) * env->obs_dim) return B200PPO_EINVAL;
  if (ws_bytes < static_cast<int64_t>(4 * env_step_layout(env->obs_dim, env->act_dim, B).total)) return B200PPO_EINVAL;
  return 0;
}
}  // namespace

extern "C" int64_t b200ppo_synth_env_step_workspace_bytes(int32_t O, int32_t A, int32_t B) {
  if (O <= 0 || A <= 0 || B <= 0) return -1;
  return static_cast<int64_t>(4 * env_step_layout(O, A, B).total);
}

extern "C" int b200ppo_synth_env_begin(void* stream, const b200ppo_synth_env* env, int32_t B, const float* env_obs,
                                       float* obs0, void* ws, int64_t ws_bytes) {
  int rc = check_env_step(env, B, ws, ws_bytes);
  if (rc) return rc;
  if (!env_obs || !obs0) return B200PPO_EINVAL;
  rc = set_attrs_tc();
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int O = env->obs_dim, A = env->act_dim;
  const EnvStepLayout L = env_step_layout(O, A, B);
  float* w = static_cast<float*>(ws);
  const int Nt = planes_ntile(O, L.n_tile);
  const int tiles = cdiv(O, Nt), nst = cdiv(O + A, RK);
  const size_t total = static_cast<size_t>(tiles) * nst * (RK / 4) * (Nt + 1);
  rg_prep_b_kernel<<<cdiv(static_cast<int64_t>(total), 256), 256, 0, s>>>(env->Wo, O, O + A, O, Nt, nst, tiles,
                                                                        reinterpret_cast<float4*>(w + L.planes));
  B200PPO_LAUNCH_CHECK();
  wide_init_kernel<<<cdiv(static_cast<int64_t>(B) * O, 256 * 8), 256, 0, s>>>(env_obs, B, O, w + L.x, L.ldx, obs0, nullptr, 0,
                                                                             nullptr, nullptr);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200ppo_synth_env_step(void* stream, const b200ppo_synth_env* env, const float* y, int32_t ldy,
                                      float min_std, float std_scale, const uint32_t* rng_state,
                                      const uint32_t* iter_keys, int32_t t, int32_t T, int32_t B, float* env_obs,
                                      int32_t* env_counter, uint32_t* env_term, float* obs, float* raw_action,
                                      float* action, float* loglik, float* reward, uint8_t* done, uint8_t* truncated,
                                      float* next_obs_last, void* ws, int64_t ws_bytes) {
  int rc = check_env_step(env, B, ws, ws_bytes);
  if (rc) return rc;
  if (!y || !rng_state || !iter_keys || !env_obs || !env_counter || !env_term || !obs || !raw_action || !action ||
      !loglik || !reward || !done || !truncated || !next_obs_last || t < 0 || t >= T || ldy < 2 * env->act_dim)
    return B200PPO_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int O = env->obs_dim, A = env->act_dim;
  const EnvStepLayout L = env_step_layout(O, A, B);
  float* w = static_cast<float*>(ws);
  float* X = w + L.x;
  const size_t row0 = static_cast<size_t>(t) * B;
  wide_sampler_kernel<<<cdiv(static_cast<int64_t>(B) * A, 256), 256, 0, s>>>(
      y, ldy, B, A, min_std, std_scale, rng_state, static_cast<uint32_t>(t), raw_action + row0 * A, action + row0 * A, X,
      L.ldx, O, w + L.llterm);
  B200PPO_LAUNCH_CHECK();
  GemmArgs g = gemm_defaults();
  g.A = X; g.lda = L.ldx; g.M = B; g.K = O + A;
  g.B = env->Wo; g.ldb = O;
  g.Bplanes = w + L.planes;
  g.C = nullptr; g.ldc = L.ldo;
  g.C2 = w + L.ynext; g.ldc2 = L.ldo; g.act_c2 = B200PPO_ACT_TANH;
  rc = launch_gemm(s, g, O, 1, L.n_tile);
  if (rc) return rc;
  WideBookArgs b;
  b.B = B; b.O = O; b.A = A; b.T = T; b.t = t; b.max_len = env->max_len; b.term_thresh16 = env->term_thresh16;
  b.ldx = L.ldx; b.ldo = L.ldo; b.iter_keys = iter_keys; b.ynext = w + L.ynext; b.llterm = w + L.llterm; b.X = X;
  b.Xn = nullptr; b.mean = nullptr; b.stdv = nullptr;
  b.obs_next = t + 1 < T ? obs + (row0 + B) * O : nullptr;
  b.loglik = loglik + row0; b.reward = reward + row0; b.done = done + row0; b.trunc = truncated + row0;
  b.next_obs_last = next_obs_last; b.env_obs = env_obs; b.env_counter = env_counter; b.env_term = env_term;
  wide_book_kernel<<<cdiv(B, 8), 256, 0, s>>>(b);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200ppo_lstm_set_persistent(int on) {
  const int prev = g_seq_persist < 0 ? -1 : g_seq_persist;
  if (on == 0 || on == 1) g_seq_persist = on;
  return prev;
}

extern "C" int b200ppo_lstm_seq_supported(const b200ppo_lstm_plan* plan) { return check_plan_tc(plan) == 0 ? 1 : 0; }

extern "C" int64_t b200ppo_lstm_seq_workspace_floats(const b200ppo_lstm_plan* plan, int32_t T, int32_t rows) {
  if (check_plan_tc(plan) || T <= 0 || rows <= 0) return -1;
  return static_cast<int64_t>(seq_layout(*plan, T, rows).total);
}

extern "C" int b200ppo_lstm_seq_num_launches(const b200ppo_lstm_plan* plan, int32_t T, int32_t rows, int32_t backward) {
  if (check_plan_tc(plan) || T <= 0 || rows <= 0) return -1;
  const size_t p_smem = 2 * static_cast<size_t>(plan->hidden / 4) * tc::plane_bytes(4 * UT) + PF_NS * 2u * (RK / 4) * tc::plane_bytes(RM);
  const int steps = (seq_persistent(*plan, T, rows) && p_smem <= PERSIST_SMEM_MAX) ? 1 : T;
  if (!backward) return 1 /* weight planes */ + 1 /* carry in */ + 2 /* pre, proj */ + steps + 1 /* post */ + 1 /* carry out */;
  return 1 /* dh_post */ + T /* element-wise */ + (T - 1) /* split-K dh_rec */ + 1 /* du */ + 3 /* TN */ + 1 /* reductions */ +
         (plan->init_c_off > 0 ? 2 : 0) /* initial-carry gradient */;
}

extern "C" int b200ppo_lstm_seq_forward(void* stream, const b200ppo_lstm_plan* plan, const float* params,
                                        const float* norm_mean, const float* norm_std, const float* x,
                                        const uint8_t* done, const int32_t* inds, int32_t B, float* c, float* h,
                                        int32_t T, int32_t rows, float* ws, float* y, int32_t keep_cache) {
  int rc = check_plan_tc(plan);
  if (rc) return rc;
  if (T <= 0 || rows <= 0 || B <= 0 || !params || !x || !c || !h || !ws || !y) return B200PPO_EINVAL;
  if ((norm_mean == nullptr) != (norm_std == nullptr)) return B200PPO_EINVAL;
  if (reinterpret_cast<uintptr_t>(ws) & 255) return B200PPO_EALIGN;
  rc = set_attrs_tc();
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const SeqLayout L = seq_layout(*plan, T, rows);
  const int O = plan->obs_dim, P = plan->pre_dim, H = plan->hidden, Y = plan->out_dim;
  const int R = T * rows, LC = P + H, planes = H / 4;
  float* z1 = ws + L.z1;
  float* cat = ws + L.cat;
  const float* Wh = params + plan->wcat_off + static_cast<size_t>(P) * 4 * H;
  // Wh -> operand planes (forward and transposed forms); keep_cache bit 1: the planes in ws are current (the
  // rollout calls this T times between two parameter updates)
  if (!(keep_cache & 2)) {
    PrepWhArgs a;
    a.Wh = Wh; a.H = H; a.NBT = L.NBT; a.whf = ws + L.whf; a.whb = ws + L.whb;
    const long long n = static_cast<long long>(H / UT) * planes * 64 + static_cast<long long>(L.nbt) * 4 * planes * L.NBT;
    lstm_prep_wh_kernel<<<cdiv(n, 256), 256, 0, s>>>(a);
    B200PPO_LAUNCH_CHECK();
  }
  keep_cache &= 1;
  // carry in: h -> A planes of step 0 (and row-major cat[0][:, P:] for the weight gradients), c -> state planes
  {
    CarryInArgs a;
    a.c = c; a.h = h; a.rows = rows; a.H = H;
    a.hp_hi = ws + L.hp; a.hp_lo = ws + L.hp + L.hp_half; a.hp_tile = static_cast<long long>(planes) * PLA;
    a.cp = ws + L.cp; a.cp_tile = static_cast<long long>(L.st_tile);
    a.cat_h = cat + P; a.ld_cat = LC;
    lstm_carry_in_kernel<<<cdiv(static_cast<int64_t>(L.tiles) * planes * RM, 256), 256, 0, s>>>(a);
    B200PPO_LAUNCH_CHECK();
  }
  // z1 = x W1 + b1;  u = act(z1) -> cat[:, :P]
  {
    GemmArgs g = gemm_defaults();
    g.A = x; g.lda = O; g.M = R; g.K = O; g.a_mean = norm_mean; g.a_std = norm_std;
    g.B = params + plan->w1_off; g.ldb = P;
    g.C = z1; g.ldc = P; g.bias = params + plan->b1_off;
    g.C2 = cat; g.ldc2 = LC; g.act_c2 = plan->act;
    rc = launch_gemm(s, g, P, 1);
    if (rc) return rc;
  }
  // gx = u Wi + bl, stored as planes over the R row space
  {
    GemmArgs g = gemm_defaults();
    g.A = cat; g.lda = LC; g.M = R; g.K = P;
    g.B = params + plan->wcat_off; g.ldb = 4 * H;
    g.C = ws + L.gxp; g.c_planes = 1; g.c_cols4 = H; g.bias = params + plan->bl_off;
    rc = launch_gemm(s, g, 4 * H, 1);
    if (rc) return rc;
  }
  // the recurrence: one persistent launch when every CTA fits on the device at once, else one launch per step
  auto step_args = [&](int t) {
    StepFwd2Args a;
    const size_t cur = static_cast<size_t>(t & 1) * L.hp_buf, nxt = static_cast<size_t>((t + 1) & 1) * L.hp_buf;
    a.hp_hi = ws + L.hp + cur; a.hp_lo = ws + L.hp + cur + L.hp_half; a.hp_tile = static_cast<long long>(planes) * PLA;
    a.whf = ws + L.whf; a.whf_tile = 2LL * planes * plb(4 * UT);
    a.gxp = ws + L.gxp; a.r0 = static_cast<long long>(t) * rows; a.gx_cols4 = H;
    a.cp = ws + L.cp; a.cp_tile = static_cast<long long>(L.st_tile);
    a.hn_hi = ws + L.hp + nxt; a.hn_lo = ws + L.hp + nxt + L.hp_half;
    a.h_next_rm = cat + static_cast<size_t>(t + 1) * rows * LC + P; a.ld_hn = LC;
    a.hn_rm = ws + L.hn + static_cast<size_t>(t) * rows * H;
    if (keep_cache) {
      const size_t so = static_cast<size_t>(t) * L.cache_step;
      a.gi = ws + L.cache[0] + so; a.gf = ws + L.cache[1] + so; a.gg = ws + L.cache[2] + so;
      a.go = ws + L.cache[3] + so; a.tcc = ws + L.cache[4] + so; a.cin = ws + L.cache[5] + so;
    } else {
      a.gi = a.gf = a.gg = a.go = a.tcc = a.cin = nullptr;
    }
    a.done = done != nullptr ? done + static_cast<size_t>(t) * B : nullptr;
    a.inds = inds;
    a.init_c = plan->init_c_off > 0 ? params + plan->init_c_off : nullptr;
    a.init_h = plan->init_h_off > 0 ? params + plan->init_h_off : nullptr;
    a.rows = rows; a.H = H;
    return a;
  };
  const size_t p_smem = 2 * static_cast<size_t>(planes) * tc::plane_bytes(4 * UT) + PF_NS * 2u * (RK / 4) * tc::plane_bytes(RM);
  if (seq_persistent(*plan, T, rows) && p_smem <= PERSIST_SMEM_MAX) {
    if (cudaMemsetAsync(ws + L.flags, 0, sizeof(int) * L.tiles * (static_cast<size_t>(T) + 1), s) != cudaSuccess)
      return static_cast<int>(cudaGetLastError());
    SeqFwdPArgs pa;
    pa.s = step_args(0);
    pa.T = T;
    pa.hp_buf = static_cast<long long>(L.hp_buf);
    pa.cat_step = static_cast<long long>(rows) * LC; pa.hn_step = static_cast<long long>(rows) * H;
    pa.cache_step = static_cast<long long>(L.cache_step); pa.done_step = B;
    pa.hp_base_hi = ws + L.hp; pa.hp_base_lo = ws + L.hp + L.hp_half;
    pa.flags = reinterpret_cast<int*>(ws + L.flags);
    pa.n_ut = H / UT;
    lstm_seq_fwd_persistent_kernel<<<dim3(L.tiles, H / UT), RT, p_smem, s>>>(pa);
    B200PPO_LAUNCH_CHECK();
  } else {
    // few CTAs (small minibatch): a cluster pair per tile, each CTA half of K (see lstm_step_fwd3_kernel)
    const bool pair = seq_pair_mode() && H % 32 == 0 && 2 * L.tiles * (H / UT) <= b200ppo_num_sms();
    for (int t = 0; t < T; ++t) {
      const StepFwd2Args a = step_args(t);
      if (pair)
        lstm_step_fwd3_kernel<<<dim3(L.tiles, H / UT, 2), RT, FWD3_XB_BYTES + SP_NS * sp_slot_bytes(4 * UT), s>>>(a);
      else
        lstm_step_fwd2_kernel<<<dim3(L.tiles, H / UT), RT, SP_NS * sp_slot_bytes(4 * UT), s>>>(a);
      B200PPO_LAUNCH_CHECK();
    }
  }
  // y = h' W2 + b2
  {
    GemmArgs g = gemm_defaults();
    g.A = ws + L.hn; g.lda = H; g.M = R; g.K = H;
    g.B = params + plan->w2_off; g.ldb = Y;
    g.C = y; g.ldc = Y; g.bias = params + plan->b2_off;
    rc = launch_gemm(s, g, Y, 1);
    if (rc) return rc;
  }
  // carry out (reset already applied by the last step)
  lstm_carry_out_kernel<<<cdiv(static_cast<int64_t>(L.tiles) * planes * RM, 256), 256, 0, s>>>(
      ws + L.cp, static_cast<long long>(L.st_tile), c, cat + static_cast<size_t>(T) * rows * LC + P, LC, h, rows, H);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200ppo_lstm_seq_backward(void* stream, const b200ppo_lstm_plan* plan, const float* params,
                                         const float* x, const float* d_y, const uint8_t* done, const int32_t* inds,
                                         int32_t B, int32_t T, int32_t rows, float* ws, float* grad) {
  int rc = check_plan_tc(plan);
  if (rc) return rc;
  if (T <= 0 || rows <= 0 || B <= 0 || !params || !x || !d_y || !ws || !grad) return B200PPO_EINVAL;
  if (reinterpret_cast<uintptr_t>(ws) & 255) return B200PPO_EALIGN;
  rc = set_attrs_tc();
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const SeqLayout L = seq_layout(*plan, T, rows);
  const int O = plan->obs_dim, P = plan->pre_dim, H = plan->hidden, Y = plan->out_dim;
  const int R = T * rows, LC = P + H, planes = H / 4;
  float* cat = ws + L.cat;
  float* da = ws + L.da;
  // dh_post = dY W2^T, as planes over the R row space
  {
    GemmArgs g = gemm_defaults();
    g.A = d_y; g.lda = Y; g.M = R; g.K = Y;
    g.B = params + plan->w2_off; g.ldb = Y; g.b_nt = 1;
    g.C = ws + L.dhpp; g.c_planes = 1; g.c_cols4 = planes;
    rc = launch_gemm(s, g, H, 1);
    if (rc) return rc;
  }
  const long long dap_slice = static_cast<long long>(planes) * PLA, dap_tile = 4 * dap_slice;
  const long long dhr_slice = static_cast<long long>(L.tiles) * L.st_tile;
  for (int t = T - 1; t >= 0; --t) {
    const size_t so = static_cast<size_t>(t) * L.cache_step;
    StepBwd2Args e;
    e.dhp = ws + L.dhpp; e.r0 = static_cast<long long>(t) * rows; e.dhp_cols4 = planes;
    e.dhr = t == T - 1 ? nullptr : ws + L.dhr; e.dhr_slice = dhr_slice; e.n_slices = 4 * L.KS;
    e.dcp = ws + L.dcp;
    e.gi = ws + L.cache[0] + so; e.gf = ws + L.cache[1] + so; e.gg = ws + L.cache[2] + so;
    e.go = ws + L.cache[3] + so; e.tcc = ws + L.cache[4] + so; e.cin = ws + L.cache[5] + so;
    e.st_tile = static_cast<long long>(L.st_tile);
    e.done = done != nullptr ? done + static_cast<size_t>(t) * B : nullptr;
    e.inds = inds;
    e.da_rm = da + static_cast<size_t>(t) * rows * 4 * H;
    e.dap_hi = ws + L.dap; e.dap_lo = ws + L.dap + L.dap_half; e.dap_tile = dap_tile; e.dap_slice = dap_slice;
    e.mdc = plan->init_c_off > 0 ? ws + L.mdc + so : nullptr;
    e.mdh = plan->init_c_off > 0 ? ws + L.mdh + so : nullptr;
    e.dc_zero = t == T - 1 ? 1 : 0;
    e.rows = rows; e.H = H;
    lstm_step_bwd2_kernel<<<cdiv(static_cast<int64_t>(L.tiles) * planes * RM, 256), 256, 0, s>>>(e);
    B200PPO_LAUNCH_CHECK();
    if (t > 0) {
      // dh_rec[q] = da_t[:, gate q] Wh[:, gate q]^T   (split-K over the four gate blocks)
      StepBwdGemmArgs g;
      g.dap_hi = e.dap_hi; g.dap_lo = e.dap_lo; g.dap_tile = dap_tile; g.dap_slice = dap_slice;
      g.whb = ws + L.whb; g.whb_slice = 2LL * planes * plb(L.NBT); g.whb_tile = 4 * g.whb_slice;
      g.dhr = ws + L.dhr; g.dhr_slice = dhr_slice; g.st_tile = static_cast<long long>(L.st_tile);
      g.rows = rows; g.H = H; g.NBT = L.NBT; g.KS = L.KS;
      lstm_step_bwd_gemm_kernel<<<dim3(L.tiles, L.nbt, 4 * L.KS), RT, SP_NS * sp_slot_bytes(L.NBT), s>>>(g);
      B200PPO_LAUNCH_CHECK();
    }
  }
  if (plan->init_c_off > 0) {     // gradient of the learned initial carry (recurrent.py:85-87 Params)
    InitGradArgs ig;
    ig.mdc = ws + L.mdc; ig.mdh = ws + L.mdh; ig.step_stride = static_cast<long long>(L.cache_step);
    ig.st_tile = static_cast<long long>(L.st_tile); ig.tiles = L.tiles; ig.rows = rows; ig.H = H; ig.T = T;
    ig.part = ws + L.ipart; ig.g_c = grad + plan->init_c_off; ig.g_h = grad + plan->init_h_off;
    lstm_init_grad_part_kernel<<<dim3(planes, T), RM, 0, s>>>(ig);
    B200PPO_LAUNCH_CHECK();
    lstm_init_grad_sum_kernel<<<cdiv(2 * H, 256), 256, 0, s>>>(ig);
    B200PPO_LAUNCH_CHECK();
  }
  // du = da Wi^T;  dz1 = du * act'(z1)
  {
    GemmArgs g = gemm_defaults();
    g.A = da; g.lda = 4 * H; g.M = R; g.K = 4 * H;
    g.B = params + plan->wcat_off; g.ldb = 4 * H; g.b_nt = 1;
    g.C = ws + L.dz1; g.ldc = P;
    g.Zmul = ws + L.z1; g.ldz = P; g.act_z = plan->act;
    rc = launch_gemm(s, g, P, 1);
    if (rc) return rc;
  }
  // weight gradients (x: the NORMALISED observations the forward pass saw); one reduction launch for all of them
  ReduceSegs segs;
  segs.n_seg = 0; segs.blk0[0] = 0;
  rc = launch_tn(s, cat, LC, 0, LC, B200PPO_ACT_NONE, da, 4 * H, 4 * H, R, L.S_cat, ws + L.part, ws + L.bpart,
                 grad + plan->wcat_off, grad + plan->bl_off, &segs);
  if (rc) return rc;
  rc = launch_tn(s, x, O, 0, O, B200PPO_ACT_NONE, ws + L.dz1, P, P, R, L.S_w1, ws + L.part1, ws + L.bpart1,
                 grad + plan->w1_off, grad + plan->b1_off, &segs);
  if (rc) return rc;
  rc = launch_tn(s, ws + L.hn, H, 0, H, B200PPO_ACT_NONE, d_y, Y, Y, R, L.S_w2, ws + L.part2, ws + L.bpart2,
                 grad + plan->w2_off, grad + plan->b2_off, &segs);
  if (rc) return rc;
  rg_reduce_multi_kernel<<<segs.blk0[segs.n_seg], 256, 0, s>>>(segs);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

// test hook: C[M][N] = A[M][K] B (b_nt = 0: B is [K][N]; 1: B is [N][K]) through the generic tile kernel
extern "C" int b200ppo_rg_gemm_test(void* stream, const float* A, const float* B, float* C, int32_t M, int32_t N,
                                    int32_t K, int32_t b_nt, int32_t k_slices) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || k_slices < 1 || K % k_slices) return B200PPO_EINVAL;
  int rc = set_attrs_tc();
  if (rc) return rc;
  GemmArgs g = gemm_defaults();
  g.A = A; g.lda = K; g.M = M; g.K = K / k_slices;
  g.B = B; g.ldb = b_nt ? K : N; g.b_nt = b_nt;
  g.C = C; g.ldc = N;
  g.k_slice = K / k_slices; g.c_slice_stride = static_cast<long long>(M) * N;
  return launch_gemm(static_cast<cudaStream_t>(stream), g, N, k_slices);
}

// test hook: W[Kdim][N] = A[rows][Kdim]^T D[rows][N], b[N] = column sums of D
extern "C" int b200ppo_rg_tn_test(void* stream, const float* A, const float* D, int32_t rows, int32_t Kdim, int32_t N,
                                  int32_t S, float* scratch, float* W, float* b) {
  if (!A || !D || !scratch || !W || rows <= 0 || Kdim <= 0 || N <= 0 || S < 1 || S > 64) return B200PPO_EINVAL;
  int rc = set_attrs_tc();
  if (rc) return rc;
  float* part = scratch;
  float* bpart = scratch + static_cast<size_t>(S) * Kdim * N;
  return launch_tn(static_cast<cudaStream_t>(stream), A, Kdim, 0, Kdim, B200PPO_ACT_NONE, D, N, N, rows, S, part, bpart, W, b);
}
