// Standalone tcgen05 GEMM (C = A B, fp32 in / fp32 out, 1xTF32 or error-compensated 3xTF32):
// the bring-up and parity harness of the tensor-core building blocks in tc.cuh.
#include "common.cuh"
#include "tc.cuh"

using namespace b200ppo;

namespace {

constexpr int TCM = 128;     // rows per CTA (UMMA M)
constexpr int TCK = 32;      // k per stage (4 MMAs of K = 8)
constexpr int TCT = 256;     // threads

// stage layout (bytes): A_hi | A_lo | B_hi | B_lo, each KC/4 planes
__host__ __device__ constexpr uint32_t stage_bytes(int n) {
  return 2u * (TCK / 4) * tc::plane_bytes(TCM) + 2u * (TCK / 4) * tc::plane_bytes(n);
}

__global__ void __launch_bounds__(TCT, 1)
tc_gemm_test_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                    int M, int N, int K, int split, int tmem_cols) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_empty[2];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * TCM;
  const uint32_t pa = tc::plane_bytes(TCM), pb = tc::plane_bytes(N);
  const uint32_t sbytes = stage_bytes(N);

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, tmem_cols);
  if (tid == 32) {
    tc::mbar_init(&bar_empty[0], 1);
    tc::mbar_init(&bar_empty[1], 1);
    tc::mbar_init(&bar_done, 1);
    tc::mbar_init_fence();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = tc::make_idesc_tf32(TCM, N);

  const int nst = (K + TCK - 1) / TCK;
  uint32_t phase[2] = {0u, 0u};
  for (int s = 0; s < nst; ++s) {
    const int buf = s & 1;
    uint8_t* st = smem + buf * sbytes;
    uint8_t* a_hi = st;
    uint8_t* a_lo = a_hi + (TCK / 4) * pa;
    uint8_t* b_hi = a_lo + (TCK / 4) * pa;
    uint8_t* b_lo = b_hi + (TCK / 4) * pb;
    if (s >= 2) {                       // the MMAs that read this buffer two stages ago are done
      tc::mbar_wait(&bar_empty[buf], phase[buf]);
      phase[buf] ^= 1u;
    }
    const int k0 = s * TCK;
    // ---- stage A: rows x 8 planes; lane -> (plane q = idx & 7, row = idx >> 3): 128 B per row ----
    for (int idx = tid; idx < TCM * (TCK / 4); idx += TCT) {
      const int q = idx & 7, r = idx >> 3;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      const int k = k0 + 4 * q;
      if (row0 + r < M) {
        const float* src = A + static_cast<size_t>(row0 + r) * K + k;
        if (k + 3 < K && (K & 3) == 0) x = *reinterpret_cast<const float4*>(src);
        else {
          if (k + 0 < K) x.x = src[0];
          if (k + 1 < K) x.y = src[1];
          if (k + 2 < K) x.z = src[2];
          if (k + 3 < K) x.w = src[3];
        }
      }
      float4 hi, lo;
      tc::split4(x, hi, lo);
      *reinterpret_cast<float4*>(a_hi + q * pa + r * 16) = hi;
      *reinterpret_cast<float4*>(a_lo + q * pa + r * 16) = lo;
    }
    // ---- stage B: B_op(n, k) = B[k][n]; lane -> consecutive n (coalesced), 4 k's per thread ----
    for (int idx = tid; idx < N * (TCK / 4); idx += TCT) {
      const int n = idx % N, q = idx / N;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      const int k = k0 + 4 * q;
      if (k + 0 < K) x.x = B[static_cast<size_t>(k + 0) * N + n];
      if (k + 1 < K) x.y = B[static_cast<size_t>(k + 1) * N + n];
      if (k + 2 < K) x.z = B[static_cast<size_t>(k + 2) * N + n];
      if (k + 3 < K) x.w = B[static_cast<size_t>(k + 3) * N + n];
      float4 hi, lo;
      tc::split4(x, hi, lo);
      *reinterpret_cast<float4*>(b_hi + q * pb + n * 16) = hi;
      *reinterpret_cast<float4*>(b_lo + q * pb + n * 16) = lo;
    }
    tc::fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
#pragma unroll
      for (int j = 0; j < TCK / 8; ++j) {
        const uint64_t ah = tc::make_desc(tc::smem_u32(a_hi + 2 * j * pa), pa, 128);
        const uint64_t al = tc::make_desc(tc::smem_u32(a_lo + 2 * j * pa), pa, 128);
        const uint64_t bh = tc::make_desc(tc::smem_u32(b_hi + 2 * j * pb), pb, 128);
        const uint64_t bl = tc::make_desc(tc::smem_u32(b_lo + 2 * j * pb), pb, 128);
        const uint32_t acc0 = (s > 0 || j > 0) ? 1u : 0u;
        if (split) {
          // small terms first, then the dominant product
          tc::mma_tf32(tmem_base, al, bh, idesc, acc0);
          tc::mma_tf32(tmem_base, ah, bl, idesc, 1u);
          tc::mma_tf32(tmem_base, ah, bh, idesc, 1u);
        } else {
          tc::mma_tf32(tmem_base, ah, bh, idesc, acc0);
        }
      }
      tc::commit(&bar_empty[buf]);
      if (s == nst - 1) tc::commit(&bar_done);
    }
  }
  // ---- epilogue: TMEM -> registers -> global ----
  tc::mbar_wait(&bar_done, 0u);
  tc::tc_fence_after();
  const int sub = warp & 3;                 // TMEM sub-partition this warp may read
  const int row = row0 + sub * 32 + lane;
  for (int c = (warp >> 2) * 16; c < N; c += 32) {
    float v[16];
    tc::tmem_ld16(tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + static_cast<uint32_t>(c), v);
    if (row < M) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c + i < N) C[static_cast<size_t>(row) * N + c + i] = v[i];
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// Microbenchmark: (a) `iters` dependent tcgen05.mma (M=128, N, K=8) on resident smem operands,
// (b) `iters` bulk copies of `bytes` into smem, each waited for.  out[0] = cycles of (a), out[1] = (b),
// out[2] = (a) with `nacc` independent accumulators round-robin.
__global__ void __launch_bounds__(128, 1)
tc_microbench_kernel(const float* __restrict__ src, long long* __restrict__ out, int N, int iters, int bytes,
                     int nacc) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 32) { tc::mbar_init(&bar, 1); tc::mbar_init(&bar2, 1); tc::mbar_init_fence(); }
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.0f;
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    const uint32_t pa = tc::plane_bytes(128), pb = tc::plane_bytes(N);
    const uint32_t idesc = tc::make_idesc_tf32(128, N);
    const uint64_t ad = tc::make_desc(tc::smem_u32(smem), pa, 128);
    const uint64_t bd = tc::make_desc(tc::smem_u32(smem + 8192), pb, 128);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) tc::mma_tf32(tmem_base, ad, bd, idesc, 1u);
    tc::commit(&bar);
    tc::mbar_wait(&bar, 0u);
    long long t1 = clock64();
    out[0] = t1 - t0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) tc::mma_tf32(tmem_base + static_cast<uint32_t>((i % nacc) * N), ad, bd, idesc, 1u);
    tc::commit(&bar);
    tc::mbar_wait(&bar, 1u);
    t1 = clock64();
    out[2] = t1 - t0;
    t0 = clock64();
    uint32_t par = 0;
    for (int i = 0; i < iters; ++i) {
      tc::mbar_arrive_expect_tx(&bar2, bytes);
      tc::bulk_g2s(smem + 65536, src, bytes, &bar2);
      tc::mbar_wait(&bar2, par);
      par ^= 1u;
    }
    t1 = clock64();
    out[1] = t1 - t0;
    // 4 copies in flight
    t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      tc::mbar_arrive_expect_tx(&bar2, 4 * (bytes / 4));
      for (int j = 0; j < 4; ++j) tc::bulk_g2s(smem + 65536 + j * (bytes / 4), src + j * (bytes / 16), bytes / 4, &bar2);
      tc::mbar_wait(&bar2, par);
      par ^= 1u;
    }
    t1 = clock64();
    out[3] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace

extern "C" int b200ppo_tc_microbench(void* stream, const float* src, long long* out, int32_t N, int32_t iters,
                                     int32_t bytes, int32_t nacc, int32_t blocks) {
  if (!src || !out || N < 16 || N > 256 || (N & 15) || bytes <= 0 || (bytes & 63) || bytes > 131072 || nacc < 1 ||
      nacc * N > 512 || blocks < 1)
    return B200PPO_EINVAL;
  const size_t smem = 65536 + 131072;
  cudaError_t e = cudaFuncSetAttribute(tc_microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  tc_microbench_kernel<<<blocks, 128, smem, static_cast<cudaStream_t>(stream)>>>(src, out, N, iters, bytes, nacc);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

// Test hook: C[M][N] = A[M][K] * B[K][N] on the tensor cores.  N multiple of 16, 16 <= N <= 256.
extern "C" int b200ppo_tc_gemm_test(void* stream, const float* A, const float* B, float* C, int32_t M,
                                    int32_t N, int32_t K, int32_t split) {
  if (!A || !B || !C || M <= 0 || K <= 0) return B200PPO_EINVAL;
  if (N < 16 || N > 256 || (N & 15)) return B200PPO_EINVAL;
  int cols = 32;
  while (cols < N) cols <<= 1;
  const size_t smem = 2 * static_cast<size_t>(stage_bytes(N));
  cudaError_t e = cudaFuncSetAttribute(tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  tc_gemm_test_kernel<<<cdiv(M, TCM), TCT, smem, static_cast<cudaStream_t>(stream)>>>(A, B, C, M, N, K, split, cols);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Bring-up / parity harness of the MN-major operand path (tc.cuh: make_desc_mn_sw128): C[K][N] = H^T D for
// row-major H [rows][K <= 128], D [rows][N <= 256], both fetched by swizzled TMA boxes and consumed by the MMA
// without a software transposition; the hi half of the 3xTF32 split is the raw fp32 tile (the tensor core reads
// the upper 19 bits), the lo half x - trunc(x) is computed element-wise at the same offsets.
// ------------------------------------------------------------------------------------------
#include "tmap.cuh"

namespace {

__global__ void __launch_bounds__(128, 1)
tc_mn_test_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmD,
                  float* __restrict__ C, int rows, int K, int N, int split, int tmem_cols, float* __restrict__ dbg) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int n32 = (N + 31) / 32, npad = n32 * 32;
  const uint32_t a_bytes = 4u * tc::MN_ATOM_BYTES, b_bytes = static_cast<uint32_t>(n32) * tc::MN_ATOM_BYTES;
  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + a_bytes;
  uint8_t* b_hi = a_lo + a_bytes;
  uint8_t* b_lo = b_hi + b_bytes;
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, tmem_cols);
  if (tid == 32) { tc::mbar_init(&bar_full, 1); tc::mbar_init(&bar_done, 1); tc::mbar_init_fence(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int nst = (rows + 15) / 16;
  const uint32_t idesc = tc::make_idesc_tf32_mn(128, npad);
  for (int s = 0; s < nst; ++s) {
    if (tid == 0) {
      tc::mbar_arrive_expect_tx(&bar_full, a_bytes + b_bytes);
      for (int i = 0; i < 4; ++i) tc::tma_load_2d(a_hi + i * tc::MN_ATOM_BYTES, &tmH, 32 * i, 16 * s, &bar_full);
      for (int j = 0; j < n32; ++j) tc::tma_load_2d(b_hi + j * tc::MN_ATOM_BYTES, &tmD, 32 * j, 16 * s, &bar_full);
    }
    tc::mbar_wait(&bar_full, static_cast<uint32_t>(s) & 1u);
    if (dbg != nullptr && s == nst - 1) {                  // the raw operand tiles of the last stage, as they landed
      for (uint32_t i = tid; i < a_bytes / 4; i += 128) dbg[i] = reinterpret_cast<const float*>(a_hi)[i];
      for (uint32_t i = tid; i < b_bytes / 4; i += 128) dbg[a_bytes / 4 + i] = reinterpret_cast<const float*>(b_hi)[i];
    }
    if (split) {
      const int nchunk = static_cast<int>((a_bytes + b_bytes) / 16);     // A and B hi tiles are adjacent only when
      for (int c = tid; c < nchunk; c += 128) {                          // a_lo is skipped: index them separately
        const bool isa = c < static_cast<int>(a_bytes / 16);
        const uint32_t off = isa ? c * 16u : (c * 16u - a_bytes);
        const uint4 x = *reinterpret_cast<const uint4*>((isa ? a_hi : b_hi) + off);
        float4 lo;
        lo.x = __uint_as_float(x.x) - __uint_as_float(x.x & 0xFFFFE000u);
        lo.y = __uint_as_float(x.y) - __uint_as_float(x.y & 0xFFFFE000u);
        lo.z = __uint_as_float(x.z) - __uint_as_float(x.z & 0xFFFFE000u);
        lo.w = __uint_as_float(x.w) - __uint_as_float(x.w & 0xFFFFE000u);
        *reinterpret_cast<float4*>((isa ? a_lo : b_lo) + off) = lo;
      }
      tc::fence_proxy_async();
    }
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      for (int j = 0; j < 2; ++j) {
        const uint64_t ah = tc::make_desc_mn_sw128(tc::smem_u32(a_hi) + j * 1024u, tc::MN_ATOM_BYTES, tc::MN_SBO_BYTES);
        const uint64_t bh = tc::make_desc_mn_sw128(tc::smem_u32(b_hi) + j * 1024u, tc::MN_ATOM_BYTES, tc::MN_SBO_BYTES);
        const uint32_t acc0 = (s > 0 || j > 0) ? 1u : 0u;
        if (split) {
          const uint64_t al = tc::make_desc_mn_sw128(tc::smem_u32(a_lo) + j * 1024u, tc::MN_ATOM_BYTES, tc::MN_SBO_BYTES);
          const uint64_t bl = tc::make_desc_mn_sw128(tc::smem_u32(b_lo) + j * 1024u, tc::MN_ATOM_BYTES, tc::MN_SBO_BYTES);
          tc::mma_tf32(tmem_base, al, bh, idesc, acc0);
          tc::mma_tf32(tmem_base, ah, bl, idesc, 1u);
          tc::mma_tf32(tmem_base, ah, bh, idesc, 1u);
        } else {
          tc::mma_tf32(tmem_base, ah, bh, idesc, acc0);
        }
      }
      tc::commit(&bar_done);
      tc::mbar_wait(&bar_done, static_cast<uint32_t>(s) & 1u);
    }
    __syncthreads();
  }
  tc::tc_fence_after();
  const int m = warp * 32 + lane;
  for (int c = 0; c < npad; c += 16) {
    float v[16];
    tc::tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(c), v);
    if (m < K)
      for (int i = 0; i < 16; ++i)
        if (c + i < N) C[static_cast<size_t>(m) * N + c + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace

// Test hook: C[K][N] = H[rows][K]^T D[rows][N] through the MN-major swizzled operand path.  K <= 128, N <= 256,
// K % 4 == 0, N % 4 == 0 (TMA row pitch); synchronous tensor-map encoding, asynchronous launch.
extern "C" int b200ppo_tc_mn_test(void* stream, const float* H, const float* D, float* C, int32_t rows, int32_t K,
                                  int32_t N, int32_t split, float* dbg) {
  if (!H || !D || !C || rows <= 0 || K <= 0 || K > 128 || (K & 3) || N <= 0 || N > 256 || (N & 3)) return B200PPO_EINVAL;
  CUtensorMap tmH, tmD;
  int rc = encode_tiled_2d(&tmH, H, K, rows, 32, 16, true);
  if (rc) return rc;
  rc = encode_tiled_2d(&tmD, D, N, rows, 32, 16, true);
  if (rc) return rc;
  const int n32 = (N + 31) / 32;
  int cols = 32;
  while (cols < n32 * 32) cols <<= 1;
  const size_t smem = 1024 + 2 * (4 + static_cast<size_t>(n32)) * tc::MN_ATOM_BYTES;
  cudaError_t e = cudaFuncSetAttribute(tc_mn_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  tc_mn_test_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(tmH, tmD, C, rows, K, N, split, cols, dbg);
  B200PPO_LAUNCH_CHECK();
  return 0;
}
