// K7 (threefry test hooks), K6 (permutation indices), K2 (standalone GAE), K5 (Normalizer
// statistics), iteration bookkeeping and the FFMA peak probe.  sm_100a.
#include "common.cuh"

using namespace b200ppo;

#include <cstdlib>

static int g_num_sms = 0;
static int g_pdl = -1;

bool b200ppo::pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = std::getenv("B200PPO_PDL");
    // Off by default: measured on B200 inside the captured iteration graph (profiles/r2_notes.md) the
    // programmatic edges made the iteration SLOWER (6.23 ms vs 5.64 ms per configs[1] iteration).
    g_pdl = (e && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : 0;
  }
  return g_pdl != 0;
}

int b200ppo::pdl_mode() {
  pdl_enabled();
  return g_pdl;
}

extern "C" int b200ppo_set_pdl(int on) {
  const int prev = b200ppo::pdl_enabled() ? 1 : 0;
  if (on >= 0 && on <= 3) g_pdl = on;
  return prev;
}

extern "C" int b200ppo_version(void) { return 100; }

extern "C" int b200ppo_num_sms(void) {
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
    g_num_sms = n > 0 ? n : 148;
  }
  return g_num_sms;
}

extern "C" const char* b200ppo_error_string(int code) {
  if (code == 0) return "ok";
  if (code == B200PPO_EINVAL) return "b200ppo: invalid argument (shape / null pointer / unsupported plan)";
  if (code == B200PPO_ELIMIT) return "b200ppo: size exceeds a compiled-in limit";
  if (code == B200PPO_EALIGN) return "b200ppo: misaligned pointer or offset";
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "b200ppo: unknown error";
}

// ------------------------------------------------------------------------------------------
// K7 test hooks
// ------------------------------------------------------------------------------------------
__global__ void random_bits_kernel(Key k, int64_t n, uint32_t* __restrict__ out) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = random_bits_at(k, static_cast<uint32_t>(i));
}
__global__ void random_normal_kernel(Key k, int64_t n, float* __restrict__ out) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = bits_to_normal(random_bits_at(k, static_cast<uint32_t>(i)));
}

extern "C" int b200ppo_random_bits(void* stream, uint32_t k0, uint32_t k1, int64_t n, uint32_t* out) {
  if (n < 0 || n > 0xFFFFFFFFll || (n > 0 && !out)) return B200PPO_EINVAL;
  if (n == 0) return 0;
  random_bits_kernel<<<cdiv(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(Key{k0, k1}, n, out);
  B200PPO_LAUNCH_CHECK();
  return 0;
}
extern "C" int b200ppo_random_normal(void* stream, uint32_t k0, uint32_t k1, int64_t n, float* out) {
  if (n < 0 || n > 0xFFFFFFFFll || (n > 0 && !out)) return B200PPO_EINVAL;
  if (n == 0) return 0;
  random_normal_kernel<<<cdiv(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(Key{k0, k1}, n, out);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// K6: jax.random.permutation(fold_in(new_key, e), n) for each epoch e — ppo.py:287-294.
// One CTA per epoch.  Each round: key, sub = split(key); sort_keys = bits(sub, n); stable
// sort_key_val.  Stability = sort the 64-bit composite (sort_key << 32 | position) with a
// bitonic network (in shared memory when it fits, else in the global scratch).
// ------------------------------------------------------------------------------------------
constexpr int PERM_THREADS = 1024;
constexpr int PERM_SMEM_MAX_ELEMS = 16384;  // 128 KiB of composites

static int next_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}
static int perm_rounds(int n) {
  // ceil(3 * ln(n) / ln(2^32 - 1)), as jax/_src/random.py _shuffle computes it in float64.
  if (n <= 1) return 0;
  double r = 3.0 * log(static_cast<double>(n)) / log(4294967295.0);
  return static_cast<int>(ceil(r));
}

__global__ void __launch_bounds__(PERM_THREADS)
perm_kernel(const uint32_t* __restrict__ new_key, int n, int n2, int rounds, int use_smem,
            int32_t* __restrict__ out, unsigned long long* __restrict__ gcomp,
            int32_t* __restrict__ xtmp) {
  extern __shared__ unsigned long long scomp[];
  const int e = blockIdx.x;
  unsigned long long* comp = use_smem ? scomp : gcomp + static_cast<size_t>(e) * n2;
  int32_t* x = out + static_cast<size_t>(e) * n;
  int32_t* xt = xtmp + static_cast<size_t>(e) * n;
  Key key = fold_in(Key{new_key[0], new_key[1]}, static_cast<uint32_t>(e));
  for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = i;
  __syncthreads();
  for (int r = 0; r < rounds; ++r) {
    const Key sub = split_at(key, 1u);
    key = split_at(key, 0u);
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      comp[i] = i < n ? ((static_cast<unsigned long long>(random_bits_at(sub, i)) << 32) |
                         static_cast<unsigned long long>(i))
                      : ~0ull;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < n2; i += blockDim.x) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const unsigned long long a = comp[i], b = comp[ixj];
            const bool up = (i & k) == 0;
            if ((a > b) == up) {
              comp[i] = b;
              comp[ixj] = a;
            }
          }
        }
        __syncthreads();
      }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      xt[i] = x[static_cast<uint32_t>(comp[i] & 0xFFFFFFFFull)];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = xt[i];
    __syncthreads();
  }
}

extern "C" int64_t b200ppo_permutation_scratch_bytes(int32_t n, int32_t n_epochs) {
  if (n <= 0 || n_epochs <= 0) return 0;
  int64_t n2 = next_pow2(n);
  return n_epochs * (n2 * 8 + static_cast<int64_t>(n) * 4) + 256;
}

extern "C" int b200ppo_permutation(void* stream, const uint32_t* new_key, int32_t n, int32_t n_epochs,
                                   int32_t* out, void* scratch) {
  if (n < 0 || n_epochs < 0 || (n > 0 && n_epochs > 0 && (!new_key || !out || !scratch))) return B200PPO_EINVAL;
  if (n == 0 || n_epochs == 0) return 0;
  if (n > (1 << 24)) return B200PPO_ELIMIT;
  if (reinterpret_cast<uintptr_t>(scratch) & 7) return B200PPO_EALIGN;
  const int n2 = next_pow2(n);
  const int use_smem = n2 <= PERM_SMEM_MAX_ELEMS;
  const size_t smem = use_smem ? static_cast<size_t>(n2) * 8 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(perm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         PERM_SMEM_MAX_ELEMS * 8);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  auto* gcomp = static_cast<unsigned long long*>(scratch);
  auto* xtmp = reinterpret_cast<int32_t*>(gcomp + static_cast<size_t>(n_epochs) * n2);
  perm_kernel<<<n_epochs, PERM_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      new_key, n, n2, perm_rounds(n), use_smem, out, gcomp, xtmp);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// K2: gae — ppo.py:351-394.  One thread per env (a warp covers 32 adjacent envs, so every load
// of the time-major [T][B] buffers is a coalesced 128-byte line); reverse scan over T with the
// reference's exact operation order (no FMA contraction: the reference's own KAT leaves ~15 %
// headroom under its 1e-6 gate).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
           const float* __restrict__ last_value, const uint8_t* __restrict__ done,
           const uint8_t* __restrict__ trunc, int T, int B, float lambda_, float gamma,
           float* __restrict__ adv) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float next_adv = 0.0f;
  float next_val = last_value[b];
  constexpr int CH = 8;   // loads of a chunk are independent of the recurrence: issue them first
  for (int t0 = T - 1; t0 >= 0; t0 -= CH) {
    float rr[CH], vv[CH];
    uint8_t dd[CH], tt[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int t = t0 - i;
      if (t >= 0) {
        const size_t gi = static_cast<size_t>(t) * B + b;
        rr[i] = rewards[gi]; vv[i] = values[gi]; dd[i] = done[gi]; tt[i] = trunc[gi];
      }
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int t = t0 - i;
      if (t >= 0) {
        const bool d = dd[i] != 0, tr = tt[i] != 0;
        const float nv = d ? 0.0f : next_val;
        float a = __fsub_rn(__fadd_rn(rr[i], __fmul_rn(gamma, nv)), vv[i]);
        a = tr ? 0.0f : a;
        const float nd = d ? 0.0f : 1.0f;
        next_adv = __fadd_rn(a, __fmul_rn(__fmul_rn(__fmul_rn(nd, gamma), lambda_), next_adv));
        adv[static_cast<size_t>(t) * B + b] = next_adv;
        next_val = vv[i];
      }
    }
  }
}

extern "C" int b200ppo_gae(void* stream, const float* rewards, const float* values_excl_last,
                           const float* last_value, const uint8_t* done, const uint8_t* truncation,
                           int32_t T, int32_t B, float lambda_, float gamma, float* advantages) {
  if (T < 0 || B < 0) return B200PPO_EINVAL;
  if (T == 0 || B == 0) return 0;
  if (!rewards || !values_excl_last || !last_value || !done || !truncation || !advantages) return B200PPO_EINVAL;
  gae_kernel<<<cdiv(B, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      rewards, values_excl_last, last_value, done, truncation, T, B, lambda_, gamma, advantages);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// K5: Normalizer — normalizer.py:63-136
// ------------------------------------------------------------------------------------------
__global__ void norm_prepare_kernel(const float* __restrict__ M2, const float* __restrict__ counter,
                                    int O, float* __restrict__ std_out) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= O) return;
  const float c = counter[0];
  std_out[o] = c > 0.0f ? sqrtf(fmaxf(__fdiv_rn(M2[o], c), 1e-6f)) : 10.0f;
}

extern "C" int b200ppo_norm_prepare(void* stream, const float* M2, const float* counter, int32_t O,
                                    float* std_out) {
  if (O <= 0 || !M2 || !counter || !std_out) return B200PPO_EINVAL;
  norm_prepare_kernel<<<cdiv(O, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(M2, counter, O, std_out);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

// Two-pass batch moments (mean, then sum of squared deviations), like the reference's
// jp.mean / jp.sum(jp.square(v - bm)).  Grid = NORM_BLOCKS row chunks x column blocks; every
// thread owns one column and a strided set of rows; partials are combined in a fixed order, so
// the result is run-to-run deterministic.  HBM-bound: 4*O bytes per row, read once from HBM
// (the second pass re-reads from L2 when the batch fits).
constexpr int NORM_THREADS = 256;
constexpr int NORM_BLOCKS = 296;  // 2 per SM

struct NormGeom {
  int ox;   // columns handled per block
  int ny;   // rows handled in parallel per block
};
__host__ __device__ inline NormGeom norm_geom(int O) {
  NormGeom g;
  g.ox = O < NORM_THREADS ? O : NORM_THREADS;
  g.ny = NORM_THREADS / g.ox;
  return g;
}

template <int PASS>
__global__ void __launch_bounds__(NORM_THREADS)
norm_pass_kernel(const float* __restrict__ x, long long n_rows, int O, float* __restrict__ part_sum,
                 float* __restrict__ part_m2, float* __restrict__ batch_stats,
                 unsigned int* __restrict__ ticket) {
  __shared__ float red[NORM_THREADS];
  __shared__ bool is_last;
  const NormGeom g = norm_geom(O);
  const int tx = threadIdx.x % g.ox, ty = threadIdx.x / g.ox;
  const int o = blockIdx.y * g.ox + tx;
  const bool active = ty < g.ny && o < O;
  const int nb = gridDim.x;
  float mean = 0.0f;
  if (PASS == 1 && active) {
    float s = 0.0f;
    for (int b = 0; b < nb; ++b) s += part_sum[static_cast<size_t>(b) * O + o];
    mean = __fdiv_rn(s, static_cast<float>(n_rows));
  }
  float acc = 0.0f;
  if (active) {
    const long long stride = static_cast<long long>(nb) * g.ny;
#pragma unroll 8
    for (long long r = static_cast<long long>(blockIdx.x) * g.ny + ty; r < n_rows; r += stride) {
      const float v = __ldg(x + r * O + o);
      if (PASS == 0) acc += v;
      else { const float d = v - mean; acc = fmaf(d, d, acc); }
    }
  }
  red[threadIdx.x] = active ? acc : 0.0f;
  __syncthreads();
  if (ty == 0 && o < O) {
    float s = 0.0f;
    for (int y = 0; y < g.ny; ++y) s += red[y * g.ox + tx];
    (PASS == 0 ? part_sum : part_m2)[static_cast<size_t>(blockIdx.x) * O + o] = s;
  }
  if (PASS == 1) {
    // last-arriving block of this column block finalises the batch statistics
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int t = atomicAdd(&ticket[blockIdx.y], 1u);
      is_last = (t == static_cast<unsigned int>(nb) - 1u);
      if (is_last) ticket[blockIdx.y] = 0u;
    }
    __syncthreads();
    if (is_last && ty == 0 && o < O) {
      __threadfence();
      float m2 = 0.0f;
      for (int b = 0; b < nb; ++b) m2 += __ldcg(&part_m2[static_cast<size_t>(b) * O + o]);
      batch_stats[o] = mean;
      batch_stats[O + o] = m2;
    }
  }
}

extern "C" int64_t b200ppo_norm_scratch_bytes(int32_t O) {
  if (O <= 0) return 0;
  return static_cast<int64_t>(2) * NORM_BLOCKS * O * 4 + 4096;
}

extern "C" int b200ppo_norm_batch_stats(void* stream, const float* x, int64_t n_rows, int32_t O,
                                        float* batch_stats, void* scratch) {
  if (O <= 0 || n_rows <= 0 || !x || !batch_stats || !scratch) return B200PPO_EINVAL;
  const NormGeom g = norm_geom(O);
  const int ncb = cdiv(O, g.ox);
  if (ncb > 1024) return B200PPO_ELIMIT;
  auto* part_sum = static_cast<float*>(scratch);
  auto* part_m2 = part_sum + static_cast<size_t>(NORM_BLOCKS) * O;
  auto* ticket = reinterpret_cast<unsigned int*>(part_m2 + static_cast<size_t>(NORM_BLOCKS) * O);
  int nb = static_cast<int>((n_rows + g.ny - 1) / g.ny);
  if (nb > NORM_BLOCKS) nb = NORM_BLOCKS;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(ticket, 0, 4096, s);
  if (e != cudaSuccess) return static_cast<int>(e);
  dim3 grid(nb, ncb);
  norm_pass_kernel<0><<<grid, NORM_THREADS, 0, s>>>(x, n_rows, O, part_sum, part_m2, batch_stats, ticket);
  B200PPO_LAUNCH_CHECK();
  norm_pass_kernel<1><<<grid, NORM_THREADS, 0, s>>>(x, n_rows, O, part_sum, part_m2, batch_stats, ticket);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

// Chan merge of `world` per-rank batch statistics (rank order) and then of the pooled batch into
// the running statistics with the reference's formula (normalizer.py:112-133).  For world == 1
// this is exactly the reference's update.
__global__ void __launch_bounds__(256)
norm_merge_kernel(const float* __restrict__ bs, int world, float n_rank, int O,
                  float* __restrict__ mean, float* __restrict__ M2, float* __restrict__ counter) {
  // single block: every thread reads the old count before thread 0 replaces it
  const float old_count = counter[0];
  float bn_total = n_rank;
  for (int r = 1; r < world; ++r) bn_total += n_rank;
  for (int o = threadIdx.x; o < O; o += blockDim.x) {
    float bm = bs[o], bM2 = bs[O + o], bn = n_rank;
    for (int r = 1; r < world; ++r) {
      const float m_r = bs[static_cast<size_t>(r) * 2 * O + o];
      const float M2_r = bs[static_cast<size_t>(r) * 2 * O + O + o];
      const float tot = bn + n_rank;
      const float d = m_r - bm;
      bm = bm + d * (n_rank / tot);
      bM2 = bM2 + M2_r + (d * d) * bn * n_rank / tot;
      bn = tot;
    }
    const float new_count = __fadd_rn(old_count, bn);
    const float frac = __fdiv_rn(bn, new_count);
    const float delta = __fsub_rn(bm, mean[o]);
    const float new_mean = __fadd_rn(mean[o], __fmul_rn(delta, frac));
    const float corr = __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(delta, delta), old_count), bn), new_count);
    mean[o] = new_mean;
    M2[o] = __fadd_rn(__fadd_rn(M2[o], bM2), corr);
  }
  __syncthreads();
  if (threadIdx.x == 0) counter[0] = __fadd_rn(old_count, bn_total);
}

extern "C" int b200ppo_norm_merge(void* stream, const float* batch_stats, int32_t world,
                                  float n_per_rank, int32_t O, float* mean, float* M2,
                                  float* counter) {
  if (O <= 0 || world <= 0 || !batch_stats || !mean || !M2 || !counter) return B200PPO_EINVAL;
  norm_merge_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(batch_stats, world, n_per_rank,
                                                                    O, mean, M2, counter);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// iteration bookkeeping
// ------------------------------------------------------------------------------------------
__global__ void iter_finalize_kernel(uint32_t* rng_state, uint32_t rng_adv, uint32_t adam_adv,
                                     uint32_t* comm_epoch) {
  rng_state[2] += rng_adv;
  rng_state[3] += adam_adv;
  if (comm_epoch != nullptr) *comm_epoch += adam_adv;
}
extern "C" int b200ppo_iter_finalize(void* stream, uint32_t* rng_state, uint32_t rng_advance,
                                     uint32_t adam_advance, uint32_t* comm_epoch) {
  if (!rng_state) return B200PPO_EINVAL;
  iter_finalize_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(rng_state, rng_advance, adam_advance,
                                                                       comm_epoch);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// FFMA peak probe: 16 independent register-resident FMA chains per thread.
// FLOPs = blocks * threads * iters * 16 * 2.
// ------------------------------------------------------------------------------------------
__global__ void ffma_peak_kernel(int iters, float* __restrict__ sink) {
  float a[16];
  const float x = 1.0f + 1e-7f * threadIdx.x, y = 1e-9f * (blockIdx.x + 1);
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = static_cast<float>(i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
extern "C" int b200ppo_ffma_peak(void* stream, int32_t iters, float* sink, int32_t blocks, int32_t threads) {
  if (!sink || iters <= 0 || blocks <= 0 || threads <= 0 || threads > 1024) return B200PPO_EINVAL;
  ffma_peak_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(iters, sink);
  B200PPO_LAUNCH_CHECK();
  return 0;
}
