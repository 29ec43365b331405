// Recurrent actor (reference networks/recurrent.py:89-161 on top of flax's
// OptimizedLSTMCell): one time step of   obs -> Normalizer -> Dense(act) -> LSTM -> Dense   for a
// tile of rows, forward (rollout / replay) and backward (one step of BPTT).  sm_100a, fp32 FFMA.
//
// First correct version of the row: the time loop lives on the host (T launches forward, T launches
// backward per minibatch), weights stream from L2; weight gradients are three batched GEMMs over all
// steps with fixed-order partial sums (b200ppo_lstm_weight_grads; an atomics mode remains).
// The persistent per-row-tile version (rows are independent, so a CTA can walk all T steps) and the
// tcgen05 gate GEMM are the follow-ups; the arithmetic and the reset semantics are the contract and
// are pinned against oracle/recurrent.py (tests/test_gpu_recurrent.py).
//
// Gate order (i, f, g, o); kernels stored as Wcat = [Wi; Wh] ([in + H, 4H], row-major), bias [4H]:
//   a = [u, h] Wcat + b;  i,f,o = sigmoid, g = tanh;  c' = f c + i g;  h' = o tanh(c');  y = h' W2 + b2
// After the step the carry of rows whose `done` flag is set is replaced by zeros (reset_state).
#include "common.cuh"

namespace {
using namespace b200ppo;

constexpr int NTH = 256;
// rows per CTA: 16 when there are enough rows to fill the GPU (each weight load feeds 16 FMAs),
// 4 for minibatch-sized calls (512 rows -> 128 CTAs instead of 32); the 4-row kernels run 1024
// threads (one output column each: 4x the weight loads in flight on a kernel that is bound by the
// latency of streaming the gate matrix from L2)

struct LstmDims {
  int O, P, H, Y, C;       // obs, lstm input, hidden, output (2A), cache floats per row
};
__host__ __device__ inline LstmDims dims_of(const b200ppo_lstm_plan& p) {
  LstmDims d;
  d.O = p.obs_dim; d.P = p.pre_dim; d.H = p.hidden; d.Y = p.out_dim;
  d.C = d.O + d.P + 7 * d.H;
  return d;
}
// cache row: x[O] | z1[P] | h_in[H] | c_in[H] | i,f,g,o [4H] | tanh(c')[H]

struct FwdArgs {
  b200ppo_lstm_plan plan;
  const float* params; const float* mean; const float* std;
  const float* obs; const int32_t* inds; const uint8_t* done;
  float* c; float* h; float* y; float* cache;
  int rows;
};

// out[r][n] (+)= sum_k in[r][k] * W[k*ldw + n]  for the RT rows of the tile; thread per column n
template <int RT, class Epi>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ in, int ldin, int K,
                                          const float* __restrict__ W, int ldw, int N, Epi epi) {
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float acc[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r] = 0.0f;
    for (int k = 0; k < K; ++k) {
      const float w = __ldg(W + static_cast<size_t>(k) * ldw + n);
#pragma unroll
      for (int r = 0; r < RT; ++r) acc[r] = fmaf(in[r * ldin + k], w, acc[r]);
    }
    epi(n, acc);
  }
}

template <int RT>
__global__ void __launch_bounds__(1024) lstm_step_fwd_kernel(const FwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const LstmDims d = dims_of(a.plan);
  const int O = d.O, P = d.P, H = d.H, Y = d.Y;
  float* xs = sm;                       // [RT][O]
  float* cat = xs + RT * O;             // [RT][P + H]  = [u, h_in]
  float* gates = cat + RT * (P + H);    // [RT][4H]
  float* hn = gates + RT * 4 * H;       // [RT][H]  new hidden
  const int row0 = blockIdx.x * RT;
  const float* Pm = a.params;
  // ---- load + normalise the observations, load the carry
  for (int idx = threadIdx.x; idx < RT * O; idx += blockDim.x) {
    const int r = idx / O, k = idx - r * O;
    const int row = row0 + r;
    float x = 0.0f;
    if (row < a.rows) {
      const int src = a.inds ? a.inds[row] : row;
      x = a.obs[static_cast<size_t>(src) * O + k];
      if (a.plan.normalize) x = __fdiv_rn(x - a.mean[k], a.std[k]);
    }
    xs[idx] = x;
  }
  for (int idx = threadIdx.x; idx < RT * H; idx += blockDim.x) {
    const int r = idx / H, k = idx - r * H;
    const int row = row0 + r;
    cat[r * (P + H) + P + k] = row < a.rows ? a.h[static_cast<size_t>(row) * H + k] : 0.0f;
  }
  __syncthreads();
  float* crow = a.cache;
  // ---- pre Dense: z1 = x W1 + b1, u = act(z1)
  tile_gemm<RT>(xs, O, O, Pm + a.plan.w1_off, P, P, [&](int n, float (&acc)[RT]) {
    const float b = Pm[a.plan.b1_off + n];
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const float z = acc[r] + b;
      cat[r * (P + H) + n] = act_fwd(z, a.plan.act);
      if (crow && row0 + r < a.rows) crow[static_cast<size_t>(row0 + r) * d.C + O + n] = z;
    }
  });
  __syncthreads();
  // ---- gates: a = [u, h] Wcat + b, activation applied
  tile_gemm<RT>(cat, P + H, P + H, Pm + a.plan.wcat_off, 4 * H, 4 * H, [&](int n, float (&acc)[RT]) {
    const float b = Pm[a.plan.bl_off + n];
    const bool is_g = n >= 2 * H && n < 3 * H;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const float v = acc[r] + b;
      gates[r * 4 * H + n] = is_g ? tanhf(v) : sigmoid_f(v);
    }
  });
  __syncthreads();
  // ---- cell update, cache, carry out (with reset)
  for (int idx = threadIdx.x; idx < RT * H; idx += blockDim.x) {
    const int r = idx / H, k = idx - r * H;
    const int row = row0 + r;
    if (row >= a.rows) { hn[idx] = 0.0f; continue; }
    const float ci = a.c[static_cast<size_t>(row) * H + k];
    const float hi = cat[r * (P + H) + P + k];
    const float* g4 = gates + r * 4 * H;
    const float i = g4[k], f = g4[H + k], g = g4[2 * H + k], o = g4[3 * H + k];
    const float c2 = __fadd_rn(__fmul_rn(f, ci), __fmul_rn(i, g));
    const float tc = tanhf(c2);
    const float h2 = __fmul_rn(o, tc);
    hn[idx] = h2;
    if (crow) {
      float* cr = crow + static_cast<size_t>(row) * d.C;
      cr[O + P + k] = hi;
      cr[O + P + H + k] = ci;
      cr[O + P + 2 * H + k] = i;
      cr[O + P + 3 * H + k] = f;
      cr[O + P + 4 * H + k] = g;
      cr[O + P + 5 * H + k] = o;
      cr[O + P + 6 * H + k] = tc;
    }
    bool dn = false;
    if (a.done) dn = a.done[a.inds ? a.inds[row] : row] != 0;
    a.c[static_cast<size_t>(row) * H + k] = dn ? 0.0f : c2;
    a.h[static_cast<size_t>(row) * H + k] = dn ? 0.0f : h2;
  }
  if (crow)
    for (int idx = threadIdx.x; idx < RT * O; idx += blockDim.x) {
      const int r = idx / O, k = idx - r * O;
      if (row0 + r < a.rows) crow[static_cast<size_t>(row0 + r) * d.C + k] = xs[idx];
    }
  __syncthreads();
  // ---- post Dense (linear): y = h' W2 + b2
  tile_gemm<RT>(hn, H, H, Pm + a.plan.w2_off, Y, Y, [&](int n, float (&acc)[RT]) {
    const float b = Pm[a.plan.b2_off + n];
#pragma unroll
    for (int r = 0; r < RT; ++r)
      if (row0 + r < a.rows) a.y[static_cast<size_t>(row0 + r) * Y + n] = acc[r] + b;
  });
}

struct BwdArgs {
  b200ppo_lstm_plan plan;
  const float* params;
  const float* d_y; const float* cache; const uint8_t* done; const int32_t* inds;
  float* dc; float* dh; float* grad;
  int rows;
  // deterministic weight gradients: instead of atomics into `grad`, this step's GEMM operands are
  // written out ([u | h_in], h', da, dz1) and b200ppo_lstm_weight_grads reduces them over all steps
  float* cat_out; float* hn_out; float* da_out; float* dz_out;
};

// grad[k][n] += sum_r A[r][k] * D[r][n]   (A, D shared-memory tiles); thread per column n
template <int RT>
__device__ __forceinline__ void tile_outer_acc(const float* __restrict__ A, int lda, int K,
                                               const float* __restrict__ D, int ldd, int N,
                                               float* __restrict__ G, float* __restrict__ gbias) {
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float dv[RT];
    float bs = 0.0f;
#pragma unroll
    for (int r = 0; r < RT; ++r) { dv[r] = D[r * ldd + n]; bs += dv[r]; }
    if (gbias) atomicAdd(gbias + n, bs);
    for (int k = 0; k < K; ++k) {
      float s = 0.0f;
#pragma unroll
      for (int r = 0; r < RT; ++r) s = fmaf(A[r * lda + k], dv[r], s);
      atomicAdd(G + static_cast<size_t>(k) * N + n, s);
    }
  }
}

// out[r][k] = sum_n D[r][n] * W[k*ldw + n]  (D shared tile [RT][N]); one warp per k, lanes over n
template <int RT, class Epi>
__device__ __forceinline__ void tile_gemm_t(const float* __restrict__ D, int ldd, int N,
                                            const float* __restrict__ W, int ldw, int K, Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < K; k += blockDim.x / 32) {
    float acc[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r] = 0.0f;
    for (int n = lane; n < N; n += 32) {
      const float w = __ldg(W + static_cast<size_t>(k) * ldw + n);
#pragma unroll
      for (int r = 0; r < RT; ++r) acc[r] = fmaf(D[r * ldd + n], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) epi(k, acc);
  }
}

template <int RT>
__global__ void __launch_bounds__(1024) lstm_step_bwd_kernel(const BwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const LstmDims d = dims_of(a.plan);
  const int O = d.O, P = d.P, H = d.H, Y = d.Y;
  float* xs = sm;                       // [RT][O]
  float* cat = xs + RT * O;             // [RT][P + H]  = [u, h_in]
  float* da = cat + RT * (P + H);       // [RT][4H]
  float* hn = da + RT * 4 * H;          // [RT][H]   h' then dh_total
  float* dy = hn + RT * H;              // [RT][Y]
  float* dz = dy + RT * Y;              // [RT][P]   d z1
  const int row0 = blockIdx.x * RT;
  const float* Pm = a.params;
  // ---- load caches
  for (int idx = threadIdx.x; idx < RT * O; idx += blockDim.x) {
    const int r = idx / O, k = idx - r * O;
    xs[idx] = row0 + r < a.rows ? a.cache[static_cast<size_t>(row0 + r) * d.C + k] : 0.0f;
  }
  for (int idx = threadIdx.x; idx < RT * P; idx += blockDim.x) {
    const int r = idx / P, k = idx - r * P;
    const float z = row0 + r < a.rows ? a.cache[static_cast<size_t>(row0 + r) * d.C + O + k] : 0.0f;
    cat[r * (P + H) + k] = act_fwd(z, a.plan.act);
  }
  for (int idx = threadIdx.x; idx < RT * H; idx += blockDim.x) {
    const int r = idx / H, k = idx - r * H;
    const bool ok = row0 + r < a.rows;
    const float* cr = a.cache + static_cast<size_t>(row0 + r) * d.C + O + P;
    cat[r * (P + H) + P + k] = ok ? cr[k] : 0.0f;
    hn[idx] = ok ? cr[5 * H + k] * cr[6 * H + k] : 0.0f;          // h' = o * tanh(c')
  }
  for (int idx = threadIdx.x; idx < RT * Y; idx += blockDim.x) {
    const int r = idx / Y, n = idx - r * Y;
    dy[idx] = row0 + r < a.rows ? a.d_y[static_cast<size_t>(row0 + r) * Y + n] : 0.0f;
  }
  __syncthreads();
  // ---- post Dense: dW2 += h'^T dY, db2 += sum dY
  const bool defer = a.da_out != nullptr;
  if (defer) {
    for (int idx = threadIdx.x; idx < RT * H; idx += blockDim.x) {
      const int r = idx / H, k = idx - r * H;
      if (row0 + r < a.rows) a.hn_out[static_cast<size_t>(row0 + r) * H + k] = hn[idx];
    }
    for (int idx = threadIdx.x; idx < RT * (P + H); idx += blockDim.x) {
      const int r = idx / (P + H), k = idx - r * (P + H);
      if (row0 + r < a.rows) a.cat_out[static_cast<size_t>(row0 + r) * (P + H) + k] = cat[idx];
    }
  } else {
    tile_outer_acc<RT>(hn, H, H, dy, Y, Y, a.grad + a.plan.w2_off, a.grad + a.plan.b2_off);
  }
  __syncthreads();
  // ---- dh_total = dY W2^T + keep * dh_next ; then gate gradients
  tile_gemm_t<RT>(dy, Y, Y, Pm + a.plan.w2_off, Y, H, [&](int k, float (&acc)[RT]) {
#pragma unroll
    for (int r = 0; r < RT; ++r) hn[r * H + k] = acc[r];
  });
  __syncthreads();
  for (int idx = threadIdx.x; idx < RT * H; idx += blockDim.x) {
    const int r = idx / H, k = idx - r * H;
    const int row = row0 + r;
    float* g4 = da + r * 4 * H;
    if (row >= a.rows) { g4[k] = g4[H + k] = g4[2 * H + k] = g4[3 * H + k] = 0.0f; continue; }
    const bool dn = a.done && a.done[a.inds ? a.inds[row] : row] != 0;
    const float keep = dn ? 0.0f : 1.0f;
    const float* cr = a.cache + static_cast<size_t>(row) * d.C + O + P;
    const float ci = cr[H + k], i = cr[2 * H + k], f = cr[3 * H + k], g = cr[4 * H + k], o = cr[5 * H + k],
                tc = cr[6 * H + k];
    const float dh = hn[idx] + keep * a.dh[static_cast<size_t>(row) * H + k];
    const float dc = dh * o * (1.0f - tc * tc) + keep * a.dc[static_cast<size_t>(row) * H + k];
    g4[k] = dc * g * i * (1.0f - i);
    g4[H + k] = dc * ci * f * (1.0f - f);
    g4[2 * H + k] = dc * i * (1.0f - g * g);
    g4[3 * H + k] = dh * tc * o * (1.0f - o);
    a.dc[static_cast<size_t>(row) * H + k] = dc * f;            // w.r.t. the carry entering this step
  }
  __syncthreads();
  // ---- dWcat += [u, h_in]^T da, dbl += sum da
  if (defer) {
    for (int idx = threadIdx.x; idx < RT * 4 * H; idx += blockDim.x) {
      const int r = idx / (4 * H), n = idx - r * 4 * H;
      if (row0 + r < a.rows) a.da_out[static_cast<size_t>(row0 + r) * 4 * H + n] = da[idx];
    }
  } else {
    tile_outer_acc<RT>(cat, P + H, P + H, da, 4 * H, 4 * H, a.grad + a.plan.wcat_off, a.grad + a.plan.bl_off);
  }
  // ---- d[u, h_in] = da Wcat^T : u part -> dz1 (x act'), h part -> dh of the previous step
  tile_gemm_t<RT>(da, 4 * H, 4 * H, Pm + a.plan.wcat_off, 4 * H, P + H, [&](int k, float (&acc)[RT]) {
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int row = row0 + r;
      if (k < P) {
        const float z = row < a.rows ? a.cache[static_cast<size_t>(row) * d.C + O + k] : 0.0f;
        dz[r * P + k] = acc[r] * act_grad(z, a.plan.act);
      } else if (row < a.rows) {
        a.dh[static_cast<size_t>(row) * H + (k - P)] = acc[r];
      }
    }
  });
  __syncthreads();
  // ---- pre Dense: dW1 += x^T dz1, db1 += sum dz1
  if (defer) {
    for (int idx = threadIdx.x; idx < RT * P; idx += blockDim.x) {
      const int r = idx / P, k = idx - r * P;
      if (row0 + r < a.rows) a.dz_out[static_cast<size_t>(row0 + r) * P + k] = dz[idx];
    }
  } else {
    tile_outer_acc<RT>(xs, O, O, dz, P, P, a.grad + a.plan.w1_off, a.grad + a.plan.b1_off);
  }
}

// ------------------------------------------------------------------------------------------
// Weight gradients of the recurrent actor as three batched GEMMs over ALL time steps:
//   out[K x N] = sum_rows A[row][k] * B[row][n],  bias[n] = sum_rows B[row][n]
// 64 x 64 output tiles x row splits; partials are reduced in fixed order => deterministic.
// ------------------------------------------------------------------------------------------
constexpr int GT = 64, GR = 32;          // output tile, rows per shared-memory slab

__global__ void __launch_bounds__(256) atb_partial_kernel(const float* __restrict__ A, int lda, int K,
                                                          const float* __restrict__ B, int ldb, int N,
                                                          int rows, int rows_per_split,
                                                          float* __restrict__ part, float* __restrict__ bpart) {
  __shared__ float As[GR][GT + 1];
  __shared__ float Bs[GR][GT + 1];
  const int k0 = blockIdx.x * GT, n0 = blockIdx.y * GT, sp = blockIdx.z;
  const int r_begin = sp * rows_per_split;
  int r_end = r_begin + rows_per_split;
  if (r_end > rows) r_end = rows;
  const int tk = threadIdx.x >> 4, tn = threadIdx.x & 15;     // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  float bs[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r0 = r_begin; r0 < r_end; r0 += GR) {
    for (int idx = threadIdx.x; idx < GR * GT; idx += 256) {
      const int r = idx / GT, c = idx - r * GT;
      const int row = r0 + r;
      As[r][c] = (row < r_end && k0 + c < K) ? A[static_cast<size_t>(row) * lda + k0 + c] : 0.0f;
      Bs[r][c] = (row < r_end && n0 + c < N) ? B[static_cast<size_t>(row) * ldb + n0 + c] : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < GR; ++r) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[r][tk * 4 + i]; bv[i] = Bs[r][tn * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      if (tk == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bs[j] += bv[j];
      }
    }
    __syncthreads();
  }
  float* po = part + static_cast<size_t>(sp) * K * N;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tk * 4 + i, n = n0 + tn * 4 + j;
      if (k < K && n < N) po[static_cast<size_t>(k) * N + n] = acc[i][j];
    }
  if (blockIdx.x == 0 && tk == 0 && bpart != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn * 4 + j;
      if (n < N) bpart[static_cast<size_t>(sp) * N + n] = bs[j];
    }
  }
}

__global__ void __launch_bounds__(256) atb_reduce_kernel(const float* __restrict__ part, int64_t n, int S,
                                                         float* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float s = 0.0f;
  for (int sp = 0; sp < S; ++sp) s += part[static_cast<size_t>(sp) * n + i];
  out[i] = s;
}

int atb_splits(int K, int N, int rows) {
  const int tiles = cdiv(K, GT) * cdiv(N, GT);
  int S = cdiv(2 * 148, tiles);
  const int smax = cdiv(rows, 4 * GR);
  if (S > smax) S = smax;
  if (S < 1) S = 1;
  if (S > 64) S = 64;
  return S;
}
size_t atb_scratch_floats(int K, int N, int rows) {
  return static_cast<size_t>(atb_splits(K, N, rows)) * (static_cast<size_t>(K) * N + N);
}
int atb_run(cudaStream_t s, const float* A, int lda, int K, const float* B, int ldb, int N, int rows, float* scratch,
            float* gw, float* gb) {
  const int S = atb_splits(K, N, rows);
  const int rps = cdiv(cdiv(rows, S), GR) * GR;
  float* part = scratch;
  float* bpart = scratch + static_cast<size_t>(S) * K * N;
  atb_partial_kernel<<<dim3(cdiv(K, GT), cdiv(N, GT), S), 256, 0, s>>>(A, lda, K, B, ldb, N, rows, rps, part, bpart);
  B200PPO_LAUNCH_CHECK();
  atb_reduce_kernel<<<cdiv(static_cast<int64_t>(K) * N, 256), 256, 0, s>>>(part, static_cast<int64_t>(K) * N, S, gw);
  B200PPO_LAUNCH_CHECK();
  atb_reduce_kernel<<<cdiv(N, 256), 256, 0, s>>>(bpart, N, S, gb);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

int check_lstm_plan(const b200ppo_lstm_plan* p) {
  if (!p || p->obs_dim <= 0 || p->pre_dim <= 0 || p->hidden <= 0 || p->out_dim <= 0 || p->n_params <= 0)
    return B200PPO_EINVAL;
  if (p->act < 0 || p->act > 3) return B200PPO_EINVAL;
  if (p->init_c_off > 0 || p->init_h_off > 0) return B200PPO_ELIMIT;     // learned initial carry: sequence API only
  const int64_t offs[6] = {p->w1_off, p->b1_off, p->wcat_off, p->bl_off, p->w2_off, p->b2_off};
  const int64_t lens[6] = {static_cast<int64_t>(p->obs_dim) * p->pre_dim, p->pre_dim,
                           static_cast<int64_t>(p->pre_dim + p->hidden) * 4 * p->hidden, 4ll * p->hidden,
                           static_cast<int64_t>(p->hidden) * p->out_dim, p->out_dim};
  for (int i = 0; i < 6; ++i)
    if (offs[i] < 0 || offs[i] + lens[i] > p->n_params) return B200PPO_EINVAL;
  return 0;
}

size_t fwd_smem(const LstmDims& d, int RT) { return sizeof(float) * RT * (d.O + (d.P + d.H) + 4 * d.H + d.H); }
size_t bwd_smem(const LstmDims& d, int RT) { return sizeof(float) * RT * (d.O + (d.P + d.H) + 4 * d.H + d.H + d.Y + d.P); }
int rows_per_cta(int rows) { return cdiv(rows, 16) >= 100 ? 16 : 4; }
constexpr size_t SMEM_MAX = 227 * 1024;

}  // namespace

extern "C" int64_t b200ppo_lstm_cache_floats(const b200ppo_lstm_plan* plan, int32_t rows) {
  if (check_lstm_plan(plan) || rows < 0) return -1;
  return static_cast<int64_t>(dims_of(*plan).C) * rows;
}

extern "C" int b200ppo_lstm_step_fwd(void* stream, const b200ppo_lstm_plan* plan, const float* params,
                                     const float* norm_mean, const float* norm_std, const float* obs,
                                     const int32_t* inds, const uint8_t* done, int32_t rows, float* c, float* h,
                                     float* y, float* cache) {
  int rc = check_lstm_plan(plan);
  if (rc) return rc;
  if (rows < 0) return B200PPO_EINVAL;
  if (rows == 0) return 0;
  if (!params || !obs || !c || !h || !y) return B200PPO_EINVAL;
  if (plan->normalize && (!norm_mean || !norm_std)) return B200PPO_EINVAL;
  const LstmDims d = dims_of(*plan);
  const int RT = rows_per_cta(rows);
  const size_t smem = fwd_smem(d, RT);
  if (smem > SMEM_MAX) return B200PPO_ELIMIT;
  cudaError_t e = cudaFuncSetAttribute(lstm_step_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(SMEM_MAX));
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(lstm_step_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(SMEM_MAX));
  if (e != cudaSuccess) return static_cast<int>(e);
  FwdArgs a;
  a.plan = *plan; a.params = params; a.mean = norm_mean; a.std = norm_std; a.obs = obs; a.inds = inds;
  a.done = done; a.c = c; a.h = h; a.y = y; a.cache = cache; a.rows = rows;
  if (RT == 16) lstm_step_fwd_kernel<16><<<cdiv(rows, 16), 1024, smem, static_cast<cudaStream_t>(stream)>>>(a);
  else lstm_step_fwd_kernel<4><<<cdiv(rows, 4), 1024, smem, static_cast<cudaStream_t>(stream)>>>(a);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200ppo_lstm_step_bwd(void* stream, const b200ppo_lstm_plan* plan, const float* params,
                                     const float* d_y, const float* cache, const int32_t* inds,
                                     const uint8_t* done, int32_t rows, float* dc, float* dh, float* grad,
                                     float* cat_out, float* hn_out, float* da_out, float* dz_out) {
  int rc = check_lstm_plan(plan);
  if (rc) return rc;
  if (rows < 0) return B200PPO_EINVAL;
  if (rows == 0) return 0;
  const bool defer = cat_out || hn_out || da_out || dz_out;
  if (defer && !(cat_out && hn_out && da_out && dz_out)) return B200PPO_EINVAL;
  if (!params || !d_y || !cache || !dc || !dh || (!grad && !defer)) return B200PPO_EINVAL;
  const LstmDims d = dims_of(*plan);
  const int RT = rows_per_cta(rows);
  const size_t smem = bwd_smem(d, RT);
  if (smem > SMEM_MAX) return B200PPO_ELIMIT;
  cudaError_t e = cudaFuncSetAttribute(lstm_step_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(SMEM_MAX));
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(lstm_step_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(SMEM_MAX));
  if (e != cudaSuccess) return static_cast<int>(e);
  BwdArgs a;
  a.plan = *plan; a.params = params; a.d_y = d_y; a.cache = cache; a.inds = inds; a.done = done;
  a.rows = rows; a.dc = dc; a.dh = dh; a.grad = grad;
  a.cat_out = cat_out; a.hn_out = hn_out; a.da_out = da_out; a.dz_out = dz_out;
  if (RT == 16) lstm_step_bwd_kernel<16><<<cdiv(rows, 16), 1024, smem, static_cast<cudaStream_t>(stream)>>>(a);
  else lstm_step_bwd_kernel<4><<<cdiv(rows, 4), 1024, smem, static_cast<cudaStream_t>(stream)>>>(a);
  B200PPO_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t b200ppo_lstm_wgrad_scratch_floats(const b200ppo_lstm_plan* plan, int32_t rows_total) {
  if (check_lstm_plan(plan) || rows_total < 0) return -1;
  const LstmDims d = dims_of(*plan);
  size_t m = atb_scratch_floats(d.P + d.H, 4 * d.H, rows_total);
  const size_t m1 = atb_scratch_floats(d.O, d.P, rows_total), m2 = atb_scratch_floats(d.H, d.Y, rows_total);
  m = m1 > m ? m1 : m;
  m = m2 > m ? m2 : m;
  return static_cast<int64_t>(m);
}

extern "C" int b200ppo_lstm_weight_grads(void* stream, const b200ppo_lstm_plan* plan, const float* cache,
                                         const float* cat, const float* hn, const float* da, const float* dz,
                                         const float* d_y, int32_t rows_total, float* grad, float* scratch) {
  int rc = check_lstm_plan(plan);
  if (rc) return rc;
  if (rows_total <= 0 || !cache || !cat || !hn || !da || !dz || !d_y || !grad || !scratch) return B200PPO_EINVAL;
  const LstmDims d = dims_of(*plan);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // every operand is one [steps * rows, width] matrix (step-major); x is the first O columns of the cache rows
  rc = atb_run(s, cat, d.P + d.H, d.P + d.H, da, 4 * d.H, 4 * d.H, rows_total, scratch, grad + plan->wcat_off,
               grad + plan->bl_off);
  if (rc) return rc;
  rc = atb_run(s, cache, d.C, d.O, dz, d.P, d.P, rows_total, scratch, grad + plan->w1_off, grad + plan->b1_off);
  if (rc) return rc;
  return atb_run(s, hn, d.H, d.H, d_y, d.Y, d.Y, rows_total, scratch, grad + plan->w2_off, grad + plan->b2_off);
}
