// K3 + K4: one PPO minibatch update — ppo.py:296-317 (update_step), 397-531 (ppo_loss),
// 555-569 (optax adam / adamw / clip_by_global_norm).  sm_100a.
//
// Reference structure replaced: nnx.grad(ppo_loss) = a T-step forward scan of tiny per-step
// GEMMs, a reverse GAE scan and the transposed scan.  For the stateless MLP plan the time scan is
// mathematically a flat batch, so one update is six launches over the R = T*mb
// gathered rows (+ mb bootstrap rows for V_{T}):
//   FWD  row-tile kernel: gather + normalise obs, both Dense stacks, pre-activations -> workspace
//   GAE  thread-per-env reverse scan on the fresh values + fp64 advantage moment sums
//   LOSS per-row sampler math, clipped surrogate / value / entropy terms, d loss / d outputs
//   BWD  row-tile dX chain (dY W^T ⊙ act'), then dW = H^T dY as split-row tiles
//   RED  fixed-order reduction of the dW partials -> flat gradient
//   ADAM fused optax update
// All GEMMs share one register-tiled fp32 FFMA routine (TM=128 rows resident in shared memory,
// 8x4 accumulators per thread, B operand streamed in double-buffered 16-deep chunks).
#include <cuda.h>

#include "common.cuh"
#include "tc.cuh"
#include "tmap.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace b200ppo;

namespace {

constexpr int TM = 128;        // rows per CTA tile
constexpr int TN = 64;         // output columns per chunk
constexpr int TK = 16;         // reduction depth of one staged B chunk
constexpr int KB = 256;        // resident reduction length of the A tile
constexpr int LDA = TM + 4;
constexpr int LDB = TN + 4;
constexpr int NTH = 256;
constexpr int GEMM_SMEM = (KB * LDA + 2 * TK * LDB) * 4;

constexpr int DW_T = 64;       // dW output tile (k x n)
constexpr int DW_R = 32;       // rows per staged chunk
constexpr int DW_LD = DW_T + 4;
constexpr int DW_SMEM = 4 * DW_R * DW_LD * 4;

constexpr int MAXL = B200PPO_MAX_LAYERS;
constexpr int MAX_PART_BLOCKS = 1024;   // GAE partial blocks
constexpr int MAX_NORM_BLOCKS = 16384;  // grad-norm partial blocks (one per 256 parameters: 4 M parameters)
constexpr int MAX_LOSS_BLOCKS = 4096;   // loss partial blocks (R <= 524288 rows per update)
constexpr int DBL_GAE_PART = 8;
constexpr int DBL_LOSS_PART = DBL_GAE_PART + 2 * MAX_PART_BLOCKS;
constexpr int NLQ = 6;                   // loss partial sums: actor, critic, reg, clipped count, sum target, sum target^2
constexpr int DBL_GN_PART = DBL_LOSS_PART + NLQ * MAX_LOSS_BLOCKS;
constexpr int DBL_TOTAL = DBL_GN_PART + MAX_NORM_BLOCKS;

// pre-split (hi / lo tf32) weight operands of one layer for the tensor-core path
struct TcLayer {
  size_t wf_hi, wf_lo, wb_hi, wb_lo;   // float offsets into the workspace
  int kpad, npad, nred_pad, kout_pad;  // K -> 32, N -> 16 (fwd operand); N -> 32, K -> 16 (dX operand)
};

struct Layout {
  int R, Rv, S, n_tiles, rows_per_split;
  int tc_ok, tc_tiles, tc_S, tc_rows_per_split, tc_dw_bulk;
  // per dW item (M-tile of a layer): number of row splits and rows per split.  Items whose operand is
  // wider than 128 columns cost ~1.2x per stage (tensor bound), so they get more, shorter splits;
  // tc_S is the maximum (rows of the partial-gradient buffer); slots an item does not use are zeroed
  int tc_item_S[32], tc_item_rps[32];
  TcLayer tca[MAXL], tcc[MAXL];
  size_t xhat, adv, gpart, grad;
  size_t za[MAXL], zc[MAXL], da[MAXL], dc[MAXL];
  size_t dbl;        // doubles: [0..1] adv_sums, [2..3] gnorm2/unused, then partial arrays
  size_t tickets;    // 8 uint32 tickets
  size_t total_floats;
};

inline size_t align64(size_t x) { return (x + 63) & ~static_cast<size_t>(63); }

// Relative cost of one 16-row stage of a dW item (the row splits are handed out so that the most expensive CTA is as
// cheap as possible).  An item on the MN-major operand path is bound by shared-memory traffic and, at 256 gradient
// columns, by the tensor pipe: measured per stage (profiles/r2c_notes.md) 0.95 / 1.0 / 2.05 at 16 / 64 / 256 columns
// (1.85 when the CTA owns only 64 columns of H), 128 columns interpolated.  An item on the transposing path (a width
// that is not a multiple of 4: the value head, a 42-column policy head) costs `thin` (<= 128 columns) or `wide`.
// B200PPO_DW_WTS="thin,wide".
double g_dw_wt_thin = 1.5, g_dw_wt_wide = 1.8;
bool dw_mn_enabled();
int dw_mn_maxn();
bool g_dw_wt_read = false;
void read_dw_weights() {
  if (g_dw_wt_read) return;
  g_dw_wt_read = true;
  const char* e = std::getenv("B200PPO_DW_WTS");
  if (e) {
    double a = 0.0, b = 0.0;
    if (std::sscanf(e, "%lf,%lf", &a, &b) == 2 && a > 0.0 && b > 0.0) { g_dw_wt_thin = a; g_dw_wt_wide = b; }
  }
}

int dw_tiles(const b200ppo_chain& c) {
  int n = 0;
  for (int l = 0; l < c.n_layers; ++l) n += cdiv(c.dims[l], DW_T) * cdiv(c.dims[l + 1], DW_T);
  return n;
}

Layout make_layout(const b200ppo_plan& p, int T, int mb) {
  Layout L;
  read_dw_weights();
  L.R = T * mb;
  L.Rv = (T + 1) * mb;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o = align64(o + n); return r; };
  L.tickets = take(64);
  L.dbl = take(2 * static_cast<size_t>(DBL_TOTAL));
  L.xhat = take(static_cast<size_t>(L.Rv) * p.obs_dim);
  L.adv = take(L.R);
  for (int l = 0; l < p.actor.n_layers; ++l) {
    L.za[l] = take(static_cast<size_t>(L.R) * p.actor.dims[l + 1]);
    L.da[l] = take(static_cast<size_t>(L.R) * p.actor.dims[l + 1]);
  }
  for (int l = 0; l < p.critic.n_layers; ++l) {
    L.zc[l] = take(static_cast<size_t>(L.Rv) * p.critic.dims[l + 1]);
    L.dc[l] = take(static_cast<size_t>(L.R) * p.critic.dims[l + 1]);
  }
  L.n_tiles = dw_tiles(p.actor) + dw_tiles(p.critic);
  const int sms = b200ppo_num_sms();
  int S = (2 * sms) / (L.n_tiles > 0 ? L.n_tiles : 1);
  if (S < 1) S = 1;
  if (S > 32) S = 32;
  int rps = cdiv(cdiv(L.R, S), DW_R) * DW_R;
  if (rps < DW_R) rps = DW_R;
  S = cdiv(L.R, rps);
  L.S = S;
  L.rows_per_split = rps;
  // tensor-core path: D rows tiled by 128, one row-range split per CTA, <= 1 CTA per SM
  L.tc_ok = 1;
  L.tc_tiles = 0;
  L.tc_dw_bulk = 1;
  for (int c = 0; c < 2; ++c) {
    const b200ppo_chain& ch = c == 0 ? p.actor : p.critic;
    TcLayer* tl = c == 0 ? L.tca : L.tcc;
    for (int l = 0; l < ch.n_layers; ++l) {
      const int K = ch.dims[l], N = ch.dims[l + 1];
      if (N > 256) L.tc_ok = 0;
      // dW v2 stages full rows of H (<= 256 wide); a wider OBSERVATION layer (l == 0, H = xhat) is fetched as 2-D
      // TMA boxes instead (16-byte row pitch needed); anything else keeps the gather kernel
      if (K > 256 && !(l == 0 && (K & 3) == 0)) L.tc_dw_bulk = 0;
      tl[l].kpad = (K + 31) & ~31;
      tl[l].npad = (N + 15) & ~15;
      tl[l].nred_pad = (N + 31) & ~31;
      tl[l].kout_pad = (K + 15) & ~15;
      // planes of (rows + 1) float4 (padded plane stride, identical to the smem stage layout)
      tl[l].wf_hi = take(static_cast<size_t>(tl[l].kpad) * (tl[l].npad + 1));
      tl[l].wf_lo = take(static_cast<size_t>(tl[l].kpad) * (tl[l].npad + 1));
      tl[l].wb_hi = take(static_cast<size_t>(tl[l].nred_pad) * (tl[l].kout_pad + 1));
      tl[l].wb_lo = take(static_cast<size_t>(tl[l].nred_pad) * (tl[l].kout_pad + 1));
      L.tc_tiles += cdiv(K, 128);
    }
  }
  // Row splits of the dW items.  `mn_on`: the operand path the cost table is for (b200ppo_set_update_paths); the table
  // of the OTHER setting is computed too, only for its largest split count: the partial-gradient buffer is sized for
  // both, so that flipping the switch on a live workspace can never overrun it.
  auto dw_splits = [&](bool mn_on, int* item_S, int* item_rps) {
    double wts[32];
    int ni = 0;
    for (int c = 0; c < 2; ++c) {
      const b200ppo_chain& ch = c == 0 ? p.actor : p.critic;
      for (int l = 0; l < ch.n_layers; ++l)
        for (int m = 0; m < cdiv(ch.dims[l], 128) && ni < 32; ++m) {
          // relative cost of one 16-row stage of the item (table above g_dw_wt_thin)
          const bool mn = mn_on && !(ch.dims[l] & 3) && !(ch.dims[l + 1] & 3) && ch.dims[l + 1] <= dw_mn_maxn();
          const int n32 = (ch.dims[l + 1] + 31) / 32;
          const int kw_m = ch.dims[l] - m * 128 < 128 ? ch.dims[l] - m * 128 : 128;
          const double w_mn = n32 <= 1 ? 0.95 : (n32 == 2 ? 1.0 : (n32 <= 4 ? 1.33 : (kw_m <= 64 ? 1.85 : 2.05)));
          wts[ni++] = mn ? w_mn : (ch.dims[l + 1] > 128 ? g_dw_wt_wide : g_dw_wt_thin);
        }
    }
    // greedy minimax: every item starts with one split; the item whose CTAs are the most expensive (cost per stage
    // x rows per split, rows in multiples of 32) gets the next one until the SMs are used up (one CTA per SM)
    int Ssum = 0;
    for (int i = 0; i < ni; ++i) {
      item_S[i] = 1;
      item_rps[i] = cdiv(L.R, 32) * 32;
      ++Ssum;
    }
    while (Ssum < sms) {
      int worst = -1;
      double wc = -1.0;
      for (int i = 0; i < ni; ++i) {
        const double c = wts[i] * item_rps[i];
        if (c > wc && item_S[i] < 64 && item_rps[i] > 32) { wc = c; worst = i; }
      }
      if (worst < 0) break;
      // the next split count that actually shortens the item's rows per split
      int Sn = item_S[worst] + 1, rps_n = item_rps[worst];
      while (Sn <= 64 && (rps_n = cdiv(cdiv(L.R, Sn), 32) * 32) >= item_rps[worst]) ++Sn;
      if (Sn > 64) break;
      const int S_real = cdiv(L.R, rps_n);
      if (Ssum - item_S[worst] + S_real > sms) break;
      Ssum += S_real - item_S[worst];
      item_S[worst] = S_real;
      item_rps[worst] = rps_n;
    }
    int smx = 1;
    for (int i = 0; i < ni; ++i)
      if (item_S[i] > smx) smx = item_S[i];
    for (int i = ni; i < 32; ++i) { item_S[i] = 0; item_rps[i] = 32; }
    return smx;
  };
  if (L.tc_tiles > 32) L.tc_ok = 0;
  L.tc_S = dw_splits(dw_mn_enabled(), L.tc_item_S, L.tc_item_rps);
  L.tc_rows_per_split = L.tc_item_rps[0];
  int other_S[32], other_rps[32];
  const int tc_S_other = dw_splits(!dw_mn_enabled(), other_S, other_rps);
  const int tc_S_cap = L.tc_S > tc_S_other ? L.tc_S : tc_S_other;
  const int smax = tc_S_cap > S ? tc_S_cap : S;
  L.gpart = take(static_cast<size_t>(smax) * p.n_params);
  L.grad = take(p.n_params);
  take(16 * 256);                                           // slack: dW v2 copies whole 16-row blocks
  L.total_floats = o;
  return L;
}

// ------------------------------------------------------------------------------------------
// peer-memory exchange (include/b200ppo.h): comm buffer layout and device helpers
// ------------------------------------------------------------------------------------------
constexpr int MAXR = B200PPO_MAX_RANKS;
// Every exchanged 32-bit word travels as ONE 8-byte store {payload, epoch} (the "LL" idea of NCCL's low-latency
// protocol): 8-byte stores are single-copy atomic, so the receiver needs no flag, no fence and no second round
// trip - it polls the word in its OWN memory until the epoch half matches.  Cost: twice the bytes (800 KB per
// rank and update at cfg 2: nothing for NVLink) for one one-way NVLink latency per exchange.  Epochs are
// monotonic and start at 1, buffers alternate on the epoch's parity, so a stale word can never match.
constexpr size_t COMM_ADV = 0;                            // uint2[2 parity][MAXR][4]: the two fp64 sums as four halves
constexpr size_t COMM_GRAD = 2 * MAXR * 4 * sizeof(uint2);   // uint2[2 parity][world][Ppad]: every rank PUSHES its gradient here
inline size_t comm_ppad(int64_t n_params) { return align64(static_cast<size_t>(n_params)); }
// Two-hop gradient exchange (reduce-scatter + all-gather inside the one Adam launch) for larger worlds: every
// rank pushing its whole gradient to every peer moves (world - 1) x P words per rank and direction (5.6 MB at
// world 8, cfg 2: measured 16 us per update, bandwidth bound for 8-byte packets); with an owner rank per
// 256-parameter block it is 2 x (world - 1) / world x P words (1.4 MB) for one more NVLink latency: measured
// 6.27 -> 6.12 ms per iteration at world 8, 5.98 -> 5.96 at world 4 (default: from 4 ranks up).  Layout
// inside COMM_GRAD: scatter uint2[2 parity][world src][chunk], then gather uint2[2 parity][Ppad].
inline size_t comm_chunk(int64_t n_params, int world) {
  const size_t c = (static_cast<size_t>(n_params) + world - 1) / world;
  return (c + 255) & ~static_cast<size_t>(255);
}
int g_twohop_min = -1;
inline bool comm_two_hop(int world) {
  if (g_twohop_min < 0) {
    const char* e = std::getenv("B200PPO_P2P_2HOP");      // smallest world size that uses the two-hop exchange
    g_twohop_min = e ? std::atoi(e) : 4;
    if (g_twohop_min < 2) g_twohop_min = 2;
  }
  return world >= g_twohop_min;
}

struct PeerComm {
  const uint64_t* table;   // device: comm base of every rank (nullptr: exchange disabled)
  int world, rank;
};
__device__ __forceinline__ uint8_t* comm_base(const PeerComm& c, int r) {
  return reinterpret_cast<uint8_t*>(static_cast<uintptr_t>(c.table[r]));
}
__device__ __forceinline__ void ll_store(uint2* p, uint32_t payload, uint32_t epoch) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(payload), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint2 ll_load(const uint2* p) {
  uint2 v;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
// poll one word of OUR buffer until the sender's store of `epoch` has landed
__device__ __forceinline__ uint32_t ll_wait(const uint2* p, uint32_t epoch, int rank, int from) {
  unsigned long long spins = 0;
  for (;;) {
    const uint2 v = ll_load(p);
    if (v.y == epoch) return v.x;
    // bounded: a peer that died must not hang this GPU forever (about a minute of polling, then trap, which
    // surfaces as a launch failure on the host)
    if (++spins > (1ull << 27)) {
      printf("b200ppo: rank %d waited too long for rank %d (epoch %u)\n", rank, from, epoch);
      __trap();
    }
  }
}

// ------------------------------------------------------------------------------------------
// the shared GEMM routine:  C[TM x N] = act_in(A)[rows x K] * B[K x N]
//   A: global row-major, row r at A + r*lda (rows >= nrows read as zero), act_in applied on load
//   B: TRANS_B ? Bm[n*ldb + k] : Bm[k*ldb + n]
//   epi(row_base, col_base, acc[8][4]) once per thread per 64-column chunk
// ------------------------------------------------------------------------------------------
template <bool TRANS_B, class Epi>
__device__ __forceinline__ void gemm_rowtile(float* __restrict__ As, float* __restrict__ Bs,
                                             const float* __restrict__ A, int lda, int row0, int nrows,
                                             int act_in, const float* __restrict__ Bm, int ldb, int K,
                                             int N, Epi epi) {
  const int tid = threadIdx.x;
  const int tn = tid & 15, tm = tid >> 4;
  const int warp = tid >> 5, lane = tid & 31;
  const int nkb = (K + KB - 1) / KB;
  for (int n0 = 0; n0 < N; n0 += TN) {
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int kb = 0; kb < nkb; ++kb) {
      const int k0 = kb * KB;
      const int kc = (K - k0) < KB ? (K - k0) : KB;
      const int kc_pad = (kc + TK - 1) / TK * TK;
      if (nkb > 1 || n0 == 0) {
        __syncthreads();
        // transposed, conflict-free staging: a warp moves 4 rows x 8 k per step
        const int nkg = kc_pad >> 3;
        const int units = (TM / 4) * nkg;
        // all AB loads of a batch are issued before the first one is consumed (memory-level
        // parallelism); the activation is applied at the shared-memory store
        constexpr int AB = 16;
        for (int u0 = warp; u0 < units; u0 += (NTH / 32) * AB) {
          float v[AB];
#pragma unroll
          for (int j = 0; j < AB; ++j) {
            const int u = u0 + j * (NTH / 32);
            const int rg = u / nkg, kg = u - rg * nkg;
            const int r = rg * 4 + (lane >> 3), k = kg * 8 + (lane & 7);
            v[j] = 0.0f;
            if (u < units && row0 + r < nrows && k < kc) v[j] = A[static_cast<size_t>(row0 + r) * lda + k0 + k];
          }
#pragma unroll
          for (int j = 0; j < AB; ++j) {
            const int u = u0 + j * (NTH / 32);
            const int rg = u / nkg, kg = u - rg * nkg;
            const int r = rg * 4 + (lane >> 3), k = kg * 8 + (lane & 7);
            if (u < units) As[k * LDA + r] = act_fwd(v[j], act_in);
          }
        }
        __syncthreads();
      }
      const int nch = kc_pad / TK;
      float breg[4];
      auto load_b = [&](int c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int kk, nn;
          if (TRANS_B) { kk = tid & 15; nn = (tid >> 4) + 16 * i; }
          else { nn = tid & 63; kk = (tid >> 6) + 4 * i; }
          const int k = k0 + c * TK + kk, n = n0 + nn;
          float v = 0.0f;
          if (k < K && n < N) v = TRANS_B ? Bm[static_cast<size_t>(n) * ldb + k] : Bm[static_cast<size_t>(k) * ldb + n];
          breg[i] = v;
        }
      };
      auto store_b = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int kk, nn;
          if (TRANS_B) { kk = tid & 15; nn = (tid >> 4) + 16 * i; }
          else { nn = tid & 63; kk = (tid >> 6) + 4 * i; }
          Bs[buf * TK * LDB + kk * LDB + nn] = breg[i];
        }
      };
      load_b(0);
      store_b(0);
      __syncthreads();
      for (int c = 0; c < nch; ++c) {
        if (c + 1 < nch) load_b(c + 1);
        const float* ap = As + (c * TK) * LDA + tm * 8;
        const float* bp = Bs + (c & 1) * TK * LDB + tn * 4;
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
          const float4 a0 = *reinterpret_cast<const float4*>(ap + kk * LDA);
          const float4 a1 = *reinterpret_cast<const float4*>(ap + kk * LDA + 4);
          const float4 b = *reinterpret_cast<const float4*>(bp + kk * LDB);
          const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (c + 1 < nch) store_b((c + 1) & 1);
        __syncthreads();
      }
    }
    epi(row0 + tm * 8, n0 + tn * 4, acc);
  }
}

// Thin-output variant (N < 16, K <= KB): one thread per (row, column).
template <class Store>
__device__ __forceinline__ void gemm_thin(float* __restrict__ As, const float* __restrict__ A, int lda,
                                          int row0, int nrows, int act_in, const float* __restrict__ W,
                                          int K, int N, Store store) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  const int kc_pad = (K + 7) & ~7;
  const int nkg = kc_pad >> 3;
  const int units = (TM / 4) * nkg;
  constexpr int AB = 16;
  for (int u0 = warp; u0 < units; u0 += (NTH / 32) * AB) {
    float v[AB];
#pragma unroll
    for (int j = 0; j < AB; ++j) {
      const int u = u0 + j * (NTH / 32);
      const int rg = u / nkg, kg = u - rg * nkg;
      const int r = rg * 4 + (lane >> 3), k = kg * 8 + (lane & 7);
      v[j] = 0.0f;
      if (u < units && row0 + r < nrows && k < K) v[j] = A[static_cast<size_t>(row0 + r) * lda + k];
    }
#pragma unroll
    for (int j = 0; j < AB; ++j) {
      const int u = u0 + j * (NTH / 32);
      const int rg = u / nkg, kg = u - rg * nkg;
      const int r = rg * 4 + (lane >> 3), k = kg * 8 + (lane & 7);
      if (u < units) As[k * LDA + r] = act_fwd(v[j], act_in);
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < TM * N; idx += NTH) {
    const int n = idx / TM, m = idx - n * TM;
    float s = 0.0f;
    for (int k = 0; k < K; ++k) s = fmaf(As[k * LDA + m], __ldg(W + static_cast<size_t>(k) * N + n), s);
    store(row0 + m, n, s);
  }
}

// ------------------------------------------------------------------------------------------
// FWD
// ------------------------------------------------------------------------------------------
struct FwdArgs {
  b200ppo_plan plan;
  Layout L;
  const float* obs; const float* next_obs_last; const int32_t* inds;
  const float* mean; const float* std; const float* params;
  float* ws;
  int T, B, mb;
};

__device__ __forceinline__ void chain_forward_tile(const b200ppo_chain& ch, const float* __restrict__ P,
                                                   float* ws, const size_t* zoff, size_t xhat_off, int O,
                                                   int row0, int nrows, float* As, float* Bs) {
  for (int l = 0; l < ch.n_layers; ++l) {
    const int K = ch.dims[l], N = ch.dims[l + 1];
    const float* A = l == 0 ? ws + xhat_off : ws + zoff[l - 1];
    const int act_in = l == 0 ? B200PPO_ACT_NONE : ch.act;
    const float* W = P + ch.w_off[l];
    const float* bias = P + ch.b_off[l];
    float* Z = ws + zoff[l];
    if (N < 16 && K <= KB) {
      gemm_thin(As, A, K, row0, nrows, act_in, W, K, N, [&](int r, int n, float s) {
        if (r < nrows) Z[static_cast<size_t>(r) * N + n] = s + __ldg(bias + n);
      });
    } else {
      gemm_rowtile<false>(As, Bs, A, K, row0, nrows, act_in, W, N, K, N,
                          [&](int rb, int cb, float (&acc)[8][4]) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                              const int r = rb + i;
                              if (r < nrows) {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                  if (cb + j < N) Z[static_cast<size_t>(r) * N + cb + j] = acc[i][j] + __ldg(bias + cb + j);
                              }
                            }
                          });
    }
    __syncthreads();  // this tile's z_l is complete before it is re-read as the next layer's input
  }
}

__global__ void __launch_bounds__(NTH, 1) upd_fwd_kernel(const FwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  pdl_launch_dependents();
  pdl_wait();
  float* As = smem;
  float* Bs = smem + KB * LDA;
  const int O = a.plan.obs_dim;
  const int row0 = blockIdx.x * TM;
  const int R = a.L.R, Rv = a.L.Rv;
  float* xhat = a.ws + a.L.xhat;
  // gather + normalise this tile's observations (ppo.py:297 gather; normalizer.py:78-80)
  // per-row source pointers once per tile, then batched coalesced copies
  const float** rowsrc = reinterpret_cast<const float**>(Bs);
  for (int m = threadIdx.x; m < TM; m += NTH) {
    const int r = row0 + m;
    const float* src = nullptr;
    if (r < R) {
      const int t = r / a.mb, j = r - t * a.mb;
      src = a.obs + (static_cast<size_t>(t) * a.B + a.inds[j]) * O;
    } else if (r < Rv) {
      src = a.next_obs_last + static_cast<size_t>(a.inds[r - R]) * O;
    }
    rowsrc[m] = src;
  }
  __syncthreads();
  constexpr int XB = 8;
  for (int i0 = threadIdx.x; i0 < TM * O; i0 += NTH * XB) {
    float xv[XB];
#pragma unroll
    for (int j = 0; j < XB; ++j) {
      const int idx = i0 + j * NTH;
      const int m = idx / O, k = idx - m * O;
      xv[j] = 0.0f;
      if (idx < TM * O && rowsrc[m] != nullptr) xv[j] = rowsrc[m][k];
    }
#pragma unroll
    for (int j = 0; j < XB; ++j) {
      const int idx = i0 + j * NTH;
      const int m = idx / O, k = idx - m * O;
      if (idx < TM * O && rowsrc[m] != nullptr) {
        float x = xv[j];
        if (a.plan.normalize) x = __fdiv_rn(x - __ldg(a.mean + k), __ldg(a.std + k));
        xhat[static_cast<size_t>(row0 + m) * O + k] = x;
      }
    }
  }
  __syncthreads();
  chain_forward_tile(a.plan.critic, a.params, a.ws, a.L.zc, a.L.xhat, O, row0, Rv, As, Bs);
  if (row0 < R) chain_forward_tile(a.plan.actor, a.params, a.ws, a.L.za, a.L.xhat, O, row0, R, As, Bs);
}

// ------------------------------------------------------------------------------------------
// GAE on the fresh values (ppo.py:447-455) + advantage moment sums for ppo.py:477-480
// ------------------------------------------------------------------------------------------
struct GaeArgs {
  Layout L;
  const float* reward; const uint8_t* done; const uint8_t* trunc; const int32_t* inds;
  float* ws; size_t v_off;
  int T, B, mb;
  float gamma, lambda_;
  const float* hpd;          // device hyper-parameter block (nullable): overrides gamma / lambda_
  PeerComm comm;
  const uint32_t* rng_state;
  const uint32_t* comm_epoch;
  int update_index;
};

// exchange epoch of this update: monotonic base (engine-owned counter, or the Adam count) + index + 1
__device__ __forceinline__ uint32_t comm_epoch_of(const uint32_t* comm_epoch, const uint32_t* rng_state,
                                                  int update_index) {
  return (comm_epoch != nullptr ? *comm_epoch : rng_state[3]) + static_cast<uint32_t>(update_index) + 1u;
}

// Block sums of (adv, adv^2) -> per-block partials -> the last block to finish adds them in block order and
// publishes the minibatch sums (and pushes them to the peers): shared tail of the two GAE kernels.
// `bid` / `nblk`: this block's index among the GAE blocks (the fused GAE + loss kernel runs them as the first blocks
// of a larger grid); `ready` (nullable): flag the last block releases once the advantages and their sums are
// complete, for the loss blocks of the same launch.
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
constexpr int TICKET_GAE_READY = 4;

template <int NW>
__device__ __forceinline__ void gae_finish(const GaeArgs& a, double s1, double s2, double (*red)[NW], bool* is_last,
                                           unsigned int bid, unsigned int nblk, unsigned int* ready) {
  __syncthreads();
  s1 = warp_sum_d(s1);
  s2 = warp_sum_d(s2);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  double* dbl = reinterpret_cast<double*>(a.ws + a.L.dbl);
  double* part = dbl + DBL_GAE_PART;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(a.ws + a.L.tickets);
  if (threadIdx.x == 0) {
    double b1 = 0.0, b2 = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) { b1 += red[0][w]; b2 += red[1][w]; }
    part[2 * bid] = b1;
    part[2 * bid + 1] = b2;
    __threadfence();
    const unsigned int tk = atomicAdd(&ticket[0], 1u);
    *is_last = tk == nblk - 1;
    if (*is_last) {
      ticket[0] = 0u;
      __threadfence();
    }
  }
  __syncthreads();
  if (!*is_last || threadIdx.x >= 32) return;
  // fixed order whatever block finishes last: lane l adds blocks l, l + 32, ... then a butterfly over the lanes
  double t1 = 0.0, t2 = 0.0;
  for (unsigned int b = threadIdx.x; b < nblk; b += 32) {
    t1 += __ldcg(&part[2 * b]);
    t2 += __ldcg(&part[2 * b + 1]);
  }
  t1 = warp_sum_d(t1);
  t2 = warp_sum_d(t2);
  if (threadIdx.x == 0) {
    dbl[0] = t1;   // adv_sums: a data-parallel caller all-reduces these two doubles
    dbl[1] = t2;
    if (a.comm.table != nullptr) {             // ... or every peer gets them through its comm buffer
      const uint32_t epoch = comm_epoch_of(a.comm_epoch, a.rng_state, a.update_index);
      const unsigned long long b1 = static_cast<unsigned long long>(__double_as_longlong(t1));
      const unsigned long long b2 = static_cast<unsigned long long>(__double_as_longlong(t2));
      const uint32_t w[4] = {static_cast<uint32_t>(b1), static_cast<uint32_t>(b1 >> 32), static_cast<uint32_t>(b2),
                             static_cast<uint32_t>(b2 >> 32)};
      for (int r = 0; r < a.comm.world; ++r) {
        uint2* slot = reinterpret_cast<uint2*>(comm_base(a.comm, r) + COMM_ADV) + ((epoch & 1u) * MAXR + a.comm.rank) * 4;
        for (int i = 0; i < 4; ++i) ll_store(slot + i, w[i], epoch);
      }
    }
    if (ready != nullptr) {                    // every block's advantages were fenced before its ticket
      __threadfence();
      st_release_u32(ready, 1u);
    }
  }
}

constexpr int GAE_THREADS = 64;
constexpr int GAE_CHUNK = 32;  // a whole T = 32 rollout in one memory round trip

__global__ void __launch_bounds__(GAE_THREADS) upd_gae_kernel(const GaeArgs a) {
  __shared__ double red[2][4];
  __shared__ bool is_last;
  pdl_launch_dependents();
  pdl_wait();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const float* v = a.ws + a.v_off;
  float* adv = a.ws + a.L.adv;
  const float gamma = a.hpd ? a.hpd[B200PPO_HP_GAMMA] : a.gamma;
  const float lambda_ = a.hpd ? a.hpd[B200PPO_HP_LAMBDA] : a.lambda_;
  double s1 = 0.0, s2 = 0.0;
  if (threadIdx.x < 4) { red[0][threadIdx.x] = 0.0; red[1][threadIdx.x] = 0.0; }
  if (j < a.mb) {
    const int e = a.inds[j];
    float next_adv = 0.0f;
    float next_val = v[static_cast<size_t>(a.L.R) + j];
    // loads of a chunk of time steps are independent of the recurrence: issue them all first
    for (int t0 = a.T - 1; t0 >= 0; t0 -= GAE_CHUNK) {
      float rr[GAE_CHUNK], vv[GAE_CHUNK];
      uint8_t dd[GAE_CHUNK], tt[GAE_CHUNK];
#pragma unroll
      for (int i = 0; i < GAE_CHUNK; ++i) {
        const int t = t0 - i;
        if (t >= 0) {
          const size_t gi = static_cast<size_t>(t) * a.B + e;
          rr[i] = a.reward[gi]; dd[i] = a.done[gi]; tt[i] = a.trunc[gi];
          vv[i] = v[static_cast<size_t>(t) * a.mb + j];
        }
      }
#pragma unroll
      for (int i = 0; i < GAE_CHUNK; ++i) {
        const int t = t0 - i;
        if (t >= 0) {
          const bool d = dd[i] != 0, tr = tt[i] != 0;
          const float nv = d ? 0.0f : next_val;
          float ad = __fsub_rn(__fadd_rn(rr[i], __fmul_rn(gamma, nv)), vv[i]);
          ad = tr ? 0.0f : ad;
          const float nd = d ? 0.0f : 1.0f;
          next_adv = __fadd_rn(ad, __fmul_rn(__fmul_rn(__fmul_rn(nd, gamma), lambda_), next_adv));
          adv[static_cast<size_t>(t) * a.mb + j] = next_adv;
          next_val = vv[i];
          s1 += next_adv;
          s2 += static_cast<double>(next_adv) * next_adv;
        }
      }
    }
  }
  gae_finish<4>(a, s1, s2, red, &is_last, blockIdx.x, gridDim.x, nullptr);
}

// The same recurrence with the time steps of an env spread over the block: every (t, env) element computes its
// TD residual and decay coefficient in parallel (one memory round trip for the whole block, coalesced in env),
// ONE thread per env then runs the T dependent multiply-adds out of shared memory (same operations in the same
// order as the serial kernel: bit-identical advantages), and all threads store the result and accumulate the
// moment sums.  ncu on the serial kernel: 8 blocks x 64 threads on 148 SMs, ~1900 dependent warp
// instructions per warp at 7 cycles each = 10.7 us per launch.
constexpr int GAE_PAR_THREADS = 256;
constexpr int GAE_PAR_EPB = 8;            // envs per block: 64 blocks for a 512-env minibatch

__device__ __forceinline__ void gae_par_body(const GaeArgs& a, float* gsm, const unsigned int bid,
                                             const unsigned int nblk, unsigned int* ready) {
  // gsm: [2][T][EPB]: residual (overwritten by the advantage) | coefficient
  __shared__ double red[2][GAE_PAR_THREADS / 32];
  __shared__ bool is_last;
  __shared__ int env_s[GAE_PAR_EPB];
  constexpr int E = GAE_PAR_EPB;
  const int j0 = static_cast<int>(bid) * E;
  const float* v = a.ws + a.v_off;
  float* adv = a.ws + a.L.adv;
  const float gamma = a.hpd ? a.hpd[B200PPO_HP_GAMMA] : a.gamma;
  const float lambda_ = a.hpd ? a.hpd[B200PPO_HP_LAMBDA] : a.lambda_;
  float* delta = gsm;
  float* coef = gsm + static_cast<size_t>(a.T) * E;
  if (threadIdx.x < E) env_s[threadIdx.x] = j0 + threadIdx.x < a.mb ? a.inds[j0 + threadIdx.x] : 0;
  __syncthreads();
  const int n = a.T * E;
  for (int idx = threadIdx.x; idx < n; idx += GAE_PAR_THREADS) {
    const int t = idx / E, jj = idx - t * E, j = j0 + jj;
    if (j < a.mb) {
      const size_t gi = static_cast<size_t>(t) * a.B + env_s[jj];
      const float rr = a.reward[gi];
      const bool d = a.done[gi] != 0, tr = a.trunc[gi] != 0;
      const float vv = v[static_cast<size_t>(t) * a.mb + j];
      const float next_val = t == a.T - 1 ? v[static_cast<size_t>(a.L.R) + j] : v[static_cast<size_t>(t + 1) * a.mb + j];
      const float nv = d ? 0.0f : next_val;
      const float ad = __fsub_rn(__fadd_rn(rr, __fmul_rn(gamma, nv)), vv);
      delta[idx] = tr ? 0.0f : ad;
      coef[idx] = __fmul_rn(__fmul_rn(d ? 0.0f : 1.0f, gamma), lambda_);
    }
  }
  __syncthreads();
  if (threadIdx.x < E && j0 + threadIdx.x < a.mb) {
    float next_adv = 0.0f;
    for (int t = a.T - 1; t >= 0; --t) {
      const int idx = t * E + threadIdx.x;
      next_adv = __fadd_rn(delta[idx], __fmul_rn(coef[idx], next_adv));
      delta[idx] = next_adv;
    }
  }
  __syncthreads();
  double s1 = 0.0, s2 = 0.0;
  for (int idx = threadIdx.x; idx < n; idx += GAE_PAR_THREADS) {
    const int t = idx / E, jj = idx - t * E, j = j0 + jj;
    if (j < a.mb) {
      const float x = delta[idx];
      adv[static_cast<size_t>(t) * a.mb + j] = x;
      s1 += x;
      s2 += static_cast<double>(x) * x;
    }
  }
  gae_finish<GAE_PAR_THREADS / 32>(a, s1, s2, red, &is_last, bid, nblk, ready);
}

__global__ void __launch_bounds__(GAE_PAR_THREADS) upd_gae_par_kernel(const GaeArgs a) {
  extern __shared__ float gsm[];
  pdl_launch_dependents();
  pdl_wait();
  gae_par_body(a, gsm, blockIdx.x, gridDim.x, nullptr);
}

// ------------------------------------------------------------------------------------------
// LOSS: per-row sampler math, loss terms and d loss / d (actor output, value) — SURVEY App. A
// ------------------------------------------------------------------------------------------
struct LossArgs {
  b200ppo_plan plan;
  Layout L;
  const float* raw_action; const float* loglik_old; const int32_t* inds;
  const uint32_t* rng_state;
  float* ws; float* metrics_out;
  size_t y_off, v_off, dy_off, dv_off;
  int T, B, mb;
  uint32_t count_offset;
  float clip, critic_w;
  const float* hpd;          // device hyper-parameter block (nullable): overrides clip / critic_w
  int normalize_adv;
  double n_global;
  PeerComm comm;
  const uint32_t* comm_epoch;
  int update_index;
  // 1: distillation head (distillation.py:160-232): the loss is the mean negative log-likelihood of the target
  // raw action in `raw_action` (the teacher's mean) under the current policy plus the entropy regulariser; no
  // advantages, no value loss (metrics: [0] = NLL, [1] = 0)
  int nll;
};

// global advantage moments: the local sums, or (peer exchange) all ranks' sums in rank order
__device__ __forceinline__ void loss_adv_stats(const LossArgs& a, float& a_mean, float& a_den) {
  a_mean = 0.0f;
  a_den = 1.0f;
  if (!a.normalize_adv || a.nll) return;
  const double* dbl = reinterpret_cast<const double*>(a.ws + a.L.dbl);
  double s1 = dbl[0], s2 = dbl[1];
  if (a.comm.table != nullptr) {
    const uint32_t epoch = comm_epoch_of(a.comm_epoch, a.rng_state, a.update_index);
    const uint2* slot = reinterpret_cast<const uint2*>(comm_base(a.comm, a.comm.rank) + COMM_ADV) + (epoch & 1u) * MAXR * 4;
    s1 = 0.0;
    s2 = 0.0;
    for (int r = 0; r < a.comm.world; ++r) {               // rank order: identical sums on every rank
      uint32_t w[4];
      for (int i = 0; i < 4; ++i) w[i] = ll_wait(slot + r * 4 + i, epoch, a.comm.rank, r);
      s1 += __longlong_as_double(static_cast<long long>((static_cast<unsigned long long>(w[1]) << 32) | w[0]));
      s2 += __longlong_as_double(static_cast<long long>((static_cast<unsigned long long>(w[3]) << 32) | w[2]));
    }
  }
  const double m = s1 / a.n_global;
  double var = s2 / a.n_global - m * m;
  var = var > 0.0 ? var : 0.0;
  a_mean = static_cast<float>(m);
  a_den = static_cast<float>(sqrt(var)) + 1e-8f;
}

// metrics_out (B200PPO_METRICS_STRIDE floats per update, include/b200ppo.h): [0] actor loss, [1] critic
// loss, [2] regularisation loss, [3] grad norm (Adam kernel), [4] clipping fraction (ppo.py:514-520),
// [5] E[target], [6] E[target^2] (for critic R^2, ppo.py:522-527), [7] E[adv], [8] E[adv^2] — all
// divided by the GLOBAL sample count, so a data-parallel SUM over ranks gives the global means.
// [7], [8]: moments of the advantages the surrogate uses.  ppo.py:477-480 REASSIGNS `advantages` to the
// normalised tensor before it is logged at :523, so with normalize_advantages these are the moments of
// (a - mean) / (std + 1e-8), derived from this rank's raw sums (GAE kernel) and the global statistics.
__device__ __forceinline__ void loss_write_metrics(const LossArgs& a, const double (&s)[NLQ], float a_mean,
                                                   float a_den) {
  const double ng = a.n_global;
  const double* dbl = reinterpret_cast<const double*>(a.ws + a.L.dbl);
  for (int q = 0; q < 3; ++q) a.metrics_out[q] = static_cast<float>(s[q] / ng);
  a.metrics_out[4] = static_cast<float>(s[3] / ng);
  a.metrics_out[5] = static_cast<float>(s[4] / ng);
  a.metrics_out[6] = static_cast<float>(s[5] / ng);
  if (a.nll) {
    a.metrics_out[7] = a.metrics_out[8] = a.metrics_out[9] = 0.0f;
    a.metrics_out[10] = static_cast<float>(static_cast<double>(a.L.R) / ng);
    return;
  }
  const double s1 = __ldcg(dbl), s2 = __ldcg(dbl + 1), n_loc = static_cast<double>(a.L.R);
  const double m = a_mean, den = a_den;
  a.metrics_out[7] = static_cast<float>((s1 - n_loc * m) / den / ng);
  a.metrics_out[8] = static_cast<float>((s2 - 2.0 * m * s1 + n_loc * m * m) / (den * den) / ng);
  // [9], [10]: the global normalisation constants (mean, std + 1e-8; 0 / 1 without normalisation), divided by
  // the number of ranks so that the caller's SUM over ranks returns them unchanged
  const double inv_world = n_loc / ng;
  a.metrics_out[9] = static_cast<float>(m * inv_world);
  a.metrics_out[10] = static_cast<float>(den * inv_world);
}

__global__ void __launch_bounds__(128) upd_loss_kernel(const LossArgs a) {
  __shared__ double red[NLQ][4];
  __shared__ float stats0_s[2];
  pdl_launch_dependents();
  pdl_wait();
  const int A = a.plan.act_dim;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const double ng = a.n_global;
  const float inv_n = static_cast<float>(1.0 / ng);
  if (threadIdx.x == 0) loss_adv_stats(a, stats0_s[0], stats0_s[1]);
  __syncthreads();
  const float a_mean = stats0_s[0], a_den = stats0_s[1];
  const float clip = a.hpd ? a.hpd[B200PPO_HP_CLIP_RANGE] : a.clip;
  const float critic_w = a.hpd ? a.hpd[B200PPO_HP_CRITIC_WEIGHT] : a.critic_w;
  double l_actor = 0.0, l_critic = 0.0, l_reg = 0.0, l_clip = 0.0, l_t1 = 0.0, l_t2 = 0.0;
  if (r < a.L.R) {
    const int t = r / a.mb, j = r - t * a.mb;
    const size_t grow = static_cast<size_t>(t) * a.B + a.inds[j];
    const float* y = a.ws + a.y_off + static_cast<size_t>(r) * 2 * A;
    float* dy = a.ws + a.dy_off + static_cast<size_t>(r) * 2 * A;
    const float* z_in = a.raw_action + grow * A;
    // entropy noise: second draw of replay step t, count = base + 2t + 1, shape (mb, A)
    const Key stream_key{a.rng_state[0], a.rng_state[1]};
    const Key k_ent = fold_in(stream_key, a.rng_state[2] + a.count_offset + 2u * static_cast<uint32_t>(t) + 1u);
    float ll = 0.0f, ent = 0.0f;
    for (int d = 0; d < A; ++d) {
      const float mu = y[d], rho = y[A + d], z = z_in[d];
      const float sigma = (softplus_f(rho) + a.plan.min_std) * a.plan.std_scale;
      const float q = (z - mu) / sigma;
      const float ls = logf(sigma);
      ll += -0.5f * q * q - (B200PPO_HALF_LOG_2PI + ls) - log_det_jac(z);
      const float eps2 = bits_to_normal(random_bits_at(k_ent, static_cast<uint32_t>(j) * A + d));
      const float zp = __fadd_rn(mu, __fmul_rn(sigma, eps2));
      ent += 0.5f + B200PPO_HALF_LOG_2PI + ls + log_det_jac(zp);
    }
    float g_ll, diff = 0.0f;
    l_reg = -static_cast<double>(a.plan.entropy_weight) * ent;
    if (a.nll) {                                               // distillation: d(-mean ll) / d ll = -1 / N
      l_actor = -static_cast<double>(ll);
      g_ll = -inv_n;
    } else {
      const float adv = a.ws[a.L.adv + r];
      const float v = a.ws[a.v_off + r];
      const float target = __fadd_rn(v, adv);                    // ppo.py:456-458
      diff = __fsub_rn(v, target);
      const float an = a.normalize_adv ? (adv - a_mean) / a_den : adv;
      const float ratio = expf(ll - a.loglik_old[grow]);
      const float lo = 1.0f - clip, hi = 1.0f + clip;
      const float c1 = __fmul_rn(ratio, an);
      const float c2 = __fmul_rn(fminf(fmaxf(ratio, lo), hi), an);
      l_actor = -static_cast<double>(fminf(c1, c2));
      l_critic = 0.5 * static_cast<double>(diff) * diff;
      l_clip = fabsf(ratio - 1.0f) > clip ? 1.0 : 0.0;
      l_t1 = target;
      l_t2 = static_cast<double>(target) * target;
      // JAX tie rules: minimum and clip split the cotangent 0.5 / 0.5 on exact ties
      const float w1 = c1 < c2 ? 1.0f : (c1 == c2 ? 0.5f : 0.0f);
      const float w2 = 1.0f - w1;
      const float dclip = (ratio > lo && ratio < hi) ? 1.0f : ((ratio == lo || ratio == hi) ? 0.5f : 0.0f);
      g_ll = -(w1 * an + w2 * an * dclip) * inv_n * ratio;
    }
    const float we = a.plan.entropy_weight * inv_n;
    for (int d = 0; d < A; ++d) {
      const float mu = y[d], rho = y[A + d], z = z_in[d];
      const float sigma = (softplus_f(rho) + a.plan.min_std) * a.plan.std_scale;
      const float eps2 = bits_to_normal(random_bits_at(k_ent, static_cast<uint32_t>(j) * A + d));
      const float th = tanhf(__fadd_rn(mu, __fmul_rn(sigma, eps2)));
      const float dm = z - mu;
      const float is = 1.0f / sigma;
      const float d_mu = g_ll * dm * is * is + we * 2.0f * th;
      const float d_sig = g_ll * (dm * dm * is * is * is - is) - we * (is - 2.0f * th * eps2);
      dy[d] = d_mu;
      dy[A + d] = d_sig * sigmoid_f(rho) * a.plan.std_scale;
    }
    a.ws[a.dv_off + r] = critic_w * diff * inv_n;
  }
  double lq[NLQ] = {l_actor, l_critic, l_reg, l_clip, l_t1, l_t2};
#pragma unroll
  for (int q = 0; q < NLQ; ++q) {
    lq[q] = warp_sum_d(lq[q]);
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = lq[q];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double* part = reinterpret_cast<double*>(a.ws + a.L.dbl) + DBL_LOSS_PART;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(a.ws + a.L.tickets);
    for (int q = 0; q < NLQ; ++q) part[NLQ * blockIdx.x + q] = red[q][0] + red[q][1] + red[q][2] + red[q][3];
    __threadfence();
    const unsigned int tk = atomicAdd(&ticket[1], 1u);
    if (tk == gridDim.x - 1) {
      ticket[1] = 0u;
      __threadfence();
      double s[NLQ] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      for (unsigned int b = 0; b < gridDim.x; ++b)
        for (int q = 0; q < NLQ; ++q) s[q] += __ldcg(&part[NLQ * b + q]);
      loss_write_metrics(a, s, a_mean, a_den);
    }
  }
}

// Same computation with one thread per (row, action dim), A <= 32 (rows padded to a power-of-two lane group): the per-element
// threefry / erfinv / transcendental chains of the A dims run in parallel and the two row sums
// (log-lik, entropy) are butterfly reductions over the A adjacent lanes.  ncu: the thread-per-row
// version ran 4 warps per SM on long dependent chains (27 us per launch).
// `bid` / `nblk`: this block's index among the loss blocks; `ready` (nullable): the fused launch's GAE blocks release
// it when the advantages and their sums are complete (the last loss block clears it again).
__device__ __forceinline__ void loss_par_body(const LossArgs& a, const unsigned int bid, const unsigned int nblk,
                                              unsigned int* ready) {
  __shared__ double red[NLQ][8];
  __shared__ float stats_s[2];
  const int A = a.plan.act_dim;
  // AP = A rounded up to a power of two lanes per row (lanes d >= A idle: Humanoid-scale A = 21 runs on 32 lanes
  // instead of falling back to the thread-per-row kernel: 70 -> 54 us per launch at configs[3])
  const int ap_log2 = A > 1 ? 32 - __clz(A - 1) : 0;
  const int AP = 1 << ap_log2;
  const int gid = static_cast<int>(bid) * blockDim.x + threadIdx.x;
  const int r = gid >> ap_log2, d = gid & (AP - 1);
  const double ng = a.n_global;
  const float inv_n = static_cast<float>(1.0 / ng);
  const bool valid = r < a.L.R && d < A;
  float llt = 0.0f, entt = 0.0f, mu = 0.f, rho = 0.f, z = 0.f, sigma = 1.f, eps2 = 0.f, th = 0.f;
  size_t grow = 0;
  int t = 0, j = 0;
  if (valid) {
    t = r / a.mb; j = r - t * a.mb;
    grow = static_cast<size_t>(t) * a.B + a.inds[j];
    const float* y = a.ws + a.y_off + static_cast<size_t>(r) * 2 * A;
    mu = y[d]; rho = y[A + d]; z = a.raw_action[grow * A + d];
    const Key stream_key{a.rng_state[0], a.rng_state[1]};
    const Key k_ent = fold_in(stream_key, a.rng_state[2] + a.count_offset + 2u * static_cast<uint32_t>(t) + 1u);
    sigma = (softplus_f(rho) + a.plan.min_std) * a.plan.std_scale;
    const float q = (z - mu) / sigma;
    const float ls = logf(sigma);
    llt = -0.5f * q * q - (B200PPO_HALF_LOG_2PI + ls) - log_det_jac(z);
    eps2 = bits_to_normal(random_bits_at(k_ent, static_cast<uint32_t>(j) * A + d));
    const float zp = __fadd_rn(mu, __fmul_rn(sigma, eps2));
    entt = 0.5f + B200PPO_HALF_LOG_2PI + ls + log_det_jac(zp);
    th = tanhf(zp);
  }
  float ll = llt, ent = entt;
  for (int o = 1; o < AP; o <<= 1) {
    ll += __shfl_xor_sync(0xffffffffu, ll, o);
    ent += __shfl_xor_sync(0xffffffffu, ent, o);
  }
  // global advantage moments (fp64 once per block).  Taken AFTER the sampler math: with the peer
  // exchange this is where a rank waits for the other ranks' GAE sums, and the NVLink latency
  // hides behind the threefry / erfinv / transcendental work above.
  if (threadIdx.x == 0) {
    // fused launch: the GAE blocks of this grid (lowest block indices: dispatched first, they wait for nobody) are
    // still at work while the sampler math above runs; everything below needs their advantages
    if (ready != nullptr)
      while (ld_acquire_u32(ready) == 0u) __nanosleep(20);
    loss_adv_stats(a, stats_s[0], stats_s[1]);
  }
  __syncthreads();
  const float a_mean = stats_s[0], a_den = stats_s[1];
  const float clip = a.hpd ? a.hpd[B200PPO_HP_CLIP_RANGE] : a.clip;
  const float critic_w = a.hpd ? a.hpd[B200PPO_HP_CRITIC_WEIGHT] : a.critic_w;
  double l_actor = 0.0, l_critic = 0.0, l_reg = 0.0, l_clip = 0.0, l_t1 = 0.0, l_t2 = 0.0;
  if (valid) {
    float g_ll, diff = 0.0f, c1 = 0.0f, c2 = 0.0f, ratio = 1.0f, target = 0.0f;
    if (a.nll) {                                               // distillation: d(-mean ll) / d ll = -1 / N
      g_ll = -inv_n;
      c1 = c2 = ll;                                            // l_actor = -min(c1, c2) = -ll below
    } else {
      const float adv = __ldcg(a.ws + a.L.adv + r);            // written by other SMs, maybe during this launch
      const float v = a.ws[a.v_off + r];
      target = __fadd_rn(v, adv);
      diff = __fsub_rn(v, target);
      const float an = a.normalize_adv ? (adv - a_mean) / a_den : adv;
      ratio = expf(ll - a.loglik_old[grow]);
      const float lo = 1.0f - clip, hi = 1.0f + clip;
      c1 = __fmul_rn(ratio, an);
      c2 = __fmul_rn(fminf(fmaxf(ratio, lo), hi), an);
      const float w1 = c1 < c2 ? 1.0f : (c1 == c2 ? 0.5f : 0.0f);
      const float w2 = 1.0f - w1;
      const float dclip = (ratio > lo && ratio < hi) ? 1.0f : ((ratio == lo || ratio == hi) ? 0.5f : 0.0f);
      g_ll = -(w1 * an + w2 * an * dclip) * inv_n * ratio;
    }
    const float we = a.plan.entropy_weight * inv_n;
    const float dm = z - mu;
    const float is = 1.0f / sigma;
    const float d_mu = g_ll * dm * is * is + we * 2.0f * th;
    const float d_sig = g_ll * (dm * dm * is * is * is - is) - we * (is - 2.0f * th * eps2);
    float* dy = a.ws + a.dy_off + static_cast<size_t>(r) * 2 * A;
    dy[d] = d_mu;
    dy[A + d] = d_sig * sigmoid_f(rho) * a.plan.std_scale;
    if (d == 0) {
      a.ws[a.dv_off + r] = critic_w * diff * inv_n;
      l_actor = -static_cast<double>(fminf(c1, c2));
      l_critic = 0.5 * static_cast<double>(diff) * diff;
      l_reg = -static_cast<double>(a.plan.entropy_weight) * ent;
      l_clip = (!a.nll && fabsf(ratio - 1.0f) > clip) ? 1.0 : 0.0;
      l_t1 = target;
      l_t2 = static_cast<double>(target) * target;
    }
  }
  double lq[NLQ] = {l_actor, l_critic, l_reg, l_clip, l_t1, l_t2};
  // only the first lane of every AP-lane row group holds a term: butterfly over those lanes only (offsets AP, 2 AP, ..:
  // 2 steps instead of 5 for A = 8; 12 shuffles + 6 fp64 adds per step saved for every warp)
#pragma unroll
  for (int q = 0; q < NLQ; ++q) {
    for (int o = AP; o < 32; o <<= 1) lq[q] += __shfl_xor_sync(0xffffffffu, lq[q], o);
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = lq[q];
  }
  __syncthreads();
  __shared__ bool is_last_s;
  double* part = reinterpret_cast<double*>(a.ws + a.L.dbl) + DBL_LOSS_PART;
  if (threadIdx.x == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(a.ws + a.L.tickets);
    for (int q = 0; q < NLQ; ++q) {
      double sacc = 0.0;
      for (int w = 0; w < 8; ++w) sacc += red[q][w];
      part[NLQ * bid + q] = sacc;
    }
    __threadfence();
    const unsigned int tk = atomicAdd(&ticket[1], 1u);
    is_last_s = tk == nblk - 1;
    if (is_last_s) {
      ticket[1] = 0u;
      if (ready != nullptr) *ready = 0u;      // every loss block passed its wait before it took a ticket
    }
  }
  __syncthreads();
  if (is_last_s) {
    // the last block sums the per-block partials with all its threads (strided loads in flight
    // together, then a fixed-order tree): deterministic, and not a serial latency chain
    __threadfence();
    double acc[NLQ] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (unsigned int b = threadIdx.x; b < nblk; b += blockDim.x)
      for (int q = 0; q < NLQ; ++q) acc[q] += __ldcg(&part[NLQ * b + q]);
    for (int q = 0; q < NLQ; ++q) acc[q] = warp_sum_d(acc[q]);
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
      for (int q = 0; q < NLQ; ++q) red[q][threadIdx.x >> 5] = acc[q];
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot[NLQ];
      for (int q = 0; q < NLQ; ++q) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[q][w];
        tot[q] = t;
      }
      loss_write_metrics(a, tot, a_mean, a_den);
    }
  }
}

__global__ void __launch_bounds__(256) upd_loss_par_kernel(const LossArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  loss_par_body(a, blockIdx.x, gridDim.x, nullptr);
}

// GAE and loss of one update in ONE launch: blocks [0, n_gae) run the time-parallel GAE, the others the loss.  The
// loss blocks' sampler math (threefry, erfinv, the transcendental chains: most of their time) needs no advantage, so
// it overlaps the GAE blocks' two dependent memory round trips; one launch boundary less per update.
static_assert(GAE_PAR_THREADS == 256, "the fused kernel runs both bodies on 256-thread blocks");
__global__ void __launch_bounds__(256) upd_gae_loss_kernel(const GaeArgs g, const LossArgs a, const unsigned int n_gae) {
  extern __shared__ float gsm[];
  pdl_launch_dependents();
  pdl_wait();
  unsigned int* ready = reinterpret_cast<unsigned int*>(a.ws + a.L.tickets) + TICKET_GAE_READY;
  if (blockIdx.x < n_gae) gae_par_body(g, gsm, blockIdx.x, n_gae, ready);
  else loss_par_body(a, blockIdx.x - n_gae, gridDim.x - n_gae, ready);
}

// ------------------------------------------------------------------------------------------
// BWD dX chain:  dpre_{l-1} = (dpre_l W_l^T) ⊙ act'(z_{l-1})
// ------------------------------------------------------------------------------------------
struct BwdArgs {
  b200ppo_plan plan;
  Layout L;
  const float* params;
  float* ws;
};

__device__ __forceinline__ void chain_backward_tile(const b200ppo_chain& ch, const float* __restrict__ P,
                                                    float* ws, const size_t* zoff, const size_t* doff,
                                                    int row0, int nrows, float* As, float* Bs) {
  for (int l = ch.n_layers - 1; l >= 1; --l) {
    const int Kl = ch.dims[l], Nl = ch.dims[l + 1];
    const float* dY = ws + doff[l];
    const float* W = P + ch.w_off[l];
    const float* zprev = ws + zoff[l - 1];
    float* dprev = ws + doff[l - 1];
    const int act = ch.act;
    gemm_rowtile<true>(As, Bs, dY, Nl, row0, nrows, B200PPO_ACT_NONE, W, Nl, Nl, Kl,
                       [&](int rb, int cb, float (&acc)[8][4]) {
#pragma unroll
                         for (int i = 0; i < 8; ++i) {
                           const int r = rb + i;
                           if (r < nrows) {
#pragma unroll
                             for (int j = 0; j < 4; ++j)
                               if (cb + j < Kl) {
                                 const size_t o = static_cast<size_t>(r) * Kl + cb + j;
                                 dprev[o] = acc[i][j] * act_grad(zprev[o], act);
                               }
                           }
                         }
                       });
    __syncthreads();
  }
}

__global__ void __launch_bounds__(NTH, 1) upd_bwd_dx_kernel(const BwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  pdl_launch_dependents();
  pdl_wait();
  float* As = smem;
  float* Bs = smem + KB * LDA;
  const int row0 = blockIdx.x * TM;
  chain_backward_tile(a.plan.critic, a.params, a.ws, a.L.zc, a.L.dc, row0, a.L.R, As, Bs);
  chain_backward_tile(a.plan.actor, a.params, a.ws, a.L.za, a.L.da, row0, a.L.R, As, Bs);
}

// ------------------------------------------------------------------------------------------
// BWD dW:  dW_l = act(z_{l-1})^T dpre_l,  db_l = colsum(dpre_l); 64x64 output tiles, rows split
// into S ranges whose partial results land in gpart[s][P] (reduced in a fixed order by RED).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTH, 2) upd_bwd_dw_kernel(const BwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  pdl_launch_dependents();
  pdl_wait();
  float* As = smem;                       // [2][DW_R][DW_LD]
  float* Bs = smem + 2 * DW_R * DW_LD;    // [2][DW_R][DW_LD]
  // decode blockIdx.x -> (chain, layer, k tile, n tile)
  int item = blockIdx.x;
  const b200ppo_chain* ch = &a.plan.actor;
  const size_t* zoff = a.L.za;
  const size_t* doff = a.L.da;
  int layer = -1, kt = 0, nt = 0;
  for (int c = 0; c < 2 && layer < 0; ++c) {
    ch = c == 0 ? &a.plan.actor : &a.plan.critic;
    zoff = c == 0 ? a.L.za : a.L.zc;
    doff = c == 0 ? a.L.da : a.L.dc;
    for (int l = 0; l < ch->n_layers; ++l) {
      const int nk = (ch->dims[l] + DW_T - 1) / DW_T, nn = (ch->dims[l + 1] + DW_T - 1) / DW_T;
      if (item < nk * nn) { layer = l; kt = item / nn; nt = item - kt * nn; break; }
      item -= nk * nn;
    }
  }
  if (layer < 0) return;
  const int K = ch->dims[layer], N = ch->dims[layer + 1];
  const int k0 = kt * DW_T, n0 = nt * DW_T;
  const float* H = layer == 0 ? a.ws + a.L.xhat : a.ws + zoff[layer - 1];
  const int act_in = layer == 0 ? B200PPO_ACT_NONE : ch->act;
  const float* D = a.ws + doff[layer];
  const int s = blockIdx.y;
  const int r_begin = s * a.L.rows_per_split;
  int r_end = r_begin + a.L.rows_per_split;
  if (r_end > a.L.R) r_end = a.L.R;
  const int tid = threadIdx.x;
  const int tk = tid >> 4, tn = tid & 15;
  const int lc = tid & 63, lr = tid >> 6;   // loader: column, row (4 rows per pass, 8 passes)
  float acc[4][4];
  float bacc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  float ra[8], rb[8];
  auto load = [&](int rc) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = rc + lr + 4 * i;
      float va = 0.0f, vb = 0.0f;
      if (r < r_end) {
        if (k0 + lc < K) va = H[static_cast<size_t>(r) * K + k0 + lc];
        if (n0 + lc < N) vb = D[static_cast<size_t>(r) * N + n0 + lc];
      }
      ra[i] = va;
      rb[i] = vb;
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      As[(buf * DW_R + lr + 4 * i) * DW_LD + lc] = act_fwd(ra[i], act_in);
      Bs[(buf * DW_R + lr + 4 * i) * DW_LD + lc] = rb[i];
    }
  };
  const int nch = (r_end - r_begin + DW_R - 1) / DW_R;
  if (nch > 0) {
    load(r_begin);
    store(0);
  }
  __syncthreads();
  for (int c = 0; c < nch; ++c) {
    if (c + 1 < nch) load(r_begin + (c + 1) * DW_R);
    const float* ap = As + (c & 1) * DW_R * DW_LD + tk * 4;
    const float* bp = Bs + (c & 1) * DW_R * DW_LD + tn * 4;
#pragma unroll
    for (int rr = 0; rr < DW_R; ++rr) {
      const float4 av = *reinterpret_cast<const float4*>(ap + rr * DW_LD);
      const float4 bv = *reinterpret_cast<const float4*>(bp + rr * DW_LD);
      const float a4[4] = {av.x, av.y, av.z, av.w};
      const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
      if (tk == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bacc[j] += b4[j];
      }
    }
    if (c + 1 < nch) store((c + 1) & 1);
    __syncthreads();
  }
  float* gp = a.ws + a.L.gpart + static_cast<size_t>(s) * a.plan.n_params;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + tk * 4 + i;
    if (k < K) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tn * 4 + j;
        if (n < N) gp[ch->w_off[layer] + static_cast<size_t>(k) * N + n] = acc[i][j];
      }
    }
  }
  if (kt == 0 && tk == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn * 4 + j;
      if (n < N) gp[ch->b_off[layer] + n] = bacc[j];
    }
  }
}

#include "update_tc.cuh"

// ------------------------------------------------------------------------------------------
// RED / exchange / grad-norm / ADAM: one kernel, block b owns parameters [256 b, 256 b + 256)
// ------------------------------------------------------------------------------------------
struct AdamArgs {
  const float* gpart; int S;        // S > 0: fixed-order reduction of the dW row-split partials first
  const float* grad;                // S == 0: the (already reduced / all-reduced) flat gradient
  float* grad_out;                  // nullable: where the gradient this launch ends up with is written
  float* params; float* mu; float* nu;
  const uint32_t* rng_state; const uint32_t* comm_epoch; const double* gnorm2; float* metrics_out;
  int64_t P;
  int update_index;
  float lr, b1, b2, eps, wd, clip;
  const float* hpd;                 // device hyper-parameter block (nullable): overrides the six values above
  // peer exchange (table != nullptr): every thread pushes its reduced gradient element, epoch attached, into
  // slot [parity][rank] of EVERY rank's buffer and sums the slots of its own buffer in rank order
  PeerComm comm;
  size_t comm_ppad;
  size_t comm_chunk;                // > 0: two-hop exchange, parameters [r * chunk, (r + 1) * chunk) are reduced by rank r
  int do_adam;                      // 0: reduce / exchange / norm only
  // nullable: squared global norm of the (summed) gradient -> *norm_out (block partials, ticket, fixed order)
  double* norm_part; double* norm_out; unsigned int* norm_ticket;
  const uint8_t* mask;  // nullable: 0 = structural zero (off-diagonal block of per-key encoders), never updated;
                        // 2 = second copy of a tied (shared-trunk) parameter: updated, not counted in the norm
  const int32_t* tie;   // nullable: index of the tied partner of parameter i, or -1 (used with S > 0 only)
  // tensor-core path: the updated weight is also re-split into the hi / lo operand planes of the next
  // update (what upd_prep_w_kernel does from scratch), so that update can skip its prep launch
  int n_seg;                                   // 0: no refresh
  float* ws;
  struct Seg { int64_t w_off; int K, N, prow_f, prow_b; size_t wf_hi, wf_lo, wb_hi, wb_lo; } seg[2 * MAXL];
};

// optax.adam / adamw (scale_by_adam -> [add_decayed_weights] -> scale(-lr)), optionally preceded by
// clip_by_global_norm — ppo.py:555-569.
template <bool TIED>       // TIED: shared-trunk plans (a.tie != nullptr); the plain instantiation carries no extra code
__global__ void __launch_bounds__(256) upd_adam_kernel(const AdamArgs a) {
  __shared__ double nred[8];
  pdl_launch_dependents();
  pdl_wait();
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const bool in = i < a.P;
  float g = 0.0f;
  if (in) {
    if (a.S > 0) {
      for (int sp = 0; sp < a.S; ++sp) g += a.gpart[static_cast<size_t>(sp) * a.P + i];
      const int32_t j = TIED ? a.tie[i] : -1;
      if (TIED && j >= 0) {  // shared trunk: d(loss)/d(w) = the actor path's + the critic path's (a + b == b + a:
        float g2 = 0.0f;     // both copies get the same bits)
        for (int sp = 0; sp < a.S; ++sp) g2 += a.gpart[static_cast<size_t>(sp) * a.P + j];
        g = __fadd_rn(g, g2);
      }
    } else {
      g = a.grad[i];
    }
  }
  if (a.comm.table != nullptr && in) {
    // push this element (with the epoch in the same 8-byte store) into slot [parity][rank] of EVERY rank's buffer,
    // then collect the slots of our own buffer in rank order: no flag, no fence, no block-wide barrier
    const uint32_t epoch = comm_epoch_of(a.comm_epoch, a.rng_state, a.update_index);
    const int W = a.comm.world, me = a.comm.rank;
    if (a.comm_chunk == 0) {
      const size_t slot = ((epoch & 1u) * W + me) * a.comm_ppad + static_cast<size_t>(i);
      for (int r = 0; r < W; ++r)
        ll_store(reinterpret_cast<uint2*>(comm_base(a.comm, r) + COMM_GRAD) + slot, __float_as_uint(g), epoch);
      const uint2* pg = reinterpret_cast<const uint2*>(comm_base(a.comm, me) + COMM_GRAD) +
                        (epoch & 1u) * W * a.comm_ppad + static_cast<size_t>(i);
      g = 0.0f;
      for (int r = 0; r < W; ++r) g += __uint_as_float(ll_wait(pg + r * a.comm_ppad, epoch, me, r));
    } else {
      // hop 1: this element to its owner rank (block-uniform: chunks are multiples of the block size); the owner
      // sums the world's words in rank order and (hop 2) hands the total to every rank - all ranks get the same bits
      const size_t chunk = a.comm_chunk;
      const int owner = static_cast<int>(static_cast<size_t>(i) / chunk);
      const size_t il = static_cast<size_t>(i) - owner * chunk;
      ll_store(reinterpret_cast<uint2*>(comm_base(a.comm, owner) + COMM_GRAD) + ((epoch & 1u) * W + me) * chunk + il,
               __float_as_uint(g), epoch);
      const size_t goff = 2u * static_cast<size_t>(W) * chunk + (epoch & 1u) * a.comm_ppad + static_cast<size_t>(i);
      if (owner == me) {
        const uint2* ps = reinterpret_cast<const uint2*>(comm_base(a.comm, me) + COMM_GRAD) + (epoch & 1u) * W * chunk + il;
        float t = 0.0f;
        for (int r = 0; r < W; ++r) t += __uint_as_float(ll_wait(ps + r * chunk, epoch, me, r));
        for (int r = 0; r < W; ++r)
          ll_store(reinterpret_cast<uint2*>(comm_base(a.comm, r) + COMM_GRAD) + goff, __float_as_uint(t), epoch);
      }
      g = __uint_as_float(ll_wait(reinterpret_cast<const uint2*>(comm_base(a.comm, me) + COMM_GRAD) + goff, epoch, me, owner));
    }
  }
  if (in && a.grad_out != nullptr) a.grad_out[i] = g;
  if (a.norm_part != nullptr) {
    double sq = (in && !(a.mask && a.mask[i] != 1)) ? static_cast<double>(g) * g : 0.0;
    sq = warp_sum_d(sq);
    if ((threadIdx.x & 31) == 0) nred[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += nred[w];
      a.norm_part[blockIdx.x] = t;
      __threadfence();
      const unsigned int tk = atomicAdd(a.norm_ticket, 1u);
      if (tk == gridDim.x - 1) {
        *a.norm_ticket = 0u;
        __threadfence();
        double tot = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) tot += __ldcg(&a.norm_part[b]);
        *a.norm_out = tot;
      }
    }
  }
  if (!a.do_adam) return;
  const float lr = a.hpd ? a.hpd[B200PPO_HP_LEARNING_RATE] : a.lr;
  const float b1 = a.hpd ? a.hpd[B200PPO_HP_ADAM_B1] : a.b1;
  const float b2 = a.hpd ? a.hpd[B200PPO_HP_ADAM_B2] : a.b2;
  const float eps = a.hpd ? a.hpd[B200PPO_HP_ADAM_EPS] : a.eps;
  const float wd = a.hpd ? a.hpd[B200PPO_HP_WEIGHT_DECAY] : a.wd;
  const float clip = a.hpd ? a.hpd[B200PPO_HP_GRAD_CLIP] : a.clip;
  float gn = 0.0f;
  if (a.clip > 0.0f) {                 // structure (clipping on / off) is the host's; the threshold is run-time data
    gn = sqrtf(static_cast<float>(*a.gnorm2));
    if (i == 0) a.metrics_out[3] = gn;
  }
  if (!in) return;
  if (a.mask && !a.mask[i]) return;
  if (a.clip > 0.0f && !(gn < clip)) g = __fmul_rn(__fdiv_rn(g, gn), clip);
  const float t = static_cast<float>(a.rng_state[3] + static_cast<uint32_t>(a.update_index) + 1u);
  const float m = __fadd_rn(__fmul_rn(1.0f - b1, g), __fmul_rn(b1, a.mu[i]));
  const float v = __fadd_rn(__fmul_rn(1.0f - b2, __fmul_rn(g, g)), __fmul_rn(b2, a.nu[i]));
  a.mu[i] = m;
  a.nu[i] = v;
  const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
  const float mh = __fdiv_rn(m, bc1), vh = __fdiv_rn(v, bc2);
  float u = __fdiv_rn(mh, __fadd_rn(sqrtf(vh), eps));
  const float p = a.params[i];
  if (a.wd >= 0.0f) u = __fadd_rn(u, __fmul_rn(wd, p));
  const float pn = __fadd_rn(p, __fmul_rn(-lr, u));
  a.params[i] = pn;
  for (int q = 0; q < a.n_seg; ++q) {
    const int64_t o = i - a.seg[q].w_off;
    if (o >= 0 && o < static_cast<int64_t>(a.seg[q].K) * a.seg[q].N) {
      const int k = static_cast<int>(o / a.seg[q].N), n = static_cast<int>(o - static_cast<int64_t>(k) * a.seg[q].N);
      const float hi = tc::tf32_round(pn), lo = tc::tf32_round(pn - hi);      // same split as upd_prep_w_kernel
      const size_t f = (static_cast<size_t>(k >> 2) * a.seg[q].prow_f + n) * 4 + (k & 3);   // Bop(n, k) = W[k][n]
      const size_t b = (static_cast<size_t>(n >> 2) * a.seg[q].prow_b + k) * 4 + (n & 3);   // Bop(k, n) = W[k][n]
      a.ws[a.seg[q].wf_hi + f] = hi;
      a.ws[a.seg[q].wf_lo + f] = lo;
      a.ws[a.seg[q].wb_hi + b] = hi;
      a.ws[a.seg[q].wb_lo + b] = lo;
      break;
    }
  }
}

int check_plan_u(const b200ppo_plan* p) {
  if (!p || p->obs_dim <= 0 || p->act_dim <= 0 || p->n_params <= 0) return B200PPO_EINVAL;
  const b200ppo_chain* cs[2] = {&p->actor, &p->critic};
  for (const b200ppo_chain* c : cs) {
    if (c->n_layers < 1 || c->n_layers > MAXL || c->dims[0] != p->obs_dim) return B200PPO_EINVAL;
    for (int l = 0; l < c->n_layers; ++l) {
      if (c->dims[l + 1] <= 0) return B200PPO_EINVAL;
      if ((c->w_off[l] & 3) || (c->b_off[l] & 3)) return B200PPO_EALIGN;
      if (c->w_off[l] < 0 || c->w_off[l] + static_cast<int64_t>(c->dims[l]) * c->dims[l + 1] > p->n_params) return B200PPO_EINVAL;
      if (c->b_off[l] < 0 || c->b_off[l] + c->dims[l + 1] > p->n_params) return B200PPO_EINVAL;
    }
  }
  if (p->actor.dims[p->actor.n_layers] != 2 * p->act_dim || p->critic.dims[p->critic.n_layers] != 1) return B200PPO_EINVAL;
  return 0;
}

bool g_attr_done = false;
int set_attrs() {
  if (g_attr_done) return 0;
  cudaError_t e;
  e = cudaFuncSetAttribute(upd_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(upd_bwd_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(upd_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(upd_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(upd_bwd_dx_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(upd_bwd_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(upd_bwd_dw_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DW2_SMEM);
  if (e != cudaSuccess) return static_cast<int>(e);
  g_attr_done = true;
  return 0;
}

// Tensor map over xhat [rows][O] (fp32, row pitch O * 4 bytes) with a 128-column x 16-row box for the dW kernel's
// wide observation layer.  cuTensorMapEncodeTiled is a host-only driver function: resolved through the runtime
// (no link dependency on libcuda).
int encode_xhat_map(CUtensorMap* tm, const float* xhat, int O, int rows) {
  return encode_tiled_2d(tm, xhat, O, rows, TCM, TCK, false);
}

// GEMM engine: 0 = fp32 FFMA (CUDA cores), 1 = tcgen05 3xTF32 (default, fp32-level accuracy),
// 2 = tcgen05 1xTF32 (JAX-GPU default matmul precision; NOT fp32 parity).  B200PPO_GEMM=ffma|tf32x3|tf32
int g_gemm_mode = -1;
int gemm_mode() {
  if (g_gemm_mode < 0) {
    const char* e = std::getenv("B200PPO_GEMM");
    g_gemm_mode = 1;
    if (e && !std::strcmp(e, "ffma")) g_gemm_mode = 0;
    if (e && !std::strcmp(e, "tf32")) g_gemm_mode = 2;
  }
  return g_gemm_mode;
}

// deepest raw ring of the dW kernel (2 .. DW2_NR_MAX; B200PPO_DW_NR=2 is the fixed two-slot ring of the first version)
int g_dw_nr = -1;
int dw_nr_max() {
  if (g_dw_nr < 0) {
    const char* e = std::getenv("B200PPO_DW_NR");
    int v = e ? std::atoi(e) : DW2_NR_MAX;
    g_dw_nr = v < 2 ? 2 : (v > DW2_NR_MAX ? DW2_NR_MAX : v);
  }
  return g_dw_nr;
}

// dW operand path: MN-major TMA boxes (default) or the transposing v2 path for every layer (B200PPO_DW_MN=0)
int g_dw_mn = -1;
bool dw_mn_enabled() {
  if (g_dw_mn < 0) {
    const char* e = std::getenv("B200PPO_DW_MN");
    g_dw_mn = (e && !std::strcmp(e, "0")) ? 0 : 1;
  }
  return g_dw_mn != 0;
}

int g_dw_mn_maxn = -1;
int dw_mn_maxn() {
  if (g_dw_mn_maxn < 0) {
    const char* e = std::getenv("B200PPO_DW_MN_MAXN");
    g_dw_mn_maxn = e ? std::atoi(e) : 256;
  }
  return g_dw_mn_maxn;
}

// GAE + loss of an update in one launch (default) or two (B200PPO_FUSE_GAE_LOSS=0)
int g_fuse_gl = -1;
bool fuse_gae_loss() {
  if (g_fuse_gl < 0) {
    const char* e = std::getenv("B200PPO_FUSE_GAE_LOSS");
    g_fuse_gl = (e && !std::strcmp(e, "0")) ? 0 : 1;
  }
  return g_fuse_gl != 0;
}

// the fused launch needs both stages in one call, both in their block-parallel form (no NLL head: it has no GAE)
bool gae_loss_fusable(const b200ppo_plan& plan, const Layout& L, int T, int mb, int stages) {
  if (!fuse_gae_loss() || !(stages & B200PPO_STAGE_GAE) || !(stages & B200PPO_STAGE_LOSS) || (stages & B200PPO_STAGE_NLL))
    return false;
  int AP = 1;
  while (AP < plan.act_dim) AP <<= 1;
  const size_t gae_smem = 2 * static_cast<size_t>(T) * GAE_PAR_EPB * sizeof(float);
  return cdiv(mb, GAE_PAR_EPB) <= MAX_PART_BLOCKS && gae_smem <= 8 * 1024 && plan.act_dim <= 32 &&
         cdiv(static_cast<int64_t>(L.R) * AP, 256) <= MAX_LOSS_BLOCKS;
}

}  // namespace

// profiling aid: phase timestamps (clock64) of CTA 0 of the most recent tensor-core update kernel
extern "C" int b200ppo_debug_timestamps(long long* out_host, int32_t max_n) {
  if (!out_host || max_n <= 0) return B200PPO_EINVAL;
  int n = 0;
  cudaError_t e = cudaMemcpyFromSymbol(&n, g_tc_nstamp, sizeof(int));
  if (e != cudaSuccess) return static_cast<int>(e);
  if (n > 128) n = 128;
  if (n > max_n) n = max_n;
  if (n > 0) {
    e = cudaMemcpyFromSymbol(out_host, g_tc_stamp, sizeof(long long) * n);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  if (n + 8 <= max_n) {     // probe accumulators (TC_PROBE builds) appended after the stamps
    e = cudaMemcpyFromSymbol(out_host + n, g_tc_acc, sizeof(long long) * 8);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  return -1000 - n;   // encodes the count: n = -(rc + 1000)
}

extern "C" int b200ppo_debug_cta_times(unsigned long long* out_host, int32_t max_ctas) {
  if (!out_host || max_ctas <= 0) return B200PPO_EINVAL;
  const int n = max_ctas < TC_MAX_CTA_T ? max_ctas : TC_MAX_CTA_T;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaMemcpyFromSymbol(out_host, g_cta_gt, sizeof(unsigned long long) * 3 * n);
  return e == cudaSuccess ? n : -static_cast<int>(e);
}

extern "C" int b200ppo_debug_select(int skip_dw) {
  cudaError_t e = cudaMemcpyToSymbol(g_tc_stamp_skip_dw, &skip_dw, sizeof(int));
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

// bit 0: GAE + loss in one launch, bit 1: dW on the MN-major operand path; argument -1 keeps a setting.  Returns
// the previous settings packed the same way.
extern "C" int b200ppo_set_update_paths(int fuse_gae_loss_on, int dw_mn_on) {
  const int prev = (fuse_gae_loss() ? 1 : 0) | (dw_mn_enabled() ? 2 : 0);
  if (fuse_gae_loss_on >= 0) g_fuse_gl = fuse_gae_loss_on ? 1 : 0;
  if (dw_mn_on >= 0) g_dw_mn = dw_mn_on ? 1 : 0;
  return prev;
}

extern "C" int b200ppo_set_gemm_mode(int mode) {
  const int prev = gemm_mode();
  if (mode >= 0 && mode <= 2) g_gemm_mode = mode;
  return prev;
}

extern "C" int b200ppo_update_num_launches(const b200ppo_plan* plan, const b200ppo_hparams* hp, int32_t T,
                                           int32_t mb, int32_t stages) {
  if (check_plan_u(plan) || !hp || T <= 0 || mb <= 0) return B200PPO_EINVAL;
  const Layout L = make_layout(*plan, T, mb);
  const bool use_tc = gemm_mode() != 0 && L.tc_ok;
  int n = 0;
  if (stages & B200PPO_STAGE_FWD) n += (use_tc && !(stages & B200PPO_STAGE_NO_PREP)) ? 2 : 1;
  if (stages & B200PPO_STAGE_GAE) n += 1;
  if (stages & B200PPO_STAGE_LOSS) n += gae_loss_fusable(*plan, L, T, mb, stages) ? 0 : 1;
  if (stages & B200PPO_STAGE_BWD) n += 2;
  else n += ((stages & B200PPO_STAGE_BWD_DX) ? 1 : 0) + ((stages & B200PPO_STAGE_BWD_DW) ? 1 : 0);
  const bool red = (stages & B200PPO_STAGE_RED) != 0, adam = (stages & B200PPO_STAGE_ADAM) != 0;
  const bool clip = hp->grad_clip > 0.0f;
  if (red && adam) n += clip ? 2 : 1;        // reduce (+ exchange) fused into the Adam launch; the norm needs its own pass
  else if (red) n += 1;
  else if (adam) n += clip ? 2 : 1;
  return n;
}

// Host-only: the weight-gradient kernel's work table for (plan, T, mb): out_splits[i] row splits of rows_per_split[i]
// rows for M-tile item i (actor layers first, one item per 128 input columns of a layer).  Returns the item count.
extern "C" int b200ppo_update_dw_splits(const b200ppo_plan* plan, int32_t T, int32_t mb, int32_t* out_splits,
                                        int32_t* out_rows_per_split, int32_t max_items) {
  if (check_plan_u(plan) || T <= 0 || mb <= 0 || !out_splits || !out_rows_per_split || max_items <= 0) return B200PPO_EINVAL;
  const Layout L = make_layout(*plan, T, mb);
  int n = L.tc_tiles < 32 ? L.tc_tiles : 32;
  if (n > max_items) n = max_items;
  for (int i = 0; i < n; ++i) {
    out_splits[i] = L.tc_item_S[i];
    out_rows_per_split[i] = L.tc_item_rps[i];
  }
  return n;
}

extern "C" int64_t b200ppo_update_workspace_bytes(const b200ppo_plan* plan, int32_t T, int32_t mb) {
  if (check_plan_u(plan) || T <= 0 || mb <= 0) return B200PPO_EINVAL;
  return static_cast<int64_t>(make_layout(*plan, T, mb).total_floats) * 4;
}

extern "C" double* b200ppo_update_adv_sums_ptr(const b200ppo_plan* plan, int32_t T, int32_t mb, void* ws) {
  if (check_plan_u(plan) || !ws) return nullptr;
  return reinterpret_cast<double*>(static_cast<float*>(ws) + make_layout(*plan, T, mb).dbl);
}
extern "C" float* b200ppo_update_grad_ptr(const b200ppo_plan* plan, int32_t T, int32_t mb, void* ws) {
  if (check_plan_u(plan) || !ws) return nullptr;
  return static_cast<float*>(ws) + make_layout(*plan, T, mb).grad;
}
extern "C" float* b200ppo_update_debug_ptr(const b200ppo_plan* plan, int32_t T, int32_t mb, void* ws, int32_t which) {
  if (check_plan_u(plan) || !ws) return nullptr;
  const Layout L = make_layout(*plan, T, mb);
  float* w = static_cast<float*>(ws);
  switch (which) {
    case 0: return w + L.adv;
    case 1: return w + L.zc[plan->critic.n_layers - 1];
    case 2: return w + L.za[plan->actor.n_layers - 1];
    case 3: return w + L.da[plan->actor.n_layers - 1];
    case 4: return w + L.dc[plan->critic.n_layers - 1];
    case 5: return w + L.xhat;
    default: return nullptr;
  }
}

extern "C" int b200ppo_update(void* stream, const b200ppo_plan* plan, const b200ppo_hparams* hp,
                              const b200ppo_update_bufs* b, int32_t T, int32_t B, int32_t mb,
                              uint32_t rng_count_offset, int32_t update_index, int32_t stages) {
  int rc = check_plan_u(plan);
  if (rc) return rc;
  if (!hp || !b || T <= 0 || B <= 0 || mb <= 0 || mb > B || update_index < 0) return B200PPO_EINVAL;
  if (!b->obs || !b->raw_action || !b->loglik_old || !b->reward || !b->done || !b->truncated ||
      !b->next_obs_last || !b->inds || !b->params || !b->adam_mu || !b->adam_nu || !b->rng_state ||
      !b->metrics_out || !b->ws)
    return B200PPO_EINVAL;
  if (plan->normalize && (!b->norm_mean || !b->norm_std)) return B200PPO_EINVAL;
  if (reinterpret_cast<uintptr_t>(b->ws) & 255) return B200PPO_EALIGN;
  if (hp->world_size < 1) return B200PPO_EINVAL;
  const bool use_comm = b->comm != nullptr && hp->world_size > 1;
  PeerComm pc{nullptr, 1, 0};
  if (use_comm) {
    if (hp->world_size > MAXR || hp->rank < 0 || hp->rank >= hp->world_size) return B200PPO_EINVAL;
    pc = PeerComm{b->comm, hp->world_size, hp->rank};
  }
  rc = set_attrs();
  if (rc) return rc;
  const Layout L = make_layout(*plan, T, mb);
  const bool use_tc = gemm_mode() != 0 && L.tc_ok;
  const int tc_split = gemm_mode() == 1 ? 1 : 0;
  if (cdiv(mb, GAE_THREADS) > MAX_PART_BLOCKS || cdiv(L.R, 128) > MAX_LOSS_BLOCKS) return B200PPO_ELIMIT;
  if (cdiv(plan->n_params, 256) > MAX_NORM_BLOCKS) return B200PPO_ELIMIT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(b->ws);
  const size_t v_off = L.zc[plan->critic.n_layers - 1];
  const size_t y_off = L.za[plan->actor.n_layers - 1];
  const size_t dv_off = L.dc[plan->critic.n_layers - 1];
  const size_t dy_off = L.da[plan->actor.n_layers - 1];
  double* dbl = reinterpret_cast<double*>(ws + L.dbl);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(ws + L.tickets);
  // rows per CTA of the row-tiled tensor-core kernels: spread the minibatch over all SMs
  auto tile_for = [](int rows) {
    int t = cdiv(rows, b200ppo_num_sms());
    t = t < 32 ? 32 : t;
    return t > TCM ? TCM : t;
  };
  const int tile_f = tile_for(L.Rv), tile_b = tile_for(L.R);
#define B200PPO_LAUNCH(...)                                           \
  do {                                                                \
    const cudaError_t le__ = launch_k(__VA_ARGS__);                   \
    if (le__ != cudaSuccess) return static_cast<int>(le__);           \
  } while (0)
#define B200PPO_LAUNCH_C(cls, ...)                                    \
  do {                                                                \
    const cudaError_t le__ = launch_kc(cls, __VA_ARGS__);             \
    if (le__ != cudaSuccess) return static_cast<int>(le__);           \
  } while (0)

  if (stages & B200PPO_STAGE_FWD) {
    FwdArgs a;
    a.plan = *plan; a.L = L; a.obs = b->obs; a.next_obs_last = b->next_obs_last; a.inds = b->inds;
    a.mean = b->norm_mean; a.std = b->norm_std; a.params = b->params; a.ws = ws;
    a.T = T; a.B = B; a.mb = mb;
    if (use_tc) {
      if (!(stages & B200PPO_STAGE_NO_PREP)) {
        PrepArgs pa;
        pa.plan = *plan; pa.L = L; pa.params = b->params; pa.ws = ws;
        B200PPO_LAUNCH(upd_prep_w_kernel, dim3(16, 2 * MAXL, 2), dim3(256), 0, s, pa);
      }
      B200PPO_LAUNCH_C(2, upd_fwd_tc_kernel, dim3(cdiv(L.Rv, tile_f)), dim3(TCT), TC_SMEM, s, a, tc_split, 3, tile_f);
    } else {
      B200PPO_LAUNCH(upd_fwd_kernel, dim3(cdiv(L.Rv, TM)), dim3(NTH), GEMM_SMEM, s, a);
    }
  }
  // GAE + loss in one launch (see upd_gae_loss_kernel) when a call runs both stages and both have their
  // block-parallel form; B200PPO_FUSE_GAE_LOSS=0 keeps the two launches (A/B switch)
  GaeArgs ga_fused;
  bool fuse_gl = false;
  if (stages & B200PPO_STAGE_GAE) {
    GaeArgs a;
    a.L = L; a.reward = b->reward; a.done = b->done; a.trunc = b->truncated; a.inds = b->inds;
    a.ws = ws; a.v_off = v_off; a.T = T; a.B = B; a.mb = mb; a.gamma = hp->gamma; a.lambda_ = hp->lambda_;
    a.hpd = b->hparams_dev;
    a.comm = hp->normalize_advantages ? pc : PeerComm{nullptr, 1, 0};
    a.rng_state = b->rng_state; a.comm_epoch = b->comm_epoch; a.update_index = update_index;
    // time-parallel version unless its partial-sum blocks or its shared-memory arrays would not fit
    const size_t gae_smem = 2 * static_cast<size_t>(T) * GAE_PAR_EPB * sizeof(float);
    fuse_gl = gae_loss_fusable(*plan, L, T, mb, stages);
    ga_fused = a;
    if (fuse_gl) {
      // launched below, together with the loss
    } else if (cdiv(mb, GAE_PAR_EPB) <= MAX_PART_BLOCKS && gae_smem <= 40 * 1024)
      B200PPO_LAUNCH_C(1, upd_gae_par_kernel, dim3(cdiv(mb, GAE_PAR_EPB)), dim3(GAE_PAR_THREADS), gae_smem, s, a);
    else
      B200PPO_LAUNCH_C(1, upd_gae_kernel, dim3(cdiv(mb, GAE_THREADS)), dim3(GAE_THREADS), 0, s, a);
  }
  if (stages & B200PPO_STAGE_LOSS) {
    LossArgs a;
    a.plan = *plan; a.L = L; a.raw_action = b->raw_action; a.loglik_old = b->loglik_old; a.inds = b->inds;
    a.rng_state = b->rng_state; a.ws = ws; a.metrics_out = b->metrics_out;
    a.y_off = y_off; a.v_off = v_off; a.dy_off = dy_off; a.dv_off = dv_off;
    a.T = T; a.B = B; a.mb = mb; a.count_offset = rng_count_offset;
    a.clip = hp->clip_range; a.critic_w = hp->critic_loss_weight; a.normalize_adv = hp->normalize_advantages;
    a.hpd = b->hparams_dev;
    a.n_global = static_cast<double>(L.R) * hp->world_size;
    a.nll = (stages & B200PPO_STAGE_NLL) ? 1 : 0;
    a.comm = (hp->normalize_advantages && !a.nll) ? pc : PeerComm{nullptr, 1, 0};
    a.comm_epoch = b->comm_epoch;
    a.update_index = update_index;
    const int A = plan->act_dim;
    int AP = 1;
    while (AP < A) AP <<= 1;
    const bool par = A <= 32 && cdiv(static_cast<int64_t>(L.R) * AP, 256) <= MAX_LOSS_BLOCKS;
    if (fuse_gl) {
      const unsigned int n_gae = cdiv(mb, GAE_PAR_EPB);
      const size_t gae_smem = 2 * static_cast<size_t>(T) * GAE_PAR_EPB * sizeof(float);
      B200PPO_LAUNCH_C(1, upd_gae_loss_kernel, dim3(n_gae + cdiv(static_cast<int64_t>(L.R) * AP, 256)), dim3(256), gae_smem, s,
                       ga_fused, a, n_gae);
    } else if (par) B200PPO_LAUNCH_C(1, upd_loss_par_kernel, dim3(cdiv(static_cast<int64_t>(L.R) * AP, 256)), dim3(256), 0, s, a);
    else B200PPO_LAUNCH_C(1, upd_loss_kernel, dim3(cdiv(L.R, 128)), dim3(128), 0, s, a);
  }
  if (stages & (B200PPO_STAGE_BWD | B200PPO_STAGE_BWD_DX | B200PPO_STAGE_BWD_DW)) {
    const bool do_dx = (stages & B200PPO_STAGE_BWD) || (stages & B200PPO_STAGE_BWD_DX);
    const bool do_dw = (stages & B200PPO_STAGE_BWD) || (stages & B200PPO_STAGE_BWD_DW);
    BwdArgs a;
    a.plan = *plan; a.L = L; a.params = b->params; a.ws = ws;
    if (use_tc) {
      if (do_dx) B200PPO_LAUNCH_C(2, upd_bwd_dx_tc_kernel, dim3(cdiv(L.R, tile_b)), dim3(TCT), TC_SMEM, s, a, tc_split, 3, tile_b);
      if (do_dw) {
        if (L.tc_dw_bulk) {
          CUtensorMap tm;
          std::memset(&tm, 0, sizeof(tm));
          if (plan->obs_dim > 256) {
            const int trc = encode_xhat_map(&tm, ws + L.xhat, plan->obs_dim, L.R);
            if (trc) return trc;
          }
          // v3 operand path (MN-major TMA boxes) for every layer whose two widths give 16-byte row pitches
          DwMaps maps;                          // passed by value: captured with the launch
          std::memset(&maps, 0, sizeof(maps));
          uint32_t mn_mask = 0u;
          if (dw_mn_enabled()) {
            for (int c = 0; c < 2; ++c) {
              const b200ppo_chain& ch = c == 0 ? plan->actor : plan->critic;
              const size_t* zo = c == 0 ? L.za : L.zc;
              const size_t* dof = c == 0 ? L.da : L.dc;
              for (int l = 0; l < ch.n_layers; ++l) {
                const int K = ch.dims[l], N = ch.dims[l + 1];
                const float* H = l == 0 ? ws + L.xhat : ws + zo[l - 1];
                const float* D = ws + dof[l];
                if ((K & 3) || (N & 3) || (reinterpret_cast<uintptr_t>(H) & 15) || (reinterpret_cast<uintptr_t>(D) & 15)) continue;
                if (N > dw_mn_maxn()) continue;
                if (encode_mn_atoms_3d(&maps.h[c * MAXL + l], H, K, L.R, TCK, 4)) continue;
                if (encode_mn_atoms_3d(&maps.d[c * MAXL + l], D, N, L.R, TCK, (N + 31) / 32)) continue;
                mn_mask |= 1u << (c * MAXL + l);
              }
            }
          }
          B200PPO_LAUNCH(upd_bwd_dw_tc2_kernel, dim3(L.tc_tiles, L.tc_S), dim3(TCT), DW2_SMEM, s, a, tc_split, 0, tm, dw_nr_max(),
                         maps, mn_mask);
        }
        else B200PPO_LAUNCH(upd_bwd_dw_tc_kernel, dim3(L.tc_tiles, L.tc_S), dim3(TCT), TC_SMEM, s, a, tc_split, 0);
      }
    } else {
      if (do_dx) B200PPO_LAUNCH(upd_bwd_dx_kernel, dim3(cdiv(L.R, TM)), dim3(NTH), GEMM_SMEM, s, a);
      if (do_dw) B200PPO_LAUNCH(upd_bwd_dw_kernel, dim3(L.n_tiles, L.S), dim3(NTH), DW_SMEM, s, a);
    }
  }
  // RED / ADAM: one kernel does the fixed-order reduction of the dW partials, the peer exchange (when
  // bufs->comm is set) and the optimizer step.  With clip_by_global_norm the squared norm of the summed
  // gradient is a grid-wide reduction that must finish first, so the work splits into two launches.
  const bool do_red = (stages & B200PPO_STAGE_RED) != 0, do_adam = (stages & B200PPO_STAGE_ADAM) != 0;
  if (do_red || do_adam) {
    const bool clip = hp->grad_clip > 0.0f;
    AdamArgs a;
    a.gpart = ws + L.gpart; a.S = do_red ? (use_tc ? L.tc_S : L.S) : 0;
    a.grad = ws + L.grad; a.grad_out = ws + L.grad;
    a.params = b->params; a.mu = b->adam_mu; a.nu = b->adam_nu;
    a.rng_state = b->rng_state; a.comm_epoch = b->comm_epoch; a.gnorm2 = dbl + 2; a.metrics_out = b->metrics_out;
    a.P = plan->n_params; a.update_index = update_index;
    a.lr = hp->learning_rate; a.b1 = hp->adam_b1; a.b2 = hp->adam_b2; a.eps = hp->adam_eps;
    a.wd = hp->weight_decay; a.clip = hp->grad_clip; a.hpd = b->hparams_dev;
    // the exchange belongs to the ADAM stage: a caller that runs RED alone gets the local gradient
    a.comm = do_adam ? pc : PeerComm{nullptr, 1, 0};
    a.comm_ppad = comm_ppad(plan->n_params);
    a.comm_chunk = (a.comm.table != nullptr && comm_two_hop(a.comm.world)) ? comm_chunk(plan->n_params, a.comm.world) : 0;
    a.norm_part = nullptr; a.norm_out = dbl + 2; a.norm_ticket = tickets + 2;
    a.mask = b->param_mask; a.tie = b->param_tie;
    a.n_seg = 0; a.ws = ws;
    if (use_tc) {
      for (int c = 0; c < 2; ++c) {
        const b200ppo_chain& ch = c == 0 ? plan->actor : plan->critic;
        const TcLayer* tl = c == 0 ? L.tca : L.tcc;
        for (int l = 0; l < ch.n_layers; ++l) {
          AdamArgs::Seg& g = a.seg[a.n_seg++];
          g.w_off = ch.w_off[l]; g.K = ch.dims[l]; g.N = ch.dims[l + 1];
          g.prow_f = tl[l].npad + 1; g.prow_b = tl[l].kout_pad + 1;
          g.wf_hi = tl[l].wf_hi; g.wf_lo = tl[l].wf_lo; g.wb_hi = tl[l].wb_hi; g.wb_lo = tl[l].wb_lo;
        }
      }
    }
    const dim3 grid(cdiv(plan->n_params, 256)), block(256);
    void (*adam_k)(const AdamArgs) = a.tie != nullptr ? upd_adam_kernel<true> : upd_adam_kernel<false>;
    if (!do_adam) {                         // RED alone
      a.do_adam = 0;
      B200PPO_LAUNCH_C(1, adam_k, grid, block, 0, s, a);
    } else if (!clip) {                     // (reduce +) (exchange +) Adam in one launch
      a.do_adam = 1;
      B200PPO_LAUNCH_C(1, adam_k, grid, block, 0, s, a);
    } else {                                // pass 1: (reduce +) (exchange +) squared norm; pass 2: clipped Adam
      AdamArgs p1 = a;
      p1.do_adam = 0; p1.norm_part = dbl + DBL_GN_PART;
      B200PPO_LAUNCH_C(1, adam_k, grid, block, 0, s, p1);
      a.S = 0; a.comm = PeerComm{nullptr, 1, 0}; a.grad_out = nullptr; a.do_adam = 1;
      B200PPO_LAUNCH_C(1, adam_k, grid, block, 0, s, a);
    }
  }
#undef B200PPO_LAUNCH
#undef B200PPO_LAUNCH_C
  return 0;
}

// ------------------------------------------------------------------------------------------
// peer-memory exchange: buffer management (CUDA IPC between the one-process-per-GPU ranks)
// ------------------------------------------------------------------------------------------
extern "C" int64_t b200ppo_comm_bytes(const b200ppo_plan* plan, int32_t world_size) {
  if (check_plan_u(plan) || world_size < 1 || world_size > MAXR) return -1;
  // one-hop layout: 2 x world x Ppad words; two-hop layout: 2 x world x chunk + 2 x Ppad words (never larger for
  // world >= 3; sized for both so that the mode is a run-time choice)
  const size_t one = 2 * static_cast<size_t>(world_size) * comm_ppad(plan->n_params);
  const size_t two = 2 * static_cast<size_t>(world_size) * comm_chunk(plan->n_params, world_size) + 2 * comm_ppad(plan->n_params);
  return static_cast<int64_t>(COMM_GRAD + (one > two ? one : two) * sizeof(uint2));
}

extern "C" int b200ppo_comm_alloc(int64_t bytes, void** out) {
  if (!out || bytes <= 0) return B200PPO_EINVAL;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, static_cast<size_t>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaMemset(p, 0, static_cast<size_t>(bytes));
  if (e != cudaSuccess) { cudaFree(p); return static_cast<int>(e); }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(p); return static_cast<int>(e); }
  *out = p;
  return 0;
}

extern "C" int b200ppo_comm_free(void* p) {
  if (!p) return 0;
  return static_cast<int>(cudaFree(p));
}

extern "C" int b200ppo_comm_ipc_get(void* p, uint8_t* handle64) {
  if (!p || !handle64) return B200PPO_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) return static_cast<int>(e);
  std::memcpy(handle64, &h, 64);
  return 0;
}

extern "C" int b200ppo_comm_ipc_open(const uint8_t* handle64, void** out) {
  if (!handle64 || !out) return B200PPO_EINVAL;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return static_cast<int>(e);
  *out = p;
  return 0;
}

extern "C" int b200ppo_comm_ipc_close(void* p) {
  if (!p) return 0;
  return static_cast<int>(cudaIpcCloseMemHandle(p));
}
