// The sequential part of the recurrent actor: per-step kernels fed by bulk copies.  Included by
// recurrent_tc.cu inside its anonymous namespace.
//
// A thread-staged GEMM (rg_gemm_kernel) spends ~2.5 K cycles per 32-k slab on loads, hi / lo splits, stores
// and a CTA barrier, against 576 cycles of tensor time for a 128 x 64 tile (measured: 24 us per recurrent
// step, profiles/r2_recurrent_notes.md).  Here every operand of a step already exists in global memory in
// the shared-memory operand layout (tc.cuh: K-major planes of [rows + 1][4] floats, hi and lo halves):
//   * the weights Wh are split once per sequence call by lstm_prep_wh_kernel (forward and transposed forms);
//   * the carry h is written in that layout by the epilogue of the previous step (lane = row, so a plane
//     is written with fully coalesced 512-byte warp stores), gate gradients by the element-wise kernel.
// A step kernel is then: one lane streams 32-k slabs of both operands into a 3-slot ring with
// cp.async.bulk (mbarrier complete_tx), one warp issues the MMAs, tcgen05.commit frees the slots, and all
// eight warps run the epilogue straight out of TMEM.  Everything element-wise that touches per-(row, unit)
// state (gates, c, the activation cache, dc, the split-K partials) uses the same "plane" layout
// [row tile][unit / 4][128 rows][4 units], in which lane = row accesses are coalesced.

constexpr int SP_NS = 3;                       // ring depth
constexpr int PLA = (RM + 1) * 4;              // floats per A operand plane (tc::plane_bytes(128) / 4)
constexpr int PLC = RM * 4;                    // floats per plane of the element-wise state layouts
__host__ __device__ inline int plb(int n) { return (n + 1) * 4; }   // floats per B operand plane of n rows

__host__ __device__ inline uint32_t sp_slot_bytes(int n) {
  return 2u * (RK / 4) * tc::plane_bytes(RM) + 2u * (RK / 4) * tc::plane_bytes(n);
}

// operands of one CTA: K / 4 consecutive planes of each half
struct BulkOperands {
  const float* a_hi; const float* a_lo;        // [K / 4][PLA]
  const float* b_hi; const float* b_lo;        // [K / 4][plb(N)]
  int K;                                       // multiple of 16
  int N;                                       // MMA N (multiple of 16, <= 256)
};

struct StepBars {
  uint64_t full[SP_NS];
  uint64_t empty[SP_NS];
  uint64_t done;
};

__device__ __forceinline__ uint32_t step_init(StepBars* bars, uint32_t* tmem_slot, int ncols, uint32_t& tmem_cols) {
  const int warp = threadIdx.x >> 5;
  tmem_cols = pow2_cols(ncols);
  if (warp == 0) tc::tmem_alloc(tmem_slot, tmem_cols);
  if (threadIdx.x == 32) {
#pragma unroll
    for (int i = 0; i < SP_NS; ++i) {
      tc::mbar_init(&bars->full[i], 1);
      tc::mbar_init(&bars->empty[i], 1);
    }
    tc::mbar_init(&bars->done, 1);
    tc::mbar_init_fence();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  return *tmem_slot;
}

__device__ __forceinline__ void step_fini(uint32_t tmem_base, uint32_t tmem_cols) {
  tc::tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// warp 1: bulk copies; warp 0: MMA issue.  Accumulates into TMEM columns [0, N).
__device__ __forceinline__ void bulk_pipeline(const BulkOperands& g, uint8_t* smem, StepBars* bars, uint32_t tmem_base) {
  const int warp = threadIdx.x >> 5;
  const uint32_t pa = tc::plane_bytes(RM), pb = tc::plane_bytes(g.N);
  const uint32_t slot_bytes = sp_slot_bytes(g.N);
  const int planes = g.K >> 2;
  const int nst = (planes + RK / 4 - 1) / (RK / 4);
  if (warp == 1) {
    if ((threadIdx.x & 31) == 0) {
      for (int s = 0; s < nst; ++s) {
        const int slot = s % SP_NS;
        if (s >= SP_NS) tc::mbar_wait(&bars->empty[slot], static_cast<uint32_t>(s / SP_NS - 1) & 1u);
        const int p0 = s * (RK / 4);
        const int np = (planes - p0) < RK / 4 ? (planes - p0) : RK / 4;
        const uint32_t ab = static_cast<uint32_t>(np) * pa, bb = static_cast<uint32_t>(np) * pb;
        uint8_t* st = smem + slot * slot_bytes;
        tc::mbar_arrive_expect_tx(&bars->full[slot], 2u * ab + 2u * bb);
        tc::bulk_g2s(st, g.a_hi + static_cast<size_t>(p0) * PLA, ab, &bars->full[slot]);
        tc::bulk_g2s(st + (RK / 4) * pa, g.a_lo + static_cast<size_t>(p0) * PLA, ab, &bars->full[slot]);
        tc::bulk_g2s(st + 2 * (RK / 4) * pa, g.b_hi + static_cast<size_t>(p0) * plb(g.N), bb, &bars->full[slot]);
        tc::bulk_g2s(st + 2 * (RK / 4) * pa + (RK / 4) * pb, g.b_lo + static_cast<size_t>(p0) * plb(g.N), bb,
                     &bars->full[slot]);
      }
    }
    __syncwarp();
  } else if (warp == 0) {
    // descriptors are warp-uniform: constant high words, low words advance by adds (see tc_issue in update_tc.cuh)
    const uint64_t da0 = tc::make_desc(tc::smem_u32(smem), pa, 128);
    const uint64_t db0 = tc::make_desc(tc::smem_u32(smem + 2 * (RK / 4) * pa), pb, 128);
    const uint32_t a_hiw = static_cast<uint32_t>(da0 >> 32), b_hiw = static_cast<uint32_t>(db0 >> 32);
    const uint32_t a_lo0 = static_cast<uint32_t>(da0), b_lo0 = static_cast<uint32_t>(db0);
    const uint32_t a_kstep = (2u * pa) >> 4, b_kstep = (2u * pb) >> 4;
    const uint32_t a_half = ((RK / 4) * pa) >> 4, b_half = ((RK / 4) * pb) >> 4;
    const uint32_t slot_step = slot_bytes >> 4;
    const uint32_t idesc = tc::make_idesc_tf32(RM, g.N);
    for (int s = 0; s < nst; ++s) {
      const int slot = s % SP_NS;
      const int p0 = s * (RK / 4);
      const int ks = ((planes - p0) < RK / 4 ? (planes - p0) : RK / 4) >> 1;     // k-steps of 8 in this slab
      tc::mbar_wait(&bars->full[slot], static_cast<uint32_t>(s / SP_NS) & 1u);
      tc::tc_fence_after();
      const uint32_t a0 = a_lo0 + static_cast<uint32_t>(slot) * slot_step;
      const uint32_t b0 = b_lo0 + static_cast<uint32_t>(slot) * slot_step;
      if (tc::elect_one()) {
        for (int j = 0; j < ks; ++j) {
          const uint64_t ah = (static_cast<uint64_t>(a_hiw) << 32) | (a0 + j * a_kstep);
          const uint64_t al = (static_cast<uint64_t>(a_hiw) << 32) | (a0 + j * a_kstep + a_half);
          const uint64_t bh = (static_cast<uint64_t>(b_hiw) << 32) | (b0 + j * b_kstep);
          const uint64_t bl = (static_cast<uint64_t>(b_hiw) << 32) | (b0 + j * b_kstep + b_half);
          const uint32_t acc0 = (s > 0 || j > 0) ? 1u : 0u;
          tc::mma_tf32(tmem_base, al, bh, idesc, acc0);        // small terms first, then the dominant product
          tc::mma_tf32(tmem_base, ah, bl, idesc, 1u);
          tc::mma_tf32(tmem_base, ah, bh, idesc, 1u);
        }
        tc::commit(&bars->empty[slot]);
        if (s == nst - 1) tc::commit(&bars->done);
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------
// Wh [H][4H] -> operand planes (hi / lo), once per sequence call:
//   forward  WhF[unit tile][half][k / 4][64 + 1][4]:  Bop(n', k)  = Wh[k][(n' >> 4) * H + 16 * tile + (n' & 15)]
//   backward WhB[n tile][gate q][half][k / 4][NBT + 1][4]:  Bop(jj, k) = Wh[NBT * tile + jj][q * H + k]
// ------------------------------------------------------------------------------------------
struct PrepWhArgs {
  const float* Wh; int H; int NBT;             // NBT: MMA N of the backward tiles
  float* whf; float* whb;
};

__global__ void __launch_bounds__(256) lstm_prep_wh_kernel(const PrepWhArgs a) {
  const int H = a.H, planes = H >> 2;
  const long long nf = static_cast<long long>(H / UT) * planes * 64;              // forward chunks
  const int nbt = (H + a.NBT - 1) / a.NBT;
  const long long nb = static_cast<long long>(nbt) * 4 * planes * a.NBT;          // backward chunks
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx < nf) {
    const int n = static_cast<int>(idx % 64);
    const int p = static_cast<int>((idx / 64) % planes);
    const int tile = static_cast<int>(idx / (64LL * planes));
    const int col = (n >> 4) * H + UT * tile + (n & 15);
    float4 v;
    v.x = a.Wh[static_cast<size_t>(4 * p + 0) * 4 * H + col];
    v.y = a.Wh[static_cast<size_t>(4 * p + 1) * 4 * H + col];
    v.z = a.Wh[static_cast<size_t>(4 * p + 2) * 4 * H + col];
    v.w = a.Wh[static_cast<size_t>(4 * p + 3) * 4 * H + col];
    float4 hi, lo;
    tc::split4(v, hi, lo);
    const size_t half = static_cast<size_t>(planes) * plb(64);
    float* dst = a.whf + static_cast<size_t>(tile) * 2 * half + static_cast<size_t>(p) * plb(64) + n * 4;
    *reinterpret_cast<float4*>(dst) = hi;
    *reinterpret_cast<float4*>(dst + half) = lo;
  } else if (idx < nf + nb) {
    const long long i = idx - nf;
    const int jj = static_cast<int>(i % a.NBT);
    const int p = static_cast<int>((i / a.NBT) % planes);
    const int q = static_cast<int>((i / (static_cast<long long>(a.NBT) * planes)) % 4);
    const int tile = static_cast<int>(i / (static_cast<long long>(a.NBT) * planes * 4));
    const int row = a.NBT * tile + jj;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < H) v = *reinterpret_cast<const float4*>(a.Wh + static_cast<size_t>(row) * 4 * H + q * H + 4 * p);
    float4 hi, lo;
    tc::split4(v, hi, lo);
    const size_t half = static_cast<size_t>(planes) * plb(a.NBT);
    float* dst = a.whb + (static_cast<size_t>(tile) * 4 + q) * 2 * half + static_cast<size_t>(p) * plb(a.NBT) + jj * 4;
    *reinterpret_cast<float4*>(dst) = hi;
    *reinterpret_cast<float4*>(dst + half) = lo;
  }
}

// row-major carry -> operand planes of h (hi / lo) and state planes of c, per 128-row tile
struct CarryInArgs {
  const float* c; const float* h; int rows, H;
  float* hp_hi; float* hp_lo; long long hp_tile;       // [tile][H / 4][PLA]
  float* cp; long long cp_tile;                        // [tile][H / 4][PLC]
  float* cat_h; int ld_cat;                            // row-major copy of h for the weight gradients (cat[0][:, P:])
};

__global__ void __launch_bounds__(256) lstm_carry_in_kernel(const CarryInArgs a) {
  const int planes = a.H >> 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int r = static_cast<int>(idx & (RM - 1));
  const int p = static_cast<int>((idx >> 7) % planes);
  const int tile = static_cast<int>((idx >> 7) / planes);
  const int row = tile * RM + r;
  if (tile * RM >= a.rows) return;
  float4 hv = make_float4(0.f, 0.f, 0.f, 0.f), cv = hv;
  if (row < a.rows) {
    hv = *reinterpret_cast<const float4*>(a.h + static_cast<size_t>(row) * a.H + 4 * p);
    cv = *reinterpret_cast<const float4*>(a.c + static_cast<size_t>(row) * a.H + 4 * p);
    *reinterpret_cast<float4*>(a.cat_h + static_cast<size_t>(row) * a.ld_cat + 4 * p) = hv;
  }
  float4 hi, lo;
  tc::split4_fast(hv, hi, lo);
  const size_t o = static_cast<size_t>(tile) * a.hp_tile + static_cast<size_t>(p) * PLA + r * 4;
  *reinterpret_cast<float4*>(a.hp_hi + o) = hi;
  *reinterpret_cast<float4*>(a.hp_lo + o) = lo;
  *reinterpret_cast<float4*>(a.cp + static_cast<size_t>(tile) * a.cp_tile + static_cast<size_t>(p) * PLC + r * 4) = cv;
}

// carry out: state planes of c -> row-major; h from the row-major slot the last step wrote (cat[T][:, P:])
__global__ void __launch_bounds__(256) lstm_carry_out_kernel(const float* __restrict__ cp, long long cp_tile, float* __restrict__ c,
                                                            const float* __restrict__ cat_h, int ld_cat, float* __restrict__ h,
                                                            int rows, int H) {
  const int planes = H >> 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int r = static_cast<int>(idx & (RM - 1));
  const int p = static_cast<int>((idx >> 7) % planes);
  const int tile = static_cast<int>((idx >> 7) / planes);
  const int row = tile * RM + r;
  if (row >= rows) return;
  *reinterpret_cast<float4*>(c + static_cast<size_t>(row) * H + 4 * p) =
      *reinterpret_cast<const float4*>(cp + static_cast<size_t>(tile) * cp_tile + static_cast<size_t>(p) * PLC + r * 4);
  *reinterpret_cast<float4*>(h + static_cast<size_t>(row) * H + 4 * p) =
      *reinterpret_cast<const float4*>(cat_h + static_cast<size_t>(row) * ld_cat + 4 * p);
}

// element (row, col) of a matrix stored as planes over its whole row space: [row / 128][cols / 4][128][4]
__device__ __forceinline__ size_t plane_off(long long row, int col4, int cols4) {
  return (static_cast<size_t>(row >> 7) * cols4 + col4) * PLC + static_cast<size_t>(row & (RM - 1)) * 4;
}

// ------------------------------------------------------------------------------------------
// forward step:  a = gx_t + h_in Wh  (this CTA: 128 rows x 16 hidden units x 4 gates), gate math, reset-on-done
// ------------------------------------------------------------------------------------------
struct StepFwd2Args {
  const float* hp_hi; const float* hp_lo; long long hp_tile;      // A planes of this step, per row tile
  const float* whf; long long whf_tile;                          // B planes per unit tile (hi then lo)
  const float* gxp; long long r0; int gx_cols4;                  // gx as planes over the R row space; first row of the step
  float* cp; long long cp_tile;                                  // c state planes (in / out, reset applied)
  float* hn_hi; float* hn_lo;                                    // A planes of the NEXT step
  float* h_next_rm; int ld_hn;                                   // row-major carry handed on (cat[t + 1] + P)
  float* hn_rm;                                                  // row-major pre-reset h' [rows][H]
  float* gi; float* gf; float* gg; float* go; float* tcc; float* cin;   // activation cache planes [tile][H / 4][PLC] (all or none)
  const uint8_t* done; const int32_t* inds;
  const float* init_c; const float* init_h;                      // learned reset carry [H] each (nullable: zeros)
  int rows, H;
};

__global__ void __launch_bounds__(RT, 1) lstm_step_fwd2_kernel(const StepFwd2Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ StepBars bars;
  __shared__ uint32_t tmem_slot;
  uint32_t tmem_cols;
  const uint32_t tmem_base = step_init(&bars, &tmem_slot, 4 * UT, tmem_cols);
  const int tile = blockIdx.x, ut = blockIdx.y;
  const int H = a.H, planes = H >> 2;
  BulkOperands g;
  g.a_hi = a.hp_hi + static_cast<size_t>(tile) * a.hp_tile;
  g.a_lo = a.hp_lo + static_cast<size_t>(tile) * a.hp_tile;
  g.b_hi = a.whf + static_cast<size_t>(ut) * a.whf_tile;
  g.b_lo = g.b_hi + static_cast<size_t>(planes) * plb(4 * UT);
  g.K = H; g.N = 4 * UT;
  bulk_pipeline(g, smem, &bars, tmem_base);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = warp & 3, ub = warp >> 2;
  const int r = sub * 32 + lane;
  const int row = tile * RM + r;
  const int u0 = ut * UT + 8 * ub;                    // this thread: units u0 .. u0 + 7 of its row
  const int pl0 = u0 >> 2;                            // = planes pl0, pl0 + 1
  const bool ok = row < a.rows;
  // operands of the gate math that do not depend on the accumulator: fetched before the wait
  float4 gxv[4][2], cv[2];
  bool dn = false;
  const size_t so = static_cast<size_t>(tile) * a.cp_tile + static_cast<size_t>(pl0) * PLC + r * 4;
  if (ok) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const size_t o = plane_off(a.r0 + row, q * planes + pl0, a.gx_cols4);
      gxv[q][0] = *reinterpret_cast<const float4*>(a.gxp + o);
      gxv[q][1] = *reinterpret_cast<const float4*>(a.gxp + o + PLC);
    }
    cv[0] = *reinterpret_cast<const float4*>(a.cp + so);
    cv[1] = *reinterpret_cast<const float4*>(a.cp + so + PLC);
    if (a.done != nullptr) dn = a.done[a.inds ? a.inds[row] : row] != 0;
  }
  tc::mbar_wait(&bars.done, 0u);
  tc::tc_fence_after();
  float acc[4][8];
#pragma unroll
  for (int q = 0; q < 4; ++q)
    tmem_ld8(tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + static_cast<uint32_t>(q * UT + 8 * ub), acc[q]);
  if (ok) {
    float c2[8], h2[8], iv[8], fv[8], gv[8], ov[8], tv[8];
    const float gx0[4][8] = {
        {gxv[0][0].x, gxv[0][0].y, gxv[0][0].z, gxv[0][0].w, gxv[0][1].x, gxv[0][1].y, gxv[0][1].z, gxv[0][1].w},
        {gxv[1][0].x, gxv[1][0].y, gxv[1][0].z, gxv[1][0].w, gxv[1][1].x, gxv[1][1].y, gxv[1][1].z, gxv[1][1].w},
        {gxv[2][0].x, gxv[2][0].y, gxv[2][0].z, gxv[2][0].w, gxv[2][1].x, gxv[2][1].y, gxv[2][1].z, gxv[2][1].w},
        {gxv[3][0].x, gxv[3][0].y, gxv[3][0].z, gxv[3][0].w, gxv[3][1].x, gxv[3][1].y, gxv[3][1].z, gxv[3][1].w}};
    const float c0[8] = {cv[0].x, cv[0].y, cv[0].z, cv[0].w, cv[1].x, cv[1].y, cv[1].z, cv[1].w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      iv[k] = sigmoid_f(acc[0][k] + gx0[0][k]);
      fv[k] = sigmoid_f(acc[1][k] + gx0[1][k]);
      gv[k] = tanhf(acc[2][k] + gx0[2][k]);
      ov[k] = sigmoid_f(acc[3][k] + gx0[3][k]);
      c2[k] = __fadd_rn(__fmul_rn(fv[k], c0[k]), __fmul_rn(iv[k], gv[k]));
      tv[k] = tanhf(c2[k]);
      h2[k] = __fmul_rn(ov[k], tv[k]);
    }
    auto stp = [&](float* base, const float (&v)[8]) {              // two planes, lane = row: coalesced
      *reinterpret_cast<float4*>(base + so) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(base + so + PLC) = make_float4(v[4], v[5], v[6], v[7]);
    };
    if (a.gi != nullptr) {
      stp(a.gi, iv); stp(a.gf, fv); stp(a.gg, gv); stp(a.go, ov); stp(a.tcc, tv); stp(a.cin, c0);
    }
    if (a.hn_rm != nullptr) {
      float* d = a.hn_rm + static_cast<size_t>(row) * H + u0;
      *reinterpret_cast<float4*>(d) = make_float4(h2[0], h2[1], h2[2], h2[3]);
      *reinterpret_cast<float4*>(d + 4) = make_float4(h2[4], h2[5], h2[6], h2[7]);
    }
    if (dn) {                       // reset_state: zeros, or the learned initial carry (recurrent.py:154-157)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        c2[k] = a.init_c != nullptr ? a.init_c[u0 + k] : 0.0f;
        h2[k] = a.init_h != nullptr ? a.init_h[u0 + k] : 0.0f;
      }
    }
    stp(a.cp, c2);
    {
      float* d = a.h_next_rm + static_cast<size_t>(row) * a.ld_hn + u0;
      *reinterpret_cast<float4*>(d) = make_float4(h2[0], h2[1], h2[2], h2[3]);
      *reinterpret_cast<float4*>(d + 4) = make_float4(h2[4], h2[5], h2[6], h2[7]);
    }
    // the carry as the next step's A operand
    const size_t ho = static_cast<size_t>(tile) * a.hp_tile + static_cast<size_t>(pl0) * PLA + r * 4;
    float4 hi, lo;
    tc::split4_fast(make_float4(h2[0], h2[1], h2[2], h2[3]), hi, lo);
    *reinterpret_cast<float4*>(a.hn_hi + ho) = hi;
    *reinterpret_cast<float4*>(a.hn_lo + ho) = lo;
    tc::split4_fast(make_float4(h2[4], h2[5], h2[6], h2[7]), hi, lo);
    *reinterpret_cast<float4*>(a.hn_hi + ho + PLA) = hi;
    *reinterpret_cast<float4*>(a.hn_lo + ho + PLA) = lo;
  }
  step_fini(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------
// forward step for SMALL grids (a 512-row minibatch is 4 row tiles x 16 unit tiles = 64 CTAs on 148 SMs, and a
// step is a latency chain: stream 300 KB of operands at ~64 B/clk, 96 dependent MMAs, the gate epilogue).  A
// cluster of two CTAs shares one (row tile, unit tile): each streams and multiplies one HALF of K, the partial
// accumulators are exchanged through distributed shared memory, and each CTA runs the gate math for 8 of the
// 16 units - 4 units per thread instead of 8.  Sum order: k-low half + k-high half (fixed: deterministic).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_map(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

constexpr uint32_t FWD3_XB_BYTES = 2u * 2u * 4u * RM * 16u;      // [source CTA][unit quad][gate][row] float4

__global__ void __cluster_dims__(1, 1, 2) __launch_bounds__(RT, 1) lstm_step_fwd3_kernel(const StepFwd2Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ StepBars bars;
  __shared__ uint32_t tmem_slot;
  uint32_t tmem_cols;
  const uint32_t tmem_base = step_init(&bars, &tmem_slot, 4 * UT, tmem_cols);
  cluster_arrive();                                   // "my shared memory exists"; waited for before the exchange
  const int tile = blockIdx.x, ut = blockIdx.y;
  const uint32_t z = cluster_ctarank();               // which half of K, and which 8 units this CTA finalises
  const int H = a.H, planes = H >> 2, pp = planes >> 1;
  uint8_t* ring = smem + FWD3_XB_BYTES;
  BulkOperands g;
  g.a_hi = a.hp_hi + static_cast<size_t>(tile) * a.hp_tile + static_cast<size_t>(z) * pp * PLA;
  g.a_lo = a.hp_lo + static_cast<size_t>(tile) * a.hp_tile + static_cast<size_t>(z) * pp * PLA;
  g.b_hi = a.whf + static_cast<size_t>(ut) * a.whf_tile + static_cast<size_t>(z) * pp * plb(4 * UT);
  g.b_lo = g.b_hi + static_cast<size_t>(planes) * plb(4 * UT);
  g.K = H >> 1; g.N = 4 * UT;
  bulk_pipeline(g, ring, &bars, tmem_base);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = warp & 3, ub = warp >> 2;
  const int r = sub * 32 + lane;
  const int row = tile * RM + r;
  const int u0 = ut * UT + 8 * static_cast<int>(z) + 4 * ub;      // this thread finalises units u0 .. u0 + 3 of its row
  const int pl0 = u0 >> 2;
  const bool ok = row < a.rows;
  float4 gxv[4], cv = make_float4(0.f, 0.f, 0.f, 0.f);
  bool dn = false;
  const size_t so = static_cast<size_t>(tile) * a.cp_tile + static_cast<size_t>(pl0) * PLC + r * 4;
  if (ok) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      gxv[q] = *reinterpret_cast<const float4*>(a.gxp + plane_off(a.r0 + row, q * planes + pl0, a.gx_cols4));
    cv = *reinterpret_cast<const float4*>(a.cp + so);
    if (a.done != nullptr) dn = a.done[a.inds ? a.inds[row] : row] != 0;
  }
  tc::mbar_wait(&bars.done, 0u);
  tc::tc_fence_after();
  float acc[4][8];                                    // partial sums: gate q, local units 8 ub .. 8 ub + 7
#pragma unroll
  for (int q = 0; q < 4; ++q)
    tmem_ld8(tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + static_cast<uint32_t>(q * UT + 8 * ub), acc[q]);
  // exchange: units 8 ub .. belong to CTA `ub` of the pair; there, quad h (units 8 ub + 4 h ..) goes to the thread
  // with the same row in warp group h
  cluster_wait();
  {
    const uint32_t xb = tc::smem_u32(smem);
    const uint32_t dst_cta = static_cast<uint32_t>(ub);
    const uint32_t base = dst_cta == z ? xb : cluster_map(xb, dst_cta);
#pragma unroll
    for (int hq = 0; hq < 2; ++hq)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t off = (((z * 2u + hq) * 4u + q) * RM + r) * 16u;
        const float4 v = make_float4(acc[q][4 * hq], acc[q][4 * hq + 1], acc[q][4 * hq + 2], acc[q][4 * hq + 3]);
        if (dst_cta == z) *reinterpret_cast<float4*>(smem + off) = v;
        else st_cluster_f4(base + off, v);
      }
  }
  cluster_arrive();
  cluster_wait();
  if (ok) {
    float pre[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 p0 = *reinterpret_cast<const float4*>(smem + (((0u * 2u + ub) * 4u + q) * RM + r) * 16u);
      const float4 p1 = *reinterpret_cast<const float4*>(smem + (((1u * 2u + ub) * 4u + q) * RM + r) * 16u);
      pre[q][0] = __fadd_rn(__fadd_rn(p0.x, p1.x), gxv[q].x); pre[q][1] = __fadd_rn(__fadd_rn(p0.y, p1.y), gxv[q].y);
      pre[q][2] = __fadd_rn(__fadd_rn(p0.z, p1.z), gxv[q].z); pre[q][3] = __fadd_rn(__fadd_rn(p0.w, p1.w), gxv[q].w);
    }
    const float c0[4] = {cv.x, cv.y, cv.z, cv.w};
    float c2[4], h2[4], iv[4], fv[4], gv[4], ov[4], tv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      iv[k] = sigmoid_f(pre[0][k]);
      fv[k] = sigmoid_f(pre[1][k]);
      gv[k] = tanhf(pre[2][k]);
      ov[k] = sigmoid_f(pre[3][k]);
      c2[k] = __fadd_rn(__fmul_rn(fv[k], c0[k]), __fmul_rn(iv[k], gv[k]));
      tv[k] = tanhf(c2[k]);
      h2[k] = __fmul_rn(ov[k], tv[k]);
    }
    auto f4 = [](const float (&v)[4]) { return make_float4(v[0], v[1], v[2], v[3]); };
    if (a.gi != nullptr) {
      *reinterpret_cast<float4*>(a.gi + so) = f4(iv); *reinterpret_cast<float4*>(a.gf + so) = f4(fv);
      *reinterpret_cast<float4*>(a.gg + so) = f4(gv); *reinterpret_cast<float4*>(a.go + so) = f4(ov);
      *reinterpret_cast<float4*>(a.tcc + so) = f4(tv); *reinterpret_cast<float4*>(a.cin + so) = cv;
    }
    if (a.hn_rm != nullptr) *reinterpret_cast<float4*>(a.hn_rm + static_cast<size_t>(row) * H + u0) = f4(h2);
    if (dn) {                       // reset_state: zeros, or the learned initial carry
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        c2[k] = a.init_c != nullptr ? a.init_c[u0 + k] : 0.0f;
        h2[k] = a.init_h != nullptr ? a.init_h[u0 + k] : 0.0f;
      }
    }
    *reinterpret_cast<float4*>(a.cp + so) = f4(c2);
    *reinterpret_cast<float4*>(a.h_next_rm + static_cast<size_t>(row) * a.ld_hn + u0) = f4(h2);
    const size_t ho = static_cast<size_t>(tile) * a.hp_tile + static_cast<size_t>(pl0) * PLA + r * 4;
    float4 hi, lo;
    tc::split4_fast(f4(h2), hi, lo);
    *reinterpret_cast<float4*>(a.hn_hi + ho) = hi;
    *reinterpret_cast<float4*>(a.hn_lo + ho) = lo;
  }
  step_fini(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------
// backward step, element-wise part (oracle/recurrent.py ppo_loss_and_grads, BPTT loop), lane = row:
//   dh = (dY W2^T)_t + keep * sum_q dh_rec[q];  dc = dh o (1 - tc^2) + keep * dc_next;  da = gate gradients
// writes da row-major (batched GEMMs) and as the A operand planes of the split-K GEMM that follows
// ------------------------------------------------------------------------------------------
constexpr int BWD_MAX_SLICES = 8;            // 4 gate blocks x at most 2 k halves
struct StepBwd2Args {
  const float* dhp; long long r0; int dhp_cols4;   // dY W2^T as planes over the R row space; first row of the step
  const float* dhr; long long dhr_slice; int n_slices;   // split-K partial planes [slice][tile][H / 4][PLC] (nullable: last step)
  float* dcp;                                  // dc state planes [tile][H / 4][PLC] (in: dc_next, out)
  const float* gi; const float* gf; const float* gg; const float* go; const float* tcc; const float* cin;
  long long st_tile;                           // tile stride of the state plane layouts (H / 4 * PLC)
  const uint8_t* done; const int32_t* inds;
  float* da_rm;                                // [rows][4H]
  float* dap_hi; float* dap_lo; long long dap_tile; long long dap_slice;   // A planes [tile][slice q][H / 4][PLA]
  // learned initial carry (nullable): where this step's `done` is set, the gradient that would have flowed into the
  // carry handed on belongs to the initial-carry parameters instead - kept per step as planes [tile][H / 4][PLC]
  // (zeros elsewhere) and summed in fixed order by lstm_init_grad_kernel
  float* mdc; float* mdh;
  int dc_zero;                                 // last step of the sequence: the incoming dc is zero (dcp not read)
  int rows, H;
};

__global__ void __launch_bounds__(256) lstm_step_bwd2_kernel(const StepBwd2Args a) {
  const int planes = a.H >> 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int r = static_cast<int>(idx & (RM - 1));
  const int p = static_cast<int>((idx >> 7) % planes);
  const int tile = static_cast<int>((idx >> 7) / planes);
  if (tile * RM >= a.rows) return;
  const int row = tile * RM + r;
  const size_t so = static_cast<size_t>(tile) * a.st_tile + static_cast<size_t>(p) * PLC + r * 4;
  float4 di = make_float4(0.f, 0.f, 0.f, 0.f), df = di, dg = di, dO = di;
  if (row < a.rows) {
    const bool dn = a.done != nullptr && a.done[a.inds ? a.inds[row] : row] != 0;
    const float keep = dn ? 0.0f : 1.0f;
    const float4 dp = *reinterpret_cast<const float4*>(a.dhp + plane_off(a.r0 + row, p, a.dhp_cols4));
    auto ld = [&](const float* q) { return *reinterpret_cast<const float4*>(q + so); };
    const float4 i4 = ld(a.gi), f4 = ld(a.gf), g4 = ld(a.gg), o4 = ld(a.go), t4 = ld(a.tcc), c4 = ld(a.cin);
    const float4 dcn = a.dc_zero ? make_float4(0.f, 0.f, 0.f, 0.f) : ld(a.dcp);
    float dh[4] = {dp.x, dp.y, dp.z, dp.w};
    if (a.dhr != nullptr) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 pv[BWD_MAX_SLICES];                               // all partials in flight at once (one round trip)
#pragma unroll
      for (int q = 0; q < BWD_MAX_SLICES; ++q)
        pv[q] = q < a.n_slices ? *reinterpret_cast<const float4*>(a.dhr + q * a.dhr_slice + so) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < BWD_MAX_SLICES; ++q) {               // fixed order: deterministic
        s.x += pv[q].x; s.y += pv[q].y; s.z += pv[q].z; s.w += pv[q].w;
      }
      dh[0] += keep * s.x; dh[1] += keep * s.y; dh[2] += keep * s.z; dh[3] += keep * s.w;
      if (a.mdh != nullptr) *reinterpret_cast<float4*>(a.mdh + so) = dn ? s : make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (a.mdh != nullptr) {
      *reinterpret_cast<float4*>(a.mdh + so) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float iv[4] = {i4.x, i4.y, i4.z, i4.w}, fv[4] = {f4.x, f4.y, f4.z, f4.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w};
    const float ov[4] = {o4.x, o4.y, o4.z, o4.w}, tv[4] = {t4.x, t4.y, t4.z, t4.w}, cv[4] = {c4.x, c4.y, c4.z, c4.w};
    const float dcv[4] = {dcn.x, dcn.y, dcn.z, dcn.w};
    if (a.mdc != nullptr) *reinterpret_cast<float4*>(a.mdc + so) = dn ? dcn : make_float4(0.f, 0.f, 0.f, 0.f);
    float ri[4], rf[4], rg[4], ro[4], dco[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float dc = dh[j] * ov[j] * (1.0f - tv[j] * tv[j]) + keep * dcv[j];
      ri[j] = dc * gv[j] * iv[j] * (1.0f - iv[j]);
      rf[j] = dc * cv[j] * fv[j] * (1.0f - fv[j]);
      rg[j] = dc * iv[j] * (1.0f - gv[j] * gv[j]);
      ro[j] = dh[j] * tv[j] * ov[j] * (1.0f - ov[j]);
      dco[j] = dc * fv[j];
    }
    di = make_float4(ri[0], ri[1], ri[2], ri[3]); df = make_float4(rf[0], rf[1], rf[2], rf[3]);
    dg = make_float4(rg[0], rg[1], rg[2], rg[3]); dO = make_float4(ro[0], ro[1], ro[2], ro[3]);
    *reinterpret_cast<float4*>(a.dcp + so) = make_float4(dco[0], dco[1], dco[2], dco[3]);
    float* dr = a.da_rm + static_cast<size_t>(row) * 4 * a.H + 4 * p;
    *reinterpret_cast<float4*>(dr) = di;
    *reinterpret_cast<float4*>(dr + a.H) = df;
    *reinterpret_cast<float4*>(dr + 2 * a.H) = dg;
    *reinterpret_cast<float4*>(dr + 3 * a.H) = dO;
  }
  // A operand planes of dh_rec = da Wh^T: [tile][gate slice][plane p][row r]  (rows beyond the batch: zeros)
  const size_t po = static_cast<size_t>(tile) * a.dap_tile + static_cast<size_t>(p) * PLA + r * 4;
  const float4 v4[4] = {di, df, dg, dO};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 hi, lo;
    tc::split4_fast(v4[q], hi, lo);
    *reinterpret_cast<float4*>(a.dap_hi + po + q * a.dap_slice) = hi;
    *reinterpret_cast<float4*>(a.dap_lo + po + q * a.dap_slice) = lo;
  }
}

// ------------------------------------------------------------------------------------------
// backward step, split-K GEMM: dh_rec[q] = da_t[:, gate q] Wh[:, gate q]^T, output tile 128 rows x NBT units
// ------------------------------------------------------------------------------------------
struct StepBwdGemmArgs {
  const float* dap_hi; const float* dap_lo; long long dap_tile; long long dap_slice;
  const float* whb; long long whb_tile; long long whb_slice;     // [n tile][q][half][H / 4][plb(NBT)]
  float* dhr; long long dhr_slice; long long st_tile;            // partial planes [slice][tile][H / 4][PLC]
  int rows, H, NBT, KS;
};

__global__ void __launch_bounds__(RT, 1) lstm_step_bwd_gemm_kernel(const StepBwdGemmArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ StepBars bars;
  __shared__ uint32_t tmem_slot;
  uint32_t tmem_cols;
  const uint32_t tmem_base = step_init(&bars, &tmem_slot, a.NBT, tmem_cols);
  // blockIdx.z = gate block q x k part kp: the K = H columns of a gate block are cut into a.KS parts, so that a
  // small minibatch still spreads over the device (each part streams 1 / KS of both operands)
  const int tile = blockIdx.x, nt = blockIdx.y, q = blockIdx.z / a.KS, kp = blockIdx.z % a.KS;
  const int planes = a.H >> 2, pp = planes / a.KS;
  BulkOperands g;
  g.a_hi = a.dap_hi + static_cast<size_t>(tile) * a.dap_tile + static_cast<size_t>(q) * a.dap_slice + static_cast<size_t>(kp) * pp * PLA;
  g.a_lo = a.dap_lo + static_cast<size_t>(tile) * a.dap_tile + static_cast<size_t>(q) * a.dap_slice + static_cast<size_t>(kp) * pp * PLA;
  g.b_hi = a.whb + static_cast<size_t>(nt) * a.whb_tile + static_cast<size_t>(q) * a.whb_slice + static_cast<size_t>(kp) * pp * plb(a.NBT);
  g.b_lo = g.b_hi + static_cast<size_t>(planes) * plb(a.NBT);
  g.K = a.H / a.KS; g.N = a.NBT;
  bulk_pipeline(g, smem, &bars, tmem_base);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = warp & 3, cg = warp >> 2;
  const int r = sub * 32 + lane;
  tc::mbar_wait(&bars.done, 0u);
  tc::tc_fence_after();
  float* out = a.dhr + static_cast<size_t>(blockIdx.z) * a.dhr_slice + static_cast<size_t>(tile) * a.st_tile;
  for (int c = cg * 16; c < a.NBT; c += 32) {
    float v[16];
    tc::tmem_ld16(tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + static_cast<uint32_t>(c), v);
    const int u = nt * a.NBT + c;                     // first unit of this chunk
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (u + 4 * i < a.H)
        *reinterpret_cast<float4*>(out + static_cast<size_t>((u >> 2) + i) * PLC + r * 4) =
            make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  step_fini(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------
// gradient of the learned initial carry: column sums of the masked per-step planes, two fixed-order stages
//   stage 1: block (plane p, step t) sums its 4 units over every row of the step -> part[t][which][4p .. 4p+3]
//   stage 2: thread j sums part[.][which][j] over the steps
// ------------------------------------------------------------------------------------------
struct InitGradArgs {
  const float* mdc; const float* mdh; long long step_stride;   // [T][tiles][H / 4][PLC]
  long long st_tile; int tiles, rows, H, T;
  float* part;                                                // [T][2][H]
  float* g_c; float* g_h;                                     // [H] each (stage 2)
};

__global__ void __launch_bounds__(RM) lstm_init_grad_part_kernel(const InitGradArgs a) {
  __shared__ float4 red[2][RM / 32];
  const int p = blockIdx.x, t = blockIdx.y, r = threadIdx.x;
  float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sh = sc;
  for (int tile = 0; tile < a.tiles; ++tile) {
    if (tile * RM + r < a.rows) {
      const size_t o = static_cast<size_t>(t) * a.step_stride + static_cast<size_t>(tile) * a.st_tile + static_cast<size_t>(p) * PLC + r * 4;
      const float4 vc = *reinterpret_cast<const float4*>(a.mdc + o), vh = *reinterpret_cast<const float4*>(a.mdh + o);
      sc.x += vc.x; sc.y += vc.y; sc.z += vc.z; sc.w += vc.w;
      sh.x += vh.x; sh.y += vh.y; sh.z += vh.z; sh.w += vh.w;
    }
  }
  auto wsum = [](float4 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
      v.z += __shfl_xor_sync(0xffffffffu, v.z, o); v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
    }
    return v;
  };
  sc = wsum(sc); sh = wsum(sh);
  if ((r & 31) == 0) { red[0][r >> 5] = sc; red[1][r >> 5] = sh; }
  __syncthreads();
  if (r < 2) {
    float4 v = red[r][0];
    for (int w = 1; w < RM / 32; ++w) { v.x += red[r][w].x; v.y += red[r][w].y; v.z += red[r][w].z; v.w += red[r][w].w; }
    *reinterpret_cast<float4*>(a.part + (static_cast<size_t>(t) * 2 + r) * a.H + 4 * p) = v;
  }
}

__global__ void __launch_bounds__(256) lstm_init_grad_sum_kernel(const InitGradArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= 2 * a.H) return;
  const int which = j / a.H, u = j - which * a.H;
  float v = 0.0f;
  for (int t = 0; t < a.T; ++t) v += a.part[(static_cast<size_t>(t) * 2 + which) * a.H + u];
  (which == 0 ? a.g_c : a.g_h)[u] = v;
}

// ------------------------------------------------------------------------------------------
// The whole forward recurrence in ONE launch (replay: T > 1).  A CTA keeps its 16-unit slice of Wh resident in
// shared memory for all T steps and only streams the carry planes; the CTAs of a 128-row tile hand the carry
// on through global memory and a per-(tile, step) arrival counter (release: bar.sync, __threadfence, atomicAdd
// by one thread; acquire: one polling lane).  Needs every CTA resident at once (grid <= SMs; one CTA per SM
// because of the resident weights): the host falls back to one lstm_step_fwd2_kernel launch per step otherwise.
// What the launch saves per step: the kernel boundary, TMEM / barrier set-up and re-reading 133 KB of weights.
// ------------------------------------------------------------------------------------------
constexpr int PF_NS = 2;                                   // A ring depth (the weights take 133 KB of the 227)

struct SeqFwdPArgs {
  StepFwd2Args s;              // the arguments of step 0; the per-step pointers advance by the strides below
  int T;
  long long hp_buf;            // floats between the two carry-plane buffers
  long long cat_step, hn_step, cache_step, done_step;
  float* hp_base_hi; float* hp_base_lo;   // carry planes, buffer 0 (step t reads buffer t & 1, writes the other)
  int* flags;                  // [tiles][T + 1] zero-initialised arrival counters
  int n_ut;                    // CTAs per row tile (H / 16)
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(RT, 1) lstm_seq_fwd_persistent_kernel(const SeqFwdPArgs pa) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[PF_NS], empty[PF_NS], bdone, bfull;
  __shared__ uint32_t tmem_slot;
  const StepFwd2Args& a = pa.s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, ut = blockIdx.y;
  const int H = a.H, planes = H >> 2;
  const uint32_t pa_b = tc::plane_bytes(RM), pb_b = tc::plane_bytes(4 * UT);
  const uint32_t b_half_bytes = static_cast<uint32_t>(planes) * pb_b;
  const uint32_t a_slot_bytes = 2u * (RK / 4) * pa_b;
  uint8_t* b_smem = smem;                                   // [hi | lo][planes][plb(64)]
  uint8_t* a_ring = smem + 2u * b_half_bytes;
  const uint32_t tmem_cols = pow2_cols(4 * UT);
  if (warp == 0) tc::tmem_alloc(&tmem_slot, tmem_cols);
  if (threadIdx.x == 32) {
    for (int i = 0; i < PF_NS; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(&bdone, 1);
    tc::mbar_init(&bfull, 1);
    tc::mbar_init_fence();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int nst = (planes + RK / 4 - 1) / (RK / 4);
  if (warp == 1 && lane == 0) {                             // the weights: once
    const float* whf = a.whf + static_cast<size_t>(ut) * a.whf_tile;
    tc::mbar_arrive_expect_tx(&bfull, 2u * b_half_bytes);
    tc::bulk_g2s(b_smem, whf, b_half_bytes, &bfull);
    tc::bulk_g2s(b_smem + b_half_bytes, whf + static_cast<size_t>(planes) * plb(4 * UT), b_half_bytes, &bfull);
  }
  const int sub = warp & 3, ub = warp >> 2;
  const int r = sub * 32 + lane;
  const int row = tile * RM + r;
  const int u0 = ut * UT + 8 * ub;
  const int pl0 = u0 >> 2;
  const bool ok = row < a.rows;
  const size_t so = static_cast<size_t>(tile) * a.cp_tile + static_cast<size_t>(pl0) * PLC + r * 4;
  const size_t ho = static_cast<size_t>(tile) * a.hp_tile + static_cast<size_t>(pl0) * PLA + r * 4;
  int* flags = pa.flags + static_cast<size_t>(tile) * (pa.T + 1);
  uint32_t gstage = 0;                                      // ring position, counted over all steps (copier / issuer)
  for (int t = 0; t < pa.T; ++t) {
    const size_t cur = static_cast<size_t>(t & 1) * pa.hp_buf, nxt = static_cast<size_t>((t + 1) & 1) * pa.hp_buf;
    if (warp == 1) {
      if (lane == 0) {
        if (t > 0) {                                        // every CTA of this row tile has written its part of h_{t-1}
          unsigned long long spins = 0;
          while (ld_acquire_gpu(flags + t) < pa.n_ut) {
            if (++spins > (1ull << 26)) { printf("b200ppo: recurrent forward, tile %d step %d: peers missing\n", tile, t); __trap(); }
          }
          fence_proxy_async_all();
        }
        const float* ah = pa.hp_base_hi + cur + static_cast<size_t>(tile) * a.hp_tile;
        const float* al = pa.hp_base_lo + cur + static_cast<size_t>(tile) * a.hp_tile;
        for (int s = 0; s < nst; ++s) {
          const uint32_t gs = gstage + s;
          const int slot = gs % PF_NS;
          if (gs >= PF_NS) tc::mbar_wait(&empty[slot], (gs / PF_NS - 1) & 1u);
          const int p0 = s * (RK / 4);
          const int np = (planes - p0) < RK / 4 ? (planes - p0) : RK / 4;
          const uint32_t ab = static_cast<uint32_t>(np) * pa_b;
          uint8_t* st = a_ring + slot * a_slot_bytes;
          tc::mbar_arrive_expect_tx(&full[slot], 2u * ab);
          tc::bulk_g2s(st, ah + static_cast<size_t>(p0) * PLA, ab, &full[slot]);
          tc::bulk_g2s(st + (RK / 4) * pa_b, al + static_cast<size_t>(p0) * PLA, ab, &full[slot]);
        }
      }
      __syncwarp();
    } else if (warp == 0) {
      if (t == 0) tc::mbar_wait(&bfull, 0u);
      const uint64_t da0 = tc::make_desc(tc::smem_u32(a_ring), pa_b, 128);
      const uint64_t db0 = tc::make_desc(tc::smem_u32(b_smem), pb_b, 128);
      const uint32_t a_hiw = static_cast<uint32_t>(da0 >> 32), b_hiw = static_cast<uint32_t>(db0 >> 32);
      const uint32_t a_lo0 = static_cast<uint32_t>(da0), b_lo0 = static_cast<uint32_t>(db0);
      const uint32_t a_kstep = (2u * pa_b) >> 4, b_kstep = (2u * pb_b) >> 4;
      const uint32_t a_half = ((RK / 4) * pa_b) >> 4, b_half = b_half_bytes >> 4;
      const uint32_t idesc = tc::make_idesc_tf32(RM, 4 * UT);
      for (int s = 0; s < nst; ++s) {
        const uint32_t gs = gstage + s;
        const int slot = gs % PF_NS;
        const int p0 = s * (RK / 4);
        const int ks = ((planes - p0) < RK / 4 ? (planes - p0) : RK / 4) >> 1;
        tc::mbar_wait(&full[slot], (gs / PF_NS) & 1u);
        tc::tc_fence_after();
        const uint32_t a0 = a_lo0 + static_cast<uint32_t>(slot) * (a_slot_bytes >> 4);
        const uint32_t b0 = b_lo0 + static_cast<uint32_t>(p0) * (pb_b >> 4);     // resident weights: plane p0 of the slice
        if (tc::elect_one()) {
          for (int j = 0; j < ks; ++j) {
            const uint64_t ahd = (static_cast<uint64_t>(a_hiw) << 32) | (a0 + j * a_kstep);
            const uint64_t ald = (static_cast<uint64_t>(a_hiw) << 32) | (a0 + j * a_kstep + a_half);
            const uint64_t bhd = (static_cast<uint64_t>(b_hiw) << 32) | (b0 + j * b_kstep);
            const uint64_t bld = (static_cast<uint64_t>(b_hiw) << 32) | (b0 + j * b_kstep + b_half);
            const uint32_t acc0 = (s > 0 || j > 0) ? 1u : 0u;
            tc::mma_tf32(tmem_base, ald, bhd, idesc, acc0);
            tc::mma_tf32(tmem_base, ahd, bld, idesc, 1u);
            tc::mma_tf32(tmem_base, ahd, bhd, idesc, 1u);
          }
          tc::commit(&empty[slot]);
          if (s == nst - 1) tc::commit(&bdone);
        }
        __syncwarp();
      }
    }
    gstage += static_cast<uint32_t>(nst);
    // ---- epilogue of step t (all eight warps) ----
    float4 gxv[4][2], cv[2];
    bool dn = false;
    if (ok) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const size_t o = plane_off(a.r0 + static_cast<long long>(t) * a.rows + row, q * planes + pl0, a.gx_cols4);
        gxv[q][0] = *reinterpret_cast<const float4*>(a.gxp + o);
        gxv[q][1] = *reinterpret_cast<const float4*>(a.gxp + o + PLC);
      }
      cv[0] = *reinterpret_cast<const float4*>(a.cp + so);
      cv[1] = *reinterpret_cast<const float4*>(a.cp + so + PLC);
      if (a.done != nullptr) dn = a.done[static_cast<size_t>(t) * pa.done_step + (a.inds ? a.inds[row] : row)] != 0;
    }
    tc::mbar_wait(&bdone, static_cast<uint32_t>(t) & 1u);
    tc::tc_fence_after();
    float acc[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      tmem_ld8(tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + static_cast<uint32_t>(q * UT + 8 * ub), acc[q]);
    if (ok) {
      float c2[8], h2[8], iv[8], fv[8], gv[8], ov[8], tv[8];
      const float gx0[4][8] = {
          {gxv[0][0].x, gxv[0][0].y, gxv[0][0].z, gxv[0][0].w, gxv[0][1].x, gxv[0][1].y, gxv[0][1].z, gxv[0][1].w},
          {gxv[1][0].x, gxv[1][0].y, gxv[1][0].z, gxv[1][0].w, gxv[1][1].x, gxv[1][1].y, gxv[1][1].z, gxv[1][1].w},
          {gxv[2][0].x, gxv[2][0].y, gxv[2][0].z, gxv[2][0].w, gxv[2][1].x, gxv[2][1].y, gxv[2][1].z, gxv[2][1].w},
          {gxv[3][0].x, gxv[3][0].y, gxv[3][0].z, gxv[3][0].w, gxv[3][1].x, gxv[3][1].y, gxv[3][1].z, gxv[3][1].w}};
      const float c0[8] = {cv[0].x, cv[0].y, cv[0].z, cv[0].w, cv[1].x, cv[1].y, cv[1].z, cv[1].w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        iv[k] = sigmoid_f(acc[0][k] + gx0[0][k]);
        fv[k] = sigmoid_f(acc[1][k] + gx0[1][k]);
        gv[k] = tanhf(acc[2][k] + gx0[2][k]);
        ov[k] = sigmoid_f(acc[3][k] + gx0[3][k]);
        c2[k] = __fadd_rn(__fmul_rn(fv[k], c0[k]), __fmul_rn(iv[k], gv[k]));
        tv[k] = tanhf(c2[k]);
        h2[k] = __fmul_rn(ov[k], tv[k]);
      }
      const size_t co = static_cast<size_t>(t) * pa.cache_step + so;
      auto stp = [&](float* base, size_t off, const float (&v)[8]) {
        *reinterpret_cast<float4*>(base + off) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(base + off + PLC) = make_float4(v[4], v[5], v[6], v[7]);
      };
      if (a.gi != nullptr) {
        stp(a.gi, co, iv); stp(a.gf, co, fv); stp(a.gg, co, gv); stp(a.go, co, ov); stp(a.tcc, co, tv); stp(a.cin, co, c0);
      }
      if (a.hn_rm != nullptr) {
        float* d = a.hn_rm + static_cast<size_t>(t) * pa.hn_step + static_cast<size_t>(row) * H + u0;
        *reinterpret_cast<float4*>(d) = make_float4(h2[0], h2[1], h2[2], h2[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(h2[4], h2[5], h2[6], h2[7]);
      }
      if (dn) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          c2[k] = a.init_c != nullptr ? a.init_c[u0 + k] : 0.0f;
          h2[k] = a.init_h != nullptr ? a.init_h[u0 + k] : 0.0f;
        }
      }
      stp(a.cp, so, c2);
      {
        float* d = a.h_next_rm + static_cast<size_t>(t) * pa.cat_step + static_cast<size_t>(row) * a.ld_hn + u0;
        *reinterpret_cast<float4*>(d) = make_float4(h2[0], h2[1], h2[2], h2[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(h2[4], h2[5], h2[6], h2[7]);
      }
      // the carry as the next step's A operand
      float4 hi, lo;
      tc::split4_fast(make_float4(h2[0], h2[1], h2[2], h2[3]), hi, lo);
      *reinterpret_cast<float4*>(pa.hp_base_hi + nxt + ho) = hi;
      *reinterpret_cast<float4*>(pa.hp_base_lo + nxt + ho) = lo;
      tc::split4_fast(make_float4(h2[4], h2[5], h2[6], h2[7]), hi, lo);
      *reinterpret_cast<float4*>(pa.hp_base_hi + nxt + ho + PLA) = hi;
      *reinterpret_cast<float4*>(pa.hp_base_lo + nxt + ho + PLA) = lo;
    }
    // hand the carry on: our writes are ordered before the arrival (generic -> async proxy, CTA barrier, device fence)
    fence_proxy_async_all();
    tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(flags + t + 1, 1);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, tmem_cols);
}
