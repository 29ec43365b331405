// Tensor-core (tcgen05, kind::tf32) versions of the three GEMM kernels of the PPO update.
// Included by update.cu inside its anonymous namespace (needs Layout / FwdArgs / BwdArgs).
//
// Every GEMM is D[128 x Npad] (fp32, TMEM) += A[128 x K] * Bop[Npad x K]^T with both operands
// K-major in shared memory (layout: tc.cuh).  fp32 parity is kept by the error-compensated split
// x = hi + lo (both tf32): D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (3xTF32), measured at 2-4x the
// error of a cuBLAS fp32 GEMM (tests/test_gpu_tensorcore.py).  Weights are split once per update
// by upd_prep_w_kernel and reach shared memory with cp.async.bulk; activations / gradients are
// split while they are staged.  The weight-gradient kernel (dW = H^T D, reduction over the rows of two
// row-major matrices) instead takes both operands MN-major, straight from swizzled TMA boxes with the hi
// half of the split = the tile as it landed (see upd_bwd_dw_tc2_kernel and tc.cuh).
//
// Warp-specialised pipeline over a 4-deep shared-memory ring of 16-k stages (no __syncthreads in
// the main loop; all hand-offs are mbarriers):
//   warps 0-15  producers: global -> registers (3 stages ahead) -> hi/lo split -> st.shared ->
//               fence.proxy.async -> arrive(full_a[stage]); later the tcgen05.ld epilogue
//   warp 16     one thread waits full_a/full_b, issues 2 k-steps x 3 MMAs, tcgen05.commit ->
//               empty[stage]
//   warp 17     one thread streams the pre-split weight planes with cp.async.bulk -> full_b[stage]
// Measured on B200 (profiles/r1_tc_notes.md): a dependent tcgen05.mma (M=128, K=8, tf32) costs
// 47 / 48 / 65 / 128 cycles for N = 16 / 64 / 128 / 256; cp.async.bulk sustains ~65 B/clk/SM with
// ~450 cycles latency; issuing one 16-k stage (6 MMAs + commit) costs the issuer ~550 cycles.

constexpr int TCM = 128;
constexpr int TCK = 16;                 // k per stage: 2 MMA k-steps of 8
constexpr int TC_NS = 4;                // ring depth
constexpr int TC_NPROD = 512;           // producer / epilogue threads (16 warps)
constexpr int TCT = TC_NPROD + 64;      // + issuer warp + bulk-copy warp
constexpr int TC_MAXN = 256;
constexpr uint32_t TC_A_BYTES = (TCK / 4) * tc::plane_bytes(TCM);        // one half (hi or lo)
constexpr uint32_t TC_B_BYTES = (TCK / 4) * tc::plane_bytes(TC_MAXN);
constexpr uint32_t TC_STAGE_BYTES = 2u * TC_A_BYTES + 2u * TC_B_BYTES;
constexpr uint32_t TC_SMEM = TC_NS * TC_STAGE_BYTES;
constexpr int TC_NBARS = 3 * TC_NS + 1;

// phase timestamps of CTA 0 (bring-up / profiling aid; read back with b200ppo_debug_timestamps)
__device__ long long g_tc_stamp[128];
__device__ int g_tc_nstamp;
__device__ int g_tc_stamp_skip_dw;
__device__ long long g_tc_acc[8];     // scratch for ad-hoc clock64 probes
__device__ __forceinline__ void tc_stamp(int& n) {
  if (threadIdx.x == 0 && n >= 0 && n < 128) g_tc_stamp[n] = clock64();
  ++n;
}

#ifdef TC_PROBE
#define TC_PROBE_DECL long long pb_t0 = 0, pb_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const bool pb_on = cx.nstamp >= 0 && (threadIdx.x & 31) == 0;
#define TC_PROBE_START() do { if (pb_on) pb_t0 = clock64(); } while (0)
#define TC_PROBE_LAP(i) do { if (pb_on) { const long long t_ = clock64(); pb_acc[i] += t_ - pb_t0; pb_t0 = t_; } } while (0)
#define TC_PROBE_FLUSH(base, n) do { if (pb_on) for (int i_ = 0; i_ < (n); ++i_) g_tc_acc[(base) + i_] = pb_acc[i_]; } while (0)
#else
#define TC_PROBE_DECL
#define TC_PROBE_START() do { } while (0)
#define TC_PROBE_LAP(i) do { } while (0)
#define TC_PROBE_FLUSH(base, n) do { } while (0)
#endif

// wall-clock (globaltimer, ns) of every CTA of the most recent tensor-core update kernel: entry, set-up done
// (TMEM / barriers), exit — read back with b200ppo_debug_cta_times; three stores per CTA
constexpr int TC_MAX_CTA_T = 1024;
__device__ unsigned long long g_cta_gt[3 * TC_MAX_CTA_T];
__device__ __forceinline__ void tc_cta_time(int which) {
  const unsigned int cta = blockIdx.x + blockIdx.y * gridDim.x;
  if (threadIdx.x == 0 && cta < TC_MAX_CTA_T) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_cta_gt[3 * cta + which] = t;
  }
}

struct TcCtx {
  uint8_t* smem;
  uint64_t* bar_empty;   // [NS] stage free (tcgen05.commit of the MMAs that read it)
  uint64_t* bar_full_a;  // [NS] producers finished writing the stage (count = producer warps)
  uint64_t* bar_full_b;  // [NS] weight planes landed (cp.async.bulk complete_tx)
  uint64_t* bar_done;
  uint32_t tmem_base;
  uint32_t g;            // sequence number of the next stage use (ring position = g % NS)
  uint32_t done_uses;    // GEMMs finished so far (selects the bar_done phase and the TMEM half)
  uint32_t b_base;       // byte offset of the B stages (after all A stages)
  uint32_t b_half;       // bytes from the hi planes of a B stage to its lo planes
  uint32_t ebuf_off;     // byte offset of the epilogue transpose buffer
  uint32_t tmem_cols;    // allocated TMEM columns (2 x TC_MAXN: consecutive GEMMs ping-pong)
  uint32_t acc_col;      // accumulator column offset of the current GEMM
  bool chain_ok;         // the epilogue buffer is outside the ring: an epilogue may feed the next GEMM's A stages
  int nstamp;
};

__device__ __forceinline__ void tc_ctx_init(TcCtx& cx, uint8_t* smem, uint64_t* bars, uint32_t* tmem_slot,
                                            uint32_t tmem_cols = TC_MAXN) {
  const int warp = threadIdx.x >> 5;
  tc_cta_time(0);
  if (warp == 0) tc::tmem_alloc(tmem_slot, tmem_cols);
  if (threadIdx.x == 32) {
#pragma unroll
    for (int i = 0; i < TC_NS; ++i) {
      tc::mbar_init(&bars[i], 1);
      tc::mbar_init(&bars[TC_NS + i], TC_NPROD / 32);   // one arrive per producer warp
      tc::mbar_init(&bars[2 * TC_NS + i], 1);
    }
    tc::mbar_init(&bars[3 * TC_NS], 1);
    tc::mbar_init_fence();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  cx.smem = smem;
  cx.bar_empty = bars;
  cx.bar_full_a = bars + TC_NS;
  cx.bar_full_b = bars + 2 * TC_NS;
  cx.bar_done = bars + 3 * TC_NS;
  cx.tmem_base = *tmem_slot;
  cx.g = 0u;
  cx.done_uses = 0u;
  cx.b_base = TC_NS * 2u * TC_A_BYTES;
  cx.b_half = TC_B_BYTES;
  cx.ebuf_off = 0u;
  cx.tmem_cols = tmem_cols;
  cx.acc_col = 0u;
  cx.chain_ok = false;
  cx.nstamp = (static_cast<int>(blockIdx.x) == (g_tc_stamp_skip_dw >> 8) && blockIdx.y == 0) ? 0 : -100000;
  tc_stamp(cx.nstamp);
  tc_cta_time(1);
}

__device__ __forceinline__ void tc_ctx_fini(TcCtx& cx) {
  tc_stamp(cx.nstamp);
  if (threadIdx.x == 0 && cx.nstamp > 0) g_tc_nstamp = cx.nstamp;
  tc::tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc(cx.tmem_base, cx.tmem_cols);
  tc_cta_time(2);
}

struct TcStage {
  uint8_t *a_hi, *a_lo, *b_hi, *b_lo;
};
__device__ __forceinline__ TcStage tc_stage(const TcCtx& cx, int buf) {
  TcStage s;
  // all A stages first (the idle A area doubles as the epilogue transpose buffer), then all B stages
  s.a_hi = cx.smem + buf * 2u * TC_A_BYTES;
  s.a_lo = s.a_hi + TC_A_BYTES;
  s.b_hi = cx.smem + cx.b_base + buf * 2u * cx.b_half;
  s.b_lo = s.b_hi + cx.b_half;
  return s;
}

// wait until ring slot `buf` may be overwritten by its `use`-th user (use counts from 0)
__device__ __forceinline__ void tc_wait_empty(const TcCtx& cx, int buf, uint32_t use) {
  if (use > 0u) tc::mbar_wait(&cx.bar_empty[buf], (use - 1u) & 1u);
}

// Issue the MMAs of one staged block (TCK/8 k-steps) and commit.  The issuing thread is a single
// dependent instruction stream, so descriptors are not rebuilt per MMA (that cost ~640 cycles per
// stage): the constant high words and the slot-0 low words are computed once per GEMM, and a
// descriptor is one 32-bit add (the start-address field holds bytes >> 4 and cannot overflow for
// shared-memory addresses).
struct TcIssue {
  uint32_t a_lo0, b_lo0;        // low descriptor words of slot 0, k-step 0 (hi operand half)
  uint32_t a_hiw, b_hiw;        // constant high words (SBO, version)
  uint32_t a_kstep, b_kstep;    // low-word increments per k-step
  uint32_t b_half;              // low-word increment from the hi half to the lo half of B
  uint32_t b_slot;              // low-word increment per ring slot of B
  uint32_t idesc;
};
__device__ __forceinline__ TcIssue tc_issue_prepare(const TcCtx& cx, int npad) {
  const uint32_t pa = tc::plane_bytes(TCM), pb = tc::plane_bytes(npad);
  const TcStage st0 = tc_stage(cx, 0);
  const uint64_t da = tc::make_desc(tc::smem_u32(st0.a_hi), pa, 128);
  const uint64_t db = tc::make_desc(tc::smem_u32(st0.b_hi), pb, 128);
  TcIssue t;
  t.a_lo0 = static_cast<uint32_t>(da); t.a_hiw = static_cast<uint32_t>(da >> 32);
  t.b_lo0 = static_cast<uint32_t>(db); t.b_hiw = static_cast<uint32_t>(db >> 32);
  t.a_kstep = (2u * pa) >> 4;
  t.b_kstep = (2u * pb) >> 4;
  t.b_half = cx.b_half >> 4;
  t.b_slot = (2u * cx.b_half) >> 4;
  t.idesc = tc::make_idesc_tf32(TCM, npad);
  return t;
}
__device__ __forceinline__ uint64_t tc_desc(uint32_t hiw, uint32_t low) {
  return (static_cast<uint64_t>(hiw) << 32) | low;
}
// Called by the WHOLE (converged) issuer warp: every operand is warp-uniform, so the descriptors
// live in uniform registers and one elected lane issues.  (Issued from inside an `if (lane == 0)`
// region the compiler wraps every tcgen05.mma in an ELECT / R2UR.BROADCAST loop: ~10 dependent
// instructions per MMA, ~600 cycles per 16-k stage.)
__device__ __forceinline__ void tc_issue(TcCtx& cx, const TcIssue& t, int buf, int split, bool first, bool last) {
  const uint32_t a0 = t.a_lo0 + static_cast<uint32_t>(buf) * ((2u * TC_A_BYTES) >> 4);
  const uint32_t b0 = t.b_lo0 + static_cast<uint32_t>(buf) * t.b_slot;
  const uint32_t dcol = cx.tmem_base + cx.acc_col;
  tc::tc_fence_after();
  if (tc::elect_one()) {
#pragma unroll
    for (int j = 0; j < TCK / 8; ++j) {
      const uint64_t ah = tc_desc(t.a_hiw, a0 + j * t.a_kstep);
      const uint64_t bh = tc_desc(t.b_hiw, b0 + j * t.b_kstep);
      const uint32_t acc0 = (!first || j > 0) ? 1u : 0u;
      if (split) {
        const uint64_t al = tc_desc(t.a_hiw, a0 + j * t.a_kstep + (TC_A_BYTES >> 4));
        const uint64_t bl = tc_desc(t.b_hiw, b0 + j * t.b_kstep + t.b_half);
        tc::mma_tf32(dcol, al, bh, t.idesc, acc0);
        tc::mma_tf32(dcol, ah, bl, t.idesc, 1u);
        tc::mma_tf32(dcol, ah, bh, t.idesc, 1u);
      } else {
        tc::mma_tf32(dcol, ah, bh, t.idesc, acc0);
      }
    }
    tc::commit(&cx.bar_empty[buf]);
    if (last) tc::commit(cx.bar_done);
  }
  __syncwarp();
}

// Epilogue of the producer warps, in steps of 64 accumulator columns:
//   phase A  warp (sub, cg) reads its 32 rows x 16 columns from TMEM, `fa(r, c, v)` may modify v
//            (bias) and accumulate thread-private results, v goes to the smem transpose buffer
//            (the idle A area of the ring; row pitch 272 B = conflict-free for both phases);
//   phase B  all 512 threads stream the 128 x 64 block out: `fb(r, col, float4)` does the global
//            I/O with full 256-byte row segments per half-warp (the direct TMEM->global version
//            wrote 16 B per lane at a 1 KB stride: 9 K cycles of LSU time per 256-wide layer).
constexpr int TC_EP_PITCH = 68;   // floats
struct TcNoPre {
  __device__ __forceinline__ float4 operator()(int, int) const { return make_float4(0.f, 0.f, 0.f, 0.f); }
};
// `pre(r, col)` (optional) fetches a per-element global operand of phase B (e.g. z of the previous
// layer for act'); it is issued BEFORE the TMEM read and the barrier so that its latency is hidden
// (loading it inside phase B cost one memory round trip per element: 22 K cycles on a 256-wide layer).
// Chained output (CHAIN): phase A also converts the next GEMM's A operand (`h` from `fa`: act(z)
// in the forward chain, dpre in the backward chain) to hi / lo planes and stores it straight into
// the ring stages of the NEXT GEMM (sequence numbers g_next ..): a phase-A thread owns one row and
// 16 consecutive columns = the four 16-byte chunks of one 16-k stage, and consecutive lanes own
// consecutive rows (conflict-free).  The next GEMM therefore starts right after phase A, with no
// global-memory round trip and no producer pass, while phase B streams this layer's result out.
struct TcChainOut {
  uint32_t g_next;       // sequence number of the next GEMM's first stage
  int nst_next;          // its number of 16-k stages
  int split;
};
__device__ __forceinline__ void tc_wait_acc(TcCtx& cx);
// CHAIN = 0: plain epilogue.  CHAIN = 1: chained operand h[16] produced by `fa` in phase A (forward
// chain).  CHAIN = 2: chained operand returned by `fb` in phase B (backward chain: it needs the
// coalesced act'(z) operand of phase B; a phase-B thread always owns the same 4-column chunk
// f4 = tid & 15, i.e. one (stage, plane) of the 4 stages a 64-column step covers; the stores are
// 2-way bank conflicted).  `pre(row, col)` prefetches a phase-B operand before the accumulator wait.
template <int CHAIN, class FA, class PRE, class FB>
__device__ __forceinline__ void tc_epilogue(TcCtx& cx, int npad, const TcChainOut co, FA fa, PRE pre, FB fb) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = warp & 3, cg = warp >> 2;
  float* ebuf = reinterpret_cast<float*>(cx.smem + cx.ebuf_off);
  const int r = sub * 32 + lane;
  constexpr int PB = TCM * 16 / TC_NPROD;                 // phase-B float4 per thread per step = 4
  const uint32_t pa = tc::plane_bytes(TCM);
  for (int c0 = 0; c0 < npad; c0 += 64) {
    const int ncol4 = ((npad - c0) < 64 ? (npad - c0) : 64) >> 2;        // float4 per row in this step
    float4 pz[PB];
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int idx = tid + j * TC_NPROD;
      const int row = idx >> 4, f4 = idx & 15;
      pz[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f4 < ncol4) pz[j] = pre(row, c0 + 4 * f4);
    }
    if (c0 == 0) {
      tc_wait_acc(cx);
      tc_stamp(cx.nstamp);                                               // accumulator complete
    }
    const int c = c0 + cg * 16;
    if (c < npad) {
      float v[16], h[16];
      tc::tmem_ld16(cx.tmem_base + cx.acc_col + (static_cast<uint32_t>(sub * 32) << 16) + static_cast<uint32_t>(c), v);
      fa(r, c, v, h);
      float4* dst = reinterpret_cast<float4*>(ebuf + r * TC_EP_PITCH + cg * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      if (CHAIN == 1) {
        const int st = c >> 4;
        if (st < co.nst_next) {
          const uint32_t gs = co.g_next + static_cast<uint32_t>(st);
          const int slot = gs % TC_NS;
          tc_wait_empty(cx, slot, gs / TC_NS);
          uint8_t* cdst = cx.smem + slot * 2u * TC_A_BYTES + r * 16;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 hi, lo;
            tc::split4_fast(make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]), hi, lo);
            *reinterpret_cast<float4*>(cdst + q * pa) = hi;
            if (co.split) *reinterpret_cast<float4*>(cdst + TC_A_BYTES + q * pa) = lo;
          }
          tc::tc_fence_before();
          tc::fence_proxy_async();
          __syncwarp();
          // 4 of the 16 producer warps feed one stage: each arrives for 4
          if (lane == 0) tc::mbar_arrive_cnt(&cx.bar_full_a[slot], TC_NPROD / 32 / 4);
        }
      }
    }
#ifdef TC_PROBE
    tc_stamp(cx.nstamp);
#endif
    asm volatile("bar.sync 1, %0;" ::"n"(TC_NPROD) : "memory");
#ifdef TC_PROBE
    tc_stamp(cx.nstamp);
#endif
    const int f4t = tid & 15;
    const int cstage = (c0 >> 4) + (f4t >> 2);
    const bool cwrite = CHAIN == 2 && f4t < ncol4 && cstage < co.nst_next;
    uint8_t* cdst2 = nullptr;
    if (cwrite) {
      const uint32_t gs = co.g_next + static_cast<uint32_t>(cstage);
      const int slot = gs % TC_NS;
      tc_wait_empty(cx, slot, gs / TC_NS);
      cdst2 = cx.smem + slot * 2u * TC_A_BYTES + (f4t & 3) * pa;
    }
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int idx = tid + j * TC_NPROD;
      const int row = idx >> 4, f4 = idx & 15;
      if (f4 < ncol4) {
        const float4 nxt = fb(row, c0 + 4 * f4, *reinterpret_cast<const float4*>(ebuf + row * TC_EP_PITCH + 4 * f4), pz[j]);
        if (cwrite) {
          float4 hi, lo;
          tc::split4_fast(nxt, hi, lo);
          *reinterpret_cast<float4*>(cdst2 + row * 16) = hi;
          if (co.split) *reinterpret_cast<float4*>(cdst2 + TC_A_BYTES + row * 16) = lo;
        }
      }
    }
    if (CHAIN == 2) {
      tc::tc_fence_before();
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int st = (c0 >> 4) + q;
          if (4 * q < ncol4 && st < co.nst_next) tc::mbar_arrive(&cx.bar_full_a[(co.g_next + st) % TC_NS]);
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(TC_NPROD) : "memory");
  }
}

__device__ __forceinline__ float4 act4(float4 x, int act) {
  if (act == B200PPO_ACT_RELU) {
    x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
  } else if (act != B200PPO_ACT_NONE) {
    x.x = act_fwd(x.x, act); x.y = act_fwd(x.y, act); x.z = act_fwd(x.z, act); x.w = act_fwd(x.w, act);
  }
  return x;
}

// Main loop of a row-tile GEMM (no epilogue).  A[row][k] (row-major, leading dimension lda, rows
// row0..row0+127, rows >= nrows read as zero, activation `act` applied on load); B from the
// pre-split planes Bhi/Blo laid out [K16/4][npad + 1][4] in global memory (same padded plane
// stride as in shared memory, so a stage of B is ONE contiguous bulk copy per half).
__device__ __forceinline__ void tc_mainloop(TcCtx& cx, const float* __restrict__ A, int lda, int row0,
                                            int nrows, int K, int act, const float* __restrict__ Bhi,
                                            const float* __restrict__ Blo, int npad, int split,
                                            bool skip_a = false) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t pa = tc::plane_bytes(TCM), pb = tc::plane_bytes(npad);
  const int nst = (K + TCK - 1) / TCK;
  const uint32_t g0 = cx.g;
  // consecutive GEMMs alternate between the two halves of the TMEM allocation, so the MMAs of the
  // next GEMM never touch the accumulator the epilogue is still reading
  cx.acc_col = (cx.tmem_cols > TC_MAXN && (cx.done_uses & 1u)) ? TC_MAXN : 0u;
  if (warp < TC_NPROD / 32 && skip_a) {
    // the previous epilogue already produced this GEMM's A stages on chip
  } else if (warp < TC_NPROD / 32) {
    // ---------------- producers ----------------
    const int q = tid & 3, r = tid >> 2;                  // one 16-byte chunk per thread per stage
    const bool rv = row0 + r < nrows;
    const float* ap = A + static_cast<size_t>(rv ? row0 + r : 0) * lda + 4 * q;
    const uint32_t soff = q * pa + r * 16;
    const bool vec = (lda & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
    auto load_a = [&](int s) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      const int k = s * TCK + 4 * q;
      if (rv && s < nst) {
        if (vec && k + 3 < K) {
          x = *reinterpret_cast<const float4*>(ap + s * TCK);
        } else if (k < K) {
          const float* src = ap + s * TCK;
          x.x = src[0];
          if (k + 1 < K) x.y = src[1];
          if (k + 2 < K) x.z = src[2];
          if (k + 3 < K) x.w = src[3];
        }
      }
      return x;
    };
    // 4 stages in flight: a 64-wide layer is one memory round trip.  The loop is unrolled by the
    // prefetch depth so that every in-flight stage has its own statically named registers: a
    // rotating `a0 = a1; a1 = a2; ...` makes the register moves wait for the newest load, i.e. one
    // full memory round trip per stage (found in the ncu source view: the stall sat on the moves).
    float4 areg[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) areg[j] = load_a(j);
    for (int s0 = 0; s0 < nst; s0 += 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int s = s0 + j;
        if (s < nst) {
          const uint32_t gs = g0 + s;
          const int buf = gs % TC_NS;
          const TcStage st = tc_stage(cx, buf);
          tc_wait_empty(cx, buf, gs / TC_NS);
          float4 hi, lo;
          tc::split4_fast(act4(areg[j], act), hi, lo);
          areg[j] = load_a(s + 4);
          *reinterpret_cast<float4*>(st.a_hi + soff) = hi;
          if (split) *reinterpret_cast<float4*>(st.a_lo + soff) = lo;
          // 512 per-thread arrives on one mbarrier word serialise (~1-2 K cycles per stage, measured):
          // every lane fences its own writes, the warp converges, one lane arrives for the warp
          tc::fence_proxy_async();
          __syncwarp();
          if ((tid & 31) == 0) tc::mbar_arrive(&cx.bar_full_a[buf]);
        }
      }
    }
  } else if (warp == TC_NPROD / 32) {
    // ---------------- MMA issuer (whole warp, converged; see tc_issue) ----------------
    const TcIssue ti = tc_issue_prepare(cx, npad);
    for (int s = 0; s < nst; ++s) {
      const uint32_t gs = g0 + s;
      const int buf = gs % TC_NS;
      const uint32_t par = (gs / TC_NS) & 1u;
      tc::mbar_wait(&cx.bar_full_a[buf], par);
      tc::mbar_wait(&cx.bar_full_b[buf], par);
      tc_issue(cx, ti, buf, split, s == 0, s == nst - 1);
    }
  } else {
    // ------- weight (B operand) bulk-copy producer: never joins the per-layer barriers, so it
    // ------- runs ahead into the next layer as soon as ring slots free up
    if ((tid & 31) == 0) {
      const uint32_t bbytes = (TCK / 4) * pb;
      for (int s = 0; s < nst; ++s) {
        const uint32_t gs = g0 + s;
        const int buf = gs % TC_NS;
        tc_wait_empty(cx, buf, gs / TC_NS);
        const TcStage st = tc_stage(cx, buf);
        const size_t off = static_cast<size_t>(s) * (bbytes / 4);     // floats
        tc::mbar_arrive_expect_tx(&cx.bar_full_b[buf], split ? 2u * bbytes : bbytes);
        tc::bulk_g2s(st.b_hi, Bhi + off, bbytes, &cx.bar_full_b[buf]);
        if (split) tc::bulk_g2s(st.b_lo, Blo + off, bbytes, &cx.bar_full_b[buf]);
      }
    }
    __syncwarp();
  }
  cx.g = g0 + nst;
}

// wait for the accumulator (producer warps) / finish a GEMM (everybody but the weight-copy warp)
__device__ __forceinline__ void tc_wait_acc(TcCtx& cx) {
  tc::mbar_wait(cx.bar_done, cx.done_uses & 1u);
  tc::tc_fence_after();
}
__device__ __forceinline__ void tc_gemm_end(TcCtx& cx) {
  cx.done_uses++;
  // producers only: their global writes of this layer become visible to each other.  The issuer
  // does not take part: it is gated by the full_a barriers of the next GEMM and writes the other
  // TMEM half.
  if ((threadIdx.x >> 5) < TC_NPROD / 32) asm volatile("bar.sync 1, %0;" ::"n"(TC_NPROD) : "memory");
}

// ------------------------------------------------------------------------------------------
// weight pre-split: Wf (forward operand, Bop(n, k) = W[k][n]) and Wb (dX operand,
// Bop(kout, nred) = W[kout][nred]), each as hi / lo planes [red16/4][rows_pad + 1][4]
// ------------------------------------------------------------------------------------------
struct PrepArgs {
  b200ppo_plan plan;
  Layout L;
  const float* params;
  float* ws;
};

__global__ void __launch_bounds__(256) upd_prep_w_kernel(const PrepArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  // blockIdx.y = chain * MAXL + layer ; blockIdx.z = 0 (Wf) / 1 (Wb)
  const int chain = blockIdx.y / MAXL, l = blockIdx.y % MAXL;
  const b200ppo_chain& ch = chain == 0 ? a.plan.actor : a.plan.critic;
  if (l >= ch.n_layers) return;
  const TcLayer& t = chain == 0 ? a.L.tca[l] : a.L.tcc[l];
  const int K = ch.dims[l], N = ch.dims[l + 1];
  const float* W = a.params + ch.w_off[l];
  const bool fwd = blockIdx.z == 0;
  const int rows = fwd ? t.npad : t.kout_pad;          // operand rows
  const int red = fwd ? t.kpad : t.nred_pad;           // reduction length (padded to TCK)
  float4* hi = reinterpret_cast<float4*>(a.ws + (fwd ? t.wf_hi : t.wb_hi));
  float4* lo = reinterpret_cast<float4*>(a.ws + (fwd ? t.wf_lo : t.wb_lo));
  const int prow = rows + 1;                           // padded plane stride (float4), see tc.cuh
  const int total = (red / 4) * prow;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int row = idx % prow, q = idx / prow;
    float x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 4 * q + j;                          // reduction index
      float v = 0.0f;
      if (fwd) { if (r < K && row < N) v = W[static_cast<size_t>(r) * N + row]; }
      else { if (r < N && row < K) v = W[static_cast<size_t>(row) * N + r]; }
      x[j] = v;
    }
    float4 h, lw;
    tc::split4(make_float4(x[0], x[1], x[2], x[3]), h, lw);
    hi[idx] = h;
    lo[idx] = lw;
  }
}

// ------------------------------------------------------------------------------------------
// FWD (tensor cores)
// ------------------------------------------------------------------------------------------
struct TcFwdSmem {
  __align__(16) float bias[MAXL][TC_MAXN];   // all layers of the current chain, staged once
  __align__(16) float thin_w[TC_MAXN * 4];   // weights of a fused thin (<= 4 outputs) last layer
  __align__(16) float thin_part[4 * TCM * 4];   // [column group][row][output] partial dots
};

// hand-off between layers that do not run a GEMM (element-wise thin layers): producers only
__device__ __forceinline__ void tc_layer_sync() {
  if ((threadIdx.x >> 5) < TC_NPROD / 32) asm volatile("bar.sync 1, %0;" ::"n"(TC_NPROD) : "memory");
}

// Per-chain shared-memory layout: B slots sized for the chain's widest operand; when that leaves
// room for the epilogue transpose buffer outside the ring, epilogues may feed the next GEMM on chip.
// The weight streamer runs ahead of everybody else, so before the layout changes it waits for the
// last GEMM issued under the old layout.
constexpr uint32_t TC_EBUF_BYTES = TCM * TC_EP_PITCH * 4u;
__device__ __forceinline__ void tc_chain_setup(TcCtx& cx, const b200ppo_chain& ch, bool backward) {
  int maxn = 16;
  for (int l = 0; l < ch.n_layers; ++l) {
    const int n = backward ? ch.dims[l] : ch.dims[l + 1];
    if ((!backward || l >= 1) && n > maxn) maxn = n;
  }
  maxn = (maxn + 15) & ~15;
  const uint32_t bh = (TCK / 4) * tc::plane_bytes(maxn);
  if (bh != cx.b_half && (threadIdx.x >> 5) > TC_NPROD / 32 && cx.done_uses > 0u)
    tc::mbar_wait(cx.bar_done, (cx.done_uses - 1u) & 1u);
  cx.b_half = bh;
  const uint32_t ring_end = cx.b_base + TC_NS * 2u * bh;
  // The transpose buffer sits at the TOP of the allocation whatever the ring size: the weight streamer of the
  // next chain runs while the producers are still in the previous chain's last epilogue, so a buffer placed
  // at ring_end was overwritten by the next chain's (larger) weight stages (actor wider than the critic: wrong
  // actor outputs whenever the epilogue was slow enough to lose the race, i.e. tanh / swish).  A new ring
  // reaches the old buffer only when the new chain has no room for one; then the streamer waits for the
  // producers to leave the old epilogue.
  const bool ok = ring_end + TC_EBUF_BYTES <= TC_SMEM;
  if (cx.done_uses > 0u && cx.chain_ok && !ok && (threadIdx.x >> 5) != TC_NPROD / 32)
    asm volatile("bar.sync 2, %0;" ::"n"(TC_NPROD + 32) : "memory");
  cx.chain_ok = ok;
  cx.ebuf_off = ok ? TC_SMEM - TC_EBUF_BYTES : 0u;
}

__device__ __forceinline__ void tc_chain_forward(TcCtx& cx, const b200ppo_chain& ch, const TcLayer* tl,
                                                 const float* __restrict__ P, float* ws, const size_t* zoff,
                                                 size_t xhat_off, int row0, int nrows, int split, TcFwdSmem& sm,
                                                 bool chain_in0 = false) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const int L = ch.n_layers;
  // a thin last layer (the critic's 256 -> 1 head) is a per-row dot product: fused into the
  // epilogue of the layer before it instead of running a 16-column padded GEMM
  const bool fuse_thin = L >= 2 && ch.dims[L] <= 4 && ch.dims[L - 1] <= TC_MAXN;
  const int Lg = fuse_thin ? L - 1 : L;
  tc_chain_setup(cx, ch, false);
  if (tid < TC_NPROD) {
    asm volatile("bar.sync 1, %0;" ::"n"(TC_NPROD) : "memory");         // previous chain done with sm.bias
    for (int i = tid; i < Lg * TC_MAXN; i += TC_NPROD) {
      const int l = i / TC_MAXN, n = i - l * TC_MAXN;
      sm.bias[l][n] = n < ch.dims[l + 1] ? __ldg(P + ch.b_off[l] + n) : 0.0f;
    }
  }
  bool chain_in = chain_in0;             // this GEMM's A stages are already staged (previous epilogue / obs gather)
  for (int l = 0; l < Lg; ++l) {
    const int K = ch.dims[l], N = ch.dims[l + 1];
    const float* A = l == 0 ? ws + xhat_off : ws + zoff[l - 1];
    const int act_in = l == 0 ? B200PPO_ACT_NONE : ch.act;
    float* Z = ws + zoff[l];
    const int npad = tl[l].npad;
    const bool thin = fuse_thin && l == L - 2;
    const int NT = thin ? ch.dims[L] : 0;
    if (tid < TC_NPROD) {
      if (thin) {
        const float* Wt = P + ch.w_off[L - 1];
        for (int i = tid; i < N * NT; i += TC_NPROD) sm.thin_w[i] = __ldg(Wt + i);
      }
    }
    tc_mainloop(cx, A, K, row0, nrows, K, act_in, ws + tl[l].wf_hi, ws + tl[l].wf_lo, npad, split, chain_in);
    tc_stamp(cx.nstamp);                                                 // producer loop done
    const bool chain_out = cx.chain_ok && l + 1 < Lg;
    if (warp < TC_NPROD / 32) {
      asm volatile("bar.sync 1, %0;" ::"n"(TC_NPROD) : "memory");       // sm.bias / sm.thin_w visible
      float tp0 = 0.f, tp1 = 0.f, tp2 = 0.f, tp3 = 0.f;    // scalars: an array here ends up in local memory
      const int act = ch.act;
      const TcChainOut co{cx.g, npad / TCK, split};
      const float* lbias = sm.bias[l];
      auto fa =
          [&](int r, int c, float (&v)[16], float (&h)[16]) {
            {
              const float4* b4 = reinterpret_cast<const float4*>(lbias + c);       // c is a multiple of 16
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 b = b4[i];
                v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
              }
            }
            if (thin) {
              if (NT == 1 && c + 16 <= N && (act == B200PPO_ACT_RELU || act == B200PPO_ACT_NONE)) {
                // branch-free common case (one output, relu / identity): the per-element version
                // below compiles to a branch + dependent LDS -> FFMA per element (~220 cycles each)
                const float4* w4 = reinterpret_cast<const float4*>(sm.thin_w + c);
                const float4 w0 = w4[0], w1 = w4[1], w2 = w4[2], w3 = w4[3];
                const float ww[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                                      w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
                float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float h = act == B200PPO_ACT_RELU ? fmaxf(v[i], 0.0f) : v[i];
                  s4[i & 3] = fmaf(h, ww[i], s4[i & 3]);
                }
                tp0 += (s4[0] + s4[1]) + (s4[2] + s4[3]);
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  if (c + i < N) {
                    const float h = act_fwd(v[i], act);
                    const float* w = sm.thin_w + (c + i) * NT;
                    tp0 = fmaf(h, w[0], tp0);
                    if (NT > 1) tp1 = fmaf(h, w[1], tp1);
                    if (NT > 2) tp2 = fmaf(h, w[2], tp2);
                    if (NT > 3) tp3 = fmaf(h, w[3], tp3);
                  }
                }
              }
            }
            if (chain_out) {                               // next layer's A operand (padding columns: act(0) = 0)
              if (act == B200PPO_ACT_RELU) {
#pragma unroll
                for (int i = 0; i < 16; ++i) h[i] = fmaxf(v[i], 0.0f);
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) h[i] = act_fwd(v[i], act);
              }
            }
          };
      auto fb = [&](int r, int col, const float4 val, const float4) {
        const int row = row0 + r;
        if (row < nrows && col < N) {
          float* dst = Z + static_cast<size_t>(row) * N + col;
          if ((N & 3) == 0) {
            *reinterpret_cast<float4*>(dst) = val;
          } else {
            dst[0] = val.x;
            if (col + 1 < N) dst[1] = val.y;
            if (col + 2 < N) dst[2] = val.z;
            if (col + 3 < N) dst[3] = val.w;
          }
        }
        return val;
      };
      if (chain_out) tc_epilogue<1>(cx, npad, co, fa, TcNoPre(), fb);
      else tc_epilogue<0>(cx, npad, co, fa, TcNoPre(), fb);
      tc_stamp(cx.nstamp);                                               // epilogue body done
      if (thin) {
        const int r = (warp & 3) * 32 + (tid & 31), cg = warp >> 2;
        *reinterpret_cast<float4*>(&sm.thin_part[(cg * TCM + r) * 4]) = make_float4(tp0, tp1, tp2, tp3);
        asm volatile("bar.sync 1, %0;" ::"n"(TC_NPROD) : "memory");
        const float* bt = P + ch.b_off[L - 1];
        float* Zt = ws + zoff[L - 1];
        for (int i = tid; i < TCM * NT; i += TC_NPROD) {
          const int rr = i / NT, j = i - rr * NT;
          if (row0 + rr < nrows) {
            float acc = 0.0f;
            for (int g = 0; g < 4; ++g) acc += sm.thin_part[(g * TCM + rr) * 4 + j];
            Zt[static_cast<size_t>(row0 + rr) * NT + j] = acc + __ldg(bt + j);
          }
        }
      }
    }
    else cx.nstamp += 2;
    tc_gemm_end(cx);
    chain_in = chain_out;
    tc_stamp(cx.nstamp);                                                 // epilogue + hand-off done
  }
}

// chains: bit 0 = critic, bit 1 = actor.  The host launches the two chains as two kernels on forked
// streams (they are independent until the loss), so GAE overlaps the actor chain and the block
// scheduler back-fills SMs the other kernel leaves idle; both kernels write the same xhat values.
// tile_rows <= 128: rows of the minibatch per CTA.  The MMAs always run M = 128; a CTA just owns fewer
// live rows, so that the grid covers all SMs (16 384 rows / 128 = 128 CTAs would leave 20 of 148 SMs idle).
__global__ void __launch_bounds__(TCT, 1) upd_fwd_tc_kernel(const FwdArgs a, const int split, const int chains,
                                                            const int tile_rows) {
  extern __shared__ __align__(128) uint8_t tsmem[];
  __shared__ uint64_t bars[TC_NBARS];
  __shared__ uint32_t tmem_slot;
  __shared__ const float* rowsrc[TCM];
  __shared__ __align__(16) TcFwdSmem sm;
  pdl_launch_dependents();
  TcCtx cx;
  tc_ctx_init(cx, tsmem, bars, &tmem_slot, 2 * TC_MAXN);
  const int O = a.plan.obs_dim;
  pdl_wait();                  // TMEM / barriers are set up; everything below reads the previous kernels' results
  const int row0 = blockIdx.x * tile_rows;
  const int rend = row0 + tile_rows;
  const int R = a.L.R < rend ? a.L.R : rend, Rv = a.L.Rv < rend ? a.L.Rv : rend;   // clipped to this tile
  float* xhat = a.ws + a.L.xhat;
  for (int m = threadIdx.x; m < TCM; m += TCT) {
    const int r = row0 + m;
    const float* src = nullptr;
    if (r < R) {
      const int t = r / a.mb, j = r - t * a.mb;
      src = a.obs + (static_cast<size_t>(t) * a.B + a.inds[j]) * O;
    } else if (r >= a.L.R && r < Rv) {
      src = a.next_obs_last + static_cast<size_t>(a.inds[r - a.L.R]) * O;
    }
    rowsrc[m] = src;
  }
  __syncthreads();
  // gather + normalise this tile's observations (ppo.py:297 gather; normalizer.py:78-80)
  // For obs_dim <= 64 the normalised chunk also goes straight into the first critic GEMM's A stages
  // (hi / lo split, ring slots 0..nst-1: the ring is still empty), so that GEMM needs no producer pass
  // and no global round trip; xhat is still written for the actor's first layer and for dW.
  const bool fuse0 = (chains & 1) && (O & 3) == 0 && O <= TC_NS * TCK;
  if (fuse0) {
    if (threadIdx.x < TC_NPROD) {
      const int O4 = O >> 2;
      const int nst0 = (O + TCK - 1) / TCK, C4 = nst0 * 4;
      const uint32_t pa0 = tc::plane_bytes(TCM);
#pragma unroll 4
      for (int idx = threadIdx.x; idx < TCM * C4; idx += TC_NPROD) {
        const int m = idx / C4, c4 = idx - m * C4;
        const float* src = rowsrc[m];
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src != nullptr && c4 < O4) {
          x = *reinterpret_cast<const float4*>(src + 4 * c4);
          if (a.plan.normalize) {
            const float4 mu = __ldg(reinterpret_cast<const float4*>(a.mean) + c4);
            const float4 sd = __ldg(reinterpret_cast<const float4*>(a.std) + c4);
            x.x = __fdiv_rn(x.x - mu.x, sd.x); x.y = __fdiv_rn(x.y - mu.y, sd.y);
            x.z = __fdiv_rn(x.z - mu.z, sd.z); x.w = __fdiv_rn(x.w - mu.w, sd.w);
          }
          *reinterpret_cast<float4*>(xhat + static_cast<size_t>(row0 + m) * O + 4 * c4) = x;
        }
        float4 hi, lo;
        tc::split4_fast(x, hi, lo);
        uint8_t* dst = cx.smem + (c4 >> 2) * 2u * TC_A_BYTES + (c4 & 3) * pa0 + m * 16;
        *reinterpret_cast<float4*>(dst) = hi;
        if (split) *reinterpret_cast<float4*>(dst + TC_A_BYTES) = lo;
      }
      tc::fence_proxy_async();
      __syncwarp();
      if ((threadIdx.x & 31) == 0)
        for (int st = 0; st < nst0; ++st) tc::mbar_arrive(&cx.bar_full_a[st]);
    }
  } else if ((O & 3) == 0) {
    const int O4 = O >> 2;
#pragma unroll 4
    for (int idx = threadIdx.x; idx < TCM * O4; idx += TCT) {
      const int m = idx / O4, c4 = idx - m * O4;
      const float* src = rowsrc[m];
      if (src != nullptr) {
        float4 x = *reinterpret_cast<const float4*>(src + 4 * c4);
        if (a.plan.normalize) {
          const float4 mu = __ldg(reinterpret_cast<const float4*>(a.mean) + c4);
          const float4 sd = __ldg(reinterpret_cast<const float4*>(a.std) + c4);
          x.x = __fdiv_rn(x.x - mu.x, sd.x); x.y = __fdiv_rn(x.y - mu.y, sd.y);
          x.z = __fdiv_rn(x.z - mu.z, sd.z); x.w = __fdiv_rn(x.w - mu.w, sd.w);
        }
        *reinterpret_cast<float4*>(xhat + static_cast<size_t>(row0 + m) * O + 4 * c4) = x;
      }
    }
  } else {
    for (int idx = threadIdx.x; idx < TCM * O; idx += TCT) {
      const int m = idx / O, k = idx - m * O;
      if (rowsrc[m] != nullptr) {
        float x = rowsrc[m][k];
        if (a.plan.normalize) x = __fdiv_rn(x - __ldg(a.mean + k), __ldg(a.std + k));
        xhat[static_cast<size_t>(row0 + m) * O + k] = x;
      }
    }
  }
  __syncthreads();
  tc_stamp(cx.nstamp);
  if (chains & 1) tc_chain_forward(cx, a.plan.critic, a.L.tcc, a.params, a.ws, a.L.zc, a.L.xhat, row0, Rv, split, sm, fuse0);
  if ((chains & 2) && row0 < R) tc_chain_forward(cx, a.plan.actor, a.L.tca, a.params, a.ws, a.L.za, a.L.xhat, row0, R, split, sm);
  tc_ctx_fini(cx);
}

// ------------------------------------------------------------------------------------------
// BWD dX (tensor cores): dpre_{l-1} = (dpre_l W_l^T) ⊙ act'(z_{l-1})
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_chain_backward(TcCtx& cx, const b200ppo_chain& ch, const TcLayer* tl,
                                                  const float* __restrict__ P, float* ws, const size_t* zoff,
                                                  const size_t* doff, int row0, int nrows, int split) {
  const int tid = threadIdx.x, warp = tid >> 5;
  tc_chain_setup(cx, ch, true);
  bool chain_in = false;
  for (int l = ch.n_layers - 1; l >= 1; --l) {
    const int Kl = ch.dims[l], Nl = ch.dims[l + 1];
    const float* dY = ws + doff[l];
    const float* zprev = ws + zoff[l - 1];
    float* dprev = ws + doff[l - 1];
    const int act = ch.act;
    if (Nl <= 4) {
      // thin layer (value head): dX is a rank-Nl outer product, done element-wise, coalesced, with
      // all loads of a batch in flight together (a dependent-load loop here cost 80 K cycles)
      if (tid < TC_NPROD) {
        const float* W = P + ch.w_off[l];
        const int K4 = Kl >> 2;
        if ((Kl & 3) == 0 && K4 <= TC_NPROD && TC_NPROD % K4 == 0 && Nl == 1) {
          // one output (value head): a thread keeps its 4 columns' weights in registers and walks
          // down the rows, 8 independent 16-byte loads in flight
          const int k4 = tid % K4, r0 = tid / K4, rstep = TC_NPROD / K4;
          const float4 w = make_float4(__ldg(W + 4 * k4), __ldg(W + 4 * k4 + 1), __ldg(W + 4 * k4 + 2),
                                       __ldg(W + 4 * k4 + 3));                    // W is [Kl][1]
          constexpr int TB = 8;
          for (int rb = r0; rb < TCM; rb += rstep * TB) {
            float4 zz[TB];
            float dy[TB];
#pragma unroll
            for (int b = 0; b < TB; ++b) {
              const int rr = rb + b * rstep, row = row0 + rr;
              zz[b] = make_float4(0.f, 0.f, 0.f, 0.f);
              dy[b] = 0.f;
              if (rr < TCM && row < nrows) {
                zz[b] = *reinterpret_cast<const float4*>(zprev + static_cast<size_t>(row) * Kl + 4 * k4);
                dy[b] = dY[row];
              }
            }
#pragma unroll
            for (int b = 0; b < TB; ++b) {
              const int rr = rb + b * rstep, row = row0 + rr;
              if (rr < TCM && row < nrows) {
                float4 g;
                if (act == B200PPO_ACT_RELU) {
                  g = make_float4(zz[b].x > 0.f ? dy[b] * w.x : 0.f, zz[b].y > 0.f ? dy[b] * w.y : 0.f,
                                  zz[b].z > 0.f ? dy[b] * w.z : 0.f, zz[b].w > 0.f ? dy[b] * w.w : 0.f);
                } else {
                  g = make_float4(dy[b] * w.x * act_grad(zz[b].x, act), dy[b] * w.y * act_grad(zz[b].y, act),
                                  dy[b] * w.z * act_grad(zz[b].z, act), dy[b] * w.w * act_grad(zz[b].w, act));
                }
                *reinterpret_cast<float4*>(dprev + static_cast<size_t>(row) * Kl + 4 * k4) = g;
              }
            }
          }
        } else {
          for (int idx = tid; idx < TCM * Kl; idx += TC_NPROD) {
            const int rr = idx / Kl, k = idx - rr * Kl;
            const int row = row0 + rr;
            if (row < nrows) {
              float acc = 0.0f;
              for (int j = 0; j < Nl; ++j) acc = fmaf(dY[static_cast<size_t>(row) * Nl + j], __ldg(W + static_cast<size_t>(k) * Nl + j), acc);
              const size_t o = static_cast<size_t>(row) * Kl + k;
              dprev[o] = acc * act_grad(zprev[o], act);
            }
          }
        }
      }
      tc_layer_sync();
      tc_stamp(cx.nstamp);
      continue;
    }
    const int npad = tl[l].kout_pad;
    tc_mainloop(cx, dY, Nl, row0, nrows, Nl, B200PPO_ACT_NONE, ws + tl[l].wb_hi, ws + tl[l].wb_lo, npad, split,
                chain_in);
    tc_stamp(cx.nstamp);
    const bool kvec = (Kl & 3) == 0;
    const bool chain_out = cx.chain_ok && l >= 2 && kvec;  // the next GEMM (layer l-1) reduces over these Kl columns
    if (warp < TC_NPROD / 32) {
      const TcChainOut co{cx.g, npad / TCK, split};
      auto pre = [&](int r, int col) {
        const int row = row0 + r;
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kvec && row < nrows && col < Kl) z = *reinterpret_cast<const float4*>(zprev + static_cast<size_t>(row) * Kl + col);
        return z;
      };
      auto fa = [&](int, int, float (&)[16], float (&)[16]) {};
      auto fb = [&](int r, int col, const float4 val, const float4 z) {
        const int row = row0 + r;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < nrows && col < Kl) {
          const size_t o = static_cast<size_t>(row) * Kl + col;
          if (kvec) {
            if (act == B200PPO_ACT_RELU) {
              g = make_float4(z.x > 0.f ? val.x : 0.f, z.y > 0.f ? val.y : 0.f, z.z > 0.f ? val.z : 0.f,
                              z.w > 0.f ? val.w : 0.f);
            } else {
              g = make_float4(val.x * act_grad(z.x, act), val.y * act_grad(z.y, act),
                              val.z * act_grad(z.z, act), val.w * act_grad(z.w, act));
            }
            *reinterpret_cast<float4*>(dprev + o) = g;
          } else {
            const float vv[4] = {val.x, val.y, val.z, val.w};
            for (int i = 0; i < 4; ++i)
              if (col + i < Kl) dprev[o + i] = vv[i] * act_grad(zprev[o + i], act);
          }
        }
        return g;                                        // the next GEMM's A operand (zero outside the tile)
      };
      if (chain_out) tc_epilogue<2>(cx, npad, co, fa, pre, fb);
      else tc_epilogue<0>(cx, npad, co, fa, pre, fb);
    }
    else ++cx.nstamp;
    tc_gemm_end(cx);
    chain_in = chain_out;
    tc_stamp(cx.nstamp);
  }
}

__global__ void __launch_bounds__(TCT, 1) upd_bwd_dx_tc_kernel(const BwdArgs a, const int split, const int chains,
                                                               const int tile_rows) {
  extern __shared__ __align__(128) uint8_t tsmem[];
  __shared__ uint64_t bars[TC_NBARS];
  __shared__ uint32_t tmem_slot;
  pdl_launch_dependents();
  TcCtx cx;
  tc_ctx_init(cx, tsmem, bars, &tmem_slot, 2 * TC_MAXN);
  pdl_wait();
  const int row0 = blockIdx.x * tile_rows;
  tc_stamp(cx.nstamp);
  const int rend = (row0 + tile_rows) < a.L.R ? (row0 + tile_rows) : a.L.R;        // this tile's live rows
  if (chains & 1) tc_chain_backward(cx, a.plan.critic, a.L.tcc, a.params, a.ws, a.L.zc, a.L.dc, row0, rend, split);
  if (chains & 2) tc_chain_backward(cx, a.plan.actor, a.L.tca, a.params, a.ws, a.L.za, a.L.da, row0, rend, split);
  tc_ctx_fini(cx);
}

// ------------------------------------------------------------------------------------------
// BWD dW (tensor cores): dW_l[k][n] = sum_r act(z_{l-1})[r][k] * dpre_l[r][n]
// D rows = k (tiles of 128), D cols = n (padded to 16), reduction over this CTA's row range; both
// operands are transposed gathers (a 16-byte chunk = the same column of 4 consecutive data rows)
// staged by the producer warps.  The threads that stage column n of dpre also accumulate its sum:
// the bias gradient.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TCT, 1) upd_bwd_dw_tc_kernel(const BwdArgs a, const int split, const int item_base) {
  extern __shared__ __align__(128) uint8_t tsmem[];
  __shared__ uint64_t bars[TC_NBARS];
  __shared__ uint32_t tmem_slot;
  __shared__ float bred[TC_NPROD];
  pdl_launch_dependents();
  int item = blockIdx.x + item_base;   // items: actor M-tiles first, then critic (the host may launch the chains separately)
  const b200ppo_chain* ch = &a.plan.actor;
  const size_t* zoff = a.L.za;
  const size_t* doff = a.L.da;
  int layer = -1, mt = 0;
  for (int c = 0; c < 2 && layer < 0; ++c) {
    ch = c == 0 ? &a.plan.actor : &a.plan.critic;
    zoff = c == 0 ? a.L.za : a.L.zc;
    doff = c == 0 ? a.L.da : a.L.dc;
    for (int l = 0; l < ch->n_layers; ++l) {
      const int nm = (ch->dims[l] + TCM - 1) / TCM;
      if (item < nm) { layer = l; mt = item; break; }
      item -= nm;
    }
  }
  if (layer < 0) return;                                   // uniform per CTA
  if (static_cast<int>(blockIdx.y) >= a.L.tc_item_S[blockIdx.x + item_base]) return;   // this item has fewer row splits
  TcCtx cx;
  tc_ctx_init(cx, tsmem, bars, &tmem_slot);
  pdl_wait();
  if (g_tc_stamp_skip_dw & 1) cx.nstamp = -100000;
  const int K = ch->dims[layer], N = ch->dims[layer + 1];
  const int npad = (N + 15) & ~15;
  const int m0 = mt * TCM;
  const float* H = layer == 0 ? a.ws + a.L.xhat : a.ws + zoff[layer - 1];
  const int act_in = layer == 0 ? B200PPO_ACT_NONE : ch->act;
  const float* D = a.ws + doff[layer];
  const int sp = blockIdx.y;
  const int item_id = blockIdx.x + item_base;
  const int Si = a.L.tc_item_S[item_id];                   // this item's number of row splits (<= gridDim.y)
  const int rps = a.L.tc_item_rps[item_id];
  const int r_begin = sp * rps;
  int r_end = r_begin + rps;
  if (r_end > a.L.R) r_end = a.L.R;
  // partial-gradient slots this item leaves unused (sp + Si, sp + 2 Si, ... < tc_S) are zeroed here,
  // so the fixed-order reduction over all tc_S slots stays valid whatever ran on this workspace before
  const int S_all = a.L.tc_S;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t pa = tc::plane_bytes(TCM), pb = tc::plane_bytes(npad);
  const int nst = (r_end - r_begin + TCK - 1) / TCK;
  float bsum = 0.0f;
  if (warp < TC_NPROD / 32) {
    // A: thread -> (m = tid & 127, plane q = tid >> 7); B: thread -> (column n = tid & 255, planes 2h, 2h+1)
    const int am = tid & (TCM - 1), aq = tid >> 7;
    const bool av = m0 + am < K;
    const float* ap = H + static_cast<size_t>(r_begin + 4 * aq) * K + (av ? m0 + am : 0);
    const uint32_t aoff = aq * pa + am * 16;
    const int bn = tid & 255, bh = tid >> 8;
    const bool bv = bn < N, bstage = bn < npad;
    const float* bp = D + static_cast<size_t>(r_begin + 8 * bh) * N + (bv ? bn : 0);
    struct Regs { float4 a, b0, b1; };
    auto load_stage = [&](int s) {
      Regs x;
      x.a = x.b0 = x.b1 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s >= nst) return x;
      const int rs = r_begin + s * TCK;
      const size_t so = static_cast<size_t>(s) * TCK;
      if (rs + TCK <= r_end) {                                  // full stage: no row bounds checks
        if (av) {
          const float* src = ap + so * K;
          x.a = make_float4(src[0], src[K], src[2 * static_cast<size_t>(K)], src[3 * static_cast<size_t>(K)]);
        }
        if (bv) {
          const float* src = bp + so * N;
          x.b0 = make_float4(src[0], src[N], src[2 * static_cast<size_t>(N)], src[3 * static_cast<size_t>(N)]);
          src += 4 * static_cast<size_t>(N);
          x.b1 = make_float4(src[0], src[N], src[2 * static_cast<size_t>(N)], src[3 * static_cast<size_t>(N)]);
        }
      } else {
        if (av) {
          const float* src = ap + so * K;
          const int r = rs + 4 * aq;
          if (r + 0 < r_end) x.a.x = src[0];
          if (r + 1 < r_end) x.a.y = src[K];
          if (r + 2 < r_end) x.a.z = src[2 * static_cast<size_t>(K)];
          if (r + 3 < r_end) x.a.w = src[3 * static_cast<size_t>(K)];
        }
        if (bv) {
          const float* src = bp + so * N;
          int r = rs + 8 * bh;
          if (r + 0 < r_end) x.b0.x = src[0];
          if (r + 1 < r_end) x.b0.y = src[N];
          if (r + 2 < r_end) x.b0.z = src[2 * static_cast<size_t>(N)];
          if (r + 3 < r_end) x.b0.w = src[3 * static_cast<size_t>(N)];
          src += 4 * static_cast<size_t>(N);
          r += 4;
          if (r + 0 < r_end) x.b1.x = src[0];
          if (r + 1 < r_end) x.b1.y = src[N];
          if (r + 2 < r_end) x.b1.z = src[2 * static_cast<size_t>(N)];
          if (r + 3 < r_end) x.b1.w = src[3 * static_cast<size_t>(N)];
        }
      }
      return x;
    };
    // statically named prefetch ring (see tc_mainloop): DW_PF stages of loads in flight
    constexpr int DW_PF = 4;
    Regs x[DW_PF];
#pragma unroll
    for (int j = 0; j < DW_PF; ++j) x[j] = load_stage(j);
    for (int s0 = 0; s0 < nst; s0 += DW_PF) {
#pragma unroll
      for (int j = 0; j < DW_PF; ++j) {
        const int s = s0 + j;
        if (s < nst) {
          const uint32_t gs = cx.g + s;
          const int buf = gs % TC_NS;
          const TcStage st = tc_stage(cx, buf);
          tc_wait_empty(cx, buf, gs / TC_NS);
          float4 hi, lo;
          tc::split4(act4(x[j].a, act_in), hi, lo);             // act(0) == 0 for relu / tanh / swish
          *reinterpret_cast<float4*>(st.a_hi + aoff) = hi;
          if (split) *reinterpret_cast<float4*>(st.a_lo + aoff) = lo;
          if (bstage) {
            bsum += ((x[j].b0.x + x[j].b0.y) + (x[j].b0.z + x[j].b0.w)) + ((x[j].b1.x + x[j].b1.y) + (x[j].b1.z + x[j].b1.w));
            tc::split4(x[j].b0, hi, lo);
            uint32_t o = (2 * bh) * pb + bn * 16;
            *reinterpret_cast<float4*>(st.b_hi + o) = hi;
            if (split) *reinterpret_cast<float4*>(st.b_lo + o) = lo;
            tc::split4(x[j].b1, hi, lo);
            o += pb;
            *reinterpret_cast<float4*>(st.b_hi + o) = hi;
            if (split) *reinterpret_cast<float4*>(st.b_lo + o) = lo;
          }
          x[j] = load_stage(s + DW_PF);
          tc::fence_proxy_async();
          __syncwarp();
          if ((tid & 31) == 0) tc::mbar_arrive(&cx.bar_full_a[buf]);
        }
      }
    }
  } else if (warp == TC_NPROD / 32) {
    const TcIssue ti = tc_issue_prepare(cx, npad);
    for (int s = 0; s < nst; ++s) {
      const uint32_t gs = cx.g + s;
      const int buf = gs % TC_NS;
      tc::mbar_wait(&cx.bar_full_a[buf], (gs / TC_NS) & 1u);
      tc_issue(cx, ti, buf, split, s == 0, s == nst - 1);
    }
  }
  cx.g += nst;
  float* gpart = a.ws + a.L.gpart + static_cast<size_t>(sp) * a.plan.n_params;
  float* gp = gpart + ch->w_off[layer];
  tc_stamp(cx.nstamp);
  if (warp < TC_NPROD / 32 && nst > 0) {
    tc_epilogue<0>(
        cx, npad, TcChainOut{0u, 0, 0}, [&](int, int, float (&)[16], float (&)[16]) {}, TcNoPre(),
        [&](int r, int col, const float4 val, const float4) {
          const int k = m0 + r;
          if (k < K && col < N) {
            float* dst = gp + static_cast<size_t>(k) * N + col;
            if ((N & 3) == 0) {
              *reinterpret_cast<float4*>(dst) = val;
            } else {
              dst[0] = val.x;
              if (col + 1 < N) dst[1] = val.y;
              if (col + 2 < N) dst[2] = val.z;
              if (col + 3 < N) dst[3] = val.w;
            }
            for (int s2 = sp + Si; s2 < S_all; s2 += Si) {
              float* dz = dst + static_cast<size_t>(s2 - sp) * a.plan.n_params;
              for (int i = 0; i < 4; ++i)
                if (col + i < N) dz[i] = 0.0f;
            }
          }
          return val;
        });
  }
  cx.done_uses++;
  if (mt == 0) {                                           // bias gradient: fixed-order column sums
    if (tid < TC_NPROD) bred[tid] = bsum;
    __syncthreads();
    if (tid < N && tid < 256) {
      gpart[ch->b_off[layer] + tid] = bred[tid] + bred[tid + 256];
      for (int s2 = sp + Si; s2 < S_all; s2 += Si)
        gpart[static_cast<size_t>(s2 - sp) * a.plan.n_params + ch->b_off[layer] + tid] = 0.0f;
    }
  }
  tc_ctx_fini(cx);
}

// ------------------------------------------------------------------------------------------
// BWD dW, v2: the 16-row blocks of both operands are CONTIGUOUS in global memory (row-major
// [rows][K] / [rows][N]), so one thread streams them raw into a shared-memory ring with
// cp.async.bulk and the producer warps transpose shared -> shared (conflict-free LDS.32 down the
// rows, hi/lo split, STS.128 into the UMMA planes).  The v1 kernel gathered the transposed chunks
// straight from global memory: ~200 instructions per warp and stage, mostly 64-bit address and
// predicate arithmetic, and the SM's issue slots were the bottleneck (ncu: issue active 49 % of the
// whole launch, tensor pipe 16 %; profiles/r1_tc_notes.md).
// Needs K <= 128 or K % 4 == 0 for every layer (bulk copies are 16-byte granular); otherwise the
// host keeps v1.
// ------------------------------------------------------------------------------------------
constexpr int DW2_NS = 3;                                    // UMMA stage ring
constexpr int DW2_NR = 2;                                    // raw ring: slots of the widest item (sizes the area)
constexpr int DW2_NR_MAX = 8;                                // narrower items cut the same area into more, smaller slots
constexpr int DW2_MAXK = 256;                                // widest H the raw ring holds (full rows)
constexpr uint32_t DW2_RAW_A = TCK * DW2_MAXK * 4u;          // 16 full rows of H
constexpr uint32_t DW2_RAW_B = TCK * TC_MAXN * 4u;           // 16 rows x <= 256 columns
constexpr uint32_t DW2_RAW_BYTES = DW2_RAW_A + DW2_RAW_B;
constexpr uint32_t DW2_SMEM = DW2_NS * TC_STAGE_BYTES + DW2_NR * DW2_RAW_BYTES;

// Observation layers wider than DW2_MAXK (layer 0 of a chain: H = xhat [rows][obs_dim]; dict-observation plans,
// BASELINE configs[3]) cannot stage full rows: their CTA's 16-row x 128-column block is fetched by ONE 2-D tiled TMA
// load (`tm_xhat`: tensor map over xhat, box 128 x 16, zero fill beyond obs_dim / the rows) and lands densely.
// dW v3 (MN-major operands, tc.cuh): both operands of dW = H^T D already have the reduction index (the row) as their
// slow dimension in global memory, which IS the MN-major operand layout — a {32 columns, 16 rows} TMA box with the
// 32-byte-atom 128-byte swizzle lands as a ready UMMA operand, no transposition.  The hi half of the 3xTF32 split is
// the tile as it landed (the tensor core reads the upper 19 bits of an fp32 word: hi = trunc(x)); the producer warps
// only apply the activation in place and write lo = x - trunc(x) at the same offsets of a second tile: ~15 instead of
// ~100 instructions per thread and 16-row stage (clock64 probes of v2, profiles/r2c_notes.md: the transposing warps
// were busy 630 of the 1 140 cycles of a stage and the issuer waited for them 57 % of the time).  Layers whose widths
// are not multiples of 4 (TMA row pitch; the value head's [rows][1] gradient) keep the v2 path below, in the same
// launch.  One tensor map per layer and operand, indexed actor layers first.
struct DwMaps {
  CUtensorMap h[2 * MAXL];
  CUtensorMap d[2 * MAXL];
};
constexpr int DW3_NS = 4;                                     // stage ring (A hi | A lo | B hi | B lo, TMA-fed): slots of the widest item
constexpr int DW3_NG = 4;                                     // producer groups (4 warps each): group g owns stages s = g (mod 4)
constexpr int DW3_NS_MAX = 8;                                 // narrower items cut the same area into more slots (latency-bound stream)
constexpr uint32_t DW3_A_BYTES = 4u * tc::MN_ATOM_BYTES;      // 128 columns of H x 16 rows
__host__ __device__ constexpr uint32_t dw3_stage_bytes(int n32) {
  return 2u * DW3_A_BYTES + 2u * static_cast<uint32_t>(n32) * tc::MN_ATOM_BYTES;
}
static_assert(DW3_NS * dw3_stage_bytes(8) + 1024u <= DW2_SMEM, "v3 ring must fit the launch's shared memory");

__global__ void __launch_bounds__(TCT, 1) upd_bwd_dw_tc2_kernel(const BwdArgs a, const int split, const int item_base,
                                                                const __grid_constant__ CUtensorMap tm_xhat,
                                                                const int nr_max, const __grid_constant__ DwMaps maps,
                                                                const uint32_t mn_mask) {
  extern __shared__ __align__(128) uint8_t tsmem[];
  __shared__ uint64_t bars[TC_NBARS];
  __shared__ uint64_t rbar[2 * DW2_NR_MAX];                  // [0..NRMAX) raw full, [NRMAX..2 NRMAX) raw empty
  __shared__ uint64_t bar3[3 * DW3_NS_MAX];                  // v3 ring: TMA landed | lo halves written | MMAs retired
  __shared__ uint32_t tmem_slot;
  __shared__ float bred[2][TC_NPROD];
  pdl_launch_dependents();
  int item = blockIdx.x + item_base;   // items: actor M-tiles first, then critic (the host may launch the chains separately)
  const b200ppo_chain* ch = &a.plan.actor;
  const size_t* zoff = a.L.za;
  const size_t* doff = a.L.da;
  int layer = -1, mt = 0, glayer = 0;                       // glayer: index into `maps` / bit of `mn_mask`
  for (int c = 0; c < 2 && layer < 0; ++c) {
    ch = c == 0 ? &a.plan.actor : &a.plan.critic;
    zoff = c == 0 ? a.L.za : a.L.zc;
    doff = c == 0 ? a.L.da : a.L.dc;
    for (int l = 0; l < ch->n_layers; ++l) {
      const int nm = (ch->dims[l] + TCM - 1) / TCM;
      if (item < nm) { layer = l; mt = item; glayer = c * MAXL + l; break; }
      item -= nm;
    }
  }
  if (layer < 0) return;                                   // uniform per CTA
  const bool use_mn = (mn_mask >> glayer) & 1u;
  if (static_cast<int>(blockIdx.y) >= a.L.tc_item_S[blockIdx.x + item_base]) return;   // this item has fewer row splits
  if (threadIdx.x == 64) {
#pragma unroll
    for (int i = 0; i < DW2_NR_MAX; ++i) {
      tc::mbar_init(&rbar[i], 1);
      tc::mbar_init(&rbar[DW2_NR_MAX + i], TC_NPROD / 32);
    }
#pragma unroll
    for (int i = 0; i < DW3_NS_MAX; ++i) {
      tc::mbar_init(&bar3[i], 1);
      tc::mbar_init(&bar3[DW3_NS_MAX + i], TC_NPROD / 32 / DW3_NG);
      tc::mbar_init(&bar3[2 * DW3_NS_MAX + i], 1);
    }
    tc::mbar_init_fence();
  }
  TcCtx cx;
  tc_ctx_init(cx, tsmem, bars, &tmem_slot);
  pdl_wait();
  cx.b_base = DW2_NS * 2u * TC_A_BYTES;
  if (g_tc_stamp_skip_dw & 1) cx.nstamp = -100000;
  const int K = ch->dims[layer], N = ch->dims[layer + 1];
  const int npad = (N + 15) & ~15;
  const int m0 = mt * TCM;
  const int kw = (K - m0) < TCM ? (K - m0) : TCM;          // columns of H this CTA owns
  const float* H = layer == 0 ? a.ws + a.L.xhat : a.ws + zoff[layer - 1];
  const int act_in = layer == 0 ? B200PPO_ACT_NONE : ch->act;
  const float* D = a.ws + doff[layer];
  const int sp = blockIdx.y;
  const int item_id = blockIdx.x + item_base;
  const int Si = a.L.tc_item_S[item_id];                   // this item's number of row splits (<= gridDim.y)
  const int rps = a.L.tc_item_rps[item_id];
  const int r_begin = sp * rps;
  int r_end = r_begin + rps;
  if (r_end > a.L.R) r_end = a.L.R;
  // partial-gradient slots this item leaves unused (sp + Si, sp + 2 Si, ... < tc_S) are zeroed here,
  // so the fixed-order reduction over all tc_S slots stays valid whatever ran on this workspace before
  const int S_all = a.L.tc_S;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t pa = tc::plane_bytes(TCM), pb = tc::plane_bytes(npad);
  const int nst = r_end > r_begin ? (r_end - r_begin + TCK - 1) / TCK : 0;
  uint8_t* raw = tsmem + DW2_NS * TC_STAGE_BYTES;
  uint64_t* raw_full = rbar;
  uint64_t* raw_empty = rbar + DW2_NR_MAX;
  // Raw ring of this item: a stage is 16 rows of H (full rows, or the CTA's 128 columns through the tensor map) and
  // of D.  The streamer is latency bound (a slot is refilled only after the transposing warps have read it, and the
  // refill comes from DRAM), so the depth of the ring is what sets the stage rate: narrow items (64 + 64 columns: 8 KB
  // per stage) get 8 slots out of the area that holds 2 stages of the widest item (256 + 256 columns).
  const uint32_t raw_a_bytes = static_cast<uint32_t>(TCK) * static_cast<uint32_t>(K > DW2_MAXK ? TCM : K) * 4u;
  const uint32_t raw_b_off = (raw_a_bytes + 127u) & ~127u;
  const uint32_t raw_slot = (raw_b_off + static_cast<uint32_t>(TCK) * static_cast<uint32_t>(N) * 4u + 127u) & ~127u;
  int NR = static_cast<int>((DW2_NR * DW2_RAW_BYTES) / raw_slot);
  NR = NR > nr_max ? nr_max : NR;
  float bsum0 = 0.0f, bsum1 = 0.0f;
  const bool spin = (g_tc_stamp_skip_dw & 2) != 0;
  auto wait = [&](uint64_t* bar, uint32_t par) {
    if (spin) tc::mbar_wait_spin(bar, par);
    else tc::mbar_wait(bar, par);
  };
  float bs[8][4];                                            // v3: column sums of this thread's D chunks (bias gradient)
#pragma unroll
  for (int i = 0; i < 8; ++i) bs[i][0] = bs[i][1] = bs[i][2] = bs[i][3] = 0.0f;
  if (use_mn) {
    // ---------------- v3: TMA-fed MN-major operands ----------------
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tsmem) + 1023) & ~static_cast<uintptr_t>(1023));
    const int n32 = (N + 31) >> 5;                           // 32-column atoms of D (zero-filled beyond N)
    const int na = (kw + 31) >> 5;                           // atoms of H this CTA owns (<= 4)
    const uint32_t b_bytes = static_cast<uint32_t>(n32) * tc::MN_ATOM_BYTES;
    const uint32_t stage_b = dw3_stage_bytes(n32);
    uint64_t* full_raw = bar3;                               // TMA landed (tx count)
    uint64_t* full_lo = bar3 + DW3_NS_MAX;                   // one arrive per producer warp
    uint64_t* empty3 = bar3 + 2 * DW3_NS_MAX;                // tcgen05.commit of the MMAs that read the slot
    int NS3 = static_cast<int>((DW2_SMEM - 1024u) / stage_b);
    NS3 = NS3 > DW3_NS_MAX ? DW3_NS_MAX : NS3;
    if (warp < TC_NPROD / 32) {
      // A producer warp's work on a stage is one dependent chain (wait -> LDS -> STS -> proxy fence -> arrive, ~500
      // cycles whatever the width), so the 16 warps form DW3_NG groups that take the stages round-robin: four stages
      // are in the producers' hands at any time.  Thread lt of a group owns, in EVERY 2 KB atom of the stage, the
      // 16-byte chunk at offset lt * 16 (row lt >> 3): its column sums (bias gradient) stay in registers.
      const int grp = warp >> 2, lt = tid & 127;
      const bool do_bias = mt == 0;
      const uint32_t coff = static_cast<uint32_t>(lt) * 16u;
      for (int s = grp; s < nst; s += DW3_NG) {
        const int slot = s % NS3;
        const uint32_t par = static_cast<uint32_t>(s / NS3) & 1u;
        uint8_t* st = ring + slot * stage_b;
        uint8_t* a_hi = st;
        uint8_t* a_lo = st + DW3_A_BYTES;
        uint8_t* b_hi = st + 2u * DW3_A_BYTES;
        uint8_t* b_lo = b_hi + b_bytes;
        wait(&full_raw[slot], par);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < na) {
            const uint32_t off = static_cast<uint32_t>(i) * tc::MN_ATOM_BYTES + coff;
            float4 x = *reinterpret_cast<const float4*>(a_hi + off);
            if (act_in != B200PPO_ACT_NONE) {
              x = act4(x, act_in);
              *reinterpret_cast<float4*>(a_hi + off) = x;
            }
            if (split) {
              float4 lo;
              lo.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
              lo.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
              lo.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
              lo.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
              *reinterpret_cast<float4*>(a_lo + off) = lo;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < n32) {
            const uint32_t off = static_cast<uint32_t>(j) * tc::MN_ATOM_BYTES + coff;
            const float4 x = *reinterpret_cast<const float4*>(b_hi + off);
            if (do_bias) { bs[j][0] += x.x; bs[j][1] += x.y; bs[j][2] += x.z; bs[j][3] += x.w; }
            if (split) {
              float4 lo;
              lo.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
              lo.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
              lo.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
              lo.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
              *reinterpret_cast<float4*>(b_lo + off) = lo;
            }
          }
        }
        tc::fence_proxy_async();
        __syncwarp();
        if ((tid & 31) == 0) tc::mbar_arrive(&full_lo[slot]);
      }
    } else if (warp == TC_NPROD / 32) {
      // issuer: descriptors of slot 0 once; a slot / k-step / half is an add on the 16-byte start-address field
      const uint32_t idesc = tc::make_idesc_tf32_mn(TCM, n32 * 32);
      const uint64_t d0 = tc::make_desc_mn_sw128(tc::smem_u32(ring), tc::MN_ATOM_BYTES, tc::MN_SBO_BYTES);
      const uint32_t hiw = static_cast<uint32_t>(d0 >> 32), low0 = static_cast<uint32_t>(d0);
      const uint32_t dcol = cx.tmem_base;
      int slot = 0;
      uint32_t par = 0u;
      for (int s = 0; s < nst; ++s) {
        wait(&full_lo[slot], par);
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint32_t a0 = low0 + ((static_cast<uint32_t>(slot) * stage_b) >> 4);
          const uint32_t b0 = a0 + ((2u * DW3_A_BYTES) >> 4);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint64_t ah = tc_desc(hiw, a0 + j * (1024u >> 4));
            const uint64_t bh = tc_desc(hiw, b0 + j * (1024u >> 4));
            const uint32_t acc0 = (s > 0 || j > 0) ? 1u : 0u;
            if (split) {
              const uint64_t al = tc_desc(hiw, a0 + j * (1024u >> 4) + (DW3_A_BYTES >> 4));
              const uint64_t bl = tc_desc(hiw, b0 + j * (1024u >> 4) + (b_bytes >> 4));
              tc::mma_tf32(dcol, al, bh, idesc, acc0);
              tc::mma_tf32(dcol, ah, bl, idesc, 1u);
              tc::mma_tf32(dcol, ah, bh, idesc, 1u);
            } else {
              tc::mma_tf32(dcol, ah, bh, idesc, acc0);
            }
          }
          tc::commit(&empty3[slot]);
          if (s == nst - 1) tc::commit(cx.bar_done);
        }
        __syncwarp();
        if (++slot == NS3) { slot = 0; par ^= 1u; }
      }
    } else {
      // TMA streamer: two instructions per 16-row stage (3-D maps, tmap.cuh): the CTA's 4 atoms of H and every atom
      // of D land in the slot's hi tiles
      if ((tid & 31) == 0) {
        const CUtensorMap* mh = &maps.h[glayer];
        const CUtensorMap* md = &maps.d[glayer];
        const uint32_t bytes = static_cast<uint32_t>(4 + n32) * tc::MN_ATOM_BYTES;   // whole boxes, zero fill included
        int slot = 0;
        uint32_t epar = 1u;
        bool first = true;
        for (int s = 0; s < nst; ++s) {
          uint8_t* st = ring + slot * stage_b;
          const int rs = r_begin + s * TCK;
          if (!first) wait(&empty3[slot], epar);
          tc::mbar_arrive_expect_tx(&full_raw[slot], bytes);
          tc::tma_load_3d(st, mh, 0, rs, m0 >> 5, &full_raw[slot]);                       // 4 atoms of H (zeros beyond K)
          tc::tma_load_3d(st + 2u * DW3_A_BYTES, md, 0, rs, 0, &full_raw[slot]);          // every atom of D
          if (++slot == NS3) { slot = 0; epar ^= 1u; first = false; }
        }
      }
      __syncwarp();
    }
  } else if (warp < TC_NPROD / 32) {
    // A: thread -> (m = tid & 127, plane q = tid >> 7): rows 4q..4q+3 of column m
    const int am = tid & (TCM - 1), aq = tid >> 7;
    const bool av = am < kw;
    const bool wide = K > DW2_MAXK;                       // raw block = this CTA's 128 columns only (TMA box), else full rows
    const uint32_t a_src = wide ? static_cast<uint32_t>(((4 * aq) * TCM + (av ? am : 0)) * 4)
                                : static_cast<uint32_t>(((4 * aq) * K + m0 + (av ? am : 0)) * 4);
    const uint32_t a_dst = aq * pa + am * 16;
    const uint32_t a_rs = wide ? static_cast<uint32_t>(TCM) * 4u : static_cast<uint32_t>(K) * 4u;
    // B: chunk ids tid and tid + 512 -> (plane p = id / npad, column = id % npad)
    const int c0 = tid, c1 = tid + TC_NPROD;
    const bool bs0 = c0 < 4 * npad, bs1 = c1 < 4 * npad;
    const int p0 = c0 / npad, col0 = c0 - p0 * npad;
    const int p1 = c1 / npad, col1 = c1 - p1 * npad;
    const bool bv0 = bs0 && col0 < N, bv1 = bs1 && col1 < N;
    const uint32_t b_rs = static_cast<uint32_t>(N) * 4u;
    const uint32_t b_src0 = raw_b_off + static_cast<uint32_t>(((4 * p0) * N + (bv0 ? col0 : 0)) * 4);
    const uint32_t b_src1 = raw_b_off + static_cast<uint32_t>(((4 * p1) * N + (bv1 ? col1 : 0)) * 4);
    const uint32_t b_dst0 = p0 * pb + col0 * 16, b_dst1 = p1 * pb + col1 * 16;
    int lslot = 0;                                          // raw slot / parity of the next `ld` (called in stage order)
    uint32_t lpar = 0u;
    int uslot = 0;
    uint32_t upar = 1u;                                     // parity to wait on `empty` (first lap skipped)
    bool ufirst = true;
    struct Raw { float4 a, b0, b1; };
    TC_PROBE_DECL
    // shared -> registers for stage s, then hand the raw slot straight back to the streamer
    auto ld = [&](int s, Raw& x) {
      const int rslot = lslot;
      const uint8_t* rw = raw + rslot * raw_slot;
      TC_PROBE_START();
      wait(&raw_full[rslot], lpar);
      if (++lslot == NR) { lslot = 0; lpar ^= 1u; }
      TC_PROBE_LAP(0);
      x.a = x.b0 = x.b1 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (av) {
        const uint8_t* q = rw + a_src;
        x.a = make_float4(*reinterpret_cast<const float*>(q), *reinterpret_cast<const float*>(q + a_rs),
                          *reinterpret_cast<const float*>(q + 2 * a_rs), *reinterpret_cast<const float*>(q + 3 * a_rs));
      }
      if (bv0) {
        const uint8_t* q = rw + b_src0;
        x.b0 = make_float4(*reinterpret_cast<const float*>(q), *reinterpret_cast<const float*>(q + b_rs),
                           *reinterpret_cast<const float*>(q + 2 * b_rs), *reinterpret_cast<const float*>(q + 3 * b_rs));
      }
      if (bv1) {
        const uint8_t* q = rw + b_src1;
        x.b1 = make_float4(*reinterpret_cast<const float*>(q), *reinterpret_cast<const float*>(q + b_rs),
                           *reinterpret_cast<const float*>(q + 2 * b_rs), *reinterpret_cast<const float*>(q + 3 * b_rs));
      }
      const int rs = r_begin + s * TCK;
      if (rs + TCK > r_end) {                               // ragged last block: rows >= r_end are garbage
        const int ra = rs + 4 * aq, rb0 = rs + 4 * p0, rb1 = rs + 4 * p1;
        if (ra + 0 >= r_end) x.a.x = 0.f;
        if (ra + 1 >= r_end) x.a.y = 0.f;
        if (ra + 2 >= r_end) x.a.z = 0.f;
        if (ra + 3 >= r_end) x.a.w = 0.f;
        if (rb0 + 0 >= r_end) x.b0.x = 0.f;
        if (rb0 + 1 >= r_end) x.b0.y = 0.f;
        if (rb0 + 2 >= r_end) x.b0.z = 0.f;
        if (rb0 + 3 >= r_end) x.b0.w = 0.f;
        if (rb1 + 0 >= r_end) x.b1.x = 0.f;
        if (rb1 + 1 >= r_end) x.b1.y = 0.f;
        if (rb1 + 2 >= r_end) x.b1.z = 0.f;
        if (rb1 + 3 >= r_end) x.b1.w = 0.f;
      }
      __syncwarp();
      if ((tid & 31) == 0) tc::mbar_arrive(&raw_empty[rslot]);   // release: ordered after the reads above
      TC_PROBE_LAP(1);
    };
    // registers -> activation, hi/lo split -> UMMA planes of the next ring slot
    auto proc = [&](const Raw& x) {
      TC_PROBE_START();
      if (!ufirst) wait(&cx.bar_empty[uslot], upar);
      TC_PROBE_LAP(2);
      uint8_t* ua = cx.smem + uslot * 2u * TC_A_BYTES;
      uint8_t* ub = cx.smem + cx.b_base + uslot * 2u * TC_B_BYTES;
      float4 hi, lo;
      tc::split4_fast(av ? act4(x.a, act_in) : x.a, hi, lo);
      *reinterpret_cast<float4*>(ua + a_dst) = hi;
      if (split) *reinterpret_cast<float4*>(ua + TC_A_BYTES + a_dst) = lo;
      if (bs0) {
        bsum0 += (x.b0.x + x.b0.y) + (x.b0.z + x.b0.w);
        tc::split4_fast(x.b0, hi, lo);
        *reinterpret_cast<float4*>(ub + b_dst0) = hi;
        if (split) *reinterpret_cast<float4*>(ub + TC_B_BYTES + b_dst0) = lo;
      }
      if (bs1) {
        bsum1 += (x.b1.x + x.b1.y) + (x.b1.z + x.b1.w);
        tc::split4_fast(x.b1, hi, lo);
        *reinterpret_cast<float4*>(ub + b_dst1) = hi;
        if (split) *reinterpret_cast<float4*>(ub + TC_B_BYTES + b_dst1) = lo;
      }
      TC_PROBE_LAP(3);
      tc::fence_proxy_async();
      __syncwarp();
      if ((tid & 31) == 0) tc::mbar_arrive(&cx.bar_full_a[uslot]);
      TC_PROBE_LAP(4);
      if (++uslot == DW2_NS) { uslot = 0; upar ^= 1u; ufirst = false; }
    };
    // software pipeline: the shared-memory reads of stage s+1 are in flight while stage s is split
    Raw xA, xB;
    if (nst > 0) ld(0, xA);
    for (int s = 0; s < nst; s += 2) {
      if (s + 1 < nst) ld(s + 1, xB);
      proc(xA);
      if (s + 1 < nst) {
        if (s + 2 < nst) ld(s + 2, xA);
        proc(xB);
      }
    }
    if (tid == 0) TC_PROBE_FLUSH(0, 5);
  } else if (warp == TC_NPROD / 32) {
    const TcIssue ti = tc_issue_prepare(cx, npad);
    int uslot = 0;
    uint32_t par = 0u;
    TC_PROBE_DECL
    for (int s = 0; s < nst; ++s) {
      TC_PROBE_START();
      wait(&cx.bar_full_a[uslot], par);
      TC_PROBE_LAP(0);
      tc_issue(cx, ti, uslot, split, s == 0, s == nst - 1);
      TC_PROBE_LAP(1);
      if (++uslot == DW2_NS) { uslot = 0; par ^= 1u; }
    }
    TC_PROBE_FLUSH(5, 2);
  } else {
    // raw-block streamer: 16 full rows of H and of D per stage, one bulk copy each (many small
    // copies are slow: ~100 cycles apiece through the copy engine, measured with 512-byte segments)
    if ((tid & 31) == 0) {
      const bool wide = K > DW2_MAXK;
      const uint32_t bytes_a = static_cast<uint32_t>(TCK) * (wide ? TCM : K) * 4u, bytes_b = static_cast<uint32_t>(TCK) * N * 4u;
      int rslot = 0;
      uint32_t epar = 1u;
      bool first = true;
      for (int s = 0; s < nst; ++s) {
        uint8_t* rw = raw + rslot * raw_slot;
        const size_t rs = static_cast<size_t>(r_begin) + static_cast<size_t>(s) * TCK;
        if (!first) wait(&raw_empty[rslot], epar);
        tc::mbar_arrive_expect_tx(&raw_full[rslot], bytes_a + bytes_b);
        if (wide) tc::tma_load_2d(rw, &tm_xhat, m0, static_cast<int>(rs), &raw_full[rslot]);
        else tc::bulk_g2s(rw, H + rs * K, bytes_a, &raw_full[rslot]);
        tc::bulk_g2s(rw + raw_b_off, D + rs * N, bytes_b, &raw_full[rslot]);
        if (++rslot == NR) { rslot = 0; epar ^= 1u; first = false; }
      }
    }
    __syncwarp();
  }
  cx.g += nst;
  float* gpart = a.ws + a.L.gpart + static_cast<size_t>(sp) * a.plan.n_params;
  float* gp = gpart + ch->w_off[layer];
  tc_stamp(cx.nstamp);
  if (warp < TC_NPROD / 32 && nst > 0) {
    tc_epilogue<0>(
        cx, npad, TcChainOut{0u, 0, 0}, [&](int, int, float (&)[16], float (&)[16]) {}, TcNoPre(),
        [&](int r, int col, const float4 val, const float4) {
          const int k = m0 + r;
          if (k < K && col < N) {
            float* dst = gp + static_cast<size_t>(k) * N + col;
            if ((N & 3) == 0) {
              *reinterpret_cast<float4*>(dst) = val;
            } else {
              dst[0] = val.x;
              if (col + 1 < N) dst[1] = val.y;
              if (col + 2 < N) dst[2] = val.z;
              if (col + 3 < N) dst[3] = val.w;
            }
            for (int s2 = sp + Si; s2 < S_all; s2 += Si) {
              float* dz = dst + static_cast<size_t>(s2 - sp) * a.plan.n_params;
              for (int i = 0; i < 4; ++i)
                if (col + i < N) dz[i] = 0.0f;
            }
          }
          return val;
        });
  }
  cx.done_uses++;
  if (mt == 0 && use_mn) {
    // v3 bias gradient: thread lt of producer group g holds, for every D atom, the sums of row k = lt >> 3 (over the
    // group's stages) of the 4 columns at the un-swizzled position of its chunk; the 4 x 16 partial sums of a column
    // meet in shared memory (the ring is idle now) and are added in (group, row) order
    __syncthreads();
    const int n32 = (N + 31) >> 5, ncol = n32 * 32;
    float* scr = reinterpret_cast<float*>(tsmem);
    if (tid < TC_NPROD) {
      const int grp = warp >> 2, lt = tid & 127;
      const int k = lt >> 3, colin = ((((lt >> 1) & 3) ^ (k & 3)) << 3) + ((lt & 1) << 2);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < n32)
          *reinterpret_cast<float4*>(scr + (grp * TCK + k) * ncol + j * 32 + colin) = make_float4(bs[j][0], bs[j][1], bs[j][2], bs[j][3]);
    }
    __syncthreads();
    if (tid < N) {
      float acc = 0.0f;
      for (int q = 0; q < DW3_NG * TCK; ++q) acc += scr[q * ncol + tid];
      gpart[ch->b_off[layer] + tid] = acc;
      for (int s2 = sp + Si; s2 < S_all; s2 += Si)
        gpart[static_cast<size_t>(s2 - sp) * a.plan.n_params + ch->b_off[layer] + tid] = 0.0f;
    }
  } else if (mt == 0) {                                    // bias gradient: fixed-order column sums
    if (tid < TC_NPROD) { bred[0][tid] = bsum0; bred[1][tid] = bsum1; }
    __syncthreads();
    if (tid < N && tid < TC_MAXN) {
      float acc = 0.0f;
      for (int cc = tid; cc < 4 * npad; cc += npad) acc += bred[cc / TC_NPROD][cc % TC_NPROD];
      gpart[ch->b_off[layer] + tid] = acc;
      for (int s2 = sp + Si; s2 < S_all; s2 += Si)
        gpart[static_cast<size_t>(s2 - sp) * a.plan.n_params + ch->b_off[layer] + tid] = 0.0f;
    }
  }
  tc_ctx_fini(cx);
}
