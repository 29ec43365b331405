// Tensor-core (tcgen05, kind::tf32) versions of the three GEMM kernels of the PPO update.
// Included by update.cu inside its anonymous namespace (needs Layout / FwdArgs / BwdArgs).
//
// Every GEMM is D[128 x Npad] (fp32, TMEM) += A[128 x K] * Bop[Npad x K]^T with both operands
// K-major in shared memory (layout: tc.cuh).  fp32 parity is kept by the error-compensated split
// x = hi + lo (both tf32): D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (3xTF32), measured at 2-4x the
// error of a cuBLAS fp32 GEMM (tests/test_gpu_tensorcore.py).  Weights are split once per update
// by upd_prep_w_kernel and reach shared memory with cp.async.bulk; activations / gradients are
// split while they are staged, one stage ahead in registers.
//
// Pipeline per GEMM (all 512 threads in lock step, 2 shared-memory stages of 32 k):
//   store A(s) regs -> fence.proxy.async -> __syncthreads -> one thread waits for B(s) and issues
//   4 k-steps x 3 MMAs + tcgen05.commit -> wait(other stage free) -> issue B(s+1) bulk copy and
//   A(s+1) loads (both in flight while the MMAs run) -> ... -> wait(done) -> tcgen05.ld epilogue.
// ncu (profiles/r1_ncu_summary_tc.md): the first version of these kernels was bound by staging
// INSTRUCTIONS (31 per element), hence the hoisted per-thread pointers, the fixed 8-plane stages
// and the bounds-check-free fast paths below.

constexpr int TCM = 128;
constexpr int TCK = 32;
constexpr int TCT = 512;
constexpr int TC_MAXN = 256;
constexpr int TC_APT = TCM * (TCK / 4) / TCT;      // A chunks (16 B) per thread per stage = 2
constexpr uint32_t TC_STAGE_BYTES = 2u * (TCK / 4) * tc::plane_bytes(TCM) + 2u * (TCK / 4) * tc::plane_bytes(TC_MAXN);
constexpr uint32_t TC_SMEM = 2u * TC_STAGE_BYTES;

struct TcCtx {
  uint8_t* smem;
  uint64_t* bar_empty;   // [2] stage buffer free (tcgen05.commit)
  uint64_t* bar_full;    // [2] B operand landed (cp.async.bulk complete_tx)
  uint64_t* bar_done;
  uint32_t tmem_base;
  uint32_t uses0, uses1;   // commits issued so far on stage buffer 0 / 1
  uint32_t full0, full1;   // B fills consumed per stage buffer
  uint32_t done_uses;
};
__device__ __forceinline__ uint32_t tc_full(const TcCtx& cx, int buf) { return buf ? cx.full1 : cx.full0; }
__device__ __forceinline__ void tc_full_inc(TcCtx& cx, int buf) { if (buf) cx.full1++; else cx.full0++; }
__device__ __forceinline__ uint32_t tc_uses(const TcCtx& cx, int buf) { return buf ? cx.uses1 : cx.uses0; }
__device__ __forceinline__ void tc_uses_inc(TcCtx& cx, int buf) { if (buf) cx.uses1++; else cx.uses0++; }
__device__ __forceinline__ void tc_wait_free(const TcCtx& cx, int buf) {
  if (tc_uses(cx, buf) > 0u) tc::mbar_wait(&cx.bar_empty[buf], (tc_uses(cx, buf) - 1u) & 1u);
}

__device__ __forceinline__ void tc_ctx_init(TcCtx& cx, uint8_t* smem, uint64_t* bars, uint32_t* tmem_slot) {
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tc::tmem_alloc(tmem_slot, TC_MAXN);
  if (threadIdx.x == 32) {
#pragma unroll
    for (int i = 0; i < 5; ++i) tc::mbar_init(&bars[i], 1);
    tc::mbar_init_fence();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  cx.smem = smem;
  cx.bar_empty = bars;
  cx.bar_full = bars + 2;
  cx.bar_done = bars + 4;
  cx.tmem_base = *tmem_slot;
  cx.uses0 = cx.uses1 = 0u;
  cx.full0 = cx.full1 = 0u;
  cx.done_uses = 0u;
}

__device__ __forceinline__ void tc_ctx_fini(TcCtx& cx) {
  tc::tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc(cx.tmem_base, TC_MAXN);
}

struct TcStage {
  uint8_t *a_hi, *a_lo, *b_hi, *b_lo;
};
__device__ __forceinline__ TcStage tc_stage(const TcCtx& cx, int buf) {
  TcStage s;
  s.a_hi = cx.smem + buf * TC_STAGE_BYTES;
  s.a_lo = s.a_hi + (TCK / 4) * tc::plane_bytes(TCM);
  s.b_hi = s.a_lo + (TCK / 4) * tc::plane_bytes(TCM);
  s.b_lo = s.b_hi + (TCK / 4) * tc::plane_bytes(TC_MAXN);
  return s;
}

// issue the MMAs of one staged block (4 k-steps of 8) and commit
__device__ __forceinline__ void tc_issue(TcCtx& cx, const TcStage& st, int buf, int npad, int split, bool first,
                                         bool last) {
  const uint32_t pa = tc::plane_bytes(TCM), pb = tc::plane_bytes(npad);
  const uint32_t idesc = tc::make_idesc_tf32(TCM, npad);
  tc::tc_fence_after();
#pragma unroll
  for (int j = 0; j < TCK / 8; ++j) {
    const uint64_t ah = tc::make_desc(tc::smem_u32(st.a_hi + 2 * j * pa), pa, 128);
    const uint64_t bh = tc::make_desc(tc::smem_u32(st.b_hi + 2 * j * pb), pb, 128);
    const uint32_t acc0 = (!first || j > 0) ? 1u : 0u;
    if (split) {
      const uint64_t al = tc::make_desc(tc::smem_u32(st.a_lo + 2 * j * pa), pa, 128);
      const uint64_t bl = tc::make_desc(tc::smem_u32(st.b_lo + 2 * j * pb), pb, 128);
      tc::mma_tf32(cx.tmem_base, al, bh, idesc, acc0);
      tc::mma_tf32(cx.tmem_base, ah, bl, idesc, 1u);
      tc::mma_tf32(cx.tmem_base, ah, bh, idesc, 1u);
    } else {
      tc::mma_tf32(cx.tmem_base, ah, bh, idesc, acc0);
    }
  }
  tc::commit(&cx.bar_empty[buf]);
  if (last) tc::commit(cx.bar_done);
}

// wait for the accumulator, run the epilogue functor on 16-column groups, release the accumulator
template <class Epi>
__device__ __forceinline__ void tc_epilogue(TcCtx& cx, int npad, Epi epi) {
  tc::mbar_wait(cx.bar_done, cx.done_uses & 1u);
  cx.done_uses++;
  tc::tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = warp & 3;
  for (int c = (warp >> 2) * 16; c < npad; c += (TCT / 128) * 16) {
    float v[16];
    tc::tmem_ld16(cx.tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + static_cast<uint32_t>(c), v);
    epi(sub * 32 + lane, c, v);
  }
  tc::tc_fence_before();
  __syncthreads();
}

__device__ __forceinline__ float4 act4(float4 x, int act) {
  if (act == B200PPO_ACT_RELU) {
    x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
  } else if (act != B200PPO_ACT_NONE) {
    x.x = act_fwd(x.x, act); x.y = act_fwd(x.y, act); x.z = act_fwd(x.z, act); x.w = act_fwd(x.w, act);
  }
  return x;
}

// Row-tile GEMM.  A[row][k] (row-major, leading dimension lda, rows row0..row0+127, rows >= nrows
// read as zero, activation `act` applied on load); B from the pre-split planes Bhi/Blo laid out
// [K32/4][npad + 1][4] in global memory (K32 = K rounded up to 32; same padded plane stride as in
// shared memory, so a stage of B is ONE contiguous bulk copy per half).
template <class Epi>
__device__ __forceinline__ void tc_gemm_rowtile(TcCtx& cx, const float* __restrict__ A, int lda, int row0,
                                                int nrows, int K, int act, const float* __restrict__ Bhi,
                                                const float* __restrict__ Blo, int npad, int split,
                                                const float* __restrict__ colvec, int ncol, float* colvec_s,
                                                Epi epi) {
  const int tid = threadIdx.x;
  // per-column epilogue vector (bias) staged in shared memory: visible after the first stage barrier
  if (colvec != nullptr && tid < npad) colvec_s[tid] = tid < ncol ? __ldg(colvec + tid) : 0.0f;
  const uint32_t pa = tc::plane_bytes(TCM), pb = tc::plane_bytes(npad);
  const int nst = (K + TCK - 1) / TCK;
  const uint32_t bbytes = (TCK / 4) * pb;
  // per-thread chunk coordinates are the same for every stage: hoist pointers and smem offsets
  const float* ap[TC_APT];
  uint32_t soff[TC_APT];
  int kq[TC_APT];
  bool rv[TC_APT];
#pragma unroll
  for (int i = 0; i < TC_APT; ++i) {
    const int idx = tid + i * TCT;
    const int q = idx & 7, r = idx >> 3;
    rv[i] = row0 + r < nrows;
    ap[i] = A + static_cast<size_t>(rv[i] ? row0 + r : 0) * lda + 4 * q;
    soff[i] = q * pa + r * 16;
    kq[i] = 4 * q;
  }
  const bool vec = (lda & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
  // A is prefetched TWO stages ahead in registers (global-load latency >> one stage of MMAs)
  float4 areg[2][TC_APT];
  auto load_a = [&](int s, float4 (&dst)[TC_APT]) {
    const int k0 = s * TCK;
#pragma unroll
    for (int i = 0; i < TC_APT; ++i) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rv[i]) {
        const int k = k0 + kq[i];
        if (vec && k + 3 < K) {
          x = *reinterpret_cast<const float4*>(ap[i] + k0);
        } else if (k < K) {
          const float* src = ap[i] + k0;
          x.x = src[0];
          if (k + 1 < K) x.y = src[1];
          if (k + 2 < K) x.z = src[2];
          if (k + 3 < K) x.w = src[3];
        }
      }
      dst[i] = x;
    }
  };
  auto issue_b = [&](int s, int buf) {        // one thread
    const TcStage st = tc_stage(cx, buf);
    const size_t off = static_cast<size_t>(s) * (bbytes / 4);     // floats
    tc::mbar_arrive_expect_tx(&cx.bar_full[buf], split ? 2u * bbytes : bbytes);
    tc::bulk_g2s(st.b_hi, Bhi + off, bbytes, &cx.bar_full[buf]);
    if (split) tc::bulk_g2s(st.b_lo, Blo + off, bbytes, &cx.bar_full[buf]);
  };
  tc_wait_free(cx, 0);
  if (tid == 0) issue_b(0, 0);
  load_a(0, areg[0]);
  if (nst > 1) load_a(1, areg[1]);
  for (int s = 0; s < nst; ++s) {
    const int buf = s & 1;
    const TcStage st = tc_stage(cx, buf);
#pragma unroll
    for (int i = 0; i < TC_APT; ++i) {
      float4 hi, lo;
      tc::split4(act4(buf ? areg[1][i] : areg[0][i], act), hi, lo);
      *reinterpret_cast<float4*>(st.a_hi + soff[i]) = hi;
      if (split) *reinterpret_cast<float4*>(st.a_lo + soff[i]) = lo;
    }
    if (s + 2 < nst) {                          // refill the register set just consumed
      if (buf) load_a(s + 2, areg[1]); else load_a(s + 2, areg[0]);
    }
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::mbar_wait(&cx.bar_full[buf], tc_full(cx, buf) & 1u);
      tc_issue(cx, st, buf, npad, split, s == 0, s == nst - 1);
    }
    tc_uses_inc(cx, buf);
    tc_full_inc(cx, buf);
    if (s + 1 < nst) {
      tc_wait_free(cx, buf ^ 1);
      if (tid == 0) issue_b(s + 1, buf ^ 1);
    }
  }
  tc_epilogue(cx, npad, epi);
}

// ------------------------------------------------------------------------------------------
// weight pre-split: Wf (forward operand, Bop(n, k) = W[k][n]) and Wb (dX operand,
// Bop(kout, nred) = W[kout][nred]), each as hi / lo planes [red32/4][rows_pad + 1][4]
// ------------------------------------------------------------------------------------------
struct PrepArgs {
  b200ppo_plan plan;
  Layout L;
  const float* params;
  float* ws;
};

__global__ void __launch_bounds__(256) upd_prep_w_kernel(const PrepArgs a) {
  // blockIdx.y = chain * MAXL + layer ; blockIdx.z = 0 (Wf) / 1 (Wb)
  const int chain = blockIdx.y / MAXL, l = blockIdx.y % MAXL;
  const b200ppo_chain& ch = chain == 0 ? a.plan.actor : a.plan.critic;
  if (l >= ch.n_layers) return;
  const TcLayer& t = chain == 0 ? a.L.tca[l] : a.L.tcc[l];
  const int K = ch.dims[l], N = ch.dims[l + 1];
  const float* W = a.params + ch.w_off[l];
  const bool fwd = blockIdx.z == 0;
  const int rows = fwd ? t.npad : t.kout_pad;          // operand rows
  const int red = fwd ? t.kpad : t.nred_pad;           // reduction length (padded to 32)
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
__device__ __forceinline__ void tc_chain_forward(TcCtx& cx, const b200ppo_chain& ch, const TcLayer* tl,
                                                 const float* __restrict__ P, float* ws, const size_t* zoff,
                                                 size_t xhat_off, int row0, int nrows, int split, float* bias_s) {
  for (int l = 0; l < ch.n_layers; ++l) {
    const int K = ch.dims[l], N = ch.dims[l + 1];
    const float* A = l == 0 ? ws + xhat_off : ws + zoff[l - 1];
    const int act_in = l == 0 ? B200PPO_ACT_NONE : ch.act;
    const float* bias = P + ch.b_off[l];
    float* Z = ws + zoff[l];
    tc_gemm_rowtile(cx, A, K, row0, nrows, K, act_in, ws + tl[l].wf_hi, ws + tl[l].wf_lo, tl[l].npad, split,
                    bias, N, bias_s, [&](int r, int c, const float (&v)[16]) {
                      const int row = row0 + r;
                      if (row < nrows) {
                        float* dst = Z + static_cast<size_t>(row) * N + c;
                        if ((N & 3) == 0) {
#pragma unroll
                          for (int i = 0; i < 16; i += 4)
                            if (c + i < N) {
                              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c + i);
                              *reinterpret_cast<float4*>(dst + i) =
                                  make_float4(v[i] + b4.x, v[i + 1] + b4.y, v[i + 2] + b4.z, v[i + 3] + b4.w);
                            }
                        } else {
#pragma unroll
                          for (int i = 0; i < 16; ++i)
                            if (c + i < N) dst[i] = v[i] + bias_s[c + i];
                        }
                      }
                    });
  }
}

__global__ void __launch_bounds__(TCT, 1) upd_fwd_tc_kernel(const FwdArgs a, const int split) {
  extern __shared__ __align__(128) uint8_t tsmem[];
  __shared__ uint64_t bars[5];
  __shared__ uint32_t tmem_slot;
  __shared__ const float* rowsrc[TCM];
  __shared__ __align__(16) float bias_s[TC_MAXN];
  TcCtx cx;
  tc_ctx_init(cx, tsmem, bars, &tmem_slot);
  const int O = a.plan.obs_dim;
  const int row0 = blockIdx.x * TCM;
  const int R = a.L.R, Rv = a.L.Rv;
  float* xhat = a.ws + a.L.xhat;
  for (int m = threadIdx.x; m < TCM; m += TCT) {
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
  constexpr int XB = 4;
  for (int i0 = threadIdx.x; i0 < TCM * O; i0 += TCT * XB) {
    float xv[XB];
#pragma unroll
    for (int j = 0; j < XB; ++j) {
      const int idx = i0 + j * TCT;
      const int m = idx / O, k = idx - m * O;
      xv[j] = 0.0f;
      if (idx < TCM * O && rowsrc[m] != nullptr) xv[j] = rowsrc[m][k];
    }
#pragma unroll
    for (int j = 0; j < XB; ++j) {
      const int idx = i0 + j * TCT;
      const int m = idx / O, k = idx - m * O;
      if (idx < TCM * O && rowsrc[m] != nullptr) {
        float x = xv[j];
        if (a.plan.normalize) x = __fdiv_rn(x - __ldg(a.mean + k), __ldg(a.std + k));
        xhat[static_cast<size_t>(row0 + m) * O + k] = x;
      }
    }
  }
  __syncthreads();
  tc_chain_forward(cx, a.plan.critic, a.L.tcc, a.params, a.ws, a.L.zc, a.L.xhat, row0, Rv, split, bias_s);
  if (row0 < R) tc_chain_forward(cx, a.plan.actor, a.L.tca, a.params, a.ws, a.L.za, a.L.xhat, row0, R, split, bias_s);
  tc_ctx_fini(cx);
}

// ------------------------------------------------------------------------------------------
// BWD dX (tensor cores): dpre_{l-1} = (dpre_l W_l^T) ⊙ act'(z_{l-1})
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_chain_backward(TcCtx& cx, const b200ppo_chain& ch, const TcLayer* tl,
                                                  float* ws, const size_t* zoff, const size_t* doff, int row0,
                                                  int nrows, int split) {
  for (int l = ch.n_layers - 1; l >= 1; --l) {
    const int Kl = ch.dims[l], Nl = ch.dims[l + 1];
    const float* dY = ws + doff[l];
    const float* zprev = ws + zoff[l - 1];
    float* dprev = ws + doff[l - 1];
    const int act = ch.act;
    tc_gemm_rowtile(cx, dY, Nl, row0, nrows, Nl, B200PPO_ACT_NONE, ws + tl[l].wb_hi, ws + tl[l].wb_lo,
                    tl[l].kout_pad, split, nullptr, 0, nullptr, [&](int r, int c, const float (&v)[16]) {
                      const int row = row0 + r;
                      if (row < nrows) {
                        const size_t o = static_cast<size_t>(row) * Kl + c;
                        if ((Kl & 3) == 0) {
#pragma unroll
                          for (int i = 0; i < 16; i += 4)
                            if (c + i < Kl) {
                              const float4 z = *reinterpret_cast<const float4*>(zprev + o + i);
                              float4 g;
                              if (act == B200PPO_ACT_RELU) {
                                g = make_float4(z.x > 0.f ? v[i] : 0.f, z.y > 0.f ? v[i + 1] : 0.f,
                                                z.z > 0.f ? v[i + 2] : 0.f, z.w > 0.f ? v[i + 3] : 0.f);
                              } else {
                                g = make_float4(v[i] * act_grad(z.x, act), v[i + 1] * act_grad(z.y, act),
                                                v[i + 2] * act_grad(z.z, act), v[i + 3] * act_grad(z.w, act));
                              }
                              *reinterpret_cast<float4*>(dprev + o + i) = g;
                            }
                        } else {
#pragma unroll
                          for (int i = 0; i < 16; ++i)
                            if (c + i < Kl) dprev[o + i] = v[i] * act_grad(zprev[o + i], act);
                        }
                      }
                    });
  }
}

__global__ void __launch_bounds__(TCT, 1) upd_bwd_dx_tc_kernel(const BwdArgs a, const int split) {
  extern __shared__ __align__(128) uint8_t tsmem[];
  __shared__ uint64_t bars[5];
  __shared__ uint32_t tmem_slot;
  TcCtx cx;
  tc_ctx_init(cx, tsmem, bars, &tmem_slot);
  const int row0 = blockIdx.x * TCM;
  tc_chain_backward(cx, a.plan.critic, a.L.tcc, a.ws, a.L.zc, a.L.dc, row0, a.L.R, split);
  tc_chain_backward(cx, a.plan.actor, a.L.tca, a.ws, a.L.za, a.L.da, row0, a.L.R, split);
  tc_ctx_fini(cx);
}

// ------------------------------------------------------------------------------------------
// BWD dW (tensor cores): dW_l[k][n] = sum_r act(z_{l-1})[r][k] * dpre_l[r][n]
// D rows = k (tiles of 128), D cols = n (padded to 16), reduction over this CTA's row range; both
// operands are transposed gathers (a 16-byte chunk = the same column of 4 consecutive data rows).
// The threads that stage column n of dpre also accumulate its sum: the bias gradient.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TCT, 1) upd_bwd_dw_tc_kernel(const BwdArgs a, const int split) {
  extern __shared__ __align__(128) uint8_t tsmem[];
  __shared__ uint64_t bars[5];
  __shared__ uint32_t tmem_slot;
  __shared__ float bred[TCT];
  int item = blockIdx.x;
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
  TcCtx cx;
  tc_ctx_init(cx, tsmem, bars, &tmem_slot);
  const int K = ch->dims[layer], N = ch->dims[layer + 1];
  const int npad = (N + 15) & ~15;
  const int m0 = mt * TCM;
  const float* H = layer == 0 ? a.ws + a.L.xhat : a.ws + zoff[layer - 1];
  const int act_in = layer == 0 ? B200PPO_ACT_NONE : ch->act;
  const float* D = a.ws + doff[layer];
  const int sp = blockIdx.y;
  const int r_begin = sp * a.L.tc_rows_per_split;
  int r_end = r_begin + a.L.tc_rows_per_split;
  if (r_end > a.L.R) r_end = a.L.R;
  const int tid = threadIdx.x;
  const uint32_t pa = tc::plane_bytes(TCM), pb = tc::plane_bytes(npad);
  const int nst = (r_end - r_begin + TCK - 1) / TCK;
  // A: thread -> (m = idx & 127, q = idx >> 7), idx = tid + i*512: same coordinates every stage
  const float* ap[TC_APT];
  uint32_t aoff[TC_APT];
  int aq[TC_APT];
  bool av[TC_APT];
#pragma unroll
  for (int i = 0; i < TC_APT; ++i) {
    const int idx = tid + i * TCT;
    const int m = idx & (TCM - 1), q = idx >> 7;
    av[i] = m0 + m < K;
    aq[i] = 4 * q;
    ap[i] = H + static_cast<size_t>(r_begin + 4 * q) * K + (av[i] ? m0 + m : 0);
    aoff[i] = q * pa + m * 16;
  }
  // B: thread -> column n = tid & 255 (if < npad), plane half = tid >> 8 (planes 4*half .. +3)
  const int bn = tid & 255, bh = tid >> 8;
  const bool bv = bn < N;
  const bool bstage = bn < npad;
  const float* bp = D + static_cast<size_t>(r_begin + 16 * bh) * N + (bv ? bn : 0);
  float bsum = 0.0f;
  float4 areg[TC_APT], breg[4];
  auto load_stage = [&](int s) {
    const int rs = r_begin + s * TCK;
    const size_t so = static_cast<size_t>(s) * TCK;
    if (rs + TCK <= r_end) {                                    // full stage: no row bounds checks
#pragma unroll
      for (int i = 0; i < TC_APT; ++i) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (av[i]) {
          const float* src = ap[i] + so * K;
          x = make_float4(src[0], src[K], src[2 * static_cast<size_t>(K)], src[3 * static_cast<size_t>(K)]);
        }
        areg[i] = x;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bv) {
          const float* src = bp + (so + 4 * j) * N;
          x = make_float4(src[0], src[N], src[2 * static_cast<size_t>(N)], src[3 * static_cast<size_t>(N)]);
        }
        breg[j] = x;
      }
    } else {
#pragma unroll
      for (int i = 0; i < TC_APT; ++i) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (av[i]) {
          const float* src = ap[i] + so * K;
          const int r = rs + aq[i];
          if (r + 0 < r_end) x.x = src[0];
          if (r + 1 < r_end) x.y = src[K];
          if (r + 2 < r_end) x.z = src[2 * static_cast<size_t>(K)];
          if (r + 3 < r_end) x.w = src[3 * static_cast<size_t>(K)];
        }
        areg[i] = x;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bv) {
          const float* src = bp + (so + 4 * j) * N;
          const int r = rs + 16 * bh + 4 * j;
          if (r + 0 < r_end) x.x = src[0];
          if (r + 1 < r_end) x.y = src[N];
          if (r + 2 < r_end) x.z = src[2 * static_cast<size_t>(N)];
          if (r + 3 < r_end) x.w = src[3 * static_cast<size_t>(N)];
        }
        breg[j] = x;
      }
    }
  };
  if (nst > 0) load_stage(0);
  for (int s = 0; s < nst; ++s) {
    const int buf = s & 1;
    const TcStage st = tc_stage(cx, buf);
    tc_wait_free(cx, buf);
#pragma unroll
    for (int i = 0; i < TC_APT; ++i) {
      float4 hi, lo;
      tc::split4(act4(areg[i], act_in), hi, lo);       // act(0) == 0 for relu / tanh / swish
      *reinterpret_cast<float4*>(st.a_hi + aoff[i]) = hi;
      if (split) *reinterpret_cast<float4*>(st.a_lo + aoff[i]) = lo;
    }
    if (bstage) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 hi, lo;
        tc::split4(breg[j], hi, lo);
        bsum += (breg[j].x + breg[j].y) + (breg[j].z + breg[j].w);
        const uint32_t o = (4 * bh + j) * pb + bn * 16;
        *reinterpret_cast<float4*>(st.b_hi + o) = hi;
        if (split) *reinterpret_cast<float4*>(st.b_lo + o) = lo;
      }
    }
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) tc_issue(cx, st, buf, npad, split, s == 0, s == nst - 1);
    tc_uses_inc(cx, buf);
    if (s + 1 < nst) load_stage(s + 1);
  }
  float* gpart = a.ws + a.L.gpart + static_cast<size_t>(sp) * a.plan.n_params;
  float* gp = gpart + ch->w_off[layer];
  if (nst > 0) {
    tc_epilogue(cx, npad, [&](int r, int c, const float (&v)[16]) {
      const int k = m0 + r;
      if (k < K) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c + i < N) gp[static_cast<size_t>(k) * N + c + i] = v[i];
      }
    });
  }
  if (mt == 0) {                                           // bias gradient: fixed-order column sums
    bred[tid] = bsum;
    __syncthreads();
    if (tid < N && tid < 256) gpart[ch->b_off[layer] + tid] = bred[tid] + bred[tid + 256];
  }
  tc_ctx_fini(cx);
}
