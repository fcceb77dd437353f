// CTC loss + gradient (torch.nn.CTCLoss semantics) for the dynamic-eval adapt step.
// Call sites replaced: lcasr/lib.py:492,570-579 (N=1), :310-331 (AWMC, ragged targets),
// earnings_finetune/train.py:259 (ragged input lengths).  Formulas: SURVEY.md appendix A.
//
// Two launches:
//   ctc_lattice_kernel : grid (2, N).  blockIdx.x = 0 runs alpha forward in t, = 1 runs beta
//                        backward in t, concurrently on two SMs.  beta is alpha on the
//                        time-reversed input and the reversed extended label sequence, so
//                        one code path serves both.  Values are kept in log2 units (bare
//                        MUFU.EX2 / MUFU.LG2) and are re-centred every step by the previous
//                        step's maximum; the removed offsets are summed in fp64.  Stored
//                        lattice values therefore stay O(1) and fp32 keeps ~1e-6 relative
//                        accuracy in the occupancies even when |log-likelihood| ~ 1e4.
//   ctc_grad_kernel    : one CTA per (t, n) row: dense exp(lp) minus the occupancy of the
//                        classes that occur in the label sequence.  Pure streaming.
#include <cstdlib>
#include "ctc_shared.cuh"

namespace dae {

struct LatSmem {              // byte offsets into dynamic shared memory (host and device agree)
  int bars, cx, wmx, lab, a0, a1, rows, row_stride, stages, total;
};
__host__ __device__ inline LatSmem lat_smem_layout(int C, int Lp, int Sp, int max_bytes) {
  LatSmem m;
  m.bars = 0;                                            // up to 16 mbarriers
  m.cx = 16 * 8;                                         // float2[4]: (centring constant, blank emission) per step slot
  m.wmx = m.cx + 4 * 8;                                  // int[4][32]: per-warp maxima per step slot
  m.lab = m.wmx + 4 * 32 * 4;
  m.a0 = (int)align_up((size_t)m.lab + (size_t)Lp * 4, 16);
  m.a1 = m.a0 + (Sp + 4) * 4;
  m.rows = (int)align_up((size_t)m.a1 + (size_t)(Sp + 4) * 4, 128);
  m.row_stride = (int)align_up((size_t)C * 4, 128);
  int st = (max_bytes - m.rows) / m.row_stride;
  st = st > 16 ? 16 : st;
  m.stages = st >= 4 ? (st / 4) * 4 : st;                // a multiple of four lets the consumer loop use static ring offsets
  m.total = m.rows + m.stages * m.row_stride;
  return m;
}

// One lattice step for the P state pairs of a consumer thread, branch-free.  `prev`/`cur` are the
// ping-pong lattice rows (this thread's pair j lives at index p2[j] = 2*pair), `xrow` the emission row
// of this frame in the smem ring, cx = (c, xb).  lneg[j] is kDead for pairs whose label state does
// not exist (pair >= L), which also keeps every dummy pair past the end of the lattice dead.
template <int P>
__device__ __forceinline__ void lattice_step(const float* __restrict__ prev, float* __restrict__ cur,
                                             const float* __restrict__ xrow, float2 cx, float* __restrict__ orow,
                                             const int (&xoff)[P], const bool (&skip)[P], const int (&p2)[P],
                                             const float (&lneg)[P], int* __restrict__ wmx_slot, int warp, int lane) {
  float vmax = kDead;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const float pm1 = prev[p2[j] - 1];
    const float2 pp = *reinterpret_cast<const float2*>(prev + p2[j]);       // (blank, label) of this pair
    const float xl = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(xrow) + xoff[j]);
    const float vb = cx.y + lse2_n(pp.x, pm1);
    const float vl = (fmaf(xl, kLog2e, -cx.x) + lneg[j]) + lse3_n(pp.y, pp.x, skip[j] ? pm1 : kDead);
    *reinterpret_cast<float2*>(cur + p2[j]) = make_float2(vb, vl);
    *reinterpret_cast<float2*>(orow + p2[j]) = make_float2(vb, vl);
    vmax = fmaxf(vmax, fmaxf(vb, vl));
  }
  const int wmax = __reduce_max_sync(0xffffffffu, f2ord(vmax));
  if (lane == 0) wmx_slot[warp] = wmax;
}

// One CTA per (sample, direction): `NTc` consumer threads + one producer warp (the last warp).
// Consumer thread i owns the state pairs p = i + j*NTc: the blank state 2p and the label state
// 2p+1, so the blank emission is a broadcast and only label states gather from the emission row.
// The producer warp keeps a ring of `stages` emission rows filled with TMA bulk copies (one 1-D
// copy per frame), waits for the next row so consumers never poll an mbarrier, reduces the
// per-warp maxima of step t-1 into the centring constant of step t+1 and accumulates the removed
// offsets in fp64.  One __syncthreads() per frame is the only CTA-wide synchronisation.
template <int P>
__global__ void __launch_bounds__(kLatThreads, 1)
ctc_lattice_kernel(const float* __restrict__ lp, int64_t sT, int64_t sN, int T, int C,
                   const int64_t* __restrict__ tgt, int64_t tgt_stride, int Lmax,
                   const int64_t* __restrict__ in_len, const int64_t* __restrict__ tgt_len, int blank,
                   float* __restrict__ nll, CtcScratch sc, LatSmem lay, int vec) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = blockIdx.y;
  const int N = gridDim.y;
  const int dir = blockIdx.x;             // 0 alpha, 1 beta
  const int NTc = blockDim.x - 64;        // consumer threads (the last two warps are helpers)
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const bool producer = tid >= NTc;
  const int n_cwarps = NTc >> 5;
  const int R = lay.stages;

  int L = (int)tgt_len[n];
  L = L < 0 ? 0 : (L > Lmax ? Lmax : L);
  int Tn = (int)in_len[n];
  Tn = Tn < 0 ? 0 : (Tn > T ? T : Tn);
  const int S = 2 * L + 1;

  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + lay.bars);
  float2* cxs = reinterpret_cast<float2*>(smem_raw + lay.cx);         // [4]
  int* wmx = reinterpret_cast<int*>(smem_raw + lay.wmx);              // [4][32]
  int* lab = reinterpret_cast<int*>(smem_raw + lay.lab);              // [Lp] direction-ordered labels
  float* a0 = reinterpret_cast<float*>(smem_raw + lay.a0) + 4;        // [-4 .. Sp)
  float* a1 = reinterpret_cast<float*>(smem_raw + lay.a1) + 4;
  unsigned char* rows = smem_raw + lay.rows;

  if (Tn == 0) {  // degenerate: empty input.  Feasible only for the empty target.
    if (tid == 0) {
      const double ll = (L == 0) ? 0.0 : -(double)CUDART_INF_F;
      sc.ll2[dir * N + n] = ll;
      if (dir == 0) nll[n] = (float)(-ll);
    }
    return;
  }

  for (int k = tid; k < L; k += blockDim.x) {
    const int src = dir ? (L - 1 - k) : k;
    lab[k] = (int)tgt[n * tgt_stride + src];
  }
  if (tid < 4) {
    a0[tid - 4] = kDead;
    a1[tid - 4] = kDead;
  }
  for (int i = tid; i < 128; i += blockDim.x) wmx[i] = f2ord(kDead);
  if (tid == 0) {
    for (int r = 0; r < R; ++r) mbar_init(&full[r], vec ? 1u : 32u);   // arrivals come from the tma warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const float* base = lp + n * sN;
  float* out = (dir ? sc.beta_rev : sc.alpha) + (int64_t)n * T * sc.Sp;
  const int64_t ostep = dir ? -(int64_t)sc.Sp : (int64_t)sc.Sp;      // output row advance per lattice step
  const int64_t xstep = dir ? -sT : sT;                               // input row advance per lattice step
  const int64_t t_first = dir ? (Tn - 1) : 0;

  if (producer) {
    // ------------------------------------------------------------------ two helper warps
    // warp `NTc/32`   (ctr): waits for the next emission row, turns the per-warp maxima of step t-1 into the
    //                        centring constant of step t+1, accumulates offsets, publishes (c, blank emission).
    // warp `NTc/32+1` (tma): keeps the emission ring full: after every frame barrier it re-arms the slot
    //                        that was just consumed with one TMA bulk copy.
    const bool is_tma = (tid - NTc) >= 32;
    const uint32_t row_bytes = (uint32_t)C * 4u;
    auto fill = [&](int slot, const float* src) {
      float* dst = reinterpret_cast<float*>(rows + (size_t)slot * lay.row_stride);
      if (vec) {
        if (lane == 0) {
          mbar_arrive_expect_tx(&full[slot], row_bytes);
          tma_row_g2s(dst, src, row_bytes, &full[slot]);
        }
      } else {
        for (int i = lane; i < C; i += 32) cp_async4(dst + i, src + i);
        cp_async_arrive(&full[slot]);
      }
    };
    if (is_tma) {
      const float* src_next = base + t_first * sT;    // row of the next lattice step to enqueue
      int t_fill = 0;
      for (; t_fill < R && t_fill < Tn; ++t_fill, src_next += xstep) fill(t_fill, src_next);
      int fill_slot = 0;                               // slot that frees up after the current step
      __syncthreads();                                 // "start"
      for (int t = 0; t < Tn; ++t) {
        __syncthreads();                               // end of step t
        if (t_fill < Tn) {                             // the slot consumed at step t is free for step t+R
          fill(fill_slot, src_next);
          src_next += xstep;
          ++t_fill;
        }
        if (++fill_slot == R) fill_slot = 0;
      }
      return;
    }
    double* offs = (dir ? sc.off_b : sc.off_a) + (int64_t)n * T;
    // step 0 constants
    mbar_wait(&full[0], 0);
    if (lane == 0) {
      cxs[0] = make_float2(0.0f, reinterpret_cast<const float*>(rows)[blank] * kLog2e);
      offs[t_first] = 0.0;
    }
    __syncthreads();                                   // "start": consumers may run step 0

    double off_acc = 0.0;                              // A_t: total offset removed up to step t
    float c_t = 0.0f, c_tm1 = 0.0f;                    // centring constants of steps t and t-1
    float m_tm2 = 0.0f;                                // centred maximum of step t-2
    bool have_tm2 = false;
    int wslot = 1 % R;                                 // ring slot of step t+1
    uint32_t wphase = (1 / R) & 1;
    int64_t tt_next = t_first + (dir ? -1 : 1);        // actual frame index of step t+1
    for (int t = 0; t < Tn; ++t) {
      if (t + 1 < Tn) {
        mbar_wait(&full[wslot], wphase);
        const float xbl = reinterpret_cast<const float*>(rows + (size_t)wslot * lay.row_stride)[blank];
        // Centring constant of step t+1.  The newest maxima available are those of step t-1 (lag 2), so
        // extrapolate the drift: aim A_{t+1} at M_{t-1} + 2*(M_{t-1} - M_{t-2}), all in small fp32 terms
        // relative to the running offset.  Any value is valid (it is only a shift); a good one keeps the
        // stored lattice near 0 where fp32 is densest.
        float c = 0.0f;
        if (t >= 1) {
          int v = (lane < n_cwarps) ? wmx[((t - 1) & 3) * 32 + lane] : f2ord(kDead);
          v = __reduce_max_sync(0xffffffffu, v);
          const float mt = ord2f(v);                   // centred maximum of step t-1
          if (mt > -1.0e29f) {                         // otherwise the whole lattice is dead (infeasible)
            const float drift = have_tm2 ? (mt - m_tm2 + c_tm1) : 0.0f;
            c = (mt - c_t) + 2.0f * drift;
            m_tm2 = mt;
            have_tm2 = true;
          } else {
            have_tm2 = false;
          }
        }
        c_tm1 = c_t;
        c_t = c;
        off_acc += (double)c;                          // warp-uniform
        if (lane == 0) {
          cxs[(t + 1) & 3] = make_float2(c, fmaf(xbl, kLog2e, -c));
          offs[tt_next] = off_acc;
        }
        if (++wslot == R) { wslot = 0; wphase ^= 1u; }
        tt_next += dir ? -1 : 1;
      }
      __syncthreads();                                 // end of step t
    }
    return;
  }

  // -------------------------------------------------------------------- consumer threads
  // Label grouping for the gradient pass (alpha CTA only): first occurrence + next occurrence.
  if (dir == 0) {
    for (int k = tid; k < L; k += NTc) {
      const int c = lab[k];
      int nxt = -1;
      for (int j = k + 1; j < L; ++j)
        if (lab[j] == c) { nxt = j; break; }
      int first = 1;
      for (int j = k - 1; j >= 0; --j)
        if (lab[j] == c) { first = 0; break; }
      sc.next_same[(int64_t)n * sc.Lp + k] = nxt;
      sc.leader[(int64_t)n * sc.Lp + k] = first;
    }
  }
  int xoff[P], p2[P];
  bool skip[P];
  float lneg[P];
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int p = tid + j * NTc;
    p2[j] = 2 * p;
    xoff[j] = blank * 4;
    skip[j] = false;
    lneg[j] = kDead;
    if (p < L) {
      xoff[j] = lab[p] * 4;
      skip[j] = (p >= 1) && (lab[p - 1] != lab[p]);
      lneg[j] = 0.0f;
    }
  }
  float* orow = out + t_first * sc.Sp;
  __syncthreads();                                     // "start": step-0 constants are in smem

  // step 0: only states 0 and 1 are reachable
  {
    const float* xrow = reinterpret_cast<const float*>(rows);
    const float2 cx = cxs[0];
    float vmax = kDead;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const bool first = (p2[j] == 0);
      const float vb = first ? cx.y : kDead;
      float vl = kDead;
      if (first && L > 0) vl = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(xrow) + xoff[j]) * kLog2e;
      *reinterpret_cast<float2*>(a0 + p2[j]) = make_float2(vb, vl);
      *reinterpret_cast<float2*>(orow + p2[j]) = make_float2(vb, vl);
      vmax = fmaxf(vmax, fmaxf(vb, vl));
    }
    const int wmax = __reduce_max_sync(0xffffffffu, f2ord(vmax));
    if (lane == 0) wmx[warp] = wmax;
    orow += ostep;
    __syncthreads();
  }
  // Frames 1..3 (and the tail) take the generic path; the main loop handles four frames per trip starting at a
  // frame t = 0 (mod 4), so the step slots, the ping-pong buffers and (because the ring depth is a multiple of
  // four) the four ring rows of a trip are all compile-time offsets from one pointer.
  const size_t rs = lay.row_stride;
  int t = 1;
#define DAE_LAT_STEP(PREV, CUR, SLOT, XROW)                                                               \
  lattice_step<P>(PREV, CUR, reinterpret_cast<const float*>(XROW), cxs[SLOT], orow, xoff, skip, p2, lneg, \
                  wmx + (SLOT) * 32, warp, lane);                                                          \
  orow += ostep;                                                                                           \
  __syncthreads();
  for (; t < Tn && (t & 3) != 0; ++t) {                // head: frames 1, 2, 3
    const unsigned char* xr = rows + (size_t)(t % R) * rs;
    if (t & 1) { DAE_LAT_STEP(a0, a1, t & 3, xr) } else { DAE_LAT_STEP(a1, a0, t & 3, xr) }
  }
  int sb = t % R;                                      // ring slot of frame t (a multiple of 4 from here on)
  for (; t + 3 < Tn; t += 4) {
    const unsigned char* xr = rows + (size_t)sb * rs;
    DAE_LAT_STEP(a1, a0, 0, xr)
    DAE_LAT_STEP(a0, a1, 1, xr + rs)
    DAE_LAT_STEP(a1, a0, 2, xr + 2 * rs)
    DAE_LAT_STEP(a0, a1, 3, xr + 3 * rs)
    sb += 4;
    if (sb == R) sb = 0;
  }
  for (; t < Tn; ++t) {                                // up to three tail frames
    const unsigned char* xr = rows + (size_t)(t % R) * rs;
    if (t & 1) { DAE_LAT_STEP(a0, a1, t & 3, xr) } else { DAE_LAT_STEP(a1, a0, t & 3, xr) }
  }
#undef DAE_LAT_STEP

  if (tid == 0) {
    const float* last = ((Tn - 1) & 1) ? a1 : a0;
    const float e1 = last[S - 1];
    const float e2 = (S > 1) ? last[S - 2] : kDead;
    // total offset = off_a/off_b of the last processed frame, written by the centring warp before the last barrier
    const double* offs = (dir ? sc.off_b : sc.off_a) + (int64_t)n * T;
    const double off_last = (Tn > 1) ? offs[dir ? 0 : (Tn - 1)] : 0.0;
    const float tail = lse2_n(e1, e2);
    const double ll2 = (tail < -1.0e29f) ? -(double)CUDART_INF_F : off_last + (double)tail;   // dead: infeasible
    sc.ll2[dir * N + n] = ll2;
    if (dir == 0) nll[n] = (float)(-ll2 * kLn2d);
  }
}

// One CTA per (t, n) row.
__global__ void __launch_bounds__(256)
ctc_grad_kernel(const float* __restrict__ lp, int64_t sT, int64_t sN, int T, int N, int C,
                const int64_t* __restrict__ tgt, int64_t tgt_stride, int Lmax,
                const int64_t* __restrict__ in_len, const int64_t* __restrict__ tgt_len, int blank,
                const float* __restrict__ gout, int64_t gout_stride, float* __restrict__ grad,
                CtcScratch sc, int vec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* xs = reinterpret_cast<float*>(smem_raw);            // [C] (padded to 4)
  float* gam = xs + ((C + 3) & ~3);                          // [Lp] occupancy per label position
  __shared__ float red[8];
  __shared__ float blank_sum_s;

  const int n = blockIdx.y, t = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
  float* orow = grad + ((int64_t)t * N + n) * C;
  int Tn = (int)in_len[n];
  Tn = Tn < 0 ? 0 : (Tn > T ? T : Tn);
  if (t >= Tn) {
    if (vec) for (int i = tid; i < (C >> 2); i += NT) st_stream4(orow + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
    else     for (int i = tid; i < C; i += NT) orow[i] = 0.0f;
    return;
  }
  int L = (int)tgt_len[n];
  L = L < 0 ? 0 : (L > Lmax ? Lmax : L);
  const int S = 2 * L + 1;
  const float* xrow = lp + t * sT + n * sN;
  if (vec) for (int i = tid; i < (C >> 2); i += NT) *reinterpret_cast<float4*>(xs + 4 * i) = ld_stream4(xrow + 4 * i);
  else     for (int i = tid; i < C; i += NT) xs[i] = ld_stream1(xrow + i);
  __syncthreads();

  // Occupancy of every lattice state at this frame:
  //   2^( alpha~ + beta~ + (off_a[t] + off_b[t] - ll2) - x ), all in log2 units; the
  //   bracket is formed in fp64 so the large offsets cancel exactly.
  const double ll2 = sc.ll2[n];
  const float kt = (float)(sc.off_a[(int64_t)n * T + t] + sc.off_b[(int64_t)n * T + t] - ll2);
  const float* arow = sc.alpha + ((int64_t)n * T + t) * sc.Sp;
  const float* brow = sc.beta_rev + ((int64_t)n * T + t) * sc.Sp;
  const int64_t* trow = tgt + n * tgt_stride;
  float bsum = 0.0f;
  for (int s = tid; s < S; s += NT) {
    const int c = (s & 1) ? (int)trow[s >> 1] : blank;
    const float e = (arow[s] + brow[S - 1 - s]) + (kt - xs[c] * kLog2e);
    const float g = fast_ex2(e);
    if (s & 1) gam[s >> 1] = g; else bsum += g;
  }
  bsum = warp_sum(bsum);
  if ((tid & 31) == 0) red[tid >> 5] = bsum;
  __syncthreads();
  if (tid == 0) {
    float s = 0.0f;
    for (int w = 0; w < (NT >> 5); ++w) s += red[w];
    blank_sum_s = s;
  }
  // dense part: exp(lp)
  for (int i = tid; i < C; i += NT) xs[i] = __expf(xs[i]);
  __syncthreads();
  // sparse correction, one writer per class, members added in label order
  const int32_t* nxt = sc.next_same + (int64_t)n * sc.Lp;
  const int32_t* lead = sc.leader + (int64_t)n * sc.Lp;
  for (int k = tid; k < L; k += NT) {
    if (lead[k]) {
      float s = 0.0f;
      for (int j = k; j >= 0; j = nxt[j]) s += gam[j];
      xs[(int)trow[k]] -= s;
    }
  }
  __syncthreads();
  if (tid == 0) xs[blank] -= blank_sum_s;
  __syncthreads();
  const float g = gout[n * gout_stride];
  if (vec) {
    for (int i = tid; i < (C >> 2); i += NT) {
      float4 v = *reinterpret_cast<const float4*>(xs + 4 * i);
      v.x *= g; v.y *= g; v.z *= g; v.w *= g;
      st_stream4(orow + 4 * i, v);
    }
  } else {
    for (int i = tid; i < C; i += NT) orow[i] = xs[i] * g;
  }
}

template <int P>
static int launch_lattice(int NT, const LatSmem& lay, int vec, cudaStream_t st, int N, const float* lp, int64_t sT,
                          int64_t sN, int T, int C, const int64_t* tgt, int64_t tgt_stride, int Lmax,
                          const int64_t* in_len, const int64_t* tgt_len, int blank, float* nll, const CtcScratch& sc) {
  DAE_CUDA(ensure_dyn_smem(ctc_lattice_kernel<P>, lay.total));
  ctc_lattice_kernel<P><<<dim3(2, N), NT, lay.total, st>>>(lp, sT, sN, T, C, tgt, tgt_stride, Lmax, in_len, tgt_len,
                                                          blank, nll, sc, lay, vec);
  DAE_LAUNCH_OK();
  return 0;
}

}  // namespace dae

extern "C" size_t dae_ctc_scratch_bytes(int T, int N, int Lmax) {
  if (T < 0 || N < 0 || Lmax < 0) return 0;
  dae::CtcScratch s;
  return dae::ctc_carve(s, nullptr, T, N, Lmax);
}

static int ctc_check(const float* lp, int T, int N, int C, const int64_t* tgt, int Lmax, const int64_t* in_len,
                     const int64_t* tgt_len, int blank, const void* scratch, size_t scratch_bytes) {
  if (!lp || !in_len || !tgt_len || T < 0 || N < 0 || C <= 0 || Lmax < 0 || blank < 0 || blank >= C) return DAE_E_BADARG;
  if (Lmax > 0 && !tgt) return DAE_E_BADARG;
  if (Lmax + 1 > (dae::kLatThreads - 64) * dae::kMaxPairsPerThread) return DAE_E_TOOBIG;
  if (!scratch || scratch_bytes < dae_ctc_scratch_bytes(T, N, Lmax)) return DAE_E_SCRATCH;
  if ((reinterpret_cast<uintptr_t>(scratch) & 255u) != 0) return DAE_E_ALIGN;
  return 0;
}

static int ctc_lattice_impl(const float* lp, int64_t sT, int64_t sN, int T, int N, int C, const int64_t* tgt,
                            int64_t tgt_stride, int Lmax, const int64_t* in_len, const int64_t* tgt_len, int blank,
                            float* nll, void* scratch, size_t scratch_bytes, void* stream, bool dense_follows) {
  using namespace dae;
  int rc = ctc_check(lp, T, N, C, tgt, Lmax, in_len, tgt_len, blank, scratch, scratch_bytes);
  if (rc) return rc;
  if (!nll) return DAE_E_BADARG;
  if (N == 0) return 0;
  CtcScratch sc;
  ctc_carve(sc, scratch, T, N, Lmax);
  if (sc.xfer)                                           // few samples, many frames: spread the time axis over the GPU
    return ctc_blocked_lattice(lp, sT, sN, T, N, C, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, nll, sc,
                               (cudaStream_t)stream, dense_follows);
  int P, NTc;
  lat_geometry(Lmax, P, NTc);
  const int NT = NTc + 64;                               // consumers + two helper warps
  const LatSmem lay = lat_smem_layout(C, sc.Lp, sc.Sp, kLatSmemBudget);
  if (lay.stages < 4) return DAE_E_TOOBIG;
  const int vec = aligned16(lp) && (C % 4 == 0) && (sT % 4 == 0) && (sN % 4 == 0);
  cudaStream_t st = (cudaStream_t)stream;
#define DAE_LAT(PP) return launch_lattice<PP>(NT, lay, vec, st, N, lp, sT, sN, T, C, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, nll, sc)
  if (P <= 1) DAE_LAT(1);
  if (P <= 2) DAE_LAT(2);
  DAE_LAT(4);
#undef DAE_LAT
}

extern "C" int dae_ctc_lattice(const float* lp, int64_t sT, int64_t sN, int T, int N, int C, const int64_t* tgt,
                               int64_t tgt_stride, int Lmax, const int64_t* in_len, const int64_t* tgt_len,
                               int blank, float* nll, void* scratch, size_t scratch_bytes, void* stream) {
  return ctc_lattice_impl(lp, sT, sN, T, N, C, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, nll, scratch,
                          scratch_bytes, stream, false);
}

extern "C" int dae_ctc_grad(const float* lp, int64_t sT, int64_t sN, int T, int N, int C, const int64_t* tgt,
                            int64_t tgt_stride, int Lmax, const int64_t* in_len, const int64_t* tgt_len, int blank,
                            const float* nll, const float* gout, int64_t gout_stride, float* grad,
                            const void* scratch, size_t scratch_bytes, void* stream) {
  using namespace dae;
  int rc = ctc_check(lp, T, N, C, tgt, Lmax, in_len, tgt_len, blank, scratch, scratch_bytes);
  if (rc) return rc;
  if (!nll || !gout || !grad) return DAE_E_BADARG;
  if (N == 0 || T == 0) return 0;
  CtcScratch sc;
  ctc_carve(sc, const_cast<void*>(scratch), T, N, Lmax);
  const int vec_g = aligned16(lp) && aligned16(grad) && (C % 4 == 0) && (sT % 4 == 0) && (sN % 4 == 0);
  if (sc.xfer) {                                         // blocked path: alpha/beta rows are rebuilt per block
    rc = ctc_blocked_grad(lp, sT, sN, T, N, C, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, gout, gout_stride, grad,
                          sc, vec_g, (cudaStream_t)stream);
    if (rc != 1) return rc;
  }
  const size_t smem = (size_t)((C + 3) & ~3) * 4 + (size_t)sc.Lp * 4;
  if (smem > 200 * 1024) return DAE_E_TOOBIG;
  if (smem > 48 * 1024)
    DAE_CUDA(ensure_dyn_smem(ctc_grad_kernel, (int)smem));
  const int vec = aligned16(lp) && aligned16(grad) && (C % 4 == 0) && (sT % 4 == 0) && (sN % 4 == 0);
  int work = (C / 4 > 2 * Lmax + 1) ? C / 4 : 2 * Lmax + 1;
  int NT = ((work + 31) / 32) * 32;
  NT = NT < 32 ? 32 : (NT > 256 ? 256 : NT);
  ctc_grad_kernel<<<dim3(T, N), NT, smem, (cudaStream_t)stream>>>(lp, sT, sN, T, N, C, tgt, tgt_stride, Lmax, in_len,
                                                                tgt_len, blank, gout, gout_stride, grad, sc, vec);
  DAE_LAUNCH_OK();
  return 0;
}

// Loss and gradient in one call, for callers that know the upstream gradient when they ask for the loss.
extern "C" int dae_ctc_loss_grad(const float* lp, int64_t sT, int64_t sN, int T, int N, int C, const int64_t* tgt,
                                 int64_t tgt_stride, int Lmax, const int64_t* in_len, const int64_t* tgt_len,
                                 int blank, float* nll, const float* gout, int64_t gout_stride, float* grad,
                                 void* scratch, size_t scratch_bytes, void* stream) {
  using namespace dae;
  if (!gout || !grad) return DAE_E_BADARG;
  int rc = ctc_check(lp, T, N, C, tgt, Lmax, in_len, tgt_len, blank, scratch, scratch_bytes);
  if (rc) return rc;
  CtcScratch sc;
  ctc_carve(sc, scratch, T, N, Lmax);
  const int vec_g = aligned16(lp) && aligned16(grad) && (C % 4 == 0) && (sT % 4 == 0) && (sN % 4 == 0);
  const bool split = N > 0 && T > 0 && ctc_split_fits(sc, N, vec_g);
  rc = ctc_lattice_impl(lp, sT, sN, T, N, C, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, nll, scratch, scratch_bytes,
                        stream, split);
  if (rc || N == 0 || T == 0) return rc;
  if (split) {
    // bit 1: the label-class kernel is launched as a dependent of the dense kernel, loads its inputs while the scan
    // runs and waits for the scan's CTAs itself; then the dense kernel need not outlive the scan
    const bool early = (ctc_config().overlap.load(std::memory_order_relaxed) & 2) != 0;
    rc = ctc_blocked_dense(lp, sT, sN, T, N, C, in_len, gout, gout_stride, grad, (cudaStream_t)stream, !early);
    if (rc) return rc;
    return ctc_blocked_grad(lp, sT, sN, T, N, C, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, gout, gout_stride, grad,
                            sc, vec_g, (cudaStream_t)stream, true, early ? ctc_scan_ctas(sc, N) : 0);
  }
  return dae_ctc_grad(lp, sT, sN, T, N, C, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, nll, gout, gout_stride, grad,
                      scratch, scratch_bytes, stream);
}

namespace dae {
// grad[t,n,:] *= gout[n] / hint wherever that ratio is not exactly 1 (the gradient was formed in the forward call
// with the caller's expected upstream scale; in the adapt loop the ratio is 1 and every CTA returns after one load).
__global__ void __launch_bounds__(256)
ctc_rescale_kernel(float* __restrict__ grad, int T, int N, int C, const float* __restrict__ gout, int64_t gout_stride,
                   float hint) {
  const int n = blockIdx.y;
  const float g = gout[(int64_t)n * gout_stride];
  if (g == hint) return;
  const float r = g / hint;
  for (int t = blockIdx.x; t < T; t += gridDim.x) {
    float* row = grad + ((int64_t)t * N + n) * C;
    for (int c = threadIdx.x; c < C; c += 256) row[c] *= r;
  }
}
}  // namespace dae

extern "C" int dae_ctc_rescale(float* grad, int T, int N, int C, const float* gout, int64_t gout_stride, float hint,
                               void* stream) {
  using namespace dae;
  if (!grad || !gout || T < 0 || N < 0 || C <= 0 || !(hint != 0.0f)) return DAE_E_BADARG;
  if (T == 0 || N == 0) return 0;
  const int gx = T < kNumSMs * 4 ? T : kNumSMs * 4;
  ctc_rescale_kernel<<<dim3(gx, N), 256, 0, (cudaStream_t)stream>>>(grad, T, N, C, gout, gout_stride, hint);
  DAE_LAUNCH_OK();
  return 0;
}
