// Time-blocked CTC loss + gradient for few samples and many frames (the dynamic-eval adapt step: N=1, T=2048).
// Same contract as ctc.cu behind dae_ctc_lattice / dae_ctc_grad (torch.nn.CTCLoss semantics; call sites
// lcasr/lib.py:492,570-579, AWMC :324-331).
//
// The per-frame chain needs T dependent steps on two SMs.  The recursion is linear in the (log-sum, +)
// semiring, so it is cut into blocks of K frames and spread over the whole GPU:
//   dae_ctc_lattice
//   1. ctc_xfer_kernel        every (block, source state) in parallel: the K-frame transfer band
//                             X_b[s][d] = log2 sum over paths that start in state s just before the block and
//                             end in state s+d on its last frame (d <= 2K).  One thread per source, the band
//                             lives in registers, no synchronisation.  Also: label grouping for the gradient,
//                             zeroing of the scan's hand-over words.
//   2. ctc_boundary_kernel    the only sequential part, T/K steps: boundary vectors
//                             alpha_end(b)[s'] = LSE_d X_b[s'-d][d] + alpha_end(b-1)[s'-d]   and, with the same
//                             bands read the other way, betahat_start(b)[s] = LSE_d X_b[s][d] + betahat_start(b+1)[s+d].
//                             States are split into 64-state regions, one CTA each; a region only needs the
//                             last 2K values of the region below it, handed over as tagged 8-byte words (tag in
//                             the data, no fences) through the receiver's shared memory inside a thread-block
//                             cluster and through global memory across clusters: a skewed pipeline of regions.
//                             Leaves the log-likelihood.
//   dae_ctc_grad
//   3. ctc_block_grad_kernel  every block in parallel: its K alpha rows and K beta rows are rebuilt in shared
//                             memory from the boundary vectors and turned into the K gradient rows at once.
//      (ctc_fill_kernel + ctc_grad_kernel do the same through the scratch buffer when a block's rows do not
//      fit in shared memory.)
// Values are log2 units with the finite dead-state sentinel of ctc_shared.cuh; every region / row carries an
// fp64 offset so stored fp32 values stay O(1).
#include <cooperative_groups.h>
#include "ctc_shared.cuh"

namespace dae {

// generic address space: the word lives in global memory or in the shared memory of a CTA of the same cluster
__device__ __forceinline__ int2 ld_tagged(const int2* p) {
  int2 v;
  asm volatile("ld.volatile.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_tagged(int2* p, int2 v) {
  asm volatile("st.volatile.v2.s32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu_i32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void cp_async_f32(float* dst_smem, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

__device__ __forceinline__ void clamp_lengths(const int64_t* in_len, const int64_t* tgt_len, int n, int T, int Lmax,
                                              int& Tn, int& L) {
  L = (int)tgt_len[n];
  L = L < 0 ? 0 : (L > Lmax ? Lmax : L);
  Tn = (int)in_len[n];
  Tn = Tn < 0 ? 0 : (Tn > T ? T : Tn);
}

// ------------------------------------------------------------------------------------------ 1. transfer bands
constexpr int kXferTile = 256;                    // source states per CTA

// K lattice steps from the unit vector at source state s.  v[j] is the band entry of state s+j; PAR is the
// parity of s, so the blank/label kind of every entry is a compile-time property of (PAR, j) and a warp never
// diverges.  Entries above 2(k-1) are still dead before step k; those terms are dropped at compile time.
template <int K, int PAR>
__device__ __forceinline__ void xfer_steps(float (&v)[2 * K + 1], const float* __restrict__ es, int es_stride,
                                           const unsigned char* __restrict__ skp, int jl, int kb) {
#pragma unroll
  for (int k = 1; k <= K; ++k) {
    if (k <= kb) {
      const float* e = es + (k - 1) * es_stride + jl;
#pragma unroll
      for (int j = 2 * k; j >= 0; --j) {
        const bool label = ((PAR + j) & 1) != 0;
        const bool has0 = j <= 2 * (k - 1);                       // v[j] may be alive
        const bool has1 = j >= 1 && (j - 1) <= 2 * (k - 1);       // v[j-1] may be alive
        const bool has2 = label && j >= 2;                        // skip transition exists for label states
        const float ev = e[j];
        float acc;
        if (has2) {
          const float c2 = skp[jl + j] ? v[j - 2] : kDead;
          if (has0) acc = lse3_n(v[j], v[j - 1], c2);
          else if (has1) acc = lse2_n(v[j - 1], c2);
          else acc = c2;
        } else if (has0 && has1) {
          acc = lse2_n(v[j], v[j - 1]);
        } else if (has0) {
          acc = v[j];
        } else if (has1) {
          acc = v[j - 1];
        } else {
          acc = kDead;
        }
        v[j] = ev + acc;
      }
    }
  }
}

template <int K>
__global__ void __launch_bounds__(kXferTile)
ctc_xfer_kernel(const float* __restrict__ lp, int64_t sT, int64_t sN, int T, const int64_t* __restrict__ tgt,
                int64_t tgt_stride, int Lmax, const int64_t* __restrict__ in_len,
                const int64_t* __restrict__ tgt_len, int blank, CtcScratch sc) {
  constexpr int W = 2 * K + 1, EW = kXferTile + 2 * K;
  __shared__ float es[K][EW];                    // emission (log2) of state s0+j at frame t0+k
  __shared__ unsigned char skp[EW];              // 1: state s0+j may be entered from s0+j-2
  __shared__ __align__(16) float outs[kXferTile * W];
  const int b = blockIdx.x, s0 = blockIdx.y * kXferTile, n = blockIdx.z, tid = threadIdx.x;
  {
    // the scan's global hand-over words are matched by tag (step number): they start from zero on every call
    const size_t n_words = (size_t)2 * gridDim.z * (sc.nblk + 1) * sc.G * kHaloWords;
    const size_t n_thr = (size_t)gridDim.x * gridDim.y * gridDim.z * kXferTile;
    const size_t me = ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * kXferTile + tid;
    for (size_t w = me; w < n_words; w += n_thr) sc.halo[w] = make_int2(0, 0);
    if (me < 64) sc.sync[me] = 0;
  }
  int Tn, L;
  clamp_lengths(in_len, tgt_len, n, T, Lmax, Tn, L);
  const int S = 2 * L + 1;
  const int nb = (Tn + K - 1) / K;
  if (b >= nb) return;                           // bands of unused blocks are never read
  const int t0 = b * K;
  const int kb = min(K, Tn - t0);
  const int64_t* trow = tgt + n * tgt_stride;
  const float* base = lp + n * sN + (int64_t)t0 * sT;
  for (int j = tid; j < EW; j += kXferTile) {
    const int s = s0 + j;
    int cls = blank;
    bool sk = false;
    if ((s & 1) && s < S) {
      cls = (int)trow[s >> 1];
      sk = (s >= 3) && ((int)trow[(s >> 1) - 1] != cls);
    }
    skp[j] = sk ? 1 : 0;
#pragma unroll
    for (int k = 0; k < K; ++k) es[k][j] = (k < kb && s < S) ? base[k * sT + cls] * kLog2e : kDead;
  }
  __syncthreads();
  // the gradient pass reads the emissions of its frames from here (coalesced, no label indirection)
  if (s0 + tid < sc.Sq) {
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (k < kb) sc.emis[((size_t)n * sc.nblk * K + (size_t)(t0 + k)) * sc.Sq + s0 + tid] = es[k][tid];
  }
  // warps 0-3 own the even sources of the tile, warps 4-7 the odd ones
  const int par = tid >> 7;
  const int jl = 2 * (tid & 127) + par;
  float v[W];
  v[0] = 0.0f;
#pragma unroll
  for (int j = 1; j < W; ++j) v[j] = kDead;
  if (par == 0) xfer_steps<K, 0>(v, &es[0][0], EW, skp, jl, kb);
  else          xfer_steps<K, 1>(v, &es[0][0], EW, skp, jl, kb);
  // bands leave as rows [source][d] (what the boundary scan bulk-copies), staged so the stores are 128-bit
  __syncthreads();
#pragma unroll
  for (int d = 0; d < W; ++d) outs[jl * W + d] = v[d];
  __syncthreads();
  const int ncols = min(kXferTile, sc.Sq - s0);
  float4* dst = reinterpret_cast<float4*>(sc.xfer + ((size_t)(n * sc.nblk + b) * sc.Sq + s0) * W);
  const float4* src = reinterpret_cast<const float4*>(outs);
  for (int i = tid; i < ncols * W / 4; i += kXferTile) dst[i] = src[i];
  if (blockIdx.y == 0) {
    // Label grouping for the gradient pass (first occurrence + next occurrence of each class), spread over the
    // sample's time blocks: this CTA takes label positions [k0, k1), one per warp, lanes scan the sequence.
    const int lane = tid & 31, warp = tid >> 5;
    const int per = (L + nb - 1) / nb;
    const int k0 = b * per, k1 = min(L, k0 + per);
    for (int k = k0 + warp; k < k1; k += kXferTile / 32) {
      const int c = (int)trow[k];
      int nxt = 0x7fffffff, before = 0;
      for (int j = lane; j < L; j += 32) {
        const bool same = (int)trow[j] == c;
        if (same && j > k) nxt = min(nxt, j);
        if (same && j < k) before = 1;
      }
      nxt = __reduce_min_sync(0xffffffffu, nxt);
      before = __reduce_max_sync(0xffffffffu, before);
      if (lane == 0) {
        sc.next_same[(int64_t)n * sc.Lp + k] = (nxt == 0x7fffffff) ? -1 : nxt;
        sc.leader[(int64_t)n * sc.Lp + k] = before ? 0 : 1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ 2. boundary scan
constexpr int kBndStages = 4;                     // transfer-band prefetch depth (steps)
constexpr int kTpd = 2;                           // consumer threads per destination state (4: slower, 1: no shuffles but 2x the stream)
constexpr int kTermsPerThread = (2 * kBlkK + 1 + kTpd - 1) / kTpd;   // they share the 2K+1 terms of the state
constexpr int kBndThreads = kTpd * kRegion + 96;  // consumers + band-prefetch, hand-over and frame warps

// grid (G, 2, N): region g of direction dir (0 alpha, 1 beta in reversed state order u = S-1-s) of sample n.
// Two consumer threads per destination state and three helper warps, all meeting at one barrier per step:
//   tma   warp: keeps kBndStages band chunks in flight (one TMA bulk copy per step) and waits for the next
//               step's chunk before the step barrier, so consumers never poll an mbarrier;
//   halo  warp: fetches the 2K values (and the frame) the region below produced for the same vector; the word
//               of the next step is requested a step ahead so its L2 round trip stays off the critical path;
//   frame warp: chooses the fp64 frame of reference of the vector two steps ahead and publishes it.
// Halo values stay relative to the lower region's frame and are shifted when used.
template <int K>
__global__ void __launch_bounds__(kBndThreads)
ctc_boundary_kernel(int T, int Lmax, const int64_t* __restrict__ in_len, const int64_t* __restrict__ tgt_len,
                    float* __restrict__ nll, CtcScratch sc, int release_dependents) {
  constexpr int W = 2 * K + 1, H = 2 * K, CH = (kRegion + H) * W;
  __shared__ __align__(128) float xs[kBndStages][CH];   // band rows of the states this region reads, per step
  __shared__ __align__(8) float buf[2][H + kRegion + 2];   // [halo of the region below (raw) | own region | dead cell, pad]
  __shared__ double hoff[2];                            // offset the halo values of each buffer are relative to
  __shared__ int hmaxs[2];                              // their maximum (order-preserving int)
  __shared__ float oshift[2];                           // per step: shift of own values into the new frame
  __shared__ double Fd[2];                              // per step: the new frame itself
  __shared__ uint64_t full[kBndStages];
  __shared__ __align__(8) int2 hstage[4][kHaloWords];   // cluster-boundary regions: hand-over words staged by cp.async
  const int g = blockIdx.x, dir = blockIdx.y, n = blockIdx.z, N = gridDim.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Hand-over inside a thread-block cluster goes through the receiver's shared memory (DSMEM): the region below
  // writes its tagged words straight into this mailbox and the hand-over warp polls locally.  Regions at a
  // cluster boundary (and launches without clusters) use the global mailbox.
  extern __shared__ __align__(16) int2 mbox[];          // [nblk+1][kHaloWords], slot = vector number
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks();
  const int crank = (int)cluster.block_rank();
  if (CL > 1) {
    for (int k = tid; k < (sc.nblk + 1) * kHaloWords; k += blockDim.x) mbox[k] = make_int2(0, 0);
    cluster.sync();
  }
  // Launched with programmatic stream serialization: everything above overlaps the band kernel's tail; its
  // results (bands, zeroed hand-over words, label groups) are visible from here on.
  cudaGridDependencySynchronize();
  // every CTA of the scan is resident now: the dependent launch that follows (ctc_dense_grad_kernel, which waits
  // for this grid before it completes) may take the idle SMs
  if (release_dependents) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (g >= sc.G) return;                                // grid padded to whole clusters
  const bool up_local = CL > 1 && crank > 0 && g > 0;
  const bool down_local = CL > 1 && crank + 1 < CL && g + 1 < sc.G;
  int2* mbox_down = down_local ? cluster.map_shared_rank(mbox, crank + 1) : nullptr;
  const bool consumer = tid < kTpd * kRegion;
  const bool tma_warp = warp == kTpd * kRegion / 32, halo_warp = warp == kTpd * kRegion / 32 + 1,
             frame_warp = warp == kTpd * kRegion / 32 + 2;
  int Tn, L;
  clamp_lengths(in_len, tgt_len, n, T, Lmax, Tn, L);
  const int S = 2 * L + 1;
  const int nb = (Tn + K - 1) / K;
  if (Tn == 0) {                                 // empty input: feasible only for the empty target
    if (g == 0 && tid == 0) {
      const double ll = (L == 0) ? 0.0 : -(double)CUDART_INF_F;
      sc.ll2[dir * N + n] = ll;
      if (dir == 0) nll[n] = (float)(-ll);
    }
    if (tid == 0) {
      __threadfence();
      atomicAdd(sc.sync, 1);
    }
    return;
  }
  const int Sq = sc.Sq, G = sc.G;
  const size_t vec0 = (size_t)(dir * N + n) * (sc.nblk + 1);
  float* brow0 = sc.bound + vec0 * Sq;
  double* boff0 = sc.boff + vec0 * G;
  int2* halo0 = sc.halo + vec0 * G * kHaloWords;
  const float* xf0 = sc.xfer + (size_t)n * sc.nblk * Sq * W;
  const int u0 = g * kRegion;
  const int i = (tid / kTpd) & (kRegion - 1), h = tid % kTpd;
  const int u = u0 + i;

  // Band rows this region reads each step, as one contiguous chunk of xfer[b]:
  //   alpha: sources u0-H .. u0+R-1 (region 0 has no sources below 0: its first H rows in smem stay dead);
  //   beta : rows of s = S-1-u for the region's u, i.e. s_hi-R+1 .. s_hi, clipped at 0, start aligned to 4 rows.
  int col0, ncols, dst_off;
  if (dir == 0) {
    col0 = g > 0 ? u0 - H : 0;
    ncols = g > 0 ? kRegion + H : kRegion;
    dst_off = g > 0 ? 0 : H * W;
  } else {
    const int s_hi = S - 1 - u0;
    const int s_lo = max(s_hi - (kRegion - 1), 0);
    col0 = s_lo & ~3;
    ncols = s_hi >= 0 ? ((s_hi - col0 + 1 + 3) & ~3) : 0;
    dst_off = 0;
  }
  const bool has_band = ncols > 0;               // a beta region entirely past the lattice stays dead
  const uint32_t chunk_bytes = (uint32_t)ncols * W * 4u;

  for (int k = tid; k < kBndStages * CH; k += blockDim.x) (&xs[0][0])[k] = kDead;
  if (tid == 0) {
    for (int r = 0; r < kBndStages; ++r) mbar_init(&full[r], 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    hoff[0] = 0.0;
    hoff[1] = 0.0;
    hmaxs[0] = f2ord(kDead);
    hmaxs[1] = f2ord(kDead);
  }
  // initial vector: alpha starts from the unit vector at state 0 (the virtual frame -1), betahat of the last
  // frame is 0 on the two final states (u = 0, 1).
  {
    const int idx = dir ? nb : 0;
    if (consumer) {
      float v0 = kDead;
      if (u == 0 || (dir == 1 && u == 1 && S > 1)) v0 = 0.0f;
      if (h == 0) {
        buf[0][H + i] = v0;
        brow0[(size_t)idx * Sq + u] = v0;
      }
      if (tid < H) {
        buf[0][tid] = kDead;
        buf[1][tid] = kDead;
      }
      if (tid < 2) buf[tid][H + kRegion] = kDead;
      if (tid == 0) boff0[(size_t)idx * G + g] = 0.0;
    }
  }
  __syncthreads();

  if (tma_warp) {
    // ---------------------------------------------------------------------------------- band prefetch
    auto issue = [&](int st) {
      if (st < nb && has_band && lane == 0) {
        const int b = dir ? (nb - 1 - st) : st;
        const int slot = st % kBndStages;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses of the slot
        mbar_arrive_expect_tx(&full[slot], chunk_bytes);
        tma_row_g2s(&xs[slot][dst_off], xf0 + ((size_t)b * Sq + col0) * W, chunk_bytes, &full[slot]);
      }
    };
    for (int st = 0; st < kBndStages - 1; ++st) issue(st);
    if (has_band) mbar_wait(&full[0], 0u);
    __syncthreads();                             // "start"
    for (int st = 0; st < nb; ++st) {
      issue(st + kBndStages - 1);                // its slot was consumed in step st-1
      if (st + 1 < nb && has_band) mbar_wait(&full[(st + 1) % kBndStages], (uint32_t)((st + 1) / kBndStages) & 1u);
      __syncthreads();                           // end of step st
    }
    return;
  }
  if (halo_warp) {
    // ------------------------------------------------------------------------------ hand-over from region g-1
    // Fetches the 2K values, the frame and the maximum of the halo of V_{st+1} while the consumers work on step st.
    const int64_t pstep = up_local ? (int64_t)kHaloWords : (dir ? -1 : 1) * (int64_t)G * kHaloWords;
    const int2* hp = up_local ? (mbox + kHaloWords + lane)
                              : (halo0 + ((size_t)(dir ? (nb - 1) : 1) * G + (g > 0 ? g - 1 : 0)) * kHaloWords + lane);
    const bool polls = g > 0 && lane < kHaloWords;
    if (up_local || g == 0) {
      // -- mailbox in this CTA's shared memory (written by the region below through DSMEM): polling is cheap
      __syncthreads();                           // "start"
      for (int st = 0; st < nb; ++st) {
        if (g > 0) {
          int2 w = make_int2(0, 0);
          if (polls) {
            do { w = ld_tagged(hp); } while (w.y != st + 1);
          }
          const int lo = __shfl_sync(0xffffffffu, w.x, H), hi = __shfl_sync(0xffffffffu, w.x, H + 1);
          const float hv = (lane < H) ? __int_as_float(w.x) : kDead;
          if (lane < H) buf[(st + 1) & 1][lane] = hv;
          const int hm = __reduce_max_sync(0xffffffffu, f2ord(hv));
          if (lane == 0) {
            hoff[(st + 1) & 1] = __hiloint2double(hi, lo);
            hmaxs[(st + 1) & 1] = hm;
          }
          hp += pstep;
        }
        __syncthreads();                         // end of step st
      }
      return;
    }
    // -- mailbox in global memory (cluster boundary).  An L2 round trip is longer than a step, and a load in
    // flight is waited for at the step barrier, so the words are fetched with cp.async into a small staging
    // ring kAhead steps before they are consumed (asynchronous copies are tracked by their own groups, not by
    // the barrier), and the region starts kLead steps late so that those requests always find their words.
    constexpr int kAhead = 3, kLead = 6;
    {
      const int v = nb < kLead ? nb : kLead;
      const int2* late = hp + (int64_t)(v - 1) * pstep;
      if (polls) while (ld_tagged(late).y != v) { }
      __syncwarp();
    }
    auto request = [&](int v) {                  // stage the words of vector v (1-based) into slot v & 3
      if (v <= nb && polls)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(&hstage[v & 3][lane])),
                     "l"(hp + (int64_t)(v - 1) * pstep) : "memory");
      cp_async_commit();
    };
    for (int v = 1; v <= kAhead; ++v) request(v);
    __syncthreads();                             // "start"
    for (int st = 0; st < nb; ++st) {
      request(st + 1 + kAhead);                  // its slot held vector st+1-... consumed a step ago
      cp_async_wait<kAhead>();                   // vector st+1 has landed
      int2 w = make_int2(0, 0);
      if (polls) {
        w = ld_tagged(&hstage[(st + 1) & 3][lane]);
        const int2* direct = hp + (int64_t)st * pstep;
        while (w.y != st + 1) w = ld_tagged(direct);     // not published yet when requested: poll directly
      }
      const int lo = __shfl_sync(0xffffffffu, w.x, H), hi = __shfl_sync(0xffffffffu, w.x, H + 1);
      const float hv = (lane < H) ? __int_as_float(w.x) : kDead;
      if (lane < H) buf[(st + 1) & 1][lane] = hv;
      const int hm = __reduce_max_sync(0xffffffffu, f2ord(hv));
      if (lane == 0) {
        hoff[(st + 1) & 1] = __hiloint2double(hi, lo);
        hmaxs[(st + 1) & 1] = hm;
      }
      __syncthreads();                           // end of step st
    }
    return;
  }
  if (frame_warp) {
    // ------------------------------------------------------------------------------------ frames of reference
    // Stored values of vector V_v are relative to the fp64 frame F(v) of this region.  F(st+2) is fixed here, in
    // the shadow of step st, from what is known when the step starts: the maximum of the own part of V_st and of
    // its halo, each extrapolated two steps by its drift.  Consumers never compute an offset: they read the own
    // shift F(st)-F(st+1) and form the halo shift from F(st+1).  Any frame is correct; a good one keeps stored
    // values near 0.  A region with nothing alive takes the lower region's frame, so the first values that flow
    // in are shifted by one step's drift only.
    const int64_t vstep = dir ? -1 : 1;
    double* bo = boff0 + (size_t)(dir ? (nb - 1) : 1) * G + g;          // frame of V_1
    int2* ho_words = (down_local ? (mbox_down + kHaloWords) : (halo0 + ((size_t)(dir ? (nb - 1) : 1) * G + g) * kHaloWords)) +
                     H + (lane & 1);
    const int64_t ho_step = down_local ? (int64_t)kHaloWords : vstep * G * kHaloWords;
    const bool pub = g + 1 < G && lane < 2;
    // Frames are fp64 sums of fp32 steps; all decisions are made on small fp32 quantities relative to the
    // newest frame F1 = F(st+1), so the per-step chain is a dozen fp32 instructions and two fp64 adds.
    double F1 = 0.0;                              // frame of V_{st+1}
    float s01 = 0.0f;                             // F(st) - F(st+1)
    float s10 = 0.0f;                             // F(st-1) - F(st)
    float m0p = 0.0f, hmp = 0.0f, hrelp = 0.0f;   // previous own max / halo max / halo frame relative to F(st)
    bool own_had = false, halo_had = false;
    if (lane == 0) {
      oshift[0] = 0.0f;
      Fd[0] = 0.0;
      *bo = 0.0;
    }
    if (pub) st_tagged(ho_words, make_int2(0, 1));
    __syncthreads();                             // "start"
    for (int st = 0; st < nb; ++st) {
      // own maximum of V_st (relative to F(st)), straight from the vector the consumers left in smem
      const float2 pv = *reinterpret_cast<const float2*>(&buf[st & 1][H + 2 * lane]);
      float m0 = fmaxf(pv.x, pv.y);
#pragma unroll
      for (int q = 1; q < kRegion / 64; ++q) {
        const float2 pq = *reinterpret_cast<const float2*>(&buf[st & 1][H + 64 * q + 2 * lane]);
        m0 = fmaxf(m0, fmaxf(pq.x, pq.y));
      }
      m0 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(m0)));
      const float hm = ord2f(hmaxs[st & 1]);      // halo maximum of V_st, relative to the lower region's frame
      const float hrel = (float)(hoff[st & 1] - F1);
      const bool own_alive = m0 > -1.0e29f, halo_alive = hm > -1.0e29f;
      // true maxima relative to F1, extrapolated two steps by their drift
      const float ao = m0 + s01;
      const float co = own_had ? fmaf(2.0f, (m0 - m0p) - s10, ao) : ao;
      const float ah = hm + hrel;
      const float ch = halo_had ? fmaf(2.0f, ah - (hmp + hrelp), ah) : ah;
      float delta = (g > 0) ? hrel : 0.0f;        // nothing alive: take the lower region's frame
      if (own_alive) delta = co;
      if (halo_alive) delta = own_alive ? fmaxf(co, ch) : ch;
      const double F2 = F1 + (double)delta;
      m0p = m0; own_had = own_alive;
      hmp = hm; hrelp = hrel - delta; halo_had = halo_alive;     // relative to F2, the next step's F1
      if (st + 1 < nb) {
        bo += vstep * G;
        ho_words += ho_step;
        if (lane == 0) {
          oshift[(st + 1) & 1] = -delta;
          Fd[(st + 1) & 1] = F2;
          *bo = F2;
        }
        if (pub) st_tagged(ho_words, make_int2(lane ? __double2hiint(F2) : __double2loint(F2), st + 2));
      }
      s10 = s01;
      s01 = -delta;
      F1 = F2;
      __syncthreads();                           // end of step st
    }
    return;
  }

  // -------------------------------------------------------------------------------------- consumers
  // Per-thread constants of the term loop: this thread owns terms d = h*(K+1) + k.  Smem addresses are
  // precomputed for ring slot 0 / buffer 0; the step adds a compile-time offset.  Thread h=1 has one term less:
  // its last slot points at a permanently dead cell.
  const int srow = S - 1 - u;                     // beta: band row of this destination
  const bool forced_dead = (dir == 1 && srow < 0);
  const float* xp[kTermsPerThread];
  const float* pp[kTermsPerThread];
  float hmask[kTermsPerThread];                   // 1 where the term's source lies in the halo
#pragma unroll
  for (int k = 0; k < kTermsPerThread; ++k) {
    const int d = h * kTermsPerThread + k;
    const bool valid = d <= W - 1;
    const int dd = valid ? d : 0;
    const int xo = dir ? ((forced_dead ? 0 : srow - col0) * W + dd) : ((i + H - dd) * W + dd);
    const int po = valid ? (H + i - dd) : (H + kRegion);
    xp[k] = &xs[0][xo];
    pp[k] = &buf[0][po];
    hmask[k] = (valid && po < H) ? 1.0f : 0.0f;
  }
  const bool uses_halo = i < H;                  // destinations 0..2K-1 of the region reach below it (whole warps)
  const int64_t vstep = dir ? -1 : 1;            // boundary vector index advance per step
  float* bout = brow0 + (size_t)(dir ? (nb - 1) : 1) * Sq + u;
  int2* hout = (down_local ? (mbox_down + kHaloWords) : (halo0 + ((size_t)(dir ? (nb - 1) : 1) * G + g) * kHaloWords)) +
               (i - (kRegion - H));
  const int64_t hout_step = down_local ? (int64_t)kHaloWords : vstep * G * kHaloWords;
  const bool pub_val = h == 0 && i >= kRegion - H && g + 1 < G;
  constexpr int BS = H + kRegion + 2;            // buffer stride
  __syncthreads();                               // "start"
  // one step: V_st (buffer PAR) -> V_{st+1} (buffer PAR^1), band chunk in ring slot SLOT
#define DAE_BND_STEP(SLOT, PAR)                                                                            \
  {                                                                                                        \
    const float osh = oshift[PAR];                                                                         \
    float term[kTermsPerThread];                                                                           \
    if (uses_halo) {                                                                                       \
      const float dsh = (float)(hoff[PAR] - Fd[PAR]) - osh;                                                \
      _Pragma("unroll") for (int k = 0; k < kTermsPerThread; ++k)                                          \
        term[k] = (xp[k][(SLOT) * CH] + pp[k][(PAR) * BS]) + fmaf(hmask[k], dsh, osh);                     \
    } else {                                                                                               \
      _Pragma("unroll") for (int k = 0; k < kTermsPerThread; ++k)                                          \
        term[k] = xp[k][(SLOT) * CH] + pp[k][(PAR) * BS];                                                  \
    }                                                                                                      \
    float mx = term[0];                                                                                    \
    _Pragma("unroll") for (int k = 1; k < kTermsPerThread; ++k) mx = fmaxf(mx, term[k]);                   \
    _Pragma("unroll") for (int o = 1; o < kTpd; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); \
    float s0 = 0.0f, s1 = 0.0f;                                                                            \
    _Pragma("unroll") for (int k = 0; k < kTermsPerThread; ++k) {                                          \
      const float e = fast_ex2(term[k] - mx);                                                              \
      if (k & 1) s1 += e; else s0 += e;                                                                    \
    }                                                                                                      \
    float sum = s0 + s1;                                                                                   \
    _Pragma("unroll") for (int o = 1; o < kTpd; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);   \
    float val = mx + fast_lg2(sum);                                                                        \
    if (!uses_halo) val += osh;                                                                            \
    if (forced_dead) val = kDead;                                                                          \
    if (h == 0) {                                                                                          \
      buf[(PAR) ^ 1][H + i] = val;                                                                         \
      *bout = val;                                                                                         \
    }                                                                                                      \
    if (pub_val) st_tagged(hout, make_int2(__float_as_int(val), st + 1));                                  \
    bout += vstep * Sq;                                                                                    \
    hout += hout_step;                                                                                     \
    ++st;                                                                                                  \
    __syncthreads();                                                                                       \
  }
  int st = 0;
  for (; st + 3 < nb;) {
    DAE_BND_STEP(0, 0)
    DAE_BND_STEP(1, 1)
    DAE_BND_STEP(2, 0)
    DAE_BND_STEP(3, 1)
  }
  if (st < nb) DAE_BND_STEP(0, 0)
  if (st < nb) DAE_BND_STEP(1, 1)
  if (st < nb) DAE_BND_STEP(2, 0)
#undef DAE_BND_STEP

  // log-likelihood from the final vector: alpha ends on the last two states, betahat(-1) starts on state 0 (u = S-1)
  if (h == 0 && u == S - 1) {
    const float* fin = buf[nb & 1];
    const float e1 = fin[H + i];
    const bool in_halo = i == 0;                   // state S-2 belongs to the region below
    const float e2 = (dir == 0 && S > 1) ? fin[H + i - 1] : kDead;
    // the region offset lives in warp 0; everyone can re-read it from the boundary array
    const int idx_fin = dir ? 0 : nb;
    const double off_fin = boff0[(size_t)idx_fin * G + g];
    const float e2s = in_halo ? e2 + (float)(hoff[nb & 1] - off_fin) : e2;
    const float tail = lse2_n(e1, e2s);
    const double ll2 = (tail < -1.0e29f) ? -(double)CUDART_INF_F : off_fin + (double)tail;
    sc.ll2[dir * N + n] = ll2;
    if (dir == 0) nll[n] = (float)(-ll2 * kLn2d);
  }
  // This CTA is done: count it for a gradient kernel that waits for the scan without a kernel boundary
  // (ctc_block_grad_kernel, `wait_scan`).  The consumers meet at their own barrier (the helper warps have left),
  // then one fence makes everything this CTA wrote - boundary vectors, frames, the likelihood - visible first.
  asm volatile("bar.sync 1, %0;" ::"n"(kTpd * kRegion) : "memory");
  if (tid == 0) {
    __threadfence();
    atomicAdd(sc.sync, 1);
  }
}

// ------------------------------------------------------------------------------------------ 3. block fill
struct FillSmem { int lab, a0, a1, wmx, red, total; };
__host__ __device__ inline FillSmem fill_smem_layout(int Lp, int Sp) {
  FillSmem m;
  m.red = 0;                                             // double[32]
  m.wmx = m.red + 32 * 8;                                // int[2][32]
  m.lab = m.wmx + 2 * 32 * 4;
  m.a0 = (int)align_up((size_t)m.lab + (size_t)Lp * 4, 16);
  m.a1 = m.a0 + (Sp + 4) * 4;
  m.total = m.a1 + (Sp + 4) * 4;
  return m;
}

// grid (nblk, 2, N): K ordinary lattice steps of block b in direction dir, starting from the boundary vector
// the scan left (alpha: the vector before the block; beta: betahat of the block's last frame).
template <int P, int K, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
ctc_fill_kernel(const float* __restrict__ lp, int64_t sT, int64_t sN, int T, const int64_t* __restrict__ tgt,
                int64_t tgt_stride, int Lmax, const int64_t* __restrict__ in_len,
                const int64_t* __restrict__ tgt_len, int blank, CtcScratch sc, FillSmem lay) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x, dir = blockIdx.y, n = blockIdx.z, N = gridDim.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NTc = blockDim.x, nw = NTc >> 5;
  int Tn, L;
  clamp_lengths(in_len, tgt_len, n, T, Lmax, Tn, L);
  const int nb = (Tn + K - 1) / K;
  if (b >= nb) return;
  const int t0 = b * K, kb = min(K, Tn - t0), t1 = t0 + kb - 1;

  double* red = reinterpret_cast<double*>(smem_raw + lay.red);
  int* wmx = reinterpret_cast<int*>(smem_raw + lay.wmx);
  int* lab = reinterpret_cast<int*>(smem_raw + lay.lab);
  float* a0 = reinterpret_cast<float*>(smem_raw + lay.a0) + 4;
  float* a1 = reinterpret_cast<float*>(smem_raw + lay.a1) + 4;

  for (int k = tid; k < L; k += NTc) lab[k] = (int)tgt[n * tgt_stride + (dir ? (L - 1 - k) : k)];
  if (tid < 4) {
    a0[tid - 4] = kDead;
    a1[tid - 4] = kDead;
  }
  __syncthreads();
  int p2[P];
  bool skip[P];
  float xl[P][K], xb[K];
  const float* base = lp + n * sN;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int p = tid + j * NTc;
    p2[j] = 2 * p;
    int cls = blank;
    float lneg = kDead;
    skip[j] = false;
    if (p < L) {
      cls = lab[p];
      skip[j] = (p >= 1) && (lab[p - 1] != cls);
      lneg = 0.0f;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int tk = dir ? (t1 - k) : (t0 + k);
      xl[j][k] = (k < kb) ? fmaf(base[(int64_t)tk * sT + cls], kLog2e, lneg) : kDead;
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int tk = dir ? (t1 - k) : (t0 + k);
    xb[k] = (k < kb) ? base[(int64_t)tk * sT + blank] * kLog2e : kDead;
  }

  // boundary vector: every 128-state region carries its own offset; re-base all of it on the largest value
  const int idx_in = dir ? (b + 1) : b;
  const size_t vec = (size_t)(dir * N + n) * (sc.nblk + 1) + idx_in;
  const float* brow = sc.bound + vec * sc.Sq;
  const double* boffs = sc.boff + vec * sc.G;
  float2 v[P];
  double roff[P];
  double cand = -CUDART_INF;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    v[j] = *reinterpret_cast<const float2*>(brow + p2[j]);
    roff[j] = boffs[p2[j] / kRegion];
    const float m = fmaxf(v[j].x, v[j].y);
    if (m > -1.0e29f) cand = fmax(cand, roff[j] + (double)m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cand = fmax(cand, __shfl_xor_sync(0xffffffffu, cand, o));
  if (lane == 0) red[warp] = cand;
  __syncthreads();
  double ref = -CUDART_INF;
  for (int w = 0; w < nw; ++w) ref = fmax(ref, red[w]);
  if (ref == -CUDART_INF) ref = 0.0;             // nothing alive (infeasible): any offset will do
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const float sh = (float)(roff[j] - ref);
    v[j].x = (v[j].x > -1.0e29f) ? v[j].x + sh : kDead;
    v[j].y = (v[j].y > -1.0e29f) ? v[j].y + sh : kDead;
  }

  float* out = (dir ? sc.beta_rev : sc.alpha) + (int64_t)n * T * sc.Sp;
  double* offs = (dir ? sc.off_b : sc.off_a) + (int64_t)n * T;
  double off = ref;
  float* prev = a0;
  float* cur = a1;
  int k0 = 0;
  if (dir) {
    // beta of the block's last frame = its emission + betahat: no transition
    float vmax = kDead;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const float2 r = make_float2(xb[0] + v[j].x, xl[j][0] + v[j].y);
      *reinterpret_cast<float2*>(prev + p2[j]) = r;
      *reinterpret_cast<float2*>(out + (int64_t)t1 * sc.Sp + p2[j]) = r;
      vmax = fmaxf(vmax, fmaxf(r.x, r.y));
    }
    const int wm = __reduce_max_sync(0xffffffffu, f2ord(vmax));
    if (lane == 0) wmx[warp] = wm;                 // slot 0: read by step k = 1
    if (tid == 0) offs[t1] = off;
    k0 = 1;
  } else {
    float vmax = kDead;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      *reinterpret_cast<float2*>(prev + p2[j]) = v[j];
      vmax = fmaxf(vmax, fmaxf(v[j].x, v[j].y));
    }
    const int wm = __reduce_max_sync(0xffffffffu, f2ord(vmax));
    if (lane == 0) wmx[32 + warp] = wm;            // slot 1: read by step k = 0
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if (k >= k0 && k < kb) {
      const int tk = dir ? (t1 - k) : (t0 + k);
      // centre on the maximum of the previous row
      int mv = (lane < nw) ? wmx[((k + 1) & 1) * 32 + lane] : f2ord(kDead);
      mv = __reduce_max_sync(0xffffffffu, mv);
      const float mp = ord2f(mv);
      const float c = (mp > -1.0e29f) ? mp : 0.0f;
      off += (double)c;
      float vmax = kDead;
      float* orow = out + (int64_t)tk * sc.Sp;
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const float pm1 = prev[p2[j] - 1];
        const float2 pp = *reinterpret_cast<const float2*>(prev + p2[j]);
        const float vb = (xb[k] - c) + lse2_n(pp.x, pm1);
        const float vl = (xl[j][k] - c) + lse3_n(pp.y, pp.x, skip[j] ? pm1 : kDead);
        *reinterpret_cast<float2*>(cur + p2[j]) = make_float2(vb, vl);
        *reinterpret_cast<float2*>(orow + p2[j]) = make_float2(vb, vl);
        vmax = fmaxf(vmax, fmaxf(vb, vl));
      }
      const int wm = __reduce_max_sync(0xffffffffu, f2ord(vmax));
      if (lane == 0) wmx[(k & 1) * 32 + warp] = wm;
      if (tid == 0) offs[tk] = off;
      float* tmp = prev; prev = cur; cur = tmp;
      __syncthreads();
    }
  }
}

template <int P, int MAXT, int MINB>
static int launch_fill(int NTc, cudaStream_t st, int N, const float* lp, int64_t sT, int64_t sN, int T,
                       const int64_t* tgt, int64_t tgt_stride, int Lmax, const int64_t* in_len,
                       const int64_t* tgt_len, int blank, const CtcScratch& sc) {
  const FillSmem lay = fill_smem_layout(sc.Lp, sc.Sp);
  if (lay.total > 200 * 1024) return DAE_E_TOOBIG;
  auto kern = ctc_fill_kernel<P, kBlkK, MAXT, MINB>;
  if (lay.total > 48 * 1024)
    DAE_CUDA(ensure_dyn_smem(kern, lay.total));
  kern<<<dim3(sc.nblk, 2, N), NTc, lay.total, st>>>(lp, sT, sN, T, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, sc,
                                                    lay);
  DAE_LAUNCH_OK();
  return 0;
}

// Rows of alpha / beta for every frame, from the boundary vectors (only needed when the fused gradient kernel does
// not fit: ctc_grad_kernel reads them from the scratch).
int ctc_blocked_fill(const float* lp, int64_t sT, int64_t sN, int T, int N, const int64_t* tgt, int64_t tgt_stride,
                     int Lmax, const int64_t* in_len, const int64_t* tgt_len, int blank, const CtcScratch& sc,
                     cudaStream_t st) {
  int P, NTc;
  lat_geometry(Lmax, P, NTc);
#define DAE_FILL(PP, MT, MB) \
  return launch_fill<PP, MT, MB>(NTc, st, N, lp, sT, sN, T, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, sc)
  constexpr int kMaxC = kLatThreads - 64;
  if (P <= 1 && NTc <= 320) DAE_FILL(1, 320, 4);       // short label sequences: several blocks per SM
  if (P <= 1 && NTc <= 640) DAE_FILL(1, 640, 2);
  if (P <= 1) DAE_FILL(1, kMaxC, 1);
  if (P <= 2) DAE_FILL(2, kMaxC, 1);
  DAE_FILL(4, kMaxC, 1);
#undef DAE_FILL
}

// ------------------------------------------------------------------------------------------ 4. fused block fill + gradient
struct BgSmem { int red, wmx, offs, bsum, xbs, lab, nxt, rowsA, rowsB, gam, row_stride, total; };
__host__ __device__ inline BgSmem bg_smem_layout(int Lp, int Sp, int K) {
  BgSmem m;
  m.red = 0;                                             // double[2][32]
  m.offs = m.red + 2 * 32 * 8;                           // double[2][K]
  m.wmx = m.offs + 2 * K * 8;                            // int[2 dirs][2 slots][32]
  m.bsum = m.wmx + 4 * 32 * 4;                           // float[32][K]: per-warp blank occupancy sums
  m.xbs = m.bsum + 32 * K * 4;                           // float[K]: blank emission (log2) per frame
  m.lab = m.xbs + K * 4;
  m.nxt = m.lab + Lp * 4;                                // int[Lp]: next label position with the same class
  m.row_stride = Sp + 4;                                 // floats; 4 dead cells in front of every row
  m.rowsA = (int)align_up((size_t)m.nxt + (size_t)Lp * 4, 16);
  m.rowsB = m.rowsA + K * m.row_stride * 4;
  m.gam = m.rowsB + K * m.row_stride * 4;                // float[K][Lp]: label emissions, then label occupancies
  m.total = m.gam + K * Lp * 4;
  return m;
}

// grid (nblk, N): everything the gradient of block b needs, without a round trip through HBM.
//   0. dense part of the block's K gradient rows, g*exp(lp), streamed straight from lp to grad;
//   1. the K alpha rows and K beta rows of the block, rebuilt in shared memory from the scan's boundary vectors.
//      The two recurrences are independent, so every thread advances its alpha pair(s) and its beta pair(s) in the
//      same step: twice the work per barrier on a latency-bound chain;
//   2. occupancies 2^(alpha~ + beta~ + (offA + offB - ll2) - x) of every state of every row; per class sums by
//      walking the first-occurrence lists (deterministic order), then the few classes that occur in the label
//      sequence (and blank) are overwritten with g*(exp(lp) - occupancy).
template <int P, int K, int MAXT, int MINB, bool SPARSE>
__global__ void __launch_bounds__(MAXT, MINB)
ctc_block_grad_kernel(const float* __restrict__ lp, int64_t sT, int64_t sN, int T, int C,
                      const int64_t* __restrict__ tgt, int64_t tgt_stride, int Lmax,
                      const int64_t* __restrict__ in_len, const int64_t* __restrict__ tgt_len, int blank,
                      const float* __restrict__ gout, int64_t gout_stride, float* __restrict__ grad, CtcScratch sc,
                      BgSmem lay, int vec, int wait_scan) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x, n = blockIdx.y, N = gridDim.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NTc = blockDim.x, nw = NTc >> 5;
  int Tn, L;
  clamp_lengths(in_len, tgt_len, n, T, Lmax, Tn, L);
  const int S = 2 * L + 1;
  const int t0 = b * K;
  const int kb = max(0, min(K, Tn - t0));        // frames of this block inside the sample
  const int krows = max(0, min(K, T - t0));      // gradient rows this CTA owns

  double* red = reinterpret_cast<double*>(smem_raw + lay.red);
  double* offs = reinterpret_cast<double*>(smem_raw + lay.offs);      // [0..K): alpha rows, [K..2K): beta rows
  int* wmx = reinterpret_cast<int*>(smem_raw + lay.wmx);
  float* bsum_s = reinterpret_cast<float*>(smem_raw + lay.bsum);
  float* xbs = reinterpret_cast<float*>(smem_raw + lay.xbs);
  int* lab = reinterpret_cast<int*>(smem_raw + lay.lab);
  int* nxt = reinterpret_cast<int*>(smem_raw + lay.nxt);
  float* rowsA = reinterpret_cast<float*>(smem_raw + lay.rowsA) + 4;
  float* rowsB = reinterpret_cast<float*>(smem_raw + lay.rowsB) + 4;
  float* gam = reinterpret_cast<float*>(smem_raw + lay.gam);
  const int RS = lay.row_stride, Lp = sc.Lp;
  const float* base = lp + n * sN;
  const float g = gout[n * gout_stride];

  // ---- 0. dense part (zeros for the padding frames of a short sample): all K rows of a column chunk are loaded
  // before any is stored, so a thread keeps K independent 128-bit loads in flight.
  // `occ` (shared memory, [K/2][Cq]) holds, per row of the half batch, the occupancy of every class: then the rows
  // leave complete, as 128-bit stores only; without it the caller overwrites the few classes that occur later.
  constexpr int KH = K / 2;                        // half batches: K/2 independent 128-bit loads in flight
  const int Cq = (C + 3) & ~3;
  auto dense_half = [&](int hb, const float* occ) {
    for (int i = tid; i < (C >> 2); i += NTc) {
      float4 v4[KH];
#pragma unroll
      for (int k = 0; k < KH; ++k)
        if (hb * KH + k < kb) v4[k] = __ldcs(reinterpret_cast<const float4*>(base + (int64_t)(t0 + hb * KH + k) * sT) + i);
#pragma unroll
      for (int k = 0; k < KH; ++k) {
        const int kk = hb * KH + k;
        if (kk < krows) {
          float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kk < kb) {
            float4 oc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (occ) oc = *reinterpret_cast<const float4*>(occ + k * Cq + 4 * i);
            o4 = make_float4((__expf(v4[k].x) - oc.x) * g, (__expf(v4[k].y) - oc.y) * g, (__expf(v4[k].z) - oc.z) * g,
                             (__expf(v4[k].w) - oc.w) * g);
          }
          __stcs(reinterpret_cast<float4*>(grad + ((int64_t)(t0 + kk) * N + n) * C) + i, o4);
        }
      }
    }
  };
  auto dense = [&]() {
    if (vec) {
      dense_half(0, nullptr);
      dense_half(1, nullptr);
    } else {
      for (int k = 0; k < krows; ++k) {
        float* orow = grad + ((int64_t)(t0 + k) * N + n) * C;
        const float* xrow = base + (int64_t)(t0 + k) * sT;
        for (int i = tid; i < C; i += NTc) orow[i] = (k < kb) ? __expf(xrow[i]) * g : 0.0f;
      }
    }
  };
  // With room for the occupancy table (the alpha/beta rows are dead by then) the rows are written once, at the
  // end.  Otherwise:
  // half of the grid streams first and runs its chains later, the other half the other way round, so the two
  // CTAs that share an SM overlap memory traffic with the latency-bound chains
  // SPARSE: ctc_dense_grad_kernel has written g*exp(lp) (and the zero rows) already, under the scan's shadow; only
  // the classes of the label sequence and blank are left to do.
  const bool table = !SPARSE && vec && (size_t)KH * Cq <= (size_t)2 * K * lay.row_stride;
  const bool dense_first = !SPARSE && !table && ((int)blockIdx.x * 2 < (int)gridDim.x || kb == 0);
  if (kb == 0) {
    if (!SPARSE) dense();
    return;
  }

  // ---- 1a. labels, list links, emissions of the block's frames at every label state.  All global loads of the
  // prologue are issued before the dense stream starts, so their latency hides behind it.
  int p2[P];
  bool lead_q[P];
  float xla[P][K];                               // emissions (log2) of this thread's label states, per frame
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int p = tid + j * NTc;
    p2[j] = 2 * p;
    lead_q[j] = (p < L) ? (sc.leader[(int64_t)n * Lp + p] != 0) : false;
    const float* em = sc.emis + ((size_t)n * sc.nblk * K + (size_t)t0) * sc.Sq + (2 * p + 1);   // written by ctc_xfer_kernel
#pragma unroll
    for (int k = 0; k < K; ++k)
      xla[j][k] = (k < kb && p < L) ? __ldcg(em + (size_t)k * sc.Sq) : kDead;
  }
  const float xb_mine = (tid < kb) ? __ldcg(sc.emis + ((size_t)n * sc.nblk * K + (size_t)(t0 + tid)) * sc.Sq) : kDead;
  if (dense_first) dense();
  for (int k = tid; k < L; k += NTc) {
    lab[k] = (int)tgt[n * tgt_stride + k];
    nxt[k] = sc.next_same[(int64_t)n * Lp + k];
  }
  for (int r = tid; r < 4 * 2 * K; r += NTc) {
    float* row = (r < 4 * K) ? rowsA + (r >> 2) * RS : rowsB + ((r >> 2) - K) * RS;
    row[(r & 3) - 4] = kDead;
  }
  if (tid < K) xbs[tid] = xb_mine;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int p = tid + j * NTc;
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (p < Lp) gam[k * Lp + p] = xla[j][k];   // the beta chain reads them in reversed label order
  }
  // `wait_scan` (SPARSE, launched as a programmatic dependent of the dense gradient kernel): the CTA has been
  // resident and has loaded everything above while the scan was still running.  From here on it needs the scan's
  // results, and its stores must follow the dense kernel's: wait for that grid, then for the scan's CTAs to have
  // counted themselves (one poller per CTA; the scan never waits for anything, so this ends).
  if (SPARSE && wait_scan > 0) {
    cudaGridDependencySynchronize();
    if (tid == 0) {
      int spins = 0;
      while (ld_acquire_gpu_i32(sc.sync) < wait_scan && ++spins < (1 << 22)) __nanosleep(128);
    }
    __syncthreads();
  }
  const double ll2 = sc.ll2[n];
  // ---- 1b. boundary vectors of both directions: every region carries its own frame; re-base on the largest value
  float2 v[2][P];
  double refd[2];
  {
    double roff[2][P], cand[2];
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      const size_t vecidx = (size_t)(dir * N + n) * (sc.nblk + 1) + (dir ? (b + 1) : b);
      const float* brow = sc.bound + vecidx * sc.Sq;
      const double* boffs = sc.boff + vecidx * sc.G;
      cand[dir] = -CUDART_INF;
#pragma unroll
      for (int j = 0; j < P; ++j) {
        v[dir][j] = *reinterpret_cast<const float2*>(brow + p2[j]);
        roff[dir][j] = boffs[p2[j] / kRegion];
        const float m = fmaxf(v[dir][j].x, v[dir][j].y);
        if (m > -1.0e29f) cand[dir] = fmax(cand[dir], roff[dir][j] + (double)m);
      }
    }
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cand[dir] = fmax(cand[dir], __shfl_xor_sync(0xffffffffu, cand[dir], o));
      if (lane == 0) red[dir * 32 + warp] = cand[dir];
    }
    __syncthreads();                             // also: lab, nxt, xbs, pads, label emissions are in smem
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      // maximum over the warps' candidates: one value per lane, five shuffle steps (max is exact: every warp
      // arrives at the same double)
      double ref = (lane < nw) ? red[dir * 32 + lane] : -CUDART_INF;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ref = fmax(ref, __shfl_xor_sync(0xffffffffu, ref, o));
      if (ref == -CUDART_INF) ref = 0.0;         // nothing alive (infeasible): any offset will do
      refd[dir] = ref;
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const float sh = (float)(roff[dir][j] - ref);
        v[dir][j].x = (v[dir][j].x > -1.0e29f) ? v[dir][j].x + sh : kDead;
        v[dir][j].y = (v[dir][j].y > -1.0e29f) ? v[dir][j].y + sh : kDead;
      }
    }
  }
  // ---- 1c. the two chains, one alpha step and one beta step per barrier.
  // alpha: start vector -> rowsA[0] -> ... -> rowsA[kb-1].  beta (reversed state order): the last frame's row is
  // emission + betahat (no transition), then rowsB[kb-2] ... rowsB[0].  The alpha start vector borrows rowsB[0],
  // which beta writes last; with a single frame it gets its own barrier below.
  bool skipA[P], skipB[P];
  int qb[P];                                     // label position of this thread's beta pair
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int p = p2[j] >> 1;
    skipA[j] = (p >= 1 && p < L) && (lab[p - 1] != lab[p]);
    qb[j] = L - 1 - p;
    skipB[j] = (p >= 1 && p < L) && (lab[L - p] != lab[L - 1 - p]);
  }
  float* initA = rowsB;
  double offA = refd[0], offB = refd[1];
  {
    float vmA = kDead, vmB = kDead;
    float* firstB = rowsB + (kb - 1) * RS;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      *reinterpret_cast<float2*>(initA + p2[j]) = v[0][j];
      vmA = fmaxf(vmA, fmaxf(v[0][j].x, v[0][j].y));
    }
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const float xlb = (qb[j] >= 0) ? gam[(kb - 1) * Lp + qb[j]] : kDead;
      const float2 r = make_float2(xbs[kb - 1] + v[1][j].x, xlb + v[1][j].y);
      if (kb > 1) *reinterpret_cast<float2*>(firstB + p2[j]) = r;
      v[1][j] = r;                               // kept for the kb == 1 case
      vmB = fmaxf(vmB, fmaxf(r.x, r.y));
    }
    const int wa = __reduce_max_sync(0xffffffffu, f2ord(vmA));
    const int wb = __reduce_max_sync(0xffffffffu, f2ord(vmB));
    if (lane == 0) {
      wmx[32 + warp] = wa;                       // alpha slot 1: read by step k = 0
      wmx[64 + warp] = wb;                       // beta slot 0: read by step k = 1
    }
    if (tid == 0) offs[K + kb - 1] = offB;
  }
  __syncthreads();
  const float* prevA = initA;
  const float* prevB = rowsB + (kb - 1) * RS;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if (k < kb) {
      // alpha step k (frame t0+k) and beta step k (frame t1-k; k = 0 was the emission-only row)
      const bool do_b = k >= 1;
      int mva = (lane < nw) ? wmx[((k + 1) & 1) * 32 + lane] : f2ord(kDead);
      int mvb = (lane < nw && do_b) ? wmx[64 + ((k + 1) & 1) * 32 + lane] : f2ord(kDead);
      mva = __reduce_max_sync(0xffffffffu, mva);
      mvb = __reduce_max_sync(0xffffffffu, mvb);
      const float ma = ord2f(mva), mb = ord2f(mvb);
      const float ca = (ma > -1.0e29f) ? ma : 0.0f, cb = (mb > -1.0e29f) ? mb : 0.0f;
      offA += (double)ca;
      offB += (double)cb;
      const int rb = kb - 1 - k;
      float* curA = rowsA + k * RS;
      float* curB = rowsB + rb * RS;
      const float xbA = xbs[k] - ca, xbB = xbs[do_b ? rb : 0] - cb;
      float vmA = kDead, vmB = kDead;
      float2 ra[P], rbv[P];
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const float pm1 = prevA[p2[j] - 1];
        const float2 pp = *reinterpret_cast<const float2*>(prevA + p2[j]);
        ra[j].x = xbA + lse2_n(pp.x, pm1);
        ra[j].y = (xla[j][k] - ca) + lse3_n(pp.y, pp.x, skipA[j] ? pm1 : kDead);
        vmA = fmaxf(vmA, fmaxf(ra[j].x, ra[j].y));
        if (do_b) {
          const float qm1 = prevB[p2[j] - 1];
          const float2 qq = *reinterpret_cast<const float2*>(prevB + p2[j]);
          const float xlb = (qb[j] >= 0) ? gam[rb * Lp + qb[j]] : kDead;
          rbv[j].x = xbB + lse2_n(qq.x, qm1);
          rbv[j].y = (xlb - cb) + lse3_n(qq.y, qq.x, skipB[j] ? qm1 : kDead);
          vmB = fmaxf(vmB, fmaxf(rbv[j].x, rbv[j].y));
        }
      }
      if (kb == 1) {                             // single frame: beta's row takes the place of the alpha start vector
        __syncthreads();
#pragma unroll
        for (int j = 0; j < P; ++j) *reinterpret_cast<float2*>(rowsB + p2[j]) = v[1][j];
      }
#pragma unroll
      for (int j = 0; j < P; ++j) {
        *reinterpret_cast<float2*>(curA + p2[j]) = ra[j];
        if (do_b) *reinterpret_cast<float2*>(curB + p2[j]) = rbv[j];
      }
      const int wa = __reduce_max_sync(0xffffffffu, f2ord(vmA));
      const int wb = __reduce_max_sync(0xffffffffu, f2ord(vmB));
      if (lane == 0) {
        wmx[(k & 1) * 32 + warp] = wa;
        if (do_b) wmx[64 + (k & 1) * 32 + warp] = wb;
      }
      if (tid == 0) {
        offs[k] = offA;
        if (do_b) offs[K + rb] = offB;
      }
      prevA = curA;
      if (do_b) prevB = curB;
      __syncthreads();
    }
  }

  if (!SPARSE && !dense_first && !table) dense();
  __syncthreads();                               // dense stores of every thread precede the sparse overwrite

  // ---- 2. occupancies and the sparse correction (gam: label emissions are replaced by label occupancies)
  float bs[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    bs[k] = 0.0f;
    if (k < kb) {
      const float kt = (float)(offs[k] + offs[K + k] - ll2);
      const float* arow = rowsA + k * RS;
      const float* brw = rowsB + k * RS;
      const float xbk = xbs[k];
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const int sb = p2[j];                      // blank state 2p, label state 2p+1
        if (sb < S) bs[k] += fast_ex2((arow[sb] + brw[S - 1 - sb]) + (kt - xbk));
      }
    }
    bs[k] = warp_sum(bs[k]);
    if (lane == 0) bsum_s[warp * K + k] = bs[k];
  }
  __syncthreads();                               // every beta step has read its emissions from gam
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if (k < kb) {
      const float kt = (float)(offs[k] + offs[K + k] - ll2);
      const float* arow = rowsA + k * RS;
      const float* brw = rowsB + k * RS;
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const int sb = p2[j];
        if (sb + 1 < S) gam[k * Lp + (sb >> 1)] = fast_ex2((arow[sb + 1] + brw[S - 2 - sb]) + (kt - xla[j][k]));
      }
    }
  }
  __syncthreads();
  // per class occupancy: the first occurrence of a class walks its list (members in label order)
  float acc[P][K];
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int q = p2[j] >> 1;
#pragma unroll
    for (int k = 0; k < K; ++k) acc[j][k] = 0.0f;
    if (lead_q[j]) {
      for (int m = q; m >= 0; m = nxt[m]) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[j][k] += gam[k * Lp + m];
      }
    }
  }
  if (table) {
    // alpha/beta rows are dead: their space becomes the occupancy table of a half batch of rows
    float* occ = reinterpret_cast<float*>(smem_raw + lay.rowsA);
    __syncthreads();                             // every thread is done with the rows
    for (int i = tid; i < KH * Cq; i += NTc) occ[i] = 0.0f;
    __syncthreads();
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
#pragma unroll
      for (int j = 0; j < P; ++j) {
        if (lead_q[j]) {
          const int c = lab[p2[j] >> 1];
#pragma unroll
          for (int k = 0; k < KH; ++k) occ[k * Cq + c] = acc[j][hb * KH + k];
        }
      }
      if (tid < KH) {
        float sacc = 0.0f;
        for (int w = 0; w < nw; ++w) sacc += bsum_s[w * K + hb * KH + tid];
        occ[tid * Cq + blank] = sacc;
      }
      __syncthreads();
      dense_half(hb, occ);
      __syncthreads();                           // the next half batch overwrites the same table entries
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < P; ++j) {
    if (lead_q[j]) {
      const int c = lab[p2[j] >> 1];
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (k < kb) grad[((int64_t)(t0 + k) * N + n) * C + c] = (fast_ex2(xla[j][k]) - acc[j][k]) * g;
    }
  }
  if (tid < kb) {
    float sacc = 0.0f;
    for (int w = 0; w < nw; ++w) sacc += bsum_s[w * K + tid];
    grad[((int64_t)(t0 + tid) * N + n) * C + blank] = (fast_ex2(xbs[tid]) - sacc) * g;
  }
}

template <int P, int MAXT, int MINB, bool SPARSE>
static int launch_block_grad(int NTc, cudaStream_t st, int N, const float* lp, int64_t sT, int64_t sN, int T, int C,
                             const int64_t* tgt, int64_t tgt_stride, int Lmax, const int64_t* in_len,
                             const int64_t* tgt_len, int blank, const float* gout, int64_t gout_stride, float* grad,
                             const CtcScratch& sc, const BgSmem& lay, int vec, int wait_scan) {
  auto kern = ctc_block_grad_kernel<P, kBlkK, MAXT, MINB, SPARSE>;
  DAE_CUDA(ensure_dyn_smem(kern, lay.total));
  if (SPARSE && wait_scan > 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sc.nblk, N);
    cfg.blockDim = dim3(NTc);
    cfg.dynamicSmemBytes = lay.total;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DAE_CUDA(cudaLaunchKernelEx(&cfg, kern, lp, sT, sN, T, C, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, gout,
                                gout_stride, grad, sc, lay, vec, wait_scan));
    DAE_LAUNCH_OK();
    return 0;
  }
  kern<<<dim3(sc.nblk, N), NTc, lay.total, st>>>(lp, sT, sN, T, C, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, gout,
                                                 gout_stride, grad, sc, lay, vec, 0);
  DAE_LAUNCH_OK();
  return 0;
}

// Gradient of the blocked path.  Returns 1 when the shape does not fit the fused kernel's shared memory (the
// caller then runs ctc_fill_kernel + ctc_grad_kernel).
int ctc_blocked_grad(const float* lp, int64_t sT, int64_t sN, int T, int N, int C, const int64_t* tgt,
                     int64_t tgt_stride, int Lmax, const int64_t* in_len, const int64_t* tgt_len, int blank,
                     const float* gout, int64_t gout_stride, float* grad, const CtcScratch& sc, int vec,
                     cudaStream_t st, bool sparse_only, int wait_scan) {
  int P, NTc;
  lat_geometry(Lmax, P, NTc);
  const BgSmem lay = bg_smem_layout(sc.Lp, sc.Sp, kBlkK);
  if (lay.total > 110 * 1024) {                  // two CTAs per SM or not at all
    if (sparse_only) return DAE_E_TOOBIG;        // the caller checks ctc_split_fits first
    int rc = ctc_blocked_fill(lp, sT, sN, T, N, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, sc, st);
    return rc ? rc : 1;
  }
#define DAE_BG(PP, MT, MB)                                                                                          \
  {                                                                                                                 \
    if (sparse_only)                                                                                                \
      return launch_block_grad<PP, MT, MB, true>(NTc, st, N, lp, sT, sN, T, C, tgt, tgt_stride, Lmax, in_len,       \
                                                 tgt_len, blank, gout, gout_stride, grad, sc, lay, vec, wait_scan); \
    return launch_block_grad<PP, MT, MB, false>(NTc, st, N, lp, sT, sN, T, C, tgt, tgt_stride, Lmax, in_len,        \
                                                tgt_len, blank, gout, gout_stride, grad, sc, lay, vec, 0);          \
  }
  constexpr int kMaxC = kLatThreads - 64;
  if (P <= 1 && NTc <= 640) DAE_BG(1, 640, 2);
  if (P <= 1) DAE_BG(1, kMaxC, 1);
  if (P <= 2) DAE_BG(2, kMaxC, 1);
  DAE_BG(4, kMaxC, 1);
#undef DAE_BG
}

// ------------------------------------------------------------------------------------------ 5. split gradient
// With the upstream scale known at loss time (dae_ctc_loss_grad) the class-dense part of the gradient, g*exp(lp),
// does not have to wait for the scan: it is streamed by a kernel that starts as soon as every scan CTA is
// resident (programmatic dependent launch; the scan releases its dependents right after its own wait) and runs
// on the SMs the scan leaves idle.  The launch asks for more shared memory than an SM has left beside a scan CTA,
// so no streaming CTA ever shares an SM (and issue slots) with the latency-bound scan.  Only the sparse part is
// left for after the scan (ctc_block_grad_kernel<SPARSE>).
constexpr int kDenseThreads = 1024;
constexpr int kDenseRows = 4;                     // rows per unit: independent 128-bit loads in flight per thread
constexpr int kDenseSmemReserve = 192 * 1024;     // + the scan CTA's >= 40 KB: does not fit in 227 KB
constexpr int kScanCtasMax = 48;                  // the split path is taken when the scan leaves >= 100 SMs idle

__global__ void __launch_bounds__(kDenseThreads, 1)
ctc_dense_grad_kernel(const float* __restrict__ lp, int64_t sT, int64_t sN, int T, int N, int C,
                      const int64_t* __restrict__ in_len, const float* __restrict__ gout, int64_t gout_stride,
                      float* __restrict__ grad, int wait_for_scan) {
  constexpr int R = kDenseRows;
  // the label-class gradient kernel behind this one may become resident (and load its inputs) right away
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int units_n = (T + R - 1) / R, C4 = C >> 2;
  for (int unit = blockIdx.x; unit < units_n * N; unit += gridDim.x) {
    const int n = unit / units_n, t0 = (unit - n * units_n) * R;
    int Tn = (int)in_len[n];
    Tn = Tn < 0 ? 0 : (Tn > T ? T : Tn);
    const float g = gout[n * gout_stride];
    const float* base = lp + n * sN + (int64_t)t0 * sT;
    for (int i = threadIdx.x; i < C4; i += kDenseThreads) {
      float4 v[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (t0 + r < Tn) v[r] = __ldcs(reinterpret_cast<const float4*>(base + (int64_t)r * sT) + i);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (t0 + r < T) {
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          if (t0 + r < Tn) o = make_float4(__expf(v[r].x) * g, __expf(v[r].y) * g, __expf(v[r].z) * g, __expf(v[r].w) * g);
          // plain stores: the sparse pass overwrites a few hundred words of every row while it is still in L2
          reinterpret_cast<float4*>(grad + ((int64_t)(t0 + r) * N + n) * C)[i] = o;
        }
      }
    }
  }
  // completes only after the scan has: a later launch in the stream must not overtake the scan through this kernel
  // (not needed when the kernel behind it waits for the scan itself)
  if (wait_for_scan) cudaGridDependencySynchronize();
}

int ctc_scan_ctas(const CtcScratch& sc, int N) { return 2 * sc.G * N; }

bool ctc_split_fits(const CtcScratch& sc, int N, int vec) {
  if (!sc.xfer || !vec || sm_count() < 2 * kScanCtasMax) return false;
  if ((ctc_config().overlap.load(std::memory_order_relaxed) & 1) == 0) return false;
  const BgSmem lay = bg_smem_layout(sc.Lp, sc.Sp, kBlkK);
  const int cl = 8;
  return lay.total <= 110 * 1024 && (sc.G + cl - 1) / cl * cl * 2 * N <= kScanCtasMax;
}

int ctc_blocked_dense(const float* lp, int64_t sT, int64_t sN, int T, int N, int C, const int64_t* in_len,
                      const float* gout, int64_t gout_stride, float* grad, cudaStream_t st, bool wait_for_scan) {
  DAE_CUDA(ensure_dyn_smem(ctc_dense_grad_kernel, kDenseSmemReserve));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sm_count() - kScanCtasMax);
  cfg.blockDim = dim3(kDenseThreads);
  cfg.dynamicSmemBytes = kDenseSmemReserve;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DAE_CUDA(cudaLaunchKernelEx(&cfg, ctc_dense_grad_kernel, lp, sT, sN, T, N, C, in_len, gout, gout_stride, grad,
                              wait_for_scan ? 1 : 0));
  DAE_LAUNCH_OK();
  return 0;
}

int ctc_blocked_lattice(const float* lp, int64_t sT, int64_t sN, int T, int N, int C, const int64_t* tgt,
                        int64_t tgt_stride, int Lmax, const int64_t* in_len, const int64_t* tgt_len, int blank,
                        float* nll, const CtcScratch& sc, cudaStream_t st, bool dense_follows) {
  (void)C;
  const int tiles = (sc.Sq + kXferTile - 1) / kXferTile;
  ctc_xfer_kernel<kBlkK><<<dim3(sc.nblk, tiles, N), kXferTile, 0, st>>>(lp, sT, sN, T, tgt, tgt_stride, Lmax, in_len,
                                                                       tgt_len, blank, sc);
  DAE_LAUNCH_OK();
  {
    // clusters of up to 8 regions hand over through DSMEM when the per-step mailbox fits in shared memory
    const size_t mbox_bytes = (size_t)(sc.nblk + 1) * kHaloWords * sizeof(int2);
    int cl = ctc_config().cluster.load(std::memory_order_relaxed);
    if (cl < 1 || cl > 8 || (cl & (cl - 1))) cl = 8;
    if (mbox_bytes > 96 * 1024 || sc.G < 2) cl = 1;
    while (cl > sc.G) cl >>= 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((sc.G + cl - 1) / cl * cl, 2, N);
    cfg.blockDim = dim3(kBndThreads);
    cfg.dynamicSmemBytes = cl > 1 ? mbox_bytes : 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // prologue overlaps the band kernel
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    if (cfg.dynamicSmemBytes > 24 * 1024)
      DAE_CUDA(ensure_dyn_smem(ctc_boundary_kernel<kBlkK>, (int)cfg.dynamicSmemBytes));
    DAE_CUDA(cudaLaunchKernelEx(&cfg, ctc_boundary_kernel<kBlkK>, T, Lmax, in_len, tgt_len, nll, sc, dense_follows ? 1 : 0));
    DAE_LAUNCH_OK();
  }
  return 0;
}

}  // namespace dae
