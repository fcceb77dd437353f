// Time-blocked CTC lattice for few samples and many frames (the dynamic-eval adapt step: N=1, T=2048).
// Same contract as ctc_lattice_kernel (ctc.cu): fills alpha, beta_rev, offsets, label groups, ll2 and nll
// (torch.nn.CTCLoss semantics; call sites lcasr/lib.py:492,570-579).
//
// The per-frame chain needs T dependent steps on two SMs.  The recursion is linear in the (log-sum, +)
// semiring, so it is cut into blocks of K frames and spread over the whole GPU in three launches:
//   1. ctc_xfer_kernel      every (block, source state) in parallel: the K-frame transfer band
//                           X_b[d][s] = log2 sum over paths that start in state s just before the block
//                           and end in state s+d on its last frame (d <= 2K).  One thread per source, the
//                           band lives in registers, no synchronisation.
//   2. ctc_boundary_kernel  the only sequential part, T/K steps: boundary vectors
//                           alpha_end(b)[s'] = LSE_d X_b[d][s'-d] + alpha_end(b-1)[s'-d]   and, with the same
//                           bands read the other way, betahat_start(b)[s] = LSE_d X_b[d][s] + betahat_start(b+1)[s+d].
//                           States are split into 128-state regions, one CTA each; a region only needs the last
//                           2K values of the region below it, handed over through tagged 8-byte words in global
//                           memory (tag in the data, no fences), so regions run as a skewed pipeline.
//   3. ctc_fill_kernel      every (block, direction) in parallel: K ordinary lattice steps from the block's
//                           boundary vector, written in the scratch layout ctc_grad_kernel reads.
// Values are log2 units with the finite dead-state sentinel of ctc_shared.cuh; every region / row carries an
// fp64 offset so stored fp32 values stay O(1).
#include "ctc_shared.cuh"

namespace dae {

__device__ __forceinline__ int2 ld_tagged(const int2* p) {
  int2 v;
  asm volatile("ld.volatile.global.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_tagged(int2* p, int2 v) {
  asm volatile("st.volatile.global.v2.s32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
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
  const int b = blockIdx.x, s0 = blockIdx.y * kXferTile, n = blockIdx.z, tid = threadIdx.x;
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
  // warps 0-3 own the even sources of the tile, warps 4-7 the odd ones
  const int par = tid >> 7;
  const int jl = 2 * (tid & 127) + par;
  float v[W];
  v[0] = 0.0f;
#pragma unroll
  for (int j = 1; j < W; ++j) v[j] = kDead;
  if (par == 0) xfer_steps<K, 0>(v, &es[0][0], EW, skp, jl, kb);
  else          xfer_steps<K, 1>(v, &es[0][0], EW, skp, jl, kb);
  const int s = s0 + jl;
  if (s < sc.Sq) {
    float* out = sc.xfer + ((size_t)(n * sc.nblk + b) * W) * sc.Sq + s;
#pragma unroll
    for (int d = 0; d < W; ++d) out[(size_t)d * sc.Sq] = v[d];
  }
}

// ------------------------------------------------------------------------------------------ 2. boundary scan
constexpr int kBndStages = 4;                     // transfer-band prefetch depth (steps)

// grid (G, 2, N): region g of direction dir (0 alpha, 1 beta in reversed state order u = S-1-s) of sample n.
// 128 consumer threads (one destination state each) + one helper warp (offset bookkeeping, halo hand-over).
template <int K>
__global__ void __launch_bounds__(kRegion + 32)
ctc_boundary_kernel(int T, int Lmax, const int64_t* __restrict__ in_len, const int64_t* __restrict__ tgt_len,
                    float* __restrict__ nll, CtcScratch sc) {
  constexpr int W = 2 * K + 1, H = 2 * K;
  __shared__ float buf[2][H + kRegion];          // [halo of the region below | own region]
  __shared__ float xs[kBndStages][W][kRegion];   // transfer bands of the next steps, one column per consumer
  __shared__ int wmx[2][kRegion / 32];           // per-warp maxima of the own region of each buffer
  __shared__ double off_s;
  const int g = blockIdx.x, dir = blockIdx.y, n = blockIdx.z, N = gridDim.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool helper = tid >= kRegion;
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
    return;
  }
  const int Sq = sc.Sq, G = sc.G;
  const size_t vec0 = (size_t)(dir * N + n) * (sc.nblk + 1);
  float* brow0 = sc.bound + vec0 * Sq;
  double* boff0 = sc.boff + vec0 * G;
  int2* halo0 = sc.halo + vec0 * G * kHaloWords;
  const float* xf0 = sc.xfer + (size_t)n * sc.nblk * W * Sq;
  const int u = g * kRegion + (tid & (kRegion - 1));

  // initial vector: alpha starts from the unit vector at state 0 (the virtual frame -1), betahat of the last
  // frame is 0 on the two final states (u = 0, 1).
  {
    const int idx = dir ? nb : 0;
    if (!helper) {
      float v0 = kDead;
      if (u == 0 || (dir == 1 && u == 1 && S > 1)) v0 = 0.0f;
      buf[0][H + tid] = v0;
      brow0[(size_t)idx * Sq + u] = v0;
      if (tid < H) buf[0][tid] = kDead;
      if (lane == 0) wmx[0][warp] = f2ord((g == 0 && warp == 0) ? 0.0f : kDead);
    } else if (lane == 0) {
      boff0[(size_t)idx * G + g] = 0.0;
    }
  }

  // consumer: enqueue the band column of step `st` into ring slot st % kBndStages
  auto enqueue = [&](int st) {
    if (st < nb) {
      const int b = dir ? (nb - 1 - st) : st;
      const float* xb = xf0 + (size_t)b * W * Sq;
      float* dst = &xs[st % kBndStages][0][tid];
      const int col = dir ? (S - 1 - u) : u;     // alpha: source column u-d; beta: column of s = S-1-u
#pragma unroll
      for (int d = 0; d < W; ++d) {
        const bool ok = (u - d >= 0) && (dir ? (col >= 0) : true);
        if (ok) cp_async_f32(dst + d * kRegion, xb + (size_t)d * Sq + (dir ? col : (u - d)));
        else dst[d * kRegion] = kDead;
      }
    }
    cp_async_commit();
  };
  if (!helper) {
#pragma unroll
    for (int st = 0; st < kBndStages - 1; ++st) enqueue(st);
  }
  __syncthreads();

  double off = 0.0;                              // helper: offset of the own region of the newest vector
  bool halo_prev_alive = false;
  for (int st = 0; st < nb; ++st) {
    const float* prev = buf[st & 1];
    float* cur = buf[(st + 1) & 1];
    const int idx_out = dir ? (nb - 1 - st) : (st + 1);
    // centring constant: the maximum of the own region of the previous vector (0 while the region is dead)
    float mprev = ord2f(max(max(wmx[st & 1][0], wmx[st & 1][1]), max(wmx[st & 1][2], wmx[st & 1][3])));
    const bool alive_prev = mprev > -1.0e29f;
    const float c = alive_prev ? mprev : 0.0f;
    if (!helper) {
      enqueue(st + kBndStages - 1);
      cp_async_wait<kBndStages - 1>();
      const float* xc = &xs[st % kBndStages][0][tid];
      float term[W];
      float mx = kDead;
#pragma unroll
      for (int d = 0; d < W; ++d) {
        term[d] = xc[d * kRegion] + prev[H + tid - d];
        mx = fmaxf(mx, term[d]);
      }
      float sum = 0.0f;
#pragma unroll
      for (int d = 0; d < W; ++d) sum += fast_ex2(term[d] - mx);
      const float val = (mx + fast_lg2(sum)) - c;
      cur[H + tid] = val;
      brow0[(size_t)idx_out * Sq + u] = val;
      const int wm = __reduce_max_sync(0xffffffffu, f2ord(val));
      if (lane == 0) wmx[(st + 1) & 1][warp] = wm;
      if (tid >= kRegion - H && g + 1 < G)
        st_tagged(halo0 + ((size_t)idx_out * G + g) * kHaloWords + (tid - (kRegion - H)),
                  make_int2(__float_as_int(val), st + 1));
    } else {
      // helper warp: take the halo of the region below for the vector being produced, decide this region's offset
      float hv = kDead;
      double noff = 0.0;
      bool halo_alive = false;
      if (g > 0) {
        const int2* hp = halo0 + ((size_t)idx_out * G + (g - 1)) * kHaloWords;
        int2 w = make_int2(0, 0);
        if (lane < kHaloWords) {
          do { w = ld_tagged(hp + lane); } while (w.y != st + 1);
        }
        const int lo = __shfl_sync(0xffffffffu, w.x, H), hi = __shfl_sync(0xffffffffu, w.x, H + 1);
        noff = __hiloint2double(hi, lo);
        hv = (lane < H) ? __int_as_float(w.x) : kDead;
        halo_alive = __ballot_sync(0xffffffffu, hv > -1.0e29f) != 0u;
      }
      // a region that stays dead this step adopts the offset of the region below, so the first values that
      // flow in are taken over without rounding
      const bool stays_dead = !alive_prev && !halo_prev_alive;
      const double off_new = stays_dead ? (g > 0 ? noff : off) : off + (double)c;
      if (lane < H) cur[lane] = (hv > -1.0e29f) ? hv + (float)(noff - off_new) : kDead;
      if (lane == 0) {
        boff0[(size_t)idx_out * G + g] = off_new;
        off_s = off_new;
      }
      if (g + 1 < G && lane < 2) {
        const int half = lane ? __double2hiint(off_new) : __double2loint(off_new);
        st_tagged(halo0 + ((size_t)idx_out * G + g) * kHaloWords + H + lane, make_int2(half, st + 1));
      }
      off = off_new;
      halo_prev_alive = halo_alive;
    }
    __syncthreads();
  }

  // log-likelihood from the final vector: alpha ends on the last two states, betahat(-1) starts on state 0 (u = S-1)
  if (!helper && u == S - 1) {
    const float* fin = buf[nb & 1];
    const float e1 = fin[H + tid];
    const float e2 = (dir == 0 && S > 1) ? fin[H + tid - 1] : kDead;
    const float tail = lse2_n(e1, e2);
    const double ll2 = (tail < -1.0e29f) ? -(double)CUDART_INF_F : off_s + (double)tail;
    sc.ll2[dir * N + n] = ll2;
    if (dir == 0) nll[n] = (float)(-ll2 * kLn2d);
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
template <int P, int K>
__global__ void __launch_bounds__(kLatThreads - 64)
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
  if (dir == 0 && b == 0) {                      // label grouping for the gradient pass
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

template <int P>
static int launch_fill(int NTc, cudaStream_t st, int N, const float* lp, int64_t sT, int64_t sN, int T,
                       const int64_t* tgt, int64_t tgt_stride, int Lmax, const int64_t* in_len,
                       const int64_t* tgt_len, int blank, const CtcScratch& sc) {
  const FillSmem lay = fill_smem_layout(sc.Lp, sc.Sp);
  if (lay.total > 200 * 1024) return DAE_E_TOOBIG;
  if (lay.total > 48 * 1024)
    DAE_CUDA(cudaFuncSetAttribute(ctc_fill_kernel<P, kBlkK>, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total));
  ctc_fill_kernel<P, kBlkK><<<dim3(sc.nblk, 2, N), NTc, lay.total, st>>>(lp, sT, sN, T, tgt, tgt_stride, Lmax, in_len,
                                                                        tgt_len, blank, sc, lay);
  DAE_LAUNCH_OK();
  return 0;
}

int ctc_blocked_lattice(const float* lp, int64_t sT, int64_t sN, int T, int N, int C, const int64_t* tgt,
                        int64_t tgt_stride, int Lmax, const int64_t* in_len, const int64_t* tgt_len, int blank,
                        float* nll, const CtcScratch& sc, cudaStream_t st) {
  (void)C;
  // hand-over words are matched by tag (step number), so they start from zero on every call
  DAE_CUDA(cudaMemsetAsync(sc.halo, 0, (size_t)2 * N * (sc.nblk + 1) * sc.G * kHaloWords * sizeof(int2), st));
  const int tiles = (sc.Sq + kXferTile - 1) / kXferTile;
  ctc_xfer_kernel<kBlkK><<<dim3(sc.nblk, tiles, N), kXferTile, 0, st>>>(lp, sT, sN, T, tgt, tgt_stride, Lmax, in_len,
                                                                       tgt_len, blank, sc);
  DAE_LAUNCH_OK();
  ctc_boundary_kernel<kBlkK><<<dim3(sc.G, 2, N), kRegion + 32, 0, st>>>(T, Lmax, in_len, tgt_len, nll, sc);
  DAE_LAUNCH_OK();
  int P, NTc;
  lat_geometry(Lmax, P, NTc);
  if (P <= 1) return launch_fill<1>(NTc, st, N, lp, sT, sN, T, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, sc);
  if (P <= 2) return launch_fill<2>(NTc, st, N, lp, sT, sN, T, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, sc);
  return launch_fill<4>(NTc, st, N, lp, sT, sN, T, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, sc);
}

}  // namespace dae
