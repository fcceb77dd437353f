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
#include "common.cuh"

namespace dae {

constexpr float kLog2e = 1.4426950408889634f;
constexpr double kLn2d = 0.69314718055994530942;
constexpr int kLatThreads = 1024;
constexpr int kMaxStatesPerThread = 8;            // S <= 8192  ->  Lmax <= 4095
constexpr int kPrefetch = 8;                      // time steps of lp gathers in flight (even)

struct CtcScratch {           // carved out of the caller's scratch buffer
  float* alpha;               // [N][T][Sp]  centred, log2 units
  float* beta_rev;            // [N][T][Sp]  beta stored at reversed state index S-1-s
  int32_t* next_same;         // [N][Lp]  next label position with the same class, -1 = none
  int32_t* leader;            // [N][Lp]  1 if first occurrence of its class
  double* off_a;              // [N][T]   alpha_t(s) = alpha[t][s] + off_a[t]   (log2 units)
  double* off_b;              // [N][T]   beta_t(s)  = beta_rev[t][S-1-s] + off_b[t]
  double* ll2;                // [2][N]   log2-likelihood from the alpha CTA, then from the beta CTA
  int Sp, Lp;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static inline size_t ctc_carve(CtcScratch& s, void* base, int T, int N, int Lmax) {
  s.Sp = (int)align_up((size_t)2 * Lmax + 1, 4);
  s.Lp = (int)align_up((size_t)(Lmax > 0 ? Lmax : 1), 4);
  char* p = (char*)base;
  size_t off = 0;
  const size_t lat = align_up((size_t)N * T * s.Sp * sizeof(float), 256);
  s.alpha = (float*)(p + off); off += lat;
  s.beta_rev = (float*)(p + off); off += lat;
  const size_t lab = align_up((size_t)N * s.Lp * sizeof(int32_t), 256);
  s.next_same = (int32_t*)(p + off); off += lab;
  s.leader = (int32_t*)(p + off); off += lab;
  const size_t offs = align_up((size_t)N * T * sizeof(double), 256);
  s.off_a = (double*)(p + off); off += offs;
  s.off_b = (double*)(p + off); off += offs;
  s.ll2 = (double*)(p + off); off += align_up((size_t)2 * N * sizeof(double), 256);
  return off;
}

__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// log2(2^a + 2^b + 2^c) with 3 MUFU ops: the largest term contributes exactly 1.
// -inf inputs allowed; all -inf -> -inf.
__device__ __forceinline__ float lse3_2(float a, float b, float c) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  const float m = fmaxf(hi, c), mid = fminf(hi, c);
  if (m == -CUDART_INF_F) return m;
  return m + fast_lg2(1.0f + fast_ex2(mid - m) + fast_ex2(lo - m));
}
// Order-preserving float <-> int map so a warp max can use redux.sync / smem atomicMax.
__device__ __forceinline__ int f2ord(float f) {
  const int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

template <int K>
__global__ void __launch_bounds__(kLatThreads, 1)
ctc_lattice_kernel(const float* __restrict__ lp, int64_t sT, int64_t sN, int T, int C,
                   const int64_t* __restrict__ tgt, int64_t tgt_stride, int Lmax,
                   const int64_t* __restrict__ in_len, const int64_t* __restrict__ tgt_len, int blank,
                   float* __restrict__ nll, CtcScratch sc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int mx[4];                   // rotating per-step maxima (ordered-int encoding)
  const int n = blockIdx.y;
  const int N = gridDim.y;
  const int dir = blockIdx.x;             // 0 alpha, 1 beta
  const int NT = blockDim.x;
  const int tid = threadIdx.x;

  int L = (int)tgt_len[n];
  L = L < 0 ? 0 : (L > Lmax ? Lmax : L);
  int Tn = (int)in_len[n];
  Tn = Tn < 0 ? 0 : (Tn > T ? T : Tn);
  const int S = 2 * L + 1;

  // smem: labels (direction-ordered) then two lattice rows padded on the left.
  int* lab = reinterpret_cast<int*>(smem_raw);                       // [Lp]
  float* a0 = reinterpret_cast<float*>(lab + sc.Lp) + 4;             // [-4 .. Sp)
  float* a1 = a0 + sc.Sp + 4;                                        // [-4 .. Sp)

  for (int k = tid; k < L; k += NT) {
    const int src = dir ? (L - 1 - k) : k;
    lab[k] = (int)tgt[n * tgt_stride + src];
  }
  if (tid < 4) {
    a0[tid - 4] = -CUDART_INF_F;
    a1[tid - 4] = -CUDART_INF_F;
    mx[tid] = f2ord(-CUDART_INF_F);
  }
  __syncthreads();

  // Label grouping for the gradient pass (alpha CTA only): first occurrence + next occurrence.
  if (dir == 0) {
    for (int k = tid; k < L; k += NT) {
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

  // Per-thread state constants.  State s (direction order): class and whether the s-2 skip is legal.
  int cls[K];
  bool skip[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int s = tid + j * NT;
    cls[j] = blank;
    skip[j] = false;
    if (s < S && (s & 1)) {
      const int k = s >> 1;
      cls[j] = lab[k];
      skip[j] = (k >= 1) && (lab[k - 1] != cls[j]);
    }
  }

  float* out = (dir ? sc.beta_rev : sc.alpha) + (int64_t)n * T * sc.Sp;
  double* offs = (dir ? sc.off_b : sc.off_a) + (int64_t)n * T;
  const float* base = lp + n * sN;

  if (Tn == 0) {  // degenerate: empty input.  Feasible only for the empty target.
    if (tid == 0) {
      const double ll = (L == 0) ? 0.0 : -(double)CUDART_INF_F;
      sc.ll2[dir * N + n] = ll;
      if (dir == 0) nll[n] = (float)(-ll);
    }
    return;
  }

  // Register ring of prefetched emissions: xr[u][j] = lp[time(u), cls[j]].
  float xr[kPrefetch][K];
#pragma unroll
  for (int u = 0; u < kPrefetch; ++u) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      xr[u][j] = 0.0f;
      if (u < Tn && tid + j * NT < S) {
        const int tt = dir ? (Tn - 1 - u) : u;
        xr[u][j] = __ldg(base + tt * sT + cls[j]);
      }
    }
  }

  double off_acc = 0.0;                    // thread 0: sum of the centring constants so far
  for (int t0 = 0; t0 < Tn; t0 += kPrefetch) {
#pragma unroll
    for (int u = 0; u < kPrefetch; ++u) {
      const int t = t0 + u;
      if (t < Tn) {                        // uniform across the CTA
        float* cur = (u & 1) ? a1 : a0;    // t0 is a multiple of kPrefetch (even)
        const float* prev = (u & 1) ? a0 : a1;
        const int tt = dir ? (Tn - 1 - t) : t;
        float* orow = out + (int64_t)tt * sc.Sp;
        // centring constant: the maximum of the previous step's values (0 at t = 0 / dead lattice)
        float c = 0.0f;
        if (t > 0) {
          c = ord2f(mx[t & 3]);
          if (c == -CUDART_INF_F) c = 0.0f;
        }
        if (tid == 0) {
          off_acc += (double)c;
          offs[tt] = off_acc;
          mx[(t + 2) & 3] = f2ord(-CUDART_INF_F);   // last read at step t-2, next filled at step t+1
        }
        float vmax = -CUDART_INF_F;
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const int s = tid + j * NT;
          if (s < S) {
            const float x = xr[u][j] * kLog2e;
            float v;
            if (t == 0) {
              v = (s <= 1) ? x : -CUDART_INF_F;
            } else {
              const float p0 = prev[s], p1 = prev[s - 1];
              const float p2 = skip[j] ? prev[s - 2] : -CUDART_INF_F;
              v = (x - c) + lse3_2(p0, p1, p2);
            }
            cur[s] = v;
            orow[s] = v;
            vmax = fmaxf(vmax, v);
            // refill this ring slot with the emission kPrefetch steps ahead
            const int tn = t + kPrefetch;
            if (tn < Tn) {
              const int ttn = dir ? (Tn - 1 - tn) : tn;
              xr[u][j] = __ldg(base + ttn * sT + cls[j]);
            }
          }
        }
        const int wmax = __reduce_max_sync(0xffffffffu, f2ord(vmax));
        if ((tid & 31) == 0) atomicMax(&mx[(t + 1) & 3], wmax);
        __syncthreads();
      }
    }
  }

  if (tid == 0) {
    const float* last = ((Tn - 1) & 1) ? a1 : a0;
    const float e1 = last[S - 1];
    const float e2 = (S > 1) ? last[S - 2] : -CUDART_INF_F;
    const double ll2 = off_acc + (double)lse3_2(e1, e2, -CUDART_INF_F);
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

template <int K>
static int launch_lattice(int NT, size_t smem, cudaStream_t st, int N, const float* lp, int64_t sT, int64_t sN, int T, int C,
                          const int64_t* tgt, int64_t tgt_stride, int Lmax, const int64_t* in_len,
                          const int64_t* tgt_len, int blank, float* nll, const CtcScratch& sc) {
  if (smem > 48 * 1024)
    DAE_CUDA(cudaFuncSetAttribute(ctc_lattice_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc_lattice_kernel<K><<<dim3(2, N), NT, smem, st>>>(lp, sT, sN, T, C, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, nll, sc);
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
  if (2 * Lmax + 1 > dae::kLatThreads * dae::kMaxStatesPerThread) return DAE_E_TOOBIG;
  if (!scratch || scratch_bytes < dae_ctc_scratch_bytes(T, N, Lmax)) return DAE_E_SCRATCH;
  if ((reinterpret_cast<uintptr_t>(scratch) & 255u) != 0) return DAE_E_ALIGN;
  return 0;
}

extern "C" int dae_ctc_lattice(const float* lp, int64_t sT, int64_t sN, int T, int N, int C, const int64_t* tgt,
                               int64_t tgt_stride, int Lmax, const int64_t* in_len, const int64_t* tgt_len,
                               int blank, float* nll, void* scratch, size_t scratch_bytes, void* stream) {
  using namespace dae;
  int rc = ctc_check(lp, T, N, C, tgt, Lmax, in_len, tgt_len, blank, scratch, scratch_bytes);
  if (rc) return rc;
  if (!nll) return DAE_E_BADARG;
  if (N == 0) return 0;
  CtcScratch sc;
  ctc_carve(sc, scratch, T, N, Lmax);
  const int S = 2 * Lmax + 1;
  int NT = ((S + 31) / 32) * 32;
  if (NT > kLatThreads) NT = kLatThreads;
  const int K = (S + NT - 1) / NT;
  const size_t smem = (size_t)sc.Lp * 4 + 2 * (size_t)(sc.Sp + 4) * 4;
  cudaStream_t st = (cudaStream_t)stream;
#define DAE_LAT(KK) return launch_lattice<KK>(NT, smem, st, N, lp, sT, sN, T, C, tgt, tgt_stride, Lmax, in_len, tgt_len, blank, nll, sc)
  if (K <= 1) DAE_LAT(1);
  if (K <= 2) DAE_LAT(2);
  if (K <= 4) DAE_LAT(4);
  DAE_LAT(8);
#undef DAE_LAT
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
  const size_t smem = (size_t)((C + 3) & ~3) * 4 + (size_t)sc.Lp * 4;
  if (smem > 200 * 1024) return DAE_E_TOOBIG;
  if (smem > 48 * 1024)
    DAE_CUDA(cudaFuncSetAttribute(ctc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int vec = aligned16(lp) && aligned16(grad) && (C % 4 == 0) && (sT % 4 == 0) && (sN % 4 == 0);
  int work = (C / 4 > 2 * Lmax + 1) ? C / 4 : 2 * Lmax + 1;
  int NT = ((work + 31) / 32) * 32;
  NT = NT < 32 ? 32 : (NT > 256 ? 256 : NT);
  ctc_grad_kernel<<<dim3(T, N), NT, smem, (cudaStream_t)stream>>>(lp, sT, sN, T, N, C, tgt, tgt_stride, Lmax, in_len,
                                                                tgt_len, blank, gout, gout_stride, grad, sc, vec);
  DAE_LAUNCH_OK();
  return 0;
}
