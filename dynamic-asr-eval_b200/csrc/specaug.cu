// SpecAugment band masking fused with the [augmented, clean] batch build of the adapt step.
// Call contract follows lcasr/lib.py:102-112,499,538-541: mask value = mean of the window
// (or 0 with zero_masking), n_time_masks time bands then n_freq_masks frequency bands, each
// (start,end) drawn on the host so the RNG stream stays the caller's.
#include "common.cuh"

namespace dae {

constexpr int kMaxBands = 32;
constexpr int kMaxAug = 4;
constexpr int kSumGrid = kNumSMs * 2;     // partial sums, one per CTA

struct BandSet {
  int n_aug, nf, nt;
  int2 f[kMaxAug][kMaxBands];
  int2 t[kMaxAug][kMaxBands];
};

__device__ __forceinline__ bool in_bands(const int2* b, int n, int i) {
  bool m = false;
  for (int k = 0; k < n; ++k) m |= (i >= b[k].x) & (i < b[k].y);
  return m;
}

// One cooperative grid (co-resident CTAs) does both passes with a grid barrier in between.  Pass 1 reads x once, writes the clean copies and every unmasked element of the augmented
// copies, and leaves a fp64 partial sum per CTA; after the barrier every CTA reduces the partials in the
// same fixed order and writes the fill value into the masked elements of its own slice.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int expected) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (*((volatile unsigned int*)counter) < expected) { }
    __threadfence();
  }
  __syncthreads();
}

template <bool VEC>
__global__ void __launch_bounds__(256)
specaug_fused_kernel(const float* __restrict__ x, int64_t sF, int F, int T, double* __restrict__ partials,
                     unsigned int* __restrict__ counter, int zero_masking, int n_clean,
                     const __grid_constant__ BandSet bands, float* __restrict__ out, float* __restrict__ mean_out) {
  __shared__ double red[256];
  __shared__ float fill_s;
  const int64_t plane = (int64_t)F * T;
  const int t4 = T >> 2;
  const int64_t n_items = VEC ? (int64_t)F * t4 : plane;
  const int64_t per = (n_items + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = (lo + per < n_items) ? lo + per : n_items;
  double acc = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
    if (VEC) {
      const int f = (int)(i / t4), c = 4 * (int)(i - (int64_t)f * t4);
      const float4 v = ld_stream4(x + f * sF + c);
      acc += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
      const int64_t o = (int64_t)f * T + c;
      for (int a = 0; a < bands.n_aug; ++a) {
        if (in_bands(bands.f[a], bands.nf, f)) continue;                       // whole row masked: pass 2
        if (bands.nt == 0) { st_stream4(out + a * plane + o, v); continue; }
        float* dst = out + a * plane + o;
        if (!in_bands(bands.t[a], bands.nt, c)) dst[0] = v.x;
        if (!in_bands(bands.t[a], bands.nt, c + 1)) dst[1] = v.y;
        if (!in_bands(bands.t[a], bands.nt, c + 2)) dst[2] = v.z;
        if (!in_bands(bands.t[a], bands.nt, c + 3)) dst[3] = v.w;
      }
      for (int k = 0; k < n_clean; ++k) st_stream4(out + (bands.n_aug + k) * plane + o, v);
    } else {
      const int f = (int)(i / T), c = (int)(i - (int64_t)f * T);
      const float v = x[f * sF + c];
      acc += (double)v;
      for (int a = 0; a < bands.n_aug; ++a)
        if (!(in_bands(bands.f[a], bands.nf, f) || in_bands(bands.t[a], bands.nt, c))) out[a * plane + i] = v;
      for (int k = 0; k < n_clean; ++k) out[(bands.n_aug + k) * plane + i] = v;
    }
  }
  float fill = 0.0f;
  if (!zero_masking) {
    // CTA partial in a fixed order: deterministic run to run
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double sacc = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) sacc += red[w];
      partials[blockIdx.x] = sacc;
    }
    grid_barrier(counter, gridDim.x);
    double sgl = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) sgl += __ldcg(partials + i);
    red[threadIdx.x] = sgl;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) fill_s = (float)(red[0] / ((double)F * (double)T));
    __syncthreads();
    fill = fill_s;
  }
  if (mean_out && blockIdx.x == 0 && threadIdx.x == 0) *mean_out = fill;
  // pass 2: the masked elements of this CTA's slice
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
    if (VEC) {
      const int f = (int)(i / t4), c = 4 * (int)(i - (int64_t)f * t4);
      const int64_t o = (int64_t)f * T + c;
      for (int a = 0; a < bands.n_aug; ++a) {
        float* dst = out + a * plane + o;
        if (in_bands(bands.f[a], bands.nf, f)) { st_stream4(dst, make_float4(fill, fill, fill, fill)); continue; }
        if (bands.nt == 0) continue;
        if (in_bands(bands.t[a], bands.nt, c)) dst[0] = fill;
        if (in_bands(bands.t[a], bands.nt, c + 1)) dst[1] = fill;
        if (in_bands(bands.t[a], bands.nt, c + 2)) dst[2] = fill;
        if (in_bands(bands.t[a], bands.nt, c + 3)) dst[3] = fill;
      }
    } else {
      const int f = (int)(i / T), c = (int)(i - (int64_t)f * T);
      for (int a = 0; a < bands.n_aug; ++a)
        if (in_bands(bands.f[a], bands.nf, f) || in_bands(bands.t[a], bands.nt, c)) out[a * plane + i] = fill;
    }
  }
}

// ---- window means computed once per recording + a barrier-free single-pass mask/copy kernel ------------------
// The adapt loop augments every window of a recording (lcasr/lib.py:537-541); the fill value of a window is its
// mean, a function of the spectrogram alone.  dae_window_sums leaves kWinSlices fp64 partial sums per window in ONE
// launch per recording; dae_specaug_repeat_premean then needs neither a grid barrier nor a second pass: every CTA
// adds the window's partials in the same fixed order and streams x -> [masked..., clean...] once.
constexpr int kWinSlices = 16;

__global__ void __launch_bounds__(256)
window_sums_kernel(const float* __restrict__ x, int64_t sF, int F, const int64_t* __restrict__ win_start,
                   const int64_t* __restrict__ win_len, double* __restrict__ sums) {
  // slice k of window w = its rows [F*k/16, F*(k+1)/16): whole rows, so a thread strides over contiguous memory with
  // 128-bit loads (per-element index arithmetic made this kernel 4x slower than the copy roofline)
  __shared__ double red[8];
  const int w = blockIdx.y, slice = blockIdx.x;
  const int64_t t0 = win_start[w], T = win_len[w];
  const int f_lo = (int)((int64_t)F * slice / kWinSlices), f_hi = (int)((int64_t)F * (slice + 1) / kWinSlices);
  double acc = 0.0;
  for (int f = f_lo; f < f_hi; ++f) {
    const float* row = x + (int64_t)f * sF + t0;
    if ((reinterpret_cast<uintptr_t>(row) & 15u) == 0) {
      const int64_t n4 = T >> 2;
      for (int64_t i = threadIdx.x; i < n4; i += 256) {
        const float4 v = ld_stream4(row + 4 * i);
        acc += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
      }
      for (int64_t c = 4 * n4 + threadIdx.x; c < T; c += 256) acc += (double)__ldg(row + c);
    } else {
      for (int64_t c = threadIdx.x; c < T; c += 256) acc += (double)__ldg(row + c);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k];
    sums[(int64_t)w * kWinSlices + slice] = s;
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256)
specaug_premean_kernel(const float* __restrict__ x, int64_t sF, int F, int T, const double* __restrict__ win_sums,
                       int zero_masking, int n_clean, const __grid_constant__ BandSet bands,
                       float* __restrict__ out, float* __restrict__ mean_out) {
  float fill = 0.0f;
  if (!zero_masking) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kWinSlices; ++k) s += __ldg(win_sums + k);
    fill = (float)(s / ((double)F * (double)T));
  }
  if (mean_out && blockIdx.x == 0 && threadIdx.x == 0) *mean_out = fill;
  const int64_t plane = (int64_t)F * T;
  const int t4 = T >> 2;
  const int64_t n_items = VEC ? (int64_t)F * t4 : plane;
  const int64_t stride = (int64_t)gridDim.x * 256;
  if (VEC) {
    // two items per iteration: both loads are issued before either is used
    for (int64_t i0 = (int64_t)blockIdx.x * 256 + threadIdx.x; i0 < n_items; i0 += 2 * stride) {
      const int64_t i1 = i0 + stride;
      const int f0 = (int)(i0 / t4), c0 = 4 * (int)(i0 - (int64_t)f0 * t4);
      const bool has1 = i1 < n_items;
      const int f1 = has1 ? (int)(i1 / t4) : f0, c1 = has1 ? 4 * (int)(i1 - (int64_t)f1 * t4) : c0;
      const float4 va = ld_stream4(x + f0 * sF + c0);
      const float4 vb = ld_stream4(x + f1 * sF + c1);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !has1) break;
        const int f = u ? f1 : f0, c = u ? c1 : c0;
        const float4 v = u ? vb : va;
        const int64_t o = (int64_t)f * T + c;
        for (int a = 0; a < bands.n_aug; ++a) {
          float4 m = v;
          if (in_bands(bands.f[a], bands.nf, f)) {
            m = make_float4(fill, fill, fill, fill);
          } else if (bands.nt) {
            if (in_bands(bands.t[a], bands.nt, c)) m.x = fill;
            if (in_bands(bands.t[a], bands.nt, c + 1)) m.y = fill;
            if (in_bands(bands.t[a], bands.nt, c + 2)) m.z = fill;
            if (in_bands(bands.t[a], bands.nt, c + 3)) m.w = fill;
          }
          st_stream4(out + a * plane + o, m);
        }
        for (int k = 0; k < n_clean; ++k) st_stream4(out + (bands.n_aug + k) * plane + o, v);
      }
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_items; i += stride) {
    if (VEC) {
      const int f = (int)(i / t4), c = 4 * (int)(i - (int64_t)f * t4);
      const float4 v = ld_stream4(x + f * sF + c);
      const int64_t o = (int64_t)f * T + c;
      for (int a = 0; a < bands.n_aug; ++a) {
        float4 m = v;
        if (in_bands(bands.f[a], bands.nf, f)) {
          m = make_float4(fill, fill, fill, fill);
        } else if (bands.nt) {
          if (in_bands(bands.t[a], bands.nt, c)) m.x = fill;
          if (in_bands(bands.t[a], bands.nt, c + 1)) m.y = fill;
          if (in_bands(bands.t[a], bands.nt, c + 2)) m.z = fill;
          if (in_bands(bands.t[a], bands.nt, c + 3)) m.w = fill;
        }
        st_stream4(out + a * plane + o, m);
      }
      for (int k = 0; k < n_clean; ++k) st_stream4(out + (bands.n_aug + k) * plane + o, v);
    } else {
      const int f = (int)(i / T), c = (int)(i - (int64_t)f * T);
      const float v = x[f * sF + c];
      for (int a = 0; a < bands.n_aug; ++a)
        out[a * plane + i] = (in_bands(bands.f[a], bands.nf, f) || in_bands(bands.t[a], bands.nt, c)) ? fill : v;
      for (int k = 0; k < n_clean; ++k) out[(bands.n_aug + k) * plane + i] = v;
    }
  }
}

static int fill_bands(BandSet& bs, const int32_t* fmask_host, int nf, const int32_t* tmask_host, int nt, int n_aug) {
  if (nf < 0 || nt < 0 || n_aug < 0) return DAE_E_BADARG;
  if (nf > kMaxBands || nt > kMaxBands || n_aug > kMaxAug) return DAE_E_TOOBIG;
  if ((nf && !fmask_host) || (nt && !tmask_host)) return DAE_E_BADARG;
  bs.n_aug = n_aug; bs.nf = nf; bs.nt = nt;
  for (int a = 0; a < n_aug; ++a) {
    for (int k = 0; k < nf; ++k) bs.f[a][k] = make_int2(fmask_host[(a * nf + k) * 2], fmask_host[(a * nf + k) * 2 + 1]);
    for (int k = 0; k < nt; ++k) bs.t[a][k] = make_int2(tmask_host[(a * nt + k) * 2], tmask_host[(a * nt + k) * 2 + 1]);
  }
  return 0;
}

}  // namespace dae

extern "C" int dae_window_slices(void) { return dae::kWinSlices; }

extern "C" int dae_window_sums(const float* x, int64_t sF, int F, const int64_t* win_start, const int64_t* win_len,
                               int n_win, double* sums, void* stream) {
  using namespace dae;
  if (!x || !win_start || !win_len || !sums || F <= 0 || n_win < 0) return DAE_E_BADARG;
  if (n_win == 0) return 0;
  if (n_win > 65535) return DAE_E_TOOBIG;
  window_sums_kernel<<<dim3(kWinSlices, n_win), 256, 0, (cudaStream_t)stream>>>(x, sF, F, win_start, win_len, sums);
  DAE_LAUNCH_OK();
  return 0;
}

extern "C" int dae_specaug_repeat_premean(const float* x, int64_t sF, int F, int T, const int32_t* fmask_host, int nf,
                                          const int32_t* tmask_host, int nt, int zero_masking, int n_aug, int n_clean,
                                          float* out, const double* win_sums, float* mean_out, void* stream) {
  using namespace dae;
  if (!x || !out || F <= 0 || T <= 0 || n_clean < 0) return DAE_E_BADARG;
  BandSet bs;
  const int rc = fill_bands(bs, fmask_host, nf, tmask_host, nt, n_aug);
  if (rc) return rc;
  if (!zero_masking && n_aug > 0 && !win_sums) return DAE_E_BADARG;
  if (n_aug + n_clean == 0) return 0;
  const bool vec = aligned16(x) && aligned16(out) && (T % 4 == 0) && (sF % 4 == 0);
  const int64_t n_items = vec ? (int64_t)F * (T >> 2) : (int64_t)F * T;
  int64_t want = (n_items + 511) / 512;                   // two items per thread and iteration
  const int grid = (int)(want < (int64_t)kNumSMs * 4 ? want : (int64_t)kNumSMs * 4);
  const int zm = (zero_masking || n_aug == 0) ? 1 : 0;
  if (vec)
    specaug_premean_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, sF, F, T, win_sums, zm, n_clean, bs, out, mean_out);
  else
    specaug_premean_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, sF, F, T, win_sums, zm, n_clean, bs, out, mean_out);
  DAE_LAUNCH_OK();
  return 0;
}

extern "C" size_t dae_specaug_scratch_bytes(void) { return sizeof(double) * dae::kSumGrid + 64; }

extern "C" int dae_specaug_repeat(const float* x, int64_t sF, int F, int T, const int32_t* fmask_host, int nf,
                                  const int32_t* tmask_host, int nt, int zero_masking, int n_aug, int n_clean,
                                  float* out, void* partials, float* mean_out, void* stream) {
  using namespace dae;
  if (!x || !out || F <= 0 || T <= 0 || nf < 0 || nt < 0 || n_aug < 0 || n_clean < 0) return DAE_E_BADARG;
  if (nf > kMaxBands || nt > kMaxBands || n_aug > kMaxAug) return DAE_E_TOOBIG;
  if ((nf && !fmask_host) || (nt && !tmask_host)) return DAE_E_BADARG;
  if (!zero_masking && n_aug > 0 && !partials) return DAE_E_SCRATCH;
  if (n_aug + n_clean == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  BandSet bs;
  bs.n_aug = n_aug; bs.nf = nf; bs.nt = nt;
  for (int a = 0; a < n_aug; ++a) {
    for (int k = 0; k < nf; ++k) bs.f[a][k] = make_int2(fmask_host[(a * nf + k) * 2], fmask_host[(a * nf + k) * 2 + 1]);
    for (int k = 0; k < nt; ++k) bs.t[a][k] = make_int2(tmask_host[(a * nt + k) * 2], tmask_host[(a * nt + k) * 2 + 1]);
  }
  const bool vec = aligned16(x) && aligned16(out) && (T % 4 == 0) && (sF % 4 == 0);
  const bool need_mean = !zero_masking && n_aug > 0;
  // one cooperative launch: partial sums live at partials[0 .. kSumGrid), the barrier counter after them
  double* part = (double*)partials;
  unsigned int* counter = partials ? (unsigned int*)(part + kSumGrid) : nullptr;
  if (need_mean) DAE_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
  int zm = need_mean ? 0 : 1;
  float* mo = mean_out;
  void* args[] = {(void*)&x, (void*)&sF, (void*)&F, (void*)&T, (void*)&part, (void*)&counter, (void*)&zm,
                  (void*)&n_clean, (void*)&bs, (void*)&out, (void*)&mo};
  const void* fn = vec ? (const void*)specaug_fused_kernel<true> : (const void*)specaug_fused_kernel<false>;
  DAE_CUDA(cudaLaunchCooperativeKernel(fn, dim3(kSumGrid), dim3(256), args, 0, st));
  DAE_LAUNCH_OK();
  return 0;
}
