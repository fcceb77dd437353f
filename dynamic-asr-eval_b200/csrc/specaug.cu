// SpecAugment band masking fused with the [augmented, clean] batch build of the adapt step.
// Call contract follows lcasr/lib.py:102-112,499,538-541: mask value = mean of the window
// (or 0 with zero_masking), n_time_masks time bands then n_freq_masks frequency bands, each
// (start,end) drawn on the host so the RNG stream stays the caller's.
#include "common.cuh"

namespace dae {

constexpr int kMaxBands = 32;
constexpr int kMaxAug = 4;
constexpr int kSumGrid = kNumSMs * 2;     // partial sums, one per CTA
constexpr int kSumThreads = 256;

struct BandSet {
  int n_aug, nf, nt;
  int2 f[kMaxAug][kMaxBands];
  int2 t[kMaxAug][kMaxBands];
};

// Pass 1: fp64 partial sums in a fixed order (deterministic run to run).
template <bool VEC>
__global__ void __launch_bounds__(kSumThreads)
specaug_sum_kernel(const float* __restrict__ x, int64_t sF, int F, int T, double* __restrict__ partials) {
  __shared__ double wsum[kSumThreads / 32];
  double acc = 0.0;
  if (VEC) {
    const int t4 = T >> 2;
    const int64_t n4 = (int64_t)F * t4;
    for (int64_t i = (int64_t)blockIdx.x * kSumThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kSumThreads) {
      const int f = (int)(i / t4), c = (int)(i - (int64_t)f * t4);
      const float4 v = *reinterpret_cast<const float4*>(x + f * sF + 4 * c);
      acc += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    }
  } else {
    const int64_t n = (int64_t)F * T;
    for (int64_t i = (int64_t)blockIdx.x * kSumThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kSumThreads) {
      const int f = (int)(i / T), c = (int)(i - (int64_t)f * T);
      acc += (double)x[f * sF + c];
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kSumThreads / 32; ++w) s += wsum[w];
    partials[blockIdx.x] = s;
  }
}

__device__ __forceinline__ bool in_bands(const int2* b, int n, int i) {
  bool m = false;
  for (int k = 0; k < n; ++k) m |= (i >= b[k].x) & (i < b[k].y);
  return m;
}

// Pass 2: every CTA re-reduces the (L2-resident) partials in the same fixed order, then streams
// x once and writes n_aug masked + n_clean verbatim copies.
template <bool VEC>
__global__ void __launch_bounds__(256)
specaug_apply_kernel(const float* __restrict__ x, int64_t sF, int F, int T, const double* __restrict__ partials,
                     int zero_masking, int n_clean, const __grid_constant__ BandSet bands,
                     float* __restrict__ out, float* __restrict__ mean_out) {
  __shared__ double red[256];
  __shared__ float fill_s;
  if (!zero_masking) {
    double s = 0.0;
    for (int i = threadIdx.x; i < kSumGrid; i += 256) s += partials[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) fill_s = (float)(red[0] / ((double)F * (double)T));
  } else if (threadIdx.x == 0) {
    fill_s = 0.0f;
  }
  __syncthreads();
  const float fill = fill_s;
  if (mean_out && blockIdx.x == 0 && threadIdx.x == 0) *mean_out = fill;

  const int64_t plane = (int64_t)F * T;
  if (VEC) {
    const int t4 = T >> 2;
    const int64_t n4 = (int64_t)F * t4;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
      const int f = (int)(i / t4), c = 4 * (int)(i - (int64_t)f * t4);
      const float4 v = ld_stream4(x + f * sF + c);
      const int64_t o = (int64_t)f * T + c;
      for (int a = 0; a < bands.n_aug; ++a) {
        float4 w = v;
        if (in_bands(bands.f[a], bands.nf, f)) {
          w = make_float4(fill, fill, fill, fill);
        } else if (bands.nt) {
          if (in_bands(bands.t[a], bands.nt, c)) w.x = fill;
          if (in_bands(bands.t[a], bands.nt, c + 1)) w.y = fill;
          if (in_bands(bands.t[a], bands.nt, c + 2)) w.z = fill;
          if (in_bands(bands.t[a], bands.nt, c + 3)) w.w = fill;
        }
        st_stream4(out + a * plane + o, w);
      }
      for (int k = 0; k < n_clean; ++k) st_stream4(out + (bands.n_aug + k) * plane + o, v);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < plane; i += (int64_t)gridDim.x * 256) {
      const int f = (int)(i / T), c = (int)(i - (int64_t)f * T);
      const float v = x[f * sF + c];
      for (int a = 0; a < bands.n_aug; ++a) {
        const bool m = in_bands(bands.f[a], bands.nf, f) || in_bands(bands.t[a], bands.nt, c);
        out[a * plane + i] = m ? fill : v;
      }
      for (int k = 0; k < n_clean; ++k) out[(bands.n_aug + k) * plane + i] = v;
    }
  }
}

}  // namespace dae

extern "C" size_t dae_specaug_scratch_bytes(void) { return sizeof(double) * dae::kSumGrid; }

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
  if (need_mean) {
    if (vec) specaug_sum_kernel<true><<<kSumGrid, kSumThreads, 0, st>>>(x, sF, F, T, (double*)partials);
    else     specaug_sum_kernel<false><<<kSumGrid, kSumThreads, 0, st>>>(x, sF, F, T, (double*)partials);
    DAE_LAUNCH_OK();
  }
  const int64_t work = vec ? (int64_t)F * (T / 4) : (int64_t)F * T;
  int64_t want = (work + 255) / 256;
  const int grid = (int)(want < (int64_t)kNumSMs * 8 ? (want > 0 ? want : 1) : (int64_t)kNumSMs * 8);
  if (vec)
    specaug_apply_kernel<true><<<grid, 256, 0, st>>>(x, sF, F, T, (const double*)partials, need_mean ? 0 : 1, n_clean, bs, out, mean_out);
  else
    specaug_apply_kernel<false><<<grid, 256, 0, st>>>(x, sF, F, T, (const double*)partials, need_mean ? 0 : 1, n_clean, bs, out, mean_out);
  DAE_LAUNCH_OK();
  return 0;
}
