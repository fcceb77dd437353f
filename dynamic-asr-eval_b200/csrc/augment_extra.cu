// Secondary augmentations of the augmented window, SURVEY.md §8(f)-3: frame shuffle and additive noise.
//   frame_shuffle   lcasr/lib.py:81-84   spec[:, :, randperm(T)] then spec[:, randperm(F), :]
//   add_random_noise lcasr/lib.py:379-382 spec + normal(0, spec.std()) * noise_factor
// The randomness is drawn by the caller on the HOST with the reference's own calls in the reference's order
// (torch.randperm / the standard-normal stream behind torch.normal) and arrives here as descriptors: a
// permutation per axis, or the standard-normal field z.  The kernels are gathers / streaming passes.
#include "common.cuh"

namespace dae {

// out[f, t] = x[pf[f], pt[t]]  (NULL permutation = identity).  One CTA row-block, 128-bit stores.
__global__ void __launch_bounds__(256)
frame_shuffle_kernel(const float* __restrict__ x, int64_t sF, int F, int T, const int32_t* __restrict__ pt,
                     const int32_t* __restrict__ pf, float* __restrict__ out) {
  const int64_t total = (int64_t)F * T;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
    const int f = (int)(e / T), t = (int)(e - (int64_t)f * T);
    const int sf = pf ? __ldg(pf + f) : f, st = pt ? __ldg(pt + t) : t;
    out[e] = __ldg(x + sf * sF + st);
  }
}

// fp64 partial sums of x and x*x, fixed reduction order (per-CTA partials summed in index order by the consumer)
__global__ void __launch_bounds__(256)
noise_moments_kernel(const float* __restrict__ x, int64_t sF, int F, int T, double* __restrict__ partials) {
  __shared__ double w1[8], w2[8];
  const int64_t total = (int64_t)F * T;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
    const int f = (int)(e / T), t = (int)(e - (int64_t)f * T);
    const double v = (double)__ldg(x + f * sF + t);
    s1 += v;
    s2 += v * v;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { w1[threadIdx.x >> 5] = s1; w2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += w1[k]; b += w2[k]; }
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = b;
  }
}

// x += (z * std) * noise_factor, each product and the sum rounded separately (the reference's three tensor ops);
// std = sqrt(sum (x - mean)^2 / (n - 1)) from the fp64 moments, rounded once to fp32 (torch's unbiased std()).
__global__ void __launch_bounds__(256)
noise_apply_kernel(float* __restrict__ x, int64_t sF, int F, int T, const float* __restrict__ z, float noise_factor,
                   const double* __restrict__ partials, int n_part, float* __restrict__ std_out) {
  __shared__ float s_std;
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < n_part; ++k) { a += partials[2 * k]; b += partials[2 * k + 1]; }
    const double n = (double)F * (double)T;
    double var = (b - a * a / n) / (n - 1.0);
    if (var < 0.0) var = 0.0;
    s_std = (float)sqrt(var);
    if (blockIdx.x == 0 && std_out) *std_out = s_std;
  }
  __syncthreads();
  const float sd = s_std;
  const int64_t total = (int64_t)F * T;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
    const int f = (int)(e / T), t = (int)(e - (int64_t)f * T);
    float* p = x + f * sF + t;
    *p = __fadd_rn(*p, __fmul_rn(__fmul_rn(__ldg(z + e), sd), noise_factor));
  }
}

constexpr int kNoiseParts = kNumSMs * 2;

}  // namespace dae

extern "C" int dae_frame_shuffle(const float* x, int64_t sF, int F, int T, const int32_t* perm_t,
                                 const int32_t* perm_f, float* out, void* stream) {
  using namespace dae;
  if (!x || !out || F <= 0 || T <= 0 || x == out) return DAE_E_BADARG;
  const int64_t total = (int64_t)F * T;
  const int64_t want = (total + 255) / 256;
  const int grid = (int)(want < (int64_t)kNumSMs * 8 ? want : (int64_t)kNumSMs * 8);
  frame_shuffle_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, sF, F, T, perm_t, perm_f, out);
  DAE_LAUNCH_OK();
  return 0;
}

extern "C" size_t dae_noise_scratch_bytes(void) { return (size_t)dae::kNoiseParts * 2 * sizeof(double) + 256; }

extern "C" int dae_add_noise(float* x, int64_t sF, int F, int T, const float* z, float noise_factor,
                             void* scratch, size_t scratch_bytes, float* std_out, void* stream) {
  using namespace dae;
  if (!x || !z || F <= 0 || T <= 0) return DAE_E_BADARG;
  if ((int64_t)F * T < 2) return DAE_E_BADARG;
  if (!scratch || scratch_bytes < dae_noise_scratch_bytes()) return DAE_E_SCRATCH;
  if ((reinterpret_cast<uintptr_t>(scratch) & 255u) != 0) return DAE_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  double* partials = reinterpret_cast<double*>(scratch);
  noise_moments_kernel<<<kNoiseParts, 256, 0, st>>>(x, sF, F, T, partials);
  DAE_LAUNCH_OK();
  noise_apply_kernel<<<kNoiseParts, 256, 0, st>>>(x, sF, F, T, z, noise_factor, partials, kNoiseParts, std_out);
  DAE_LAUNCH_OK();
  return 0;
}
