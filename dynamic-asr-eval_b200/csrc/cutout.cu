// Cutout augmentation (rectangles of the augmented copy overwritten with a fill value), SURVEY.md §8(f)-3.
// Semantics follow lcasr/lib.py:384-417: with cutout_val='mean' every rectangle is filled with its own
// mean taken BEFORE any rectangle is filled; 'mean_recording' uses the mean of the whole window; 'zero'
// writes 0.  Rectangles are filled in order, so where they overlap the last one wins.  Rectangle
// descriptors are drawn by the caller's RNG on the host (same torch.randint order as the reference).
#include "common.cuh"

namespace dae {

constexpr int kCutoutMaxRects = 4096;

// one CTA per rectangle (mode 1) or a grid-stride pass over the whole window (mode 2, rect = everything)
__global__ void __launch_bounds__(256)
cutout_mean_kernel(const float* __restrict__ x, int64_t sF, const int4* __restrict__ rects, int n, int whole_F,
                   int whole_T, float* __restrict__ means) {
  __shared__ double wsum[8];
  const int r = blockIdx.x;
  int4 q = (whole_T > 0) ? make_int4(0, whole_T, 0, whole_F) : rects[r];   // sx, ex, sy, ey
  const int w = q.y - q.x, h = q.w - q.z;
  const int64_t cnt = (int64_t)w * h;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < cnt; i += 256) {
    const int dy = (int)(i / w), dx = (int)(i - (int64_t)dy * w);
    acc += (double)x[(q.z + dy) * sF + q.x + dx];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += wsum[k];
    means[r] = cnt > 0 ? (float)(s / (double)cnt) : 0.0f;
  }
}

// Every element looks for the LAST rectangle that covers it; only covered elements are written.
__global__ void __launch_bounds__(256)
cutout_apply_kernel(float* __restrict__ x, int64_t sF, int F, int T, const int4* __restrict__ rects, int n, int mode,
                    const float* __restrict__ means) {
  extern __shared__ int4 srect[];
  for (int i = threadIdx.x; i < n; i += 256) srect[i] = rects[i];
  __syncthreads();
  const float whole = (mode == 2) ? means[0] : 0.0f;
  const int64_t total = (int64_t)F * T;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
    const int f = (int)(e / T), t = (int)(e - (int64_t)f * T);
    for (int i = n - 1; i >= 0; --i) {
      const int4 q = srect[i];
      if (t >= q.x && t < q.y && f >= q.z && f < q.w) {
        x[f * sF + t] = (mode == 1) ? means[i] : whole;
        break;
      }
    }
  }
}

}  // namespace dae

extern "C" size_t dae_cutout_scratch_bytes(int n_rect) {
  if (n_rect < 0) return 0;
  return 256 + (size_t)n_rect * (sizeof(int4) + sizeof(float)) + 256;
}

extern "C" int dae_cutout(float* x, int64_t sF, int F, int T, const int32_t* rects_host, int n_rect, int mode,
                          void* scratch, size_t scratch_bytes, void* stream) {
  using namespace dae;
  if (!x || F <= 0 || T <= 0 || n_rect < 0 || mode < 0 || mode > 2) return DAE_E_BADARG;
  if (n_rect == 0) return 0;
  if (!rects_host) return DAE_E_BADARG;
  if (n_rect > kCutoutMaxRects) return DAE_E_TOOBIG;
  if (!scratch || scratch_bytes < dae_cutout_scratch_bytes(n_rect)) return DAE_E_SCRATCH;
  if ((reinterpret_cast<uintptr_t>(scratch) & 255u) != 0) return DAE_E_ALIGN;
  for (int i = 0; i < n_rect; ++i) {
    const int32_t* q = rects_host + 4 * i;
    if (q[0] < 0 || q[1] > T || q[0] > q[1] || q[2] < 0 || q[3] > F || q[2] > q[3]) return DAE_E_BADARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int4* d_rects = reinterpret_cast<int4*>(scratch);
  float* d_means = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + ((size_t)n_rect * sizeof(int4) + 255) / 256 * 256);
  DAE_CUDA(cudaMemcpyAsync(d_rects, rects_host, (size_t)n_rect * sizeof(int4), cudaMemcpyHostToDevice, st));
  if (mode == 1) {
    cutout_mean_kernel<<<n_rect, 256, 0, st>>>(x, sF, d_rects, n_rect, 0, 0, d_means);
    DAE_LAUNCH_OK();
  } else if (mode == 2) {
    cutout_mean_kernel<<<1, 256, 0, st>>>(x, sF, d_rects, n_rect, F, T, d_means);
    DAE_LAUNCH_OK();
  }
  const int64_t total = (int64_t)F * T;
  int64_t want = (total + 255) / 256;
  const int grid = (int)(want < (int64_t)kNumSMs * 8 ? want : (int64_t)kNumSMs * 8);
  const size_t smem = (size_t)n_rect * sizeof(int4);
  if (smem > 48 * 1024)
    DAE_CUDA(ensure_dyn_smem(cutout_apply_kernel, (int)smem));
  cutout_apply_kernel<<<grid, 256, smem, st>>>(x, sF, F, T, d_rects, n_rect, mode, d_means);
  DAE_LAUNCH_OK();
  return 0;
}
