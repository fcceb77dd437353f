// Greedy CTC decode on device: per-frame argmax (bandwidth kernel) + collapse (tiny).
// Behaviour follows GreedyCTCDecoder as used at lcasr/lib.py:559 and
// lcasr/run_dynamic_eval_full.py:100: argmax(-1) -> unique_consecutive -> drop blank.
#include "common.cuh"

namespace dae {

// torch.argmax ordering: NaN beats everything, ties go to the lower index.
__device__ __forceinline__ bool arg_better(float v, int i, float bv, int bi) {
  const bool vn = (v != v), bn = (bv != bv);
  if (vn || bn) return vn && (!bn || i < bi);
  return v > bv || (v == bv && i < bi);
}

// One warp per row; lanes stride over the row with 128-bit loads, 8 in flight per lane.  Rows are dealt to CTAs
// round-robin (row r -> CTA r % grid) so that every SM streams the same number of bytes at any row count.
template <bool VEC>
__device__ __forceinline__ void argmax_rows(const float* __restrict__ lp, int64_t sB, int64_t sT, int B, int T, int C,
                                            const int32_t* __restrict__ lengths, int32_t* __restrict__ path) {
  const int lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  const int64_t nrows = (int64_t)B * T;
  for (int64_t row = (int64_t)blockIdx.x + (int64_t)gridDim.x * (threadIdx.x >> 5); row < nrows;
       row += (int64_t)gridDim.x * warps_per_cta) {
    const int b = (int)(row / T), t = (int)(row - (int64_t)b * T);
    if (lengths && t >= lengths[b]) {
      if (lane == 0) path[row] = -1;
      continue;
    }
    const float* x = lp + b * sB + t * sT;
    float bv = -CUDART_INF_F;
    int bi = 0x7fffffff;
    if (VEC) {
      const int n4 = C >> 2;
      constexpr int U = 16;                                // 8 KB of loads in flight per warp
      int j = lane;
      for (; j + 32 * (U - 1) < n4; j += 32 * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld_stream4(x + 4 * (j + 32 * u));
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i0 = 4 * (j + 32 * u);
          // strict '>' keeps the lowest index inside a lane (indices ascend); NaN wins once.
          if (v[u].x > bv || (v[u].x != v[u].x && bv == bv)) { bv = v[u].x; bi = i0; }
          if (v[u].y > bv || (v[u].y != v[u].y && bv == bv)) { bv = v[u].y; bi = i0 + 1; }
          if (v[u].z > bv || (v[u].z != v[u].z && bv == bv)) { bv = v[u].z; bi = i0 + 2; }
          if (v[u].w > bv || (v[u].w != v[u].w && bv == bv)) { bv = v[u].w; bi = i0 + 3; }
        }
      }
      for (; j < n4; j += 32) {
        const float4 v = ld_stream4(x + 4 * j);
        const int i0 = 4 * j;
        if (v.x > bv || (v.x != v.x && bv == bv)) { bv = v.x; bi = i0; }
        if (v.y > bv || (v.y != v.y && bv == bv)) { bv = v.y; bi = i0 + 1; }
        if (v.z > bv || (v.z != v.z && bv == bv)) { bv = v.z; bi = i0 + 2; }
        if (v.w > bv || (v.w != v.w && bv == bv)) { bv = v.w; bi = i0 + 3; }
      }
    } else {
      for (int i = lane; i < C; i += 32) {
        const float v = ld_stream1(x + i);
        if (v > bv || (v != v && bv == bv)) { bv = v; bi = i; }
      }
    }
    // A lane that saw only -inf (or nothing) still holds bi = INT_MAX; give it a real index
    // so an all -inf row resolves to class 0 like torch.argmax.
    if (bi == 0x7fffffff) bi = (VEC ? 4 * lane : lane) < C ? (VEC ? 4 * lane : lane) : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (arg_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) path[row] = bi;
  }
}

// keep[t] = path[t] != blank && path[t] != path[t-1]; ordered compaction of one batch item by one CTA of NT threads.
// `path` is read with ld.cg: in the fused kernel other CTAs have just written it.
constexpr int kCollapseThreads = 1024;
constexpr int kCollapseItems = 4;
constexpr int kGreedyThreads = 512;

template <int NT>
__device__ __forceinline__ void collapse_item(const int32_t* __restrict__ path, int T, int blank,
                                              const int32_t* __restrict__ lengths, int32_t* __restrict__ ids,
                                              int32_t* __restrict__ n_ids, int b, int* warp_tot, int* carry_p) {
  constexpr int kCollapseThreads = NT;
  int& carry = *carry_p;
  const int32_t* p = path + (int64_t)b * T;
  int32_t* out = ids + (int64_t)b * T;
  const int len = lengths ? min(T, max(0, lengths[b])) : T;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (threadIdx.x == 0) carry = 0;
  if (threadIdx.x < 32) warp_tot[threadIdx.x] = 0;
  __syncthreads();
  for (int base = 0; base < len; base += kCollapseThreads * kCollapseItems) {
    const int t0 = base + threadIdx.x * kCollapseItems;
    const int base_off = carry;  // written before the previous iteration's last barrier
    int v[kCollapseItems];
    int prev = (t0 > 0 && t0 - 1 < len) ? __ldcg(p + t0 - 1) : -2;
    int cnt = 0;
    unsigned keep = 0;
#pragma unroll
    for (int k = 0; k < kCollapseItems; ++k) {
      const int t = t0 + k;
      v[k] = (t < len) ? __ldcg(p + t) : -2;
      const bool kp = (t < len) && v[k] != blank && v[k] != prev;
      keep |= (unsigned)kp << k;
      cnt += kp;
      prev = v[k];
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = lane < kCollapseThreads / 32 ? warp_tot[lane] : 0;
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += n;
      }
      warp_tot[lane] = wi - w;  // exclusive prefix of warp totals
      if (lane == 31) carry = base_off + wi;  // everyone read the old value before the barrier above
    }
    __syncthreads();
    int off = base_off + warp_tot[warp] + incl - cnt;
#pragma unroll
    for (int k = 0; k < kCollapseItems; ++k)
      if (keep & (1u << k)) out[off++] = v[k];
    __syncthreads();  // carry / warp_tot are rewritten next iteration
  }
  if (threadIdx.x == 0) n_ids[b] = carry;
}

__global__ void __launch_bounds__(kCollapseThreads)
collapse_kernel(const int32_t* __restrict__ path, int T, int blank, const int32_t* __restrict__ lengths,
                int32_t* __restrict__ ids, int32_t* __restrict__ n_ids) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  collapse_item<kCollapseThreads>(path, T, blank, lengths, ids, n_ids, blockIdx.x, warp_tot, &carry);
}

template <bool VEC>
__global__ void __launch_bounds__(256)
argmax_rows_kernel(const float* __restrict__ lp, int64_t sB, int64_t sT, int B, int T, int C,
                   const int32_t* __restrict__ lengths, int32_t* __restrict__ path) {
  argmax_rows<VEC>(lp, sB, sT, B, T, C, lengths, path);
}

// argmax + collapse in ONE launch: the CTA that finishes last (a ticket counter in caller-provided scratch, left at
// zero again for the next call) collapses the path of every item.
template <bool VEC>
__global__ void __launch_bounds__(kGreedyThreads)
greedy_fused_kernel(const float* __restrict__ lp, int64_t sB, int64_t sT, int B, int T, int C,
                    const int32_t* __restrict__ lengths, int blank, int32_t* __restrict__ path,
                    int32_t* __restrict__ ids, int32_t* __restrict__ n_ids, unsigned int* __restrict__ counter) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  __shared__ int last_s;
  argmax_rows<VEC>(lp, sB, sT, B, T, C, lengths, path);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int ticket = atomicAdd(counter, 1u);
    last_s = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!last_s) return;
  __threadfence();
  for (int b = 0; b < B; ++b) collapse_item<kGreedyThreads>(path, T, blank, lengths, ids, n_ids, b, warp_tot, &carry);
  if (threadIdx.x == 0) *counter = 0u;
}

}  // namespace dae

extern "C" size_t dae_greedy_scratch_bytes(void) { return 256; }

extern "C" int dae_greedy_collapse(const float* lp, int64_t sB, int64_t sT, int B, int T, int C,
                                   const int32_t* lengths, int blank, int32_t* path, int32_t* ids,
                                   int32_t* n_ids, void* scratch, void* stream) {
  using namespace dae;
  if (!lp || !path || !ids || !n_ids || B < 0 || T < 0 || C <= 0) return DAE_E_BADARG;
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (T > 0 && scratch) {
    // fused single launch; two CTAs of 16 warps per SM keep ~128 KB of loads in flight per SM
    const int64_t nrows = (int64_t)B * T;
    const int grid = (int)(nrows < (int64_t)kNumSMs * 2 ? nrows : (int64_t)kNumSMs * 2);   // round-robin rows: equal bytes per SM
    const bool vec = aligned16(lp) && (C % 4 == 0) && (sT % 4 == 0) && (sB % 4 == 0);
    unsigned int* counter = (unsigned int*)scratch;
    if (vec)
      greedy_fused_kernel<true><<<grid, kGreedyThreads, 0, st>>>(lp, sB, sT, B, T, C, lengths, blank, path, ids, n_ids, counter);
    else
      greedy_fused_kernel<false><<<grid, kGreedyThreads, 0, st>>>(lp, sB, sT, B, T, C, lengths, blank, path, ids, n_ids, counter);
    DAE_LAUNCH_OK();
    return 0;
  }
  if (T > 0) {
    const int64_t nrows = (int64_t)B * T;
    const int warps = 8;
    int64_t want = (nrows + warps - 1) / warps;
    const int grid = (int)(want < (int64_t)kNumSMs * 8 ? want : (int64_t)kNumSMs * 8);
    const bool vec = aligned16(lp) && (C % 4 == 0) && (sT % 4 == 0) && (sB % 4 == 0);
    if (vec)
      argmax_rows_kernel<true><<<grid, warps * 32, 0, st>>>(lp, sB, sT, B, T, C, lengths, path);
    else
      argmax_rows_kernel<false><<<grid, warps * 32, 0, st>>>(lp, sB, sT, B, T, C, lengths, path);
    DAE_LAUNCH_OK();
  }
  collapse_kernel<<<B, kCollapseThreads, 0, st>>>(path, T, blank, lengths, ids, n_ids);
  DAE_LAUNCH_OK();
  return 0;
}

extern "C" int dae_collapse_path(const int32_t* path, int B, int T, const int32_t* lengths, int blank,
                                 int32_t* ids, int32_t* n_ids, void* stream) {
  using namespace dae;
  if (!path || !ids || !n_ids || B < 0 || T < 0) return DAE_E_BADARG;
  if (B == 0) return 0;
  collapse_kernel<<<B, kCollapseThreads, 0, (cudaStream_t)stream>>>(path, T, blank, lengths, ids, n_ids);
  DAE_LAUNCH_OK();
  return 0;
}
