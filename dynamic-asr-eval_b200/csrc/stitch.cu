// Overlap-average stitch of window posteriors with a fused per-row argmax.
// Arithmetic follows lcasr/lib.py:604-629: p = exp(lp); sum windows in start order;
// divide by the cover count; log.  The reference does this with CPU buffers of
// [spec_n//4 + seq_len, C] floats; here every output row gathers its covering windows.
#include "common.cuh"

namespace dae {

__global__ void __launch_bounds__(256)
stitch_kernel(const float* __restrict__ lp, int C, const int64_t* __restrict__ win_off,
              const int64_t* __restrict__ win_pos, const int64_t* __restrict__ win_len, int n_win,
              const int64_t* __restrict__ row_map, float* __restrict__ out, int32_t* __restrict__ path, int vec) {
  __shared__ float rv[8];
  __shared__ int ri[8];
  __shared__ int w_first_s, w_count_s;
  const int64_t r = blockIdx.x;
  const int64_t p = row_map[r];
  const int tid = threadIdx.x, NT = blockDim.x;
  if (tid == 0) {
    // windows are sorted by position; first window whose end is past p
    int lo = 0, hi = n_win;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (win_pos[mid] + win_len[mid] > p) hi = mid; else lo = mid + 1;
    }
    int cnt = 0;
    for (int w = lo; w < n_win && win_pos[w] <= p; ++w)
      if (p < win_pos[w] + win_len[w]) ++cnt;
    w_first_s = lo;
    w_count_s = cnt;
  }
  __syncthreads();
  const int w0 = w_first_s;
  const float cnt = (float)w_count_s;
  float bv = -CUDART_INF_F;
  int bi = 0x7fffffff;
  float* orow = out + r * C;
  if (vec) {
    for (int i = tid; i < (C >> 2); i += NT) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int w = w0; w < n_win && win_pos[w] <= p; ++w) {
        if (p >= win_pos[w] + win_len[w]) continue;
        const float4 v = ld_stream4(lp + (win_off[w] + (p - win_pos[w])) * C + 4 * i);
        acc.x += __expf(v.x); acc.y += __expf(v.y); acc.z += __expf(v.z); acc.w += __expf(v.w);
      }
      float4 o;
      o.x = __logf(__fdiv_rn(acc.x, cnt)); o.y = __logf(__fdiv_rn(acc.y, cnt));
      o.z = __logf(__fdiv_rn(acc.z, cnt)); o.w = __logf(__fdiv_rn(acc.w, cnt));
      st_stream4(orow + 4 * i, o);
      if (o.x > bv) { bv = o.x; bi = 4 * i; }
      if (o.y > bv) { bv = o.y; bi = 4 * i + 1; }
      if (o.z > bv) { bv = o.z; bi = 4 * i + 2; }
      if (o.w > bv) { bv = o.w; bi = 4 * i + 3; }
    }
  } else {
    for (int i = tid; i < C; i += NT) {
      float acc = 0.f;
      for (int w = w0; w < n_win && win_pos[w] <= p; ++w) {
        if (p >= win_pos[w] + win_len[w]) continue;
        acc += __expf(ld_stream1(lp + (win_off[w] + (p - win_pos[w])) * C + i));
      }
      const float o = __logf(__fdiv_rn(acc, cnt));
      orow[i] = o;
      if (o > bv) { bv = o; bi = i; }
    }
  }
  if (!path) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if ((tid & 31) == 0) { rv[tid >> 5] = bv; ri[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < (NT >> 5); ++w)
      if (rv[w] > bv || (rv[w] == bv && ri[w] < bi)) { bv = rv[w]; bi = ri[w]; }
    path[r] = (bi == 0x7fffffff) ? 0 : bi;
  }
}

}  // namespace dae

extern "C" int dae_stitch(const float* lp, int C, const int64_t* win_off, const int64_t* win_pos,
                          const int64_t* win_len, int n_win, const int64_t* row_map, int64_t n_out,
                          float* out, int32_t* path, void* stream) {
  using namespace dae;
  if (!lp || !win_off || !win_pos || !win_len || !row_map || !out || C <= 0 || n_win < 0 || n_out < 0) return DAE_E_BADARG;
  if (n_out == 0) return 0;
  if (n_out > 0x7fffffffLL) return DAE_E_TOOBIG;
  const int vec = aligned16(lp) && aligned16(out) && (C % 4 == 0);
  int NT = (((vec ? C / 4 : C) + 31) / 32) * 32;
  NT = NT < 32 ? 32 : (NT > 256 ? 256 : NT);
  stitch_kernel<<<(unsigned)n_out, NT, 0, (cudaStream_t)stream>>>(lp, C, win_off, win_pos, win_len, n_win, row_map, out, path, vec);
  DAE_LAUNCH_OK();
  return 0;
}
