// soft-DTW forward / backward as an anti-diagonal wavefront (replaces the numba kernels at
// lcasr_nemo/soft_dtw_cuda.py:33-111 and the CPU kernels at :184-239 that the reference falls back
// to above 1024 frames, :312-314).
//
// Decomposition.  The [N, M] cell grid of one sample is cut into bands of 32 rows.  One WARP owns a
// band (lane = row) and sweeps it left to right along anti-diagonals: at step p lane i is at column
// p - i, so the three predecessors are the lane's own previous value, the upper lane's previous
// value (one __shfl_up) and the upper lane's value from two steps ago (kept in a register).  There is
// no CTA-wide barrier anywhere.
//   * D (and R in the backward pass) are staged in shared memory as 32x32 tiles by cp.async, three or
//     four tiles ahead; lane i reads tile[i][(p-i)&31], which is bank-conflict free.
//   * Results are staged in a 32x32 smem tile and flushed with coalesced 128-bit stores.
//   * A band needs the bottom row of the band above.  Lane 31 publishes every bottom-row value the
//     moment it exists as an 8-byte {value, column+1} word in a zero-initialised export buffer; the
//     consumer polls the tag in the data itself, so there are no flags and no fences, and a band
//     trails its predecessor by the ideal 32 steps plus one L2 round trip.
//   * Bands are handed to warps through an atomic ticket in (band, sample) order, so a waiting warp
//     only ever waits for a warp that already started: deadlock-free for any grid size.
// The backward pass is the same sweep on the flipped grid (i' = N-1-i, j' = M-1-j) with
// E = E_dn*a + E_right*b + E_diag*c, a/b/c = exp((W[.] - R[i,j])/gamma), W = R - D (:100-108).
#include "common.cuh"

namespace dae {

constexpr int kBand = 32;                 // rows per warp
constexpr int kTile = 32;                 // columns per staged tile
// columns per export poll group: the group head (poll, prefetch, smem reads) is amortised over the group, the band
// below trails by a few more steps; measured best at 8 forward and 16 backward ([8,4096,4096])
constexpr int kGrpFwd = 8, kGrpBwd = 16;
constexpr int kDepth = 1;                 // groups between requesting band-above values and using them
constexpr float kLog2eF = 1.4426950408889634f;
constexpr float kLn2F = 0.6931471805599453f;

struct SdtwParams {
  const float* D;        // [B,N,M]
  float* R;              // [B,N,M]   fwd: out, bwd: in
  float* E;              // [B,N,M]   bwd: out (already scaled by gout[b])
  float* out;            // [B]       fwd: R[N-1,M-1]
  const float* gout;     // [B]       bwd: upstream gradient
  int64_t gout_stride;
  int2* exp_buf;         // [B][nbands][Mp] published bottom rows {value bits, column+1}
  int* ticket;           // [1]
  int B, N, M, nbands, Mp;
  float gamma, bandwidth;
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4s(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }
__device__ __forceinline__ int2 ld_volatile_int2(const int2* p) {
  int2 v;
  asm volatile("ld.volatile.global.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_int2(int2* p, int2 v) {
  asm volatile("st.volatile.global.v2.s32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ float ex2f_(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2f_(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

constexpr int kRing = 4;                  // staged input tiles per matrix (power of two: ring index = col & 127)
constexpr int kRingCols = kRing * kTile;  // 128
constexpr int kOutCols = 2 * kTile;       // result ring: two tiles

// Stage one 32-column tile (columns c0.. of the possibly flipped grid) of a [N,M] matrix into the smem
// ring sm[32][128] at ring slot (c0/32)&3.  Out-of-range elements are left untouched (never read by an
// active lane).  FLIP stages the flipped grid: smem row i' holds matrix row N-1-(r0+i'); the 32 columns
// are stored in matrix order (cp.async cannot reverse inside a 16-byte chunk) and the reader indexes
// them with (col ^ 31).
template <bool FLIP>
__device__ __forceinline__ void stage_tile(float* sm, const float* __restrict__ mat, int N, int M, int r0, int c0,
                                           int lane, bool vec) {
  const int cbase = FLIP ? (M - c0 - kTile) : c0;
  float* dst0 = sm + ((c0 >> 5) & (kRing - 1)) * kTile;
  if (vec) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {       // 8 lanes x 16 B cover one row; 4 rows per instruction
      const int ri = it * 4 + (lane >> 3);
      const int x = (lane & 7) * 4;
      const int row = FLIP ? (N - 1 - (r0 + ri)) : (r0 + ri);
      const int col = cbase + x;
      if (row >= 0 && row < N) {
        if (col >= 0 && col + 3 < M) {
          cp_async16(dst0 + ri * kRingCols + x, mat + (int64_t)row * M + col);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (col + e >= 0 && col + e < M) cp_async4s(dst0 + ri * kRingCols + x + e, mat + (int64_t)row * M + col + e);
        }
      }
    }
  } else {
    for (int ri = 0; ri < kBand; ++ri) {
      const int row = FLIP ? (N - 1 - (r0 + ri)) : (r0 + ri);
      const int col = cbase + lane;
      if (row >= 0 && row < N && col >= 0 && col < M) cp_async4s(dst0 + ri * kRingCols + lane, mat + (int64_t)row * M + col);
    }
  }
}

// Flush a finished 32-column result tile (ring sm[32][64], slot (c0/32)&1) with coalesced stores.
template <bool FLIP>
__device__ __forceinline__ void flush_tile(const float* sm, float* __restrict__ mat, int N, int M, int r0, int c0,
                                           int lane, bool vec) {
  const int cbase = FLIP ? (M - c0 - kTile) : c0;
  const float* src0 = sm + ((c0 >> 5) & 1) * kTile;
  if (vec) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int ri = it * 4 + (lane >> 3);
      const int x = (lane & 7) * 4;
      const int row = FLIP ? (N - 1 - (r0 + ri)) : (r0 + ri);
      const int col = cbase + x;
      if (row >= 0 && row < N) {
        if (col >= 0 && col + 3 < M) {
          st_stream4(mat + (int64_t)row * M + col, *reinterpret_cast<const float4*>(src0 + ri * kOutCols + x));
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (col + e >= 0 && col + e < M) mat[(int64_t)row * M + col + e] = src0[ri * kOutCols + x + e];
        }
      }
    }
  } else {
    for (int ri = 0; ri < kBand; ++ri) {
      const int row = FLIP ? (N - 1 - (r0 + ri)) : (r0 + ri);
      const int col = cbase + lane;
      if (row >= 0 && row < N && col >= 0 && col < M) mat[(int64_t)row * M + col] = src0[ri * kOutCols + lane];
    }
  }
}

template <bool BWD, bool PRUNE>
__global__ void __launch_bounds__(256)
softdtw_wave_kernel(SdtwParams P, int vec, int smem_per_warp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kGrp = BWD ? kGrpBwd : kGrpFwd;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sm = reinterpret_cast<float*>(smem_raw + (size_t)warp * smem_per_warp);
  float* sD = sm;                                            // [32][128]
  float* sR = BWD ? (sD + kBand * kRingCols) : nullptr;      // [32][128] (backward only)
  float* sO = (BWD ? sR : sD) + kBand * kRingCols;           // [32][64]
  const float* sDl = sD + lane * kRingCols;
  const float* sRl = BWD ? (sR + lane * kRingCols) : nullptr;
  float* sOl = sO + lane * kOutCols;
  const int N = P.N, M = P.M;
  const int n_agents = P.B * P.nbands;
  const int n_tiles = (M + kTile - 1) / kTile;
  const float ig2 = kLog2eF / P.gamma;                       // 1/gamma in log2 units
  const float g_ln2 = P.gamma * kLn2F;
  const float INF = CUDART_INF_F;
  const int steps = M + kBand - 1;
  const int k_last = (steps - 1) >> 5;
  const int flipx = BWD ? (kTile - 1) : 0;                   // flipped tiles are stored in matrix order

  for (;;) {
    int tk = 0;
    if (lane == 0) tk = atomicAdd(P.ticket, 1);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    if (tk >= n_agents) break;
    const int band = tk / P.B, b = tk - band * P.B;          // (band, sample) order
    const int r0 = band * kBand;                             // first (flipped) row of this band
    const int frow = r0 + lane;
    const bool row_ok = frow < N;
    const int row = BWD ? (N - 1 - frow) : frow;             // matrix row
    const float* Db = P.D + (int64_t)b * N * M;
    float* Rb = P.R + (int64_t)b * N * M;
    float* Ob = BWD ? (P.E + (int64_t)b * N * M) : Rb;
    int2* exp_mine = P.exp_buf + ((int64_t)b * P.nbands + band) * P.Mp;
    const int2* exp_up = band > 0 ? (P.exp_buf + ((int64_t)b * P.nbands + band - 1) * P.Mp) : nullptr;
    const bool publish = (lane == kBand - 1) && (band + 1 < P.nbands);
    // W of the upper band's bottom row (backward): R - D read straight from global, column index flipped
    const float* upR = BWD && band > 0 ? (Rb + (int64_t)(N - r0) * M) : nullptr;
    const float* upD = BWD && band > 0 ? (Db + (int64_t)(N - r0) * M) : nullptr;

    __syncwarp();
    // prologue: tiles 0 and 1 (tile k+2 is staged when lane 0 enters tile k)
    int staged = 0;
    for (; staged < 2 && staged < n_tiles; ++staged) {
      stage_tile<BWD>(sD, Db, N, M, r0, staged * kTile, lane, vec);
      if (BWD) stage_tile<BWD>(sR, Rb, N, M, r0, staged * kTile, lane, vec);
      cp_commit();
    }

    // Wavefront state.  Borders come from the initial values: before a lane's first cell its `left` is
    // the left border and the upper lane still holds its own initial (border) value.
    float my_v = BWD ? 0.0f : INF, my_w = -INF;              // this lane's latest cell
    float up_v = my_v, up_w = -INF;                          // upper lane's cell one step ago (= diagonal)
    if (band == 0 && lane == 0) {                            // the corner the recursion starts from
      up_v = BWD ? P.gout[b * P.gout_stride] : 0.0f;
      up_w = BWD ? Rb[(int64_t)(N - 1) * M + (M - 1)] : 0.0f;
    }
    float ab_v = BWD ? 0.0f : INF, ab_w = -INF;              // bottom row of the band above, kGrp columns
    // export words of the band above are requested kDepth groups before they are needed (tag-in-data:
    // {value, column+1}); lanes 0..kGrp-1 hold one column each
    int2 nx[kDepth];
    float nx_r[kDepth], nx_d[kDepth];
#pragma unroll
    for (int u = 0; u < kDepth; ++u) {
      nx[u] = make_int2(0, 0); nx_r[u] = 0.0f; nx_d[u] = 0.0f;
      const int col = u * kGrp + lane;
      if (exp_up && lane < kGrp && col < M) {
        nx[u] = ld_volatile_int2(exp_up + col);
        if (BWD) { nx_r[u] = __ldcg(upR + (M - 1 - col)); nx_d[u] = __ldcg(upD + (M - 1 - col)); }
      }
    }

    int p = 0;
    int j = -lane;                                           // this lane's (flipped) column at step p
    for (int k = 0; k <= k_last; ++k) {
      // ---- once per 32 steps: lane 0 enters tile k
      if (staged < n_tiles && staged <= k + 2) {             // ring slot (k+2)&3 was tile k-2's: free
        stage_tile<BWD>(sD, Db, N, M, r0, staged * kTile, lane, vec);
        if (BWD) stage_tile<BWD>(sR, Rb, N, M, r0, staged * kTile, lane, vec);
        cp_commit();
        ++staged;
      }
      if (k < n_tiles) {
        const int newer = staged - 1 - k;                    // tiles staged after tile k may stay in flight
        if (newer >= 2) cp_wait<2>(); else if (newer == 1) cp_wait<1>(); else cp_wait<0>();
        __syncwarp();
      }
      if (k >= 2 && k - 2 < n_tiles) {                       // tile k-2 is complete: flush it
        __syncwarp();
        flush_tile<BWD>(sO, Ob, N, M, r0, (k - 2) * kTile, lane, vec);
        __syncwarp();
      }
#pragma unroll 1
      for (int gq = 0; gq < kTile / kGrp; ++gq) {
        // ---- once per kGrp steps: take the kGrp band-above values requested kDepth groups ago (re-poll only
        // if they had not been published yet), then request the group needed kDepth groups from now.
        if (exp_up && p < M) {
          constexpr int us = 0;                              // kDepth == 1: a single request in flight
          const int col = p + lane;
          const bool need = lane < kGrp && col < M;
          int2 w = nx[us];
          for (;;) {
            if (__all_sync(0xffffffffu, !need || w.y == col + 1)) break;
            if (need) w = ld_volatile_int2(exp_up + col);
          }
          ab_v = __int_as_float(w.x);
          if (BWD) ab_w = (PRUNE && isinf(nx_r[us])) ? -INF : nx_r[us] - nx_d[us];
          const int col2 = col + kDepth * kGrp;
          if (lane < kGrp && col2 < M) {
            nx[us] = ld_volatile_int2(exp_up + col2);
            if (BWD) { nx_r[us] = __ldcg(upR + (M - 1 - col2)); nx_d[us] = __ldcg(upD + (M - 1 - col2)); }
          }
        }
        // inputs of the next kGrp cells leave shared memory before the dependent chain starts
        float dq[kGrp], rq[kGrp];
#pragma unroll
        for (int q = 0; q < kGrp; ++q) {
          const int sc = ((j + q) & (kRingCols - 1)) ^ flipx;
          dq[q] = sDl[sc];
          rq[q] = BWD ? sRl[sc] : 0.0f;
        }
        float pub[kGrp];
#pragma unroll
        for (int q = 0; q < kGrp; ++q, ++p, ++j) {
          // upper neighbour: the upper lane's latest cell (its column equals ours); lane 0 takes the band above
          float nu_v = __shfl_up_sync(0xffffffffu, my_v, 1);
          float nu_w = BWD ? __shfl_up_sync(0xffffffffu, my_w, 1) : 0.0f;
          {
            const float a_v = __shfl_sync(0xffffffffu, ab_v, q);
            const float a_w = BWD ? __shfl_sync(0xffffffffu, ab_w, q) : 0.0f;
            if (lane == 0) { nu_v = a_v; nu_w = a_w; }
          }
          const float d = dq[q];
          const bool active = row_ok && (unsigned)j < (unsigned)M;
          float res_v, res_w = 0.0f;
          if (!BWD) {
            // R = D + softmin_gamma(diag, up, left)  (soft_dtw_cuda.py:65-72); left = my_v, diag = up_v.
            // Without pruning every cell has a finite predecessor, so mn is finite and (mn - inf) is safe.
            const float mn = fminf(fminf(up_v, nu_v), my_v);
            const float s = ex2f_((mn - up_v) * ig2) + ex2f_((mn - nu_v) * ig2) + ex2f_((mn - my_v) * ig2);
            res_v = d + (mn - g_ln2 * lg2f_(s));
            if (PRUNE) {
              if (mn == INF || fabsf((float)(row - j)) > P.bandwidth) res_v = INF;
            }
          } else {
            // E = E_dn*a + E_right*b + E_diag*c, a,b,c = exp((W[.] - R[i,j]) / gamma)  (:100-108)
            float r = rq[q];
            if (PRUNE && isinf(r)) r = -INF;                 // :96-97 (only pruned cells hold +inf)
            res_w = r - d;
            const float a = ex2f_((nu_w - r) * ig2), bb = ex2f_((my_w - r) * ig2), c = ex2f_((up_w - r) * ig2);
            res_v = nu_v * a + my_v * bb + up_v * c;
            if (PRUNE) {
              if (r == -INF) { res_v = 0.0f; res_w = -INF; }
              if (fabsf((float)(row - (M - 1 - j))) > P.bandwidth) res_v = 0.0f;
            }
          }
          up_v = nu_v; up_w = nu_w;
          if (active) {
            my_v = res_v; my_w = res_w;
            sOl[(j & (kOutCols - 1)) ^ flipx] = res_v;
          }
          pub[q] = res_v;
        }
        // publish the band's bottom row for the band below: {value, column+1} in one 8-byte store per cell, after
        // the group so that the volatile stores do not fence the chain's shared-memory traffic
        if (publish) {
#pragma unroll
          for (int q = 0; q < kGrp; ++q) {
            const int jq = j - kGrp + q;
            if ((unsigned)jq < (unsigned)M && row_ok)
              st_volatile_int2(exp_mine + jq, make_int2(__float_as_int(pub[q]), jq + 1));
          }
        }
      }
    }
    // flush the tiles not yet flushed inside the loop (at most two)
    __syncwarp();
    for (int k = (k_last - 1 > 0 ? k_last - 1 : 0); k < n_tiles; ++k)
      flush_tile<BWD>(sO, Ob, N, M, r0, k * kTile, lane, vec);
    __syncwarp();
    if (!BWD && band == P.nbands - 1) {
      const float v = __shfl_sync(0xffffffffu, my_v, (N - 1) - r0);   // R[N-1, M-1]
      if (lane == 0) P.out[b] = v;
    }
  }
}

}  // namespace dae

extern "C" size_t dae_softdtw_scratch_bytes(int B, int N, int M) {
  if (B < 0 || N < 0 || M < 0) return 0;
  const size_t nb = (size_t)(N + dae::kBand - 1) / dae::kBand;
  const size_t Mp = ((size_t)M + 3) / 4 * 4;
  return 256 + (size_t)B * nb * Mp * sizeof(int2);
}

template <bool BWD>
static int softdtw_launch(const float* D, float* R, float* E, float* out, const float* gout, int64_t gout_stride,
                          int B, int N, int M, float gamma, float bandwidth, void* scratch, size_t scratch_bytes,
                          cudaStream_t st) {
  using namespace dae;
  if (!D || !R || B < 0 || N < 0 || M < 0 || !(gamma > 0.0f)) return DAE_E_BADARG;
  if (B == 0 || N == 0 || M == 0) return 0;
  if (!scratch || scratch_bytes < dae_softdtw_scratch_bytes(B, N, M)) return DAE_E_SCRATCH;
  if ((reinterpret_cast<uintptr_t>(scratch) & 255u) != 0) return DAE_E_ALIGN;
  SdtwParams P;
  P.D = D; P.R = R; P.E = E; P.out = out; P.gout = gout; P.gout_stride = gout_stride;
  P.B = B; P.N = N; P.M = M;
  P.nbands = (N + kBand - 1) / kBand;
  P.Mp = (M + 3) / 4 * 4;
  P.ticket = reinterpret_cast<int*>(scratch);
  P.exp_buf = reinterpret_cast<int2*>(reinterpret_cast<char*>(scratch) + 256);
  P.gamma = gamma; P.bandwidth = bandwidth;
  DAE_CUDA(cudaMemsetAsync(scratch, 0, dae_softdtw_scratch_bytes(B, N, M), st));
  const int smem_per_warp = ((BWD ? 2 : 1) * kBand * kRingCols + kBand * kOutCols) * 4;   // 24 KB / 40 KB
  const int warps = BWD ? 5 : 8;
  const int smem = smem_per_warp * warps;
  const bool prune = bandwidth > 0.0f;
  auto kern = prune ? softdtw_wave_kernel<BWD, true> : softdtw_wave_kernel<BWD, false>;
  DAE_CUDA(ensure_dyn_smem(kern, smem));
  const int agents = B * P.nbands;
  int grid = (agents + warps - 1) / warps;
  if (grid > kNumSMs) grid = kNumSMs;                     // persistent: one CTA per SM, tickets do the rest
  const int vec = aligned16(D) && aligned16(R) && (!BWD || aligned16(E)) && (M % 4 == 0);
  kern<<<grid, warps * 32, smem, st>>>(P, vec, smem_per_warp);
  DAE_LAUNCH_OK();
  return 0;
}

extern "C" int dae_softdtw_fwd(const float* D, int B, int N, int M, float gamma, float bandwidth, float* R,
                               float* out, void* scratch, size_t scratch_bytes, void* stream) {
  if (!out) return DAE_E_BADARG;
  return softdtw_launch<false>(D, R, nullptr, out, nullptr, 0, B, N, M, gamma, bandwidth, scratch, scratch_bytes,
                               (cudaStream_t)stream);
}

extern "C" int dae_softdtw_bwd(const float* D, const float* R, const float* gout, int64_t gout_stride, int B, int N,
                               int M, float gamma, float bandwidth, float* E, void* scratch, size_t scratch_bytes,
                               void* stream) {
  if (!E || !gout) return DAE_E_BADARG;
  return softdtw_launch<true>(D, const_cast<float*>(R), E, nullptr, gout, gout_stride, B, N, M, gamma, bandwidth,
                              scratch, scratch_bytes, (cudaStream_t)stream);
}
