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
constexpr int kGrp = 8;                   // columns per export poll group
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
__device__ __forceinline__ float ex2f_(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2f_(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Stage one 32x32 tile (rows r0.., columns c0..) of a [N,M] matrix into smem[32][32]; out-of-range
// elements are left untouched (never read by an active lane).  `flip` stages the flipped grid:
// smem row i' holds matrix row N-1-(r0+i'), smem column x holds matrix column cbase+x where the
// caller passes cbase = M - c0 - 32 (may be negative for the last, partial tile).
template <bool FLIP>
__device__ __forceinline__ void stage_tile(float* sm, const float* __restrict__ mat, int N, int M, int r0, int c0,
                                           int lane, bool vec) {
  const int cbase = FLIP ? (M - c0 - kTile) : c0;
  if (vec) {
    // 8 lanes x 16 B cover one row; 4 rows per instruction
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int ri = it * 4 + (lane >> 3);
      const int x = (lane & 7) * 4;
      const int row = FLIP ? (N - 1 - (r0 + ri)) : (r0 + ri);
      const int col = cbase + x;
      if (row >= 0 && row < N && col >= 0 && col + 3 < M) {
        cp_async16(sm + ri * kTile + x, mat + (int64_t)row * M + col);
      } else if (row >= 0 && row < N) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col + e >= 0 && col + e < M) cp_async4s(sm + ri * kTile + x + e, mat + (int64_t)row * M + col + e);
      }
    }
  } else {
    for (int ri = 0; ri < kBand; ++ri) {
      const int row = FLIP ? (N - 1 - (r0 + ri)) : (r0 + ri);
      const int col = cbase + lane;
      if (row >= 0 && row < N && col >= 0 && col < M) cp_async4s(sm + ri * kTile + lane, mat + (int64_t)row * M + col);
    }
  }
}

// Flush a finished 32x32 result tile to the [N,M] output with coalesced stores.
template <bool FLIP>
__device__ __forceinline__ void flush_tile(const float* sm, float* __restrict__ mat, int N, int M, int r0, int c0,
                                           int lane, bool vec) {
  const int cbase = FLIP ? (M - c0 - kTile) : c0;
  if (vec) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int ri = it * 4 + (lane >> 3);
      const int x = (lane & 7) * 4;
      const int row = FLIP ? (N - 1 - (r0 + ri)) : (r0 + ri);
      const int col = cbase + x;
      if (row >= 0 && row < N) {
        if (col >= 0 && col + 3 < M) {
          st_stream4(mat + (int64_t)row * M + col, *reinterpret_cast<const float4*>(sm + ri * kTile + x));
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (col + e >= 0 && col + e < M) mat[(int64_t)row * M + col + e] = sm[ri * kTile + x + e];
        }
      }
    }
  } else {
    for (int ri = 0; ri < kBand; ++ri) {
      const int row = FLIP ? (N - 1 - (r0 + ri)) : (r0 + ri);
      const int col = cbase + lane;
      if (row >= 0 && row < N && col >= 0 && col < M) mat[(int64_t)row * M + col] = sm[ri * kTile + lane];
    }
  }
}

template <bool BWD>
__global__ void __launch_bounds__(256)
softdtw_wave_kernel(SdtwParams P, int vec, int smem_per_warp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int RING = BWD ? 3 : 4;       // staged input tiles per matrix
  constexpr int AHEAD = RING - 2;         // tiles prefetched beyond the two in use
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sm = reinterpret_cast<float*>(smem_raw + (size_t)warp * smem_per_warp);
  float* sD = sm;                                           // [RING][32][32]
  float* sR = BWD ? (sD + RING * kBand * kTile) : nullptr;  // [RING][32][32] (backward only)
  float* sO = (BWD ? sR : sD) + RING * kBand * kTile;       // [2][32][32]
  const int N = P.N, M = P.M;
  const int n_agents = P.B * P.nbands;
  const int n_tiles = (M + kTile - 1) / kTile;
  const float ig2 = kLog2eF / P.gamma;                      // 1/gamma in log2 units
  const float g_ln2 = P.gamma * kLn2F;
  const bool prune = P.bandwidth > 0.0f;
  const float INF = CUDART_INF_F;

  for (;;) {
    int tk = 0;
    if (lane == 0) tk = atomicAdd(P.ticket, 1);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    if (tk >= n_agents) break;
    const int band = tk / P.B, b = tk - band * P.B;         // (band, sample) order
    const int r0 = band * kBand;                            // first (flipped) row of this band
    const int frow = r0 + lane;                             // this lane's (flipped) row
    const bool row_ok = frow < N;
    const int row = BWD ? (N - 1 - frow) : frow;            // matrix row
    const float* Db = P.D + (int64_t)b * N * M;
    float* Rb = P.R + (int64_t)b * N * M;
    float* Ob = BWD ? (P.E + (int64_t)b * N * M) : Rb;
    int2* exp_mine = P.exp_buf + ((int64_t)b * P.nbands + band) * P.Mp;
    const int2* exp_up = band > 0 ? (P.exp_buf + ((int64_t)b * P.nbands + band - 1) * P.Mp) : nullptr;
    const float seed = BWD ? P.gout[b * P.gout_stride] : 0.0f;

    // prologue: stage tiles 0 .. AHEAD-1 (tile k+AHEAD is staged when lane 0 enters tile k)
    for (int k = 0; k < AHEAD && k < n_tiles; ++k) {
      stage_tile<BWD>(sD + (k % RING) * kBand * kTile, Db, N, M, r0, k * kTile, lane, vec);
      if (BWD) stage_tile<BWD>(sR + (k % RING) * kBand * kTile, Rb, N, M, r0, k * kTile, lane, vec);
      cp_commit();
    }
    int staged = (AHEAD < n_tiles) ? AHEAD : n_tiles;             // tiles staged so far

    // per-lane wavefront state
    float left_v = BWD ? 0.0f : INF, left_w = -INF;         // own previous cell (value, W)
    float up_v = BWD ? 0.0f : INF, up_w = -INF;             // upper lane's previous-step cell
    float my_v = BWD ? 0.0f : INF, my_w = -INF;             // this lane's latest cell
    // bottom row of the band above, one poll group (kGrp columns) at a time, lanes 0..kGrp-1 hold it
    float ab_v = BWD ? 0.0f : INF, ab_w = -INF;             // current group
    const float r_corner = BWD ? Rb[(int64_t)(N - 1) * M + (M - 1)] : 0.0f;

    const int steps = M + kBand - 1;
    for (int p = 0; p < steps; ++p) {
      if ((p & (kTile - 1)) == 0) {
        const int k = p >> 5;                               // lane 0 enters tile k
        // slot (k+AHEAD)%RING was last used by tile k-2, which no lane touches any more
        if (staged < n_tiles && staged <= k + AHEAD) {
          stage_tile<BWD>(sD + (staged % RING) * kBand * kTile, Db, N, M, r0, staged * kTile, lane, vec);
          if (BWD) stage_tile<BWD>(sR + (staged % RING) * kBand * kTile, Rb, N, M, r0, staged * kTile, lane, vec);
          cp_commit();
          ++staged;
        }
        if (k < n_tiles) {
          // tile k must have landed; the tiles staged after it may stay in flight
          const int newer = staged - 1 - k;
          if (newer >= 2) cp_wait<2>(); else if (newer == 1) cp_wait<1>(); else cp_wait<0>();
          __syncwarp();
        }
        if (k >= 2 && k - 2 < n_tiles) {                    // tile k-2 is complete: flush it
          flush_tile<BWD>(sO + ((k - 2) & 1) * kBand * kTile, Ob, N, M, r0, (k - 2) * kTile, lane, vec);
          __syncwarp();
        }
      }
      if ((p & (kGrp - 1)) == 0 && p < M) {                 // lane 0 enters export group p/kGrp
        if (exp_up) {
          const int col = p + lane;                         // lanes 0..kGrp-1 poll their column
          const bool need = lane < kGrp && col < M;
          int2 w = make_int2(0, 0);
          for (;;) {
            if (need) w = ld_volatile_int2(exp_up + col);
            if (__all_sync(0xffffffffu, !need || w.y == col + 1)) break;
          }
          ab_v = __int_as_float(w.x);
          if (BWD && need) {                                // W of the upper band's bottom row = R - D there
            const int urow = N - 1 - (r0 - 1), ucol = M - 1 - col;
            const float ur = __ldcg(Rb + (int64_t)urow * M + ucol);
            ab_w = isinf(ur) ? -INF : ur - __ldcg(Db + (int64_t)urow * M + ucol);
          }
        } else {
          ab_v = BWD ? 0.0f : INF;
          ab_w = -INF;
        }
      }
      const int j = p - lane;                               // this lane's (flipped) column
      // upper neighbour of this step = upper lane's latest cell (its column is j as well)
      float nu_v = __shfl_up_sync(0xffffffffu, my_v, 1);
      float nu_w = BWD ? __shfl_up_sync(0xffffffffu, my_w, 1) : 0.0f;
      {
        const float a_v = __shfl_sync(0xffffffffu, ab_v, p & (kGrp - 1));
        const float a_w = BWD ? __shfl_sync(0xffffffffu, ab_w, p & (kGrp - 1)) : 0.0f;
        if (lane == 0) { nu_v = a_v; nu_w = a_w; }
      }
      float dg_v = up_v, dg_w = up_w;                       // diagonal = upper neighbour one step ago
      float lf_v = left_v, lf_w = left_w;
      if (j == 0) {                                         // left border of the grid
        lf_v = BWD ? 0.0f : INF; lf_w = -INF;
        dg_v = BWD ? 0.0f : INF; dg_w = -INF;
        if (frow == 0) { dg_v = BWD ? seed : 0.0f; dg_w = r_corner; }
      }
      const bool active = row_ok && j >= 0 && j < M;
      const int jj = active ? j : 0;
      const int tslot = (jj >> 5) % RING, tcol = jj & 31;
      const int sidx = BWD ? (kTile - 1 - tcol) : tcol;     // flipped tiles are stored unflipped along x
      const float d = sD[tslot * kBand * kTile + lane * kTile + sidx];
      float res_v, res_w = 0.0f;
      // 1-based indices differ by the same amount in both orientations
      const int ci = BWD ? (M - 1 - jj) : jj;
      const bool pruned = prune && fabsf((float)(row - ci)) > P.bandwidth;
      if (!BWD) {
        // R = D + softmin_gamma(diag, up, left)  (soft_dtw_cuda.py:65-72)
        const float mn = fminf(fminf(dg_v, nu_v), lf_v);
        float sm_;
        if (mn == INF) {
          sm_ = INF;
        } else {
          const float s = ex2f_((mn - dg_v) * ig2) + ex2f_((mn - nu_v) * ig2) + ex2f_((mn - lf_v) * ig2);
          sm_ = mn - g_ln2 * lg2f_(s);
        }
        res_v = pruned ? INF : d + sm_;
      } else {
        // E = E_dn*a + E_right*b + E_diag*c with a,b,c = exp((W[.] - R[i,j]) / gamma)  (:100-108)
        float r = sR[tslot * kBand * kTile + lane * kTile + sidx];
        if (isinf(r)) r = -INF;                             // :96-97
        res_w = r - d;
        const float a = ex2f_((nu_w - r) * ig2), bb = ex2f_((lf_w - r) * ig2), c = ex2f_((dg_w - r) * ig2);
        res_v = pruned ? 0.0f : (nu_v * a + lf_v * bb + dg_v * c);
        if (pruned || r == -INF) { res_v = 0.0f; }
        if (r == -INF) res_w = -INF;
      }
      up_v = nu_v; up_w = nu_w;
      if (active) {
        my_v = res_v; my_w = res_w;
        left_v = res_v; left_w = res_w;
        sO[((jj >> 5) & 1) * kBand * kTile + lane * kTile + sidx] = res_v;
        // publish the band's bottom row for the band below: {value, column+1} in one 8-byte store
        if (lane == kBand - 1) exp_mine[j] = make_int2(__float_as_int(res_v), j + 1);
      }
    }
    // flush the last (up to two) tiles
    __syncwarp();
    for (int k = (n_tiles >= 2 ? n_tiles - 2 : 0); k < n_tiles; ++k) {
      // tiles already flushed inside the loop: k <= (steps-1)/32 - 2
      if (k <= ((steps - 1) >> 5) - 2) continue;
      flush_tile<BWD>(sO + (k & 1) * kBand * kTile, Ob, N, M, r0, k * kTile, lane, vec);
    }
    __syncwarp();
    if (!BWD && band == P.nbands - 1) {
      // R[N-1, M-1] lives in lane (N-1) - r0 of the last band
      const float v = __shfl_sync(0xffffffffu, my_v, (N - 1) - r0);
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
  constexpr int RING = BWD ? 3 : 4;
  const int smem_per_warp = (RING * (BWD ? 2 : 1) + 2) * kBand * kTile * 4;
  const int warps = BWD ? 6 : 8;
  const int smem = smem_per_warp * warps;
  DAE_CUDA(cudaFuncSetAttribute(softdtw_wave_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int agents = B * P.nbands;
  int grid = (agents + warps - 1) / warps;
  if (grid > kNumSMs) grid = kNumSMs;                     // persistent: one CTA per SM, tickets do the rest
  const int vec = aligned16(D) && aligned16(R) && (!BWD || aligned16(E)) && (M % 4 == 0);
  softdtw_wave_kernel<BWD><<<grid, warps * 32, smem, st>>>(P, vec, smem_per_warp);
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
