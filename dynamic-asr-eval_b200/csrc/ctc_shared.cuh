// Pieces shared by the two CTC lattice paths (ctc.cu: per-frame chain; ctc_blocked.cu: time-blocked).
#pragma once
#include <cstdlib>
#include "common.cuh"

namespace dae {

constexpr float kLog2e = 1.4426950408889634f;
constexpr double kLn2d = 0.69314718055994530942;
constexpr int kLatThreads = 1024;
constexpr int kMaxPairsPerThread = 4;             // Lmax + 1 <= 4096 label/blank state pairs
constexpr int kLatSmemBudget = 200 * 1024;        // dynamic smem for labels + lattice rows + emission ring

struct CtcScratch {           // carved out of the caller's scratch buffer
  float* alpha;               // [N][T][Sp]  centred, log2 units
  float* beta_rev;            // [N][T][Sp]  beta stored at reversed state index S-1-s
  int32_t* next_same;         // [N][Lp]  next label position with the same class, -1 = none
  int32_t* leader;            // [N][Lp]  1 if first occurrence of its class
  double* off_a;              // [N][T]   alpha_t(s) = alpha[t][s] + off_a[t]   (log2 units)
  double* off_b;              // [N][T]   beta_t(s)  = beta_rev[t][S-1-s] + off_b[t]
  double* ll2;                // [4][N]   log2-likelihood from the alpha CTA, the beta CTA, then debug totals
  int* sync;                  // [64]  time-blocked path: [0] scan CTAs that have finished (zeroed by the band kernel)
  int Sp, Lp;
  // ---- time-blocked path only (ctc_blocked.cu); null / 0 when the shape is not eligible
  float* xfer;                // [N][nblk][Sq][2K+1]  K-frame transfer bands, log2 units: xfer[b][s][d] = paths s -> s+d
  float* bound;               // [2][N][nblk+1][Sq]   lattice vectors at block boundaries, centred per 128-state region
  double* boff;               // [2][N][nblk+1][G]    offset of each 128-state region of a boundary vector
  int2* halo;                 // [2][N][nblk+1][G][kHaloWords]  tagged {bits, tag} words handed to the next region
  float* emis;                // [N][nblk*K][Sq]  emission (log2 units) of every lattice state at every frame
  int nblk, G, Sq;
};

constexpr int kBlkK = 8;                          // frames per time block
constexpr int kBlkW = 2 * kBlkK + 1;              // a state moves at most two places per frame
constexpr int kRegion = 64;                       // states per boundary-scan CTA
constexpr int kHaloWords = 2 * kBlkK + 2;         // 2K values + the region offset as two 32-bit halves
constexpr int kBlkMaxN = 8;                       // batches larger than this fill the GPU with the per-frame chain
constexpr size_t kBlkMaxXferBytes = (size_t)256 << 20;

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Thread geometry of the lattice kernel: P state pairs per consumer thread, NTc consumer threads.
// Every consumer thread owns real or dummy pairs, so lattice rows are 2*NTc*P floats wide.
// Path overrides (dae_ctc_configure): process-wide, read with relaxed atomics on every call; the environment
// variables DAE_CTC_BLOCKED / DAE_CTC_CLUSTER / DAE_CTC_PAIRS / DAE_CTC_OVERLAP only seed them once, when the library
// is first used.  overlap (dae_ctc_loss_grad): bit 0 = the dense gradient streams under the scan; bit 1 = the
// label-class gradient kernel is resident and has loaded its inputs before the scan ends; -1 = all on.
struct CtcConfig { std::atomic<int> blocked{-1}, cluster{0}, pairs{0}, overlap{-1}; };
CtcConfig& ctc_config();

static inline void lat_geometry(int Lmax, int& P, int& NTc) {
  constexpr int kMaxConsumers = kLatThreads - 64;        // two helper warps: centring + TMA
  const int pairs = Lmax + 1;
  P = (pairs + kMaxConsumers - 1) / kMaxConsumers;
  P = P <= 1 ? 1 : (P <= 2 ? 2 : 4);
  const int forced = ctc_config().pairs.load(std::memory_order_relaxed);
  if ((forced == 1 || forced == 2 || forced == 4) && (pairs + forced - 1) / forced <= kMaxConsumers) P = forced;
  NTc = (((pairs + P - 1) / P + 31) / 32) * 32;
}

// The time-blocked path pays off when the per-frame chain would leave the GPU idle (few samples, many frames).
static inline bool blocked_eligible(int T, int N, int Lmax) {
  const int forced = ctc_config().blocked.load(std::memory_order_relaxed);   // 0 = never, 1 = whenever it fits
  if (forced == 0) return false;
  const int nblk = (T + kBlkK - 1) / kBlkK;
  const size_t Sq = align_up((size_t)2 * Lmax + 1, kRegion);
  const size_t xfer = (size_t)N * nblk * kBlkW * Sq * sizeof(float);
  if (N < 1 || xfer > kBlkMaxXferBytes) return false;
  if (forced == 1) return T >= 1;
  return N <= kBlkMaxN && T >= 8 * kBlkK;
}

static inline size_t ctc_carve(CtcScratch& s, void* base, int T, int N, int Lmax) {
  int P_, NTc_;
  lat_geometry(Lmax, P_, NTc_);
  s.Sp = 2 * NTc_ * P_;
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
  s.ll2 = (double*)(p + off); off += align_up((size_t)4 * N * sizeof(double), 256);
  s.sync = (int*)(p + off); off += 256;
  s.xfer = nullptr; s.bound = nullptr; s.boff = nullptr; s.halo = nullptr; s.emis = nullptr;
  s.nblk = 0; s.G = 0; s.Sq = 0;
  if (blocked_eligible(T, N, Lmax)) {
    s.nblk = (T + kBlkK - 1) / kBlkK;
    s.Sq = (int)align_up((size_t)s.Sp, kRegion);
    s.G = s.Sq / kRegion;
    s.xfer = (float*)(p + off); off += align_up((size_t)N * s.nblk * kBlkW * s.Sq * sizeof(float), 256);
    s.bound = (float*)(p + off); off += align_up((size_t)2 * N * (s.nblk + 1) * s.Sq * sizeof(float), 256);
    s.boff = (double*)(p + off); off += align_up((size_t)2 * N * (s.nblk + 1) * s.G * sizeof(double), 256);
    s.halo = (int2*)(p + off); off += align_up((size_t)2 * N * (s.nblk + 1) * s.G * kHaloWords * sizeof(int2), 256);
    s.emis = (float*)(p + off); off += align_up((size_t)N * s.nblk * kBlkK * s.Sq * sizeof(float), 256);
  }
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
// Order-preserving float <-> int map so a warp max can use redux.sync / smem atomicMax.
__device__ __forceinline__ int f2ord(float f) {
  const int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// Branch-free variants for the lattice: dead states hold the finite sentinel kDead instead of -inf
// (kDead + anything finite == kDead in fp32, and kDead - kDead == 0, so no NaN can appear).
constexpr float kDead = -1.0e30f;
__device__ __forceinline__ float lse2_n(float a, float b) {
  const float m = fmaxf(a, b), lo = fminf(a, b);
  return m + fast_lg2(1.0f + fast_ex2(lo - m));
}
__device__ __forceinline__ float lse3_n(float a, float b, float c) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  const float m = fmaxf(hi, c), mid = fminf(hi, c);
  return m + fast_lg2(1.0f + fast_ex2(mid - m) + fast_ex2(lo - m));
}

// ---- mbarrier / bulk-copy (TMA) helpers -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// 1-D bulk copy global -> shared through the TMA engine; completion is counted on `bar`.
__device__ __forceinline__ void tma_row_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// Fallback for rows that are not 16-byte aligned: per-thread 4-byte cp.async, tracked by the same mbarrier.
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ctc_blocked.cu: fills the same scratch (alpha, beta_rev, offsets, label groups, ll2) and nll as ctc_lattice_kernel.
int ctc_blocked_lattice(const float* lp, int64_t sT, int64_t sN, int T, int N, int C, const int64_t* tgt,
                        int64_t tgt_stride, int Lmax, const int64_t* in_len, const int64_t* tgt_len, int blank,
                        float* nll, const CtcScratch& sc, cudaStream_t st, bool dense_follows = false);

int ctc_blocked_fill(const float* lp, int64_t sT, int64_t sN, int T, int N, const int64_t* tgt, int64_t tgt_stride,
                     int Lmax, const int64_t* in_len, const int64_t* tgt_len, int blank, const CtcScratch& sc,
                     cudaStream_t st);
// returns 0 when the gradient has been written, 1 when only alpha/beta rows were filled (run ctc_grad_kernel next)
int ctc_blocked_grad(const float* lp, int64_t sT, int64_t sN, int T, int N, int C, const int64_t* tgt,
                     int64_t tgt_stride, int Lmax, const int64_t* in_len, const int64_t* tgt_len, int blank,
                     const float* gout, int64_t gout_stride, float* grad, const CtcScratch& sc, int vec,
                     cudaStream_t st, bool sparse_only = false, int scan_ctas_to_wait_for = 0);
// dae_ctc_loss_grad: the class-dense part of the gradient as a dependent launch of the scan (ctc_blocked.cu, 5.)
bool ctc_split_fits(const CtcScratch& sc, int N, int vec);
int ctc_blocked_dense(const float* lp, int64_t sT, int64_t sN, int T, int N, int C, const int64_t* in_len,
                      const float* gout, int64_t gout_stride, float* grad, cudaStream_t st, bool wait_for_scan);
int ctc_scan_ctas(const CtcScratch& sc, int N);      // CTAs of the scan that count themselves in sc.sync[0]

}  // namespace dae
