// CTC prefix beam search with shallow n-gram LM fusion on the GPU, one CTA per segment.
//
// Algorithm and float semantics follow lcasr/ctc_beam_search.py:212-319 exactly (see
// oracle/beam_oracle.py, which is pinned bit-for-bit to the reference class):
//   * per frame, candidate classes are i in 1..V with lp[t,i] > max(lp[t]) + top_am_threshold (:225);
//     candidates are enumerated beam-major, class ascending (:230,242), so candidate index
//     c = beam * ntop + k reproduces the reference's creation order;
//   * blank / repeat keeps the LM sequence, score = (lp + score) + penalty (:250-260); a new token appends,
//     score = (lp + (lm*alpha + beta)) + score (:261-269), every op a separate fp32 rounding;
//   * candidates with the same collapsed AM sequence merge: the first-created one survives and the
//     others are log-added to it in creation order, fp32 difference -> fp64 exp/log -> fp32 add (:157-169);
//   * stable top-beam_width by score (:152-155), then the relative prune `not score < best - val` (:202-210).
// Sequences are identified by a 64-bit rolling hash + length (+ the trailing-blank flag), which makes the
// merge key O(1); the token history and start times live in a backpointer arena.
// The LM is the flat back-off trie of dae/ngram.py resident in HBM: a beam's LM state is a trie node and
// p(w | state) walks fail links with one binary search per level (arithmetic order fixed there).
#include "common.cuh"

namespace dae {

constexpr int kBeamThreads = 256;
constexpr int kMaxBeams = 128;
constexpr int kMaxTop = 256;
constexpr int kMaxCand = 4096;
constexpr int kBtSlots = 512;

struct DaeNgram {
  const int32_t* tok; const float* logp; const float* bo; const int32_t* fail; const int32_t* cb;
  const int32_t* depth;
  int n_nodes, order, bos_state;
  float unk_lp;
  // optional dense expansion of the context nodes (dae_ngram_expand): row[state*V + w] = log p(w | state),
  // next[state*V + w] = successor state; NULL = walk the trie
  const float* row; const int32_t* next; int V;
};

struct BeamRec {            // 40 bytes
  unsigned long long hash, phash;
  float score;
  int len, last, flag, lmst, hist;
};

struct BeamHeader { int n_beams, position, arena_used, error, buf; int pad[3]; };

struct BeamParams {
  const float* lp; const int32_t* seg_off; int n_seg, C, V;
  int beam_width; float alpha, beta, top_thr, prune_val; int has_prune; float blank_pen, rep_pen;
  DaeNgram lm;
  unsigned char* scratch; size_t seg_stride; int arena_cap;
  int t_begin, t_count, finalize, n_best, out_cap;
  float* out_score; int32_t* out_len; int32_t* out_flag; int32_t* out_tok; int32_t* out_time; int32_t* out_n;
};

__host__ __device__ inline size_t beam_seg_bytes(int arena_cap) {
  size_t b = sizeof(BeamHeader) + 2 * (size_t)kMaxBeams * sizeof(BeamRec) + (size_t)arena_cap * 3 * sizeof(int32_t);
  return (b + 255) / 256 * 256;
}

__device__ __forceinline__ unsigned long long hash_push(unsigned long long h, int tok) {
  unsigned long long x = h ^ ((unsigned long long)(tok + 1) * 0x9E3779B97F4A7C15ull);
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

// per-thread probe counter (global loads issued against the LM arrays), summed per segment into out_n[.,3]
struct LmCount { int probes; };
__device__ __forceinline__ int lm_find(const DaeNgram& lm, int node, int w, LmCount* cnt = nullptr) {
  int lo = __ldg(lm.cb + node), hi = __ldg(lm.cb + node + 1);
  if (cnt) cnt->probes += 2;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int t = __ldg(lm.tok + mid);
    if (cnt) cnt->probes += 1;
    if (t == w) return mid;
    if (t < w) lo = mid + 1; else hi = mid;
  }
  return -1;
}
// log p(w | state): longest context first, fp32 adds in that order (dae/ngram.py docstring).
__device__ __forceinline__ float lm_score_walk(const DaeNgram& lm, int state, int w, LmCount* cnt = nullptr) {
  float acc = 0.0f;
  int cur = state;
  for (;;) {
    const int c = lm_find(lm, cur, w, cnt);
    if (cnt) cnt->probes += 1;
    if (c >= 0) return __fadd_rn(acc, __ldg(lm.logp + c));
    acc = __fadd_rn(acc, __ldg(lm.bo + cur));
    if (cur == 0) return __fadd_rn(acc, lm.unk_lp);
    if (cnt) cnt->probes += 1;
    cur = __ldg(lm.fail + cur);
  }
}
__device__ __forceinline__ int lm_next_state_walk(const DaeNgram& lm, int state, int w, LmCount* cnt = nullptr) {
  int cur = state;
  for (;;) {
    const int c = lm_find(lm, cur, w, cnt);
    if (cnt) cnt->probes += 1;
    if (c >= 0 && __ldg(lm.depth + c) < lm.order) return c;
    if (cur == 0) return 0;
    cur = __ldg(lm.fail + cur);
  }
}

__device__ __forceinline__ float lm_score(const DaeNgram& lm, int state, int w, LmCount* cnt) {
  if (lm.row) { cnt->probes += 1; return __ldg(lm.row + (size_t)state * lm.V + w); }
  return lm_score_walk(lm, state, w, cnt);
}
__device__ __forceinline__ int lm_next_state(const DaeNgram& lm, int state, int w, LmCount* cnt) {
  if (lm.next) { cnt->probes += 1; return __ldg(lm.next + (size_t)state * lm.V + w); }
  return lm_next_state_walk(lm, state, w, cnt);
}

// Dense expansion of the trie: one thread per (context node, token), same device functions as the search,
// so the cached scores are bit-identical to a walk.
__global__ void __launch_bounds__(256)
ngram_expand_kernel(DaeNgram lm, int n_ctx, float* __restrict__ row, int32_t* __restrict__ next) {
  const size_t n = (size_t)n_ctx * lm.V;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int node = (int)(i / lm.V), w = (int)(i - (size_t)node * lm.V);
    row[i] = lm_score_walk(lm, node, w);
    next[i] = lm_next_state_walk(lm, node, w);
  }
}

// Rows for an explicit list of LM states (the LanguageModel duck type's per-beam next-token log-probs,
// lcasr/ctc_beam_search.py:70-87): row[j*V + w] = log p(w | states[j]); same walk, bit-identical to the search.
__global__ void __launch_bounds__(256)
ngram_rows_kernel(DaeNgram lm, const int32_t* __restrict__ states, int n_states, float* __restrict__ row,
                  int32_t* __restrict__ next) {
  const size_t n = (size_t)n_states * lm.V;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i / lm.V), w = (int)(i - (size_t)j * lm.V);
    int node = __ldg(states + j);
    if (node < 0 || node >= lm.n_nodes) node = 0;
    row[i] = lm_score_walk(lm, node, w);
    if (next) next[i] = lm_next_state_walk(lm, node, w);
  }
}

// ctc_beam_search.py:157-159 under torch/NumPy-2 semantics: fp32 difference, fp64 exp/log, fp32 add.
__device__ __forceinline__ float sum_log_scores(float s1, float s2) {
  if (s1 >= s2) return __fadd_rn(s1, (float)log(1.0 + exp((double)__fsub_rn(s2, s1))));
  return __fadd_rn(s2, (float)log(1.0 + exp((double)__fsub_rn(s1, s2))));
}

__device__ __forceinline__ unsigned bt_hash(unsigned long long h, int len, int flag) {
  unsigned long long x = h + (unsigned long long)len * 0x9E3779B97F4A7C15ull + (unsigned long long)flag * 0xD6E8FEB86659FD93ull;
  x ^= x >> 29;
  return (unsigned)x & (kBtSlots - 1);
}

struct BeamSmem {
  BeamRec beams[kMaxBeams];
  BeamRec next_beams[kMaxBeams];            // next frame's beams, written at their rank
  int bt[kBtSlots];
  int top_idx[kMaxTop];
  float top_am[kMaxTop];
  float c_score[kMaxCand];
  float c_final[kMaxCand];
  int sv[kMaxCand];
  int rk_cum[257];                          // bucket starts
  unsigned char c_lead[kMaxCand];
  float red_f[kBeamThreads / 32];
  float red_g[kBeamThreads / 32];
  int red_i[kBeamThreads / 32];
  int warp_cnt[kBeamThreads / 32];
  int ntop, n_sv, arena_used, error, n_new, n_unflagged;
  float thr, best, best_nb;
};

__device__ __forceinline__ int bt_lookup(const BeamSmem& S, unsigned long long h, int len, int flag) {
  unsigned s = bt_hash(h, len, flag);
  for (int probe = 0; probe < kBtSlots; ++probe) {
    const int b = S.bt[s];
    if (b < 0) return -1;
    if (S.beams[b].hash == h && S.beams[b].len == len && S.beams[b].flag == flag) return b;
    s = (s + 1) & (kBtSlots - 1);
  }
  return -1;
}

__device__ __forceinline__ bool cand_is_stay(const BeamRec& b, int i, int blank) {
  return i == blank || (b.flag == 0 && b.last == i);
}

__global__ void __launch_bounds__(kBeamThreads)
beam_search_kernel(BeamParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BeamSmem& S = *reinterpret_cast<BeamSmem*>(smem_raw);
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int blank = P.V;
  unsigned char* base = P.scratch + (size_t)g * P.seg_stride;
  BeamHeader* hdr = reinterpret_cast<BeamHeader*>(base);
  BeamRec* gbeams = reinterpret_cast<BeamRec*>(base + sizeof(BeamHeader));
  int32_t* arena = reinterpret_cast<int32_t*>(base + sizeof(BeamHeader) + 2 * (size_t)kMaxBeams * sizeof(BeamRec));
  const int seg_lo = P.seg_off[g], seg_T = P.seg_off[g + 1] - seg_lo;

  __shared__ int nb_s, pos_s;
  __shared__ unsigned long long stat_cand, stat_probe;
  LmCount lmc{0};
  unsigned my_cand = 0;
  if (tid == 0) { stat_cand = 0ull; stat_probe = 0ull; }
  if (tid == 0) {
    if (P.t_begin == 0) {                      // initiate (:126-138): one beam, LM sequence [bos]
      BeamRec r;
      r.hash = 0x1234567887654321ull; r.phash = 0; r.score = 0.0f; r.len = 0; r.last = -1; r.flag = 0;
      r.lmst = P.lm.bos_state; r.hist = -1;
      S.beams[0] = r;
      nb_s = 1; pos_s = 0; S.arena_used = 0; S.error = 0; S.n_unflagged = 1;
    } else {
      nb_s = hdr->n_beams; pos_s = hdr->position; S.arena_used = hdr->arena_used; S.error = hdr->error;
      S.n_unflagged = 1;                       // conservative until the next general frame recounts
    }
  }
  __syncthreads();
  if (P.t_begin != 0)
    for (int b = tid; b < nb_s; b += kBeamThreads) S.beams[b] = gbeams[b];
  __syncthreads();
  int nb = nb_s;
  int t = pos_s;
  const int t_end = min(seg_T, t + P.t_count);

  // emission rows are double-buffered in smem: frame t+1 is requested (cp.async) while frame t is searched
  float* rowbuf = reinterpret_cast<float*>(smem_raw + ((sizeof(BeamSmem) + 15) / 16) * 16);
  const int Cpad = (P.C + 3) & ~3;
  const bool row16 = (P.C % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.lp) & 15u) == 0);
  auto request_row = [&](int tt, int buf) {
    if (tt >= t_end) return;
    const float* src = P.lp + (size_t)(seg_lo + tt) * P.C;
    float* dst = rowbuf + buf * Cpad;
    if (row16) {
      for (int i = tid; i < (P.C >> 2); i += kBeamThreads)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + 4 * i)), "l"(src + 4 * i) : "memory");
    } else {
      for (int i = tid; i < P.C; i += kBeamThreads)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + i)), "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  request_row(t, t & 1);
  for (; t < t_end && !S.error; ++t) {
    request_row(t + 1, (t + 1) & 1);
    if (t + 1 < t_end) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                   // frame t's row is in smem for every thread
    const float* cur = rowbuf + (t & 1) * Cpad;
    // ---- P0: frame maximum -> strict threshold (:225); also the best non-blank candidate class
    float m = -CUDART_INF_F, mnb = -CUDART_INF_F;
    for (int i = tid; i < P.C; i += kBeamThreads) {
      const float v = cur[i];
      m = fmaxf(m, v);
      if (i >= 1 && i < blank) mnb = fmaxf(mnb, v);
    }
    m = warp_max(m);
    mnb = warp_max(mnb);
    if (lane == 0) { S.red_f[warp] = m; S.red_g[warp] = mnb; }
    if (tid == 0) { S.ntop = 0; S.n_sv = 0; S.n_new = 0; }
    for (int i = tid; i < kBtSlots; i += kBeamThreads) S.bt[i] = -1;
    __syncthreads();
    if (tid == 0) {
      float mm = S.red_f[0], mb = S.red_g[0];
      for (int w = 1; w < kBeamThreads / 32; ++w) { mm = fmaxf(mm, S.red_f[w]); mb = fmaxf(mb, S.red_g[w]); }
      S.thr = __fadd_rn(mm, P.top_thr);
      S.best_nb = mb;
    }
    __syncthreads();
    const float thr = S.thr;
    // ---- fast path: blank is the only candidate class and every beam already ends in blank.  Then every
    // candidate is "stay", no two candidates share a key, and adding one constant keeps the (stable) order:
    // update the scores in place and cut the tail with the relative prune.
    {
      const float am_b = cur[blank];
      if (am_b > thr && !(S.best_nb > thr) && S.n_unflagged == 0) {
        float sc = 0.0f;
        if (tid < nb) {
          sc = __fadd_rn(__fadd_rn(am_b, S.beams[tid].score), P.blank_pen);
          S.beams[tid].score = sc;
        }
        __syncthreads();
        const float lim = P.has_prune ? __fsub_rn(S.beams[0].score, P.prune_val) : -CUDART_INF_F;
        const int kept = __syncthreads_count(tid < nb && !(sc < lim));
        nb = kept;                                   // scores are non-increasing: the kept beams are a prefix
        continue;
      }
    }
    // ---- P1: ordered compaction of candidate classes 1..V
    for (int base_i = 1; base_i <= P.V; base_i += kBeamThreads) {
      const int i = base_i + tid;
      const float v = (i <= P.V) ? cur[i] : 0.0f;
      const bool keep = (i <= P.V) && v > thr;
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) S.warp_cnt[warp] = __popc(bal);
      __syncthreads();
      int off = S.ntop;
      for (int w = 0; w < warp; ++w) off += S.warp_cnt[w];
      if (keep) {
        const int k = off + __popc(bal & ((1u << lane) - 1));
        if (k < kMaxTop) { S.top_idx[k] = i; S.top_am[k] = v; }
      }
      __syncthreads();
      if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < kBeamThreads / 32; ++w) tot += S.warp_cnt[w];
        S.ntop += tot;
      }
      __syncthreads();
    }
    const int ntop = S.ntop;
    const int ncand = nb * ntop;
    if (ntop > kMaxTop || ncand > kMaxCand) {
      if (tid == 0) S.error = DAE_E_TOOBIG;
      __syncthreads();
      break;
    }
    // ---- P2: beam table keyed by (sequence hash, length, trailing-blank flag)
    for (int b = tid; b < nb; b += kBeamThreads) {
      unsigned s = bt_hash(S.beams[b].hash, S.beams[b].len, S.beams[b].flag);
      while (atomicCAS(&S.bt[s], -1, b) != -1) s = (s + 1) & (kBtSlots - 1);
    }
    __syncthreads();
    // ---- P3: candidate scores, creation order c = beam*ntop + k
    for (int c = tid; c < ncand; c += kBeamThreads) {
      const int b = c / ntop, k = c - b * ntop;
      const int i = S.top_idx[k];
      const float am = S.top_am[k];
      const BeamRec& br = S.beams[b];
      ++my_cand;
      float sc;
      if (cand_is_stay(br, i, blank)) {
        sc = __fadd_rn(__fadd_rn(am, br.score), i == blank ? P.blank_pen : P.rep_pen);
      } else {
        const float lmv = __fadd_rn(__fmul_rn(lm_score(P.lm, br.lmst, i, &lmc), P.alpha), P.beta);
        sc = __fadd_rn(__fadd_rn(am, lmv), br.score);
      }
      S.c_score[c] = sc;
    }
    __syncthreads();
    // ---- P4: merge groups (<= 3 members), first-created member leads and folds the others in order
    for (int c = tid; c < ncand; c += kBeamThreads) {
      const int b = c / ntop, k = c - b * ntop;
      const int i = S.top_idx[k];
      const BeamRec& br = S.beams[b];
      int m1 = -1, m2 = -1;                      // candidate indices of the other members
      if (i == blank) {                          // key (seq, blank-terminated)
        const int s = bt_lookup(S, br.hash, br.len, 1 - br.flag);
        if (s >= 0) m1 = s * ntop + k;
      } else if (br.flag == 0 && br.last == i) { // repeat: key (seq, no blank)
        if (br.len >= 1) {
          const int p1 = bt_lookup(S, br.phash, br.len - 1, 1);
          if (p1 >= 0) m1 = p1 * ntop + k;
          const int p0 = bt_lookup(S, br.phash, br.len - 1, 0);
          if (p0 >= 0 && S.beams[p0].last != i) m2 = p0 * ntop + k;
        }
      } else {                                   // new token: key (seq + i, no blank)
        const int s = bt_lookup(S, br.hash, br.len, 1 - br.flag);
        if (s >= 0 && !cand_is_stay(S.beams[s], i, blank)) m1 = s * ntop + k;
        const int ch = bt_lookup(S, hash_push(br.hash, i), br.len + 1, 0);
        if (ch >= 0) m2 = ch * ntop + k;
      }
      int lo = c;
      if (m1 >= 0 && m1 < lo) lo = m1;
      if (m2 >= 0 && m2 < lo) lo = m2;
      unsigned char lead = (lo == c);
      if (lead) {
        float acc = S.c_score[c];
        int a = m1, bb = m2;
        if (a < 0 || (bb >= 0 && bb < a)) { const int tmp = a; a = bb; bb = tmp; }   // a = smaller valid index
        if (a >= 0) acc = sum_log_scores(S.c_score[a], acc);
        if (bb >= 0) acc = sum_log_scores(S.c_score[bb], acc);
        S.c_final[c] = acc;
      }
      S.c_lead[c] = lead;
    }
    __syncthreads();
    // ---- P5: best leader, relative prune, survivors
    float bs = -CUDART_INF_F;
    for (int c = tid; c < ncand; c += kBeamThreads)
      if (S.c_lead[c]) bs = fmaxf(bs, S.c_final[c]);
    bs = warp_max(bs);
    if (lane == 0) S.red_f[warp] = bs;
    __syncthreads();
    if (tid == 0) {
      float mm = S.red_f[0];
      for (int w = 1; w < kBeamThreads / 32; ++w) mm = fmaxf(mm, S.red_f[w]);
      S.best = mm;
    }
    __syncthreads();
    const float lim = P.has_prune ? __fsub_rn(S.best, P.prune_val) : -CUDART_INF_F;
    for (int c = tid; c < ncand; c += kBeamThreads) {
      if (S.c_lead[c] && !(S.c_final[c] < lim)) S.sv[atomicAdd(&S.n_sv, 1)] = c;
    }
    __syncthreads();
    const int nsv = S.n_sv;
    const int nb_new = min(nsv, P.beam_width);
    if (tid == 0) S.n_unflagged = 0;
    __syncthreads();
    BeamRec* nxt = S.next_beams;
    // rank by (score desc, creation index asc); rank < beam_width survives at position rank.  Large frames first
    // sort the survivors into 256 score buckets (bucket index is monotone in the score, equal scores share a bucket),
    // so a survivor's rank is the population of the better buckets plus an exact count inside its own bucket, and
    // survivors whose better buckets already hold beam_width entries are dropped without any comparison.
    const bool bucketed = nsv > 128;                       // CTA-uniform
    int* sv2 = reinterpret_cast<int*>(S.c_score);          // survivors grouped by bucket (c_score is dead after P4)
    if (bucketed) {
      float lo = lim;
      if (!P.has_prune) {                                  // no relative prune: the range is best - min
        float mnv = CUDART_INF_F;
        for (int j = tid; j < nsv; j += kBeamThreads) mnv = fminf(mnv, S.c_final[S.sv[j]]);
        mnv = -warp_max(-mnv);
        if (lane == 0) S.red_g[warp] = mnv;
        __syncthreads();
        lo = S.red_g[0];
        for (int w = 1; w < kBeamThreads / 32; ++w) lo = fminf(lo, S.red_g[w]);
      }
      const float range = __fsub_rn(S.best, lo);
      const float scale = (range > 0.0f && range < CUDART_INF_F) ? 256.0f / range : 0.0f;
      for (int i = tid; i < kBtSlots; i += kBeamThreads) S.bt[i] = 0;        // [0,256) histogram, [256,512) cursors
      __syncthreads();
      for (int j = tid; j < nsv; j += kBeamThreads) {
        const int c = S.sv[j];
        int b = (int)(__fsub_rn(S.best, S.c_final[c]) * scale);
        b = b < 0 ? 0 : (b > 255 ? 255 : b);
        S.c_lead[c] = (unsigned char)b;                    // the leader flags are no longer needed
        atomicAdd(&S.bt[b], 1);
      }
      __syncthreads();
      {                                                    // exclusive scan of the 256 bucket counts
        const int v = S.bt[tid];
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int n_ = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += n_;
        }
        if (lane == 31) S.warp_cnt[warp] = incl;
        __syncthreads();
        int off = 0;
        for (int w = 0; w < warp; ++w) off += S.warp_cnt[w];
        S.rk_cum[tid] = off + incl - v;
        if (tid == kBeamThreads - 1) S.rk_cum[256] = off + incl;
      }
      __syncthreads();
      for (int j = tid; j < nsv; j += kBeamThreads) {
        const int c = S.sv[j];
        const int b = S.c_lead[c];
        sv2[S.rk_cum[b] + atomicAdd(&S.bt[256 + b], 1)] = c;
      }
      __syncthreads();
    }
    for (int j = tid; j < nsv; j += kBeamThreads) {
      const int c = S.sv[j];
      const float f = S.c_final[c];
      int rank = 0;
      if (bucketed) {
        const int b = S.c_lead[c];
        rank = S.rk_cum[b];
        if (rank < P.beam_width) {
          const int hi = S.rk_cum[b + 1];
          for (int l = S.rk_cum[b]; l < hi; ++l) {
            const int c2 = sv2[l];
            const float f2 = S.c_final[c2];
            rank += (f2 > f) || (f2 == f && c2 < c);
          }
        }
      } else {
        for (int l = 0; l < nsv; ++l) {
          const int c2 = S.sv[l];
          const float f2 = S.c_final[c2];
          rank += (f2 > f) || (f2 == f && c2 < c);
        }
      }
      if (rank < P.beam_width) {
        const int b = c / ntop, k = c - b * ntop;
        const int i = S.top_idx[k];
        const BeamRec& br = S.beams[b];
        BeamRec r = br;
        r.score = f;
        if (cand_is_stay(br, i, blank)) {
          if (i == blank) r.flag = 1;
        } else {
          r.phash = br.hash;
          r.hash = hash_push(br.hash, i);
          r.len = br.len + 1;
          r.last = i;
          r.flag = 0;
          r.lmst = lm_next_state(P.lm, br.lmst, i, &lmc);
          const int e = atomicAdd(&S.arena_used, 1);
          if (e < P.arena_cap) {
            arena[3 * e] = br.hist; arena[3 * e + 1] = i; arena[3 * e + 2] = t;
            r.hist = e;
          } else {
            S.error = DAE_E_SCRATCH;
            r.hist = br.hist;
          }
        }
        nxt[rank] = r;
        if (r.flag == 0) atomicAdd(&S.n_unflagged, 1);
      }
    }
    __syncthreads();
    for (int b = tid; b < nb_new; b += kBeamThreads) S.beams[b] = nxt[b];
    nb = nb_new;
    __syncthreads();
  }

  // persist state
  for (int b = tid; b < nb; b += kBeamThreads) gbeams[b] = S.beams[b];
  if (tid == 0) {
    hdr->n_beams = nb; hdr->position = t; hdr->arena_used = S.arena_used; hdr->error = S.error;
  }
  __syncthreads();
  atomicAdd(&stat_cand, (unsigned long long)my_cand);
  atomicAdd(&stat_probe, (unsigned long long)lmc.probes);
  __syncthreads();
  if (P.finalize) {
    if (tid == 0) {
      P.out_n[4 * g] = nb; P.out_n[4 * g + 1] = S.error;
      // statistics of this launch: candidates scored and loads issued against the LM arrays (saturating int32)
      P.out_n[4 * g + 2] = (int)(stat_cand > 0x7fffffffull ? 0x7fffffffull : stat_cand);
      P.out_n[4 * g + 3] = (int)(stat_probe > 0x7fffffffull ? 0x7fffffffull : stat_probe);
    }
    for (int r = tid; r < P.n_best; r += kBeamThreads) {
      const size_t o = (size_t)g * P.n_best + r;
      if (r < nb) {
        const BeamRec& br = S.beams[r];
        P.out_score[o] = br.score; P.out_len[o] = br.len; P.out_flag[o] = br.flag;
        int e = br.hist;
        for (int k = br.len - 1; k >= 0 && e >= 0; --k) {
          if (k < P.out_cap) {
            P.out_tok[o * P.out_cap + k] = arena[3 * e + 1];
            P.out_time[o * P.out_cap + k] = arena[3 * e + 2];
          }
          e = arena[3 * e];
        }
      } else {
        P.out_score[o] = -CUDART_INF_F; P.out_len[o] = -1; P.out_flag[o] = 0;
      }
    }
  }
}

}  // namespace dae

extern "C" size_t dae_beam_scratch_bytes(int n_seg, int arena_cap) {
  if (n_seg < 0 || arena_cap < 0) return 0;
  return (size_t)n_seg * dae::beam_seg_bytes(arena_cap);
}

extern "C" int dae_beam_search(const float* lp, const int32_t* seg_offsets, int n_seg, int C, int blank,
                               int beam_width, float alpha, float beta, float top_am_threshold,
                               float prune_less_than_val, int has_prune, float blank_penalty, float repetition_penalty,
                               const int32_t* lm_tok, const float* lm_logp, const float* lm_bo, const int32_t* lm_fail,
                               const int32_t* lm_cb, const int32_t* lm_depth, int lm_nodes, int lm_order,
                               int lm_bos_state, float lm_unk_lp, const float* lm_row, const int32_t* lm_next,
                               void* scratch, size_t scratch_bytes, int arena_cap,
                               int t_begin, int t_count, int finalize, int n_best, int out_cap,
                               float* out_score, int32_t* out_len, int32_t* out_flag, int32_t* out_tok,
                               int32_t* out_time, int32_t* out_n, void* stream) {
  using namespace dae;
  if (!lp || !seg_offsets || !scratch || n_seg < 0 || C < 2 || beam_width < 1 || t_begin < 0 || t_count < 0)
    return DAE_E_BADARG;
  if (blank != C - 1) return DAE_E_BADARG;            // the reference class assumes blank_id == vocab_size
  if (!lm_tok || !lm_logp || !lm_bo || !lm_fail || !lm_cb || !lm_depth || lm_nodes < 1 || lm_order < 1) return DAE_E_BADARG;
  if (beam_width > kMaxBeams) return DAE_E_TOOBIG;
  if (finalize && (!out_score || !out_len || !out_flag || !out_tok || !out_time || !out_n || n_best < 1 || out_cap < 1))
    return DAE_E_BADARG;
  if (scratch_bytes < dae_beam_scratch_bytes(n_seg, arena_cap)) return DAE_E_SCRATCH;
  if ((reinterpret_cast<uintptr_t>(scratch) & 255u) != 0) return DAE_E_ALIGN;
  if (n_seg == 0) return 0;
  BeamParams P;
  P.lp = lp; P.seg_off = seg_offsets; P.n_seg = n_seg; P.C = C; P.V = blank;
  P.beam_width = beam_width; P.alpha = alpha; P.beta = beta; P.top_thr = top_am_threshold;
  P.prune_val = prune_less_than_val; P.has_prune = has_prune; P.blank_pen = blank_penalty; P.rep_pen = repetition_penalty;
  P.lm = DaeNgram{lm_tok, lm_logp, lm_bo, lm_fail, lm_cb, lm_depth, lm_nodes, lm_order, lm_bos_state, lm_unk_lp,
                  lm_row, lm_next, blank};
  P.scratch = (unsigned char*)scratch; P.seg_stride = beam_seg_bytes(arena_cap); P.arena_cap = arena_cap;
  P.t_begin = t_begin; P.t_count = t_count; P.finalize = finalize; P.n_best = n_best; P.out_cap = out_cap;
  P.out_score = out_score; P.out_len = out_len; P.out_flag = out_flag; P.out_tok = out_tok; P.out_time = out_time;
  P.out_n = out_n;
  const int smem = (int)(((sizeof(BeamSmem) + 15) / 16) * 16 + 2 * (size_t)((C + 3) & ~3) * sizeof(float));
  if (smem > 220 * 1024) return DAE_E_TOOBIG;
  DAE_CUDA(ensure_dyn_smem(beam_search_kernel, smem));
  beam_search_kernel<<<n_seg, kBeamThreads, smem, (cudaStream_t)stream>>>(P);
  DAE_LAUNCH_OK();
  return 0;
}

extern "C" int dae_ngram_expand(const int32_t* lm_tok, const float* lm_logp, const float* lm_bo, const int32_t* lm_fail,
                                const int32_t* lm_cb, const int32_t* lm_depth, int lm_nodes, int lm_order, float lm_unk_lp,
                                int vocab, int n_ctx, float* row, int32_t* next, void* stream) {
  using namespace dae;
  if (!lm_tok || !lm_logp || !lm_bo || !lm_fail || !lm_cb || !lm_depth || !row || !next || lm_nodes < 1 || lm_order < 1 ||
      vocab < 1 || n_ctx < 0 || n_ctx > lm_nodes)
    return DAE_E_BADARG;
  if (n_ctx == 0) return 0;
  DaeNgram lm{lm_tok, lm_logp, lm_bo, lm_fail, lm_cb, lm_depth, lm_nodes, lm_order, 0, lm_unk_lp, nullptr, nullptr, vocab};
  const size_t n = (size_t)n_ctx * vocab;
  size_t want = (n + 255) / 256;
  const int grid = (int)(want < (size_t)kNumSMs * 16 ? want : (size_t)kNumSMs * 16);
  ngram_expand_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(lm, n_ctx, row, next);
  DAE_LAUNCH_OK();
  return 0;
}

extern "C" int dae_ngram_rows(const int32_t* lm_tok, const float* lm_logp, const float* lm_bo, const int32_t* lm_fail,
                              const int32_t* lm_cb, const int32_t* lm_depth, int lm_nodes, int lm_order, float lm_unk_lp,
                              int vocab, const int32_t* states, int n_states, float* row, int32_t* next, void* stream) {
  using namespace dae;
  if (!lm_tok || !lm_logp || !lm_bo || !lm_fail || !lm_cb || !lm_depth || !row || !states || lm_nodes < 1 ||
      lm_order < 1 || vocab < 1 || n_states < 0)
    return DAE_E_BADARG;
  if (n_states == 0) return 0;
  DaeNgram lm{lm_tok, lm_logp, lm_bo, lm_fail, lm_cb, lm_depth, lm_nodes, lm_order, 0, lm_unk_lp, nullptr, nullptr, vocab};
  const size_t n = (size_t)n_states * vocab;
  size_t want = (n + 255) / 256;
  const int grid = (int)(want < (size_t)kNumSMs * 16 ? want : (size_t)kNumSMs * 16);
  ngram_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(lm, states, n_states, row, next);
  DAE_LAUNCH_OK();
  return 0;
}
