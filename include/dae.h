/*
 * dae.h — C ABI of libdae.so, the B200 (sm_100a) kernels behind the dynamic-evaluation
 * hot path of robflynnyh/dynamic-asr-eval.
 *
 * The reference has no FFI layer of its own (it is pure Python); each entry
 * point below replaces one Python-level call on the hot path and cites it as
 * "replaces: <file>:<line>" (paths relative to the reference checkout).  The host side
 * (package `dae`, directory dynamic-asr-eval_b200/) mirrors the reference's Python
 * signatures and binds these symbols with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host;
 *   - outputs and scratch are pre-allocated by the caller; no entry point allocates, frees
 *     or synchronises the device; all work is enqueued on `stream` (a cudaStream_t passed
 *     as void*, NULL = legacy default stream);
 *   - return value: 0 = enqueued, < 0 = argument error (DAE_E_*), > 0 = a cudaError_t;
 *   - strides are in ELEMENTS, not bytes; the innermost (class / column) stride is 1;
 *   - re-entrant from several host threads on distinct streams.  Process-wide state is limited to the launch
 *     counter, a mutex-guarded cache of per-kernel shared-memory attributes, and the CTC path override below.
 */
#ifndef DAE_H_
#define DAE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAE_ABI_VERSION 2

#define DAE_E_BADARG   (-1)  /* NULL pointer, negative size, ...                      */
#define DAE_E_TOOBIG   (-2)  /* a dimension exceeds what the kernel supports          */
#define DAE_E_SCRATCH  (-3)  /* scratch buffer smaller than *_scratch_bytes() says    */
#define DAE_E_ALIGN    (-4)  /* pointer / stride alignment requirement not met        */

int         dae_abi_version(void);
const char* dae_error_string(int code);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t     dae_launch_count(void);

/* ------------------------------------------------------------------------------------
 * (3b) greedy CTC decode: per-frame argmax -> collapse repeats -> drop blank.
 * replaces: lcasr.decoding.greedy.GreedyCTCDecoder.__call__ as used at
 *           lcasr/lib.py:559,565 and lcasr/run_dynamic_eval_full.py:100
 *           (there the [T,C] posteriors are first copied to the host).
 * lp      [B,T,C] fp32, strides sB,sT (elements), class stride 1
 * lengths [B] int32 valid frames per item, or NULL (= T for all)
 * path    [B,T] int32 out: argmax class of every frame (first index on ties, NaN = max,
 *                 as torch.argmax)
 * ids     [B,T] int32 out: collapsed label ids, ids[b, 0 .. n_ids[b])
 * n_ids   [B]   int32 out
 * scratch dae_greedy_scratch_bytes() bytes that are ZERO on entry (the call leaves them zero again, so one
 *         zero-initialised buffer per stream serves every call): argmax and collapse then run as ONE launch, the
 *         last CTA to finish doing the collapse.  NULL = two launches, no scratch needed.
 * ------------------------------------------------------------------------------------ */
size_t dae_greedy_scratch_bytes(void);
int dae_greedy_collapse(const float* lp, int64_t sB, int64_t sT, int B, int T, int C,
                        const int32_t* lengths, int blank,
                        int32_t* path, int32_t* ids, int32_t* n_ids, void* scratch, void* stream);

/* Collapse an already computed per-frame argmax path (e.g. dae_stitch's fused `path` output):
 * same collapse as above without the argmax pass.  path [B,T] int32, ids [B,T], n_ids [B]. */
int dae_collapse_path(const int32_t* path, int B, int T, const int32_t* lengths, int blank,
                      int32_t* ids, int32_t* n_ids, void* stream);

/* ------------------------------------------------------------------------------------
 * (3a) SpecAugment masking fused with the [augmented..., clean...] batch build.
 * replaces: lcasr.utils.augmentation.SpecAugment.__call__ + Tensor.repeat at
 *           lcasr/lib.py:538-541 (CPU tensors there).
 * x        [F,T] fp32 window of the spectrogram, row stride sF (a view into [1,F,spec_n])
 * out      [n_aug+n_clean, F, T] fp32 contiguous.  Copies 0..n_aug-1 are masked, the rest
 *          are verbatim copies of x.
 * fmask_host [n_aug][nf][2], tmask_host [n_aug][nt][2]: HOST int32 (start,end) half-open
 *          bands over the frequency / time axis, drawn by the caller's RNG (nf,nt <= 32,
 *          n_aug <= 4).
 * mask value = 0 if zero_masking else mean(x) (fp64 accumulation, fixed reduction order,
 *          rounded once to fp32).  partials: scratch of dae_specaug_scratch_bytes() bytes (partial sums +
 *          the counter of the grid barrier; reset by the call).  One cooperative launch: x is read once.
 * mean_out  [1] fp32 out (the fill value actually used), may be NULL.
 * ------------------------------------------------------------------------------------ */
size_t dae_specaug_scratch_bytes(void);
int dae_specaug_repeat(const float* x, int64_t sF, int F, int T,
                       const int32_t* fmask_host, int nf, const int32_t* tmask_host, int nt,
                       int zero_masking, int n_aug, int n_clean,
                       float* out, void* partials, float* mean_out, void* stream);

/* Per-recording variant used by the adapt loop (every window of a recording is augmented, lcasr/lib.py:537-541,
 * and a window's fill value is its mean — a function of the spectrogram alone):
 * dae_window_sums: ONE launch per recording leaves dae_window_slices() fp64 partial sums per window,
 *            sums[w][k] = sum of the k-th slice of x[:, win_start[w] : win_start[w]+win_len[w]] (fixed order).
 *            x [F, spec_n] fp32 row stride sF; win_start / win_len [n_win] int64 DEVICE arrays.
 * dae_specaug_repeat_premean: dae_specaug_repeat for a window whose partial sums are already known (win_sums =
 *            &sums[w][0], device): one ordinary launch, x read once, no grid barrier, no scratch. */
int dae_window_slices(void);
int dae_window_sums(const float* x, int64_t sF, int F, const int64_t* win_start, const int64_t* win_len, int n_win,
                    double* sums, void* stream);
int dae_specaug_repeat_premean(const float* x, int64_t sF, int F, int T,
                               const int32_t* fmask_host, int nf, const int32_t* tmask_host, int nt,
                               int zero_masking, int n_aug, int n_clean,
                               float* out, const double* win_sums, float* mean_out, void* stream);

/* ------------------------------------------------------------------------------------
 * (f-3) cutout: rectangles of the augmented window overwritten in place.
 * replaces: cutout() at lcasr/lib.py:384-417 (called at :544 on the augmented copy).
 * x          [F,T] fp32 in place, row stride sF
 * rects_host [n_rect][4] HOST int32 (start_x, end_x, start_y, end_y), half-open, x = time, y = frequency,
 *            in the order the reference draws and fills them (the last rectangle wins where they overlap)
 * mode       0 = zero, 1 = each rectangle's own mean (taken before any fill), 2 = mean of the whole window
 * scratch    dae_cutout_scratch_bytes(n_rect) bytes, 256-aligned (device copy of the table + means)
 * ------------------------------------------------------------------------------------ */
size_t dae_cutout_scratch_bytes(int n_rect);
int dae_cutout(float* x, int64_t sF, int F, int T, const int32_t* rects_host, int n_rect, int mode,
               void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * (f-3) frame shuffle and additive noise on the augmented window; randomness is host-drawn by the caller.
 * replaces: frame_shuffle() at lcasr/lib.py:81-84 (called at :542) and add_random_noise() at :379-382 (:543).
 * dae_frame_shuffle: out[f,t] = x[perm_f[f], perm_t[t]]; perm_t [T] / perm_f [F] int32 DEVICE arrays drawn with
 *            torch.randperm in the reference's order (time first, then frequency), NULL = identity; out [F,T]
 *            contiguous, must not alias x.
 * dae_add_noise:     x += (z * std(x)) * noise_factor in place, every product rounded to fp32 on its own (the
 *            reference's tensor ops); z [F,T] contiguous DEVICE = the standard-normal field behind the reference's
 *            torch.normal(0, std, size) draw (bitwise z*std, verified); std = unbiased standard deviation of x
 *            (fp64 moments, fixed order, rounded once), also written to std_out (may be NULL).
 *            scratch: dae_noise_scratch_bytes() bytes, 256-aligned.
 * ------------------------------------------------------------------------------------ */
int dae_frame_shuffle(const float* x, int64_t sF, int F, int T, const int32_t* perm_t, const int32_t* perm_f,
                      float* out, void* stream);
size_t dae_noise_scratch_bytes(void);
int dae_add_noise(float* x, int64_t sF, int F, int T, const float* z, float noise_factor,
                  void* scratch, size_t scratch_bytes, float* std_out, void* stream);

/* ------------------------------------------------------------------------------------
 * (1) CTC loss + gradient (torch.nn.CTCLoss semantics, zero_infinity=False).
 * replaces: torch.nn.CTCLoss(blank, reduction='sum')(...) and its backward at
 *           lcasr/lib.py:492,575-579 (AWMC: :250,324-331; finetune: earnings_finetune/train.py:259).
 *
 * dae_ctc_lattice : log-space alpha (forward in t) and beta (backward in t) recursions over
 *   the blank-interleaved label lattice, one CTA per (sample, direction), both directions
 *   concurrently.  Writes nll[n] and keeps alpha/beta in `scratch` for dae_ctc_grad.
 * dae_ctc_grad    : dense streaming pass
 *   grad[t,n,c] = gout[n] * ( exp(lp[t,n,c]) - exp(ab[t,n,c] + nll[n] - lp[t,n,c]) ), 0 for t >= in_len[n]
 *   (gradient w.r.t. the pre-log-softmax logits, torch's convention).
 *
 * lp      [T,N,C] fp32, strides sT,sN (elements), class stride 1 (any batch/time layout)
 * tgt     [N,Lmax] int64 padded targets, row stride tgt_stride; in_len/tgt_len [N] int64
 * nll     [N] fp32 out (per-sample negative log-likelihood; +inf when infeasible)
 * gout    per-sample upstream gradient, gout[n*gout_stride] (gout_stride 0 = one scalar)
 * grad    [T,N,C] fp32 contiguous out
 * ------------------------------------------------------------------------------------ */
size_t dae_ctc_scratch_bytes(int T, int N, int Lmax);
int dae_ctc_lattice(const float* lp, int64_t sT, int64_t sN, int T, int N, int C,
                    const int64_t* tgt, int64_t tgt_stride, int Lmax,
                    const int64_t* in_len, const int64_t* tgt_len, int blank,
                    float* nll, void* scratch, size_t scratch_bytes, void* stream);
int dae_ctc_grad(const float* lp, int64_t sT, int64_t sN, int T, int N, int C,
                 const int64_t* tgt, int64_t tgt_stride, int Lmax,
                 const int64_t* in_len, const int64_t* tgt_len, int blank,
                 const float* nll, const float* gout, int64_t gout_stride,
                 float* grad, const void* scratch, size_t scratch_bytes, void* stream);

/* The reference always follows the loss with `/ (T*N)` and `.backward()` (lcasr/lib.py:573-579), so the upstream
 * gradient is known when the loss is computed: a caller may run dae_ctc_grad right after dae_ctc_lattice with the
 * expected scale `hint` (both launches back to back, no host round trip in between) and, when the real upstream
 * gradient arrives, call dae_ctc_rescale: grad[t,n,:] *= gout[n]/hint where that ratio is not exactly 1 — a
 * kernel whose CTAs return after one load in the expected case. */
int dae_ctc_rescale(float* grad, int T, int N, int C, const float* gout, int64_t gout_stride, float hint,
                    void* stream);
/* dae_ctc_lattice followed by dae_ctc_grad as ONE call (same arguments, same results, scratch as for the pair):
 * knowing that the gradient is wanted while the lattice is still running lets the library stream the class-dense
 * part g*exp(lp) on the SMs the sequential scan leaves idle (few samples, many frames: the adapt step), so that
 * only the few hundred label classes of every frame are left to do when the scan ends.  gout/grad as in
 * dae_ctc_grad; nothing of `gout`, `lp` or `grad` may be produced by work queued on another stream. */
int dae_ctc_loss_grad(const float* lp, int64_t sT, int64_t sN, int T, int N, int C,
                      const int64_t* tgt, int64_t tgt_stride, int Lmax,
                      const int64_t* in_len, const int64_t* tgt_len, int blank,
                      float* nll, const float* gout, int64_t gout_stride,
                      float* grad, void* scratch, size_t scratch_bytes, void* stream);

/* Debug/test switch for which CTC lattice implementation dae_ctc_lattice and dae_ctc_scratch_bytes pick
 * (process-wide; seeded once from the environment variables DAE_CTC_BLOCKED / DAE_CTC_CLUSTER / DAE_CTC_PAIRS /
 * DAE_CTC_OVERLAP):
 *   blocked  -1 = by shape (default), 0 = always the per-frame chain, 1 = the time-blocked scan whenever it fits
 *   cluster  thread-block cluster size of the scan's region hand-over (1, 2, 4, 8; 0 = default 8; 1 = global memory only)
 *   pairs    state pairs per consumer thread of the per-frame chain (1, 2, 4; 0 = by label length)
 *   overlap  bit mask of what dae_ctc_loss_grad runs under the scan (-1 = everything, the default; 0 = nothing: the
 *            call behaves as dae_ctc_lattice + dae_ctc_grad); bit 0 = the dense part of the gradient; bit 1 (with
 *            bit 0) = the label-class gradient kernel is resident and has loaded its inputs before the scan ends */
void dae_ctc_configure(int blocked, int cluster, int pairs, int overlap);

/* ------------------------------------------------------------------------------------
 * (f-1) overlap-average stitch of window posteriors, with a fused greedy argmax.
 * replaces: the exp / slice-accumulate / count / divide / log passes over CPU buffers at
 *           lcasr/lib.py:594-629 (same arithmetic at :357-371 for AWMC).
 * win_lp   [n_win] windows of log-probs, window w = rows win_off[w] .. +win_len[w] of the
 *          device buffer `lp` ([rows_total, C] fp32, row stride C)
 * win_pos  [n_win] int64 first output row of window w (host arithmetic of lib.py:615-621,
 *          already sorted by window start); win_len [n_win] int64; win_off [n_win] int64
 * out      [n_out, C] fp32: log( sum_w exp(lp_w) / count ), windows added in index order
 * path     [n_out] int32 argmax of each stitched row (may be NULL)
 * Rows covered by no window are not part of the output (lib.py:624-627): the caller passes
 * n_out = number of covered rows and row_map [n_out] int64 = their buffer positions.
 * ------------------------------------------------------------------------------------ */
int dae_stitch(const float* lp, int C, const int64_t* win_off, const int64_t* win_pos,
               const int64_t* win_len, int n_win, const int64_t* row_map, int64_t n_out,
               float* out, int32_t* path, void* stream);

/* ------------------------------------------------------------------------------------
 * (2) soft-DTW forward / backward on a pairwise cost matrix.
 * replaces: _SoftDTWCUDA.forward/backward and the numba kernels compute_softdtw_cuda /
 *           compute_softdtw_backward_cuda at lcasr_nemo/soft_dtw_cuda.py:33-111,114-174, and the CPU
 *           kernels :184-239 the reference falls back to above 1024 frames (:312-314).
 * D    [B,N,M] fp32 contiguous cost matrices
 * fwd: R[b,i,j] = D[b,i,j] + softmin_gamma(R[i-1,j-1], R[i-1,j], R[i,j-1]) with the reference's borders
 *      (R[-1,-1] = 0, other border cells +inf); +inf where |i-j| > bandwidth > 0.
 *   out  [B] = R[b,N-1,M-1] (the soft-DTW value, the reference's R[:, -2, -2])
 *   W    [B,N,M,2] fp32, 16-byte aligned: per cell the softmin weights of its `up` (i-1,j) and `left` (i,j-1)
 *        predecessors; the `diag` weight is 1 - up - left.  These are the reference's backward coefficients
 *        a, b, c of :100-103 (a at cell (i,j) = W[i+1,j].up, b = W[i,j+1].left, c = 1 - both of W[i+1,j+1]);
 *        this is what the forward pass saves for the backward pass instead of the reference's fp32 R (:144).
 *   R    [B,N,M] fp32 or NULL: the interior R[:,1:N+1,1:M+1] of the reference's padded array, written only
 *        when asked for (the value and the gradient do not need it).
 * bwd: E [B,N,M] = gout[b] * dR[N-1,M-1]/dD  (the reference's grad_output * E[:,1:N+1,1:M+1], :172-174);
 *      gout[b*gout_stride] is the upstream gradient of out[b].
 * scratch: dae_softdtw_scratch_bytes() bytes, 256-byte aligned; zeroed by the call itself.
 * ------------------------------------------------------------------------------------ */
size_t dae_softdtw_scratch_bytes(int B, int N, int M);
int dae_softdtw_fwd(const float* D, int B, int N, int M, float gamma, float bandwidth,
                    float* W, float* R, float* out, void* scratch, size_t scratch_bytes, void* stream);
int dae_softdtw_bwd(const float* W, const float* gout, int64_t gout_stride, int B, int N, int M,
                    float* E, void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * (4) CTC prefix beam search with shallow n-gram LM fusion, one CTA per segment.
 * replaces: BeamSearch.run_search / step / merge / prune at lcasr/ctc_beam_search.py:152-319 and the
 *           per-frame LM calls of LanguageModel.__call__ (:77-87); the LM is a back-off n-gram trie in
 *           HBM (layout: dae/ngram.py) instead of the Transformer LM of lcasr/lib.py:37-72.
 * lp           [total_T, C] fp32 log-probs, row stride C; blank must be C-1 (the class assumes
 *              blank_id == vocab_size, lib.py:64)
 * seg_offsets  [n_seg+1] int32 (device): segment g = rows seg_offsets[g] .. seg_offsets[g+1]
 * lm_*         flat trie arrays on the device (tok/logp/bo/fail/depth [lm_nodes], cb [lm_nodes+1]);
 *              lm_row/lm_next: optional dense expansion from dae_ngram_expand (NULL = walk the trie)
 * scratch      dae_beam_scratch_bytes(n_seg, arena_cap) bytes, 256-aligned; holds the beams and the
 *              backpointer arena (arena_cap new-token entries per segment) and persists between
 *              calls, so a search can be advanced t_count frames at a time (BeamSearch.step()).
 *              t_begin == 0 (re)initialises the search; otherwise it resumes from the stored position.
 * finalize     non-zero: write the n_best best beams of every segment:
 *              out_score/out_len/out_flag [n_seg, n_best], out_tok/out_time [n_seg, n_best, out_cap]
 *              (token ids and their start frames), out_n [n_seg, 4] = (number of live beams, error code:
 *              0, DAE_E_TOOBIG = more than 4096 candidates in one frame, DAE_E_SCRATCH = arena full; candidates
 *              scored in this launch; loads issued against the LM arrays in this launch — measured LM-probe
 *              traffic = loads x 32-byte sectors).
 * ------------------------------------------------------------------------------------ */
size_t dae_beam_scratch_bytes(int n_seg, int arena_cap);
/* Dense expansion of the trie's context nodes (nodes 0..n_ctx-1, i.e. depth < order): row[s*vocab+w] =
 * log p(w | state s), next[s*vocab+w] = successor state.  Same arithmetic as the in-search walk (bit-identical);
 * trades n_ctx*vocab*8 bytes of HBM for one load per LM query.  vocab = number of LM tokens (= blank id). */
int dae_ngram_expand(const int32_t* lm_tok, const float* lm_logp, const float* lm_bo, const int32_t* lm_fail,
                     const int32_t* lm_cb, const int32_t* lm_depth, int lm_nodes, int lm_order, float lm_unk_lp,
                     int vocab, int n_ctx, float* row, int32_t* next, void* stream);
/* Rows for an explicit list of LM states: row[j*vocab + w] = log p(w | states[j]), next[j*vocab + w] = successor
 * state (next may be NULL).  replaces: LanguageModel.get_initial_state / __call__ at
 * lcasr/ctc_beam_search.py:70-87 (one Transformer-LM forward per batch of beams there).  states [n_states] int32
 * device; out-of-range states are scored from the root. */
int dae_ngram_rows(const int32_t* lm_tok, const float* lm_logp, const float* lm_bo, const int32_t* lm_fail,
                   const int32_t* lm_cb, const int32_t* lm_depth, int lm_nodes, int lm_order, float lm_unk_lp,
                   int vocab, const int32_t* states, int n_states, float* row, int32_t* next, void* stream);
int dae_beam_search(const float* lp, const int32_t* seg_offsets, int n_seg, int C, int blank,
                    int beam_width, float alpha, float beta, float top_am_threshold,
                    float prune_less_than_val, int has_prune, float blank_penalty, float repetition_penalty,
                    const int32_t* lm_tok, const float* lm_logp, const float* lm_bo, const int32_t* lm_fail,
                    const int32_t* lm_cb, const int32_t* lm_depth, int lm_nodes, int lm_order,
                    int lm_bos_state, float lm_unk_lp, const float* lm_row, const int32_t* lm_next,
                    void* scratch, size_t scratch_bytes, int arena_cap,
                    int t_begin, int t_count, int finalize, int n_best, int out_cap,
                    float* out_score, int32_t* out_len, int32_t* out_flag, int32_t* out_tok,
                    int32_t* out_time, int32_t* out_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DAE_H_ */
