"""Word error counting with the signature of ``lcasr.eval.wer.word_error_rate_detail``
(un-vendored; called at lcasr/run_dynamic_eval_full.py:112-115, lcasr/lib.py:1348-1349).

Returns ``(wer, words, ins_rate, del_rate, sub_rate)``.  The total error count and the reference
word count are alignment independent, so ``wer`` is exact; the S/D/I split depends on the
aligner's tie-break (PARITY UNPINNED against jiwer's): here the backtrace prefers
match/substitution, then deletion, then insertion.  Host-side integer work (ms per recording);
the int64 count vector is what the multi-GPU all-reduce sums.
"""
import numpy as np


def edit_counts(hyp_words, ref_words):
    """-> (S, D, I) of a minimum-cost alignment of hyp against ref."""
    n, m = len(ref_words), len(hyp_words)
    if n == 0:
        return 0, 0, m
    if m == 0:
        return 0, n, 0
    vocab = {}
    r = np.fromiter((vocab.setdefault(w, len(vocab)) for w in ref_words), dtype=np.int64, count=n)
    h = np.fromiter((vocab.setdefault(w, len(vocab)) for w in hyp_words), dtype=np.int64, count=m)
    # cost matrix rows = ref prefix, cols = hyp prefix; keep it whole for the backtrace
    d = np.zeros((n + 1, m + 1), dtype=np.int32)
    d[0, :] = np.arange(m + 1)
    d[:, 0] = np.arange(n + 1)
    ar = np.arange(m + 1, dtype=np.int32)
    for i in range(1, n + 1):
        sub = d[i - 1, :-1] + (h != r[i - 1])
        dele = d[i - 1, 1:] + 1
        best = np.minimum(sub, dele)
        # insertions chain along the row: d[i,j] = min(best[j-1], d[i,j-1] + 1) -> prefix-min trick
        row = np.empty(m + 1, dtype=np.int32)
        row[0] = i
        row[1:] = best
        row = np.minimum.accumulate(row - ar) + ar
        d[i] = row
    S = D = I = 0
    i, j = n, m
    while i > 0 or j > 0:
        if i > 0 and j > 0 and d[i, j] == d[i - 1, j - 1] + (r[i - 1] != h[j - 1]):
            S += int(r[i - 1] != h[j - 1])
            i, j = i - 1, j - 1
        elif i > 0 and d[i, j] == d[i - 1, j] + 1:
            D += 1
            i -= 1
        else:
            I += 1
            j -= 1
    return S, D, I


def word_error_counts(hypotheses, references, use_cer=False):
    """-> int64[5] = (S, D, I, ref_words, n_pairs), summed over all pairs."""
    if len(hypotheses) != len(references):
        raise ValueError(f"{len(hypotheses)} hypotheses vs {len(references)} references")
    tot = np.zeros(5, dtype=np.int64)
    for hyp, ref in zip(hypotheses, references):
        hw, rw = (list(hyp), list(ref)) if use_cer else (hyp.split(), ref.split())
        s, d, i = edit_counts(hw, rw)
        tot += np.array([s, d, i, len(rw), 1], dtype=np.int64)
    return tot


def rates_from_counts(counts):
    s, d, i, words = (int(x) for x in counts[:4])
    if words == 0:
        inf = float("inf")
        return inf, 0, inf, inf, inf
    return (s + d + i) / words, words, i / words, d / words, s / words


def word_error_rate_detail(hypotheses, references, use_cer=False):
    return rates_from_counts(word_error_counts(hypotheses, references, use_cer))
