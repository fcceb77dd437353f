"""CPU restatement (fp64, pure Python/numpy, small cases only) of the time-blocked CTC decomposition that
`dynamic-asr-eval_b200/csrc/ctc_blocked.cu` runs on the GPU.  TEST INFRASTRUCTURE ONLY.

The per-frame recursion of torch.nn.CTCLoss (call sites lcasr/lib.py:492,570-579; formulas SURVEY.md appendix A,
restated in oracle/ctc_oracle.py) is linear in the (log-sum, +) semiring, so it factors over blocks of K frames:

  X_b[d][s]            log2-sum over paths that sit in state s just before block b and in state s+d on its last
                       frame (0 <= d <= 2K), emissions of the block's frames included         (ctc_xfer_kernel)
  a[b+1][s']           = LSE_d X_b[d][s'-d] + a[b][s'-d],   a[0] = unit vector at state 0 (virtual frame -1)
  bh[b][s]             = LSE_d X_b[d][s] + bh[b+1][s+d],    bh[nblk] = 0 on the two final states
                       (alpha at block ends / beta without its own emission at block starts;  ctc_boundary_kernel)
  alpha, beta in a block: K ordinary steps from a[b] / from emission + bh[b+1]                (ctc_block_grad_kernel)

`tests/test_oracle_pins.py` checks this against oracle/ctc_oracle.py (itself pinned to torch's CPU CTCLoss).
"""
import numpy as np

NEG = -np.inf


def _lse2(*xs):
    m = max(xs)
    if m == NEG:
        return NEG
    return m + np.log2(sum(2.0 ** (x - m) for x in xs))


def blocked_alpha_beta(lp, labels, blank, K=8):
    """lp [T, C] natural-log posteriors, labels list[int].  Returns (nll, alpha, beta, E) with alpha/beta/E in log2
    units, alpha_t(s) and beta_t(s) both including the emission of frame t (torch's convention)."""
    lp = np.asarray(lp, dtype=np.float64)
    T, _ = lp.shape
    L = len(labels)
    S = 2 * L + 1
    x = lp * np.log2(np.e)
    cls = [blank if s % 2 == 0 else int(labels[s // 2]) for s in range(S)]
    skip = [s % 2 == 1 and s >= 3 and labels[s // 2] != labels[s // 2 - 1] for s in range(S)]
    E = np.array([[x[t, cls[s]] for s in range(S)] for t in range(T)]).reshape(T, S)
    nblk = (T + K - 1) // K
    W = 2 * K + 1
    X = np.full((nblk, W, S), NEG)
    for b in range(nblk):
        t0, t1 = b * K, min((b + 1) * K, T) - 1
        for s in range(S):
            v = np.full(W, NEG)
            v[0] = 0.0
            for k, t in enumerate(range(t0, t1 + 1), 1):
                for j in range(min(2 * k, W - 1), -1, -1):
                    sj = s + j
                    if sj >= S:
                        v[j] = NEG
                        continue
                    terms = [v[j]]
                    if j >= 1:
                        terms.append(v[j - 1])
                    if j >= 2 and skip[sj]:
                        terms.append(v[j - 2])
                    v[j] = E[t, sj] + _lse2(*terms)
            X[b, :, s] = v
    a = np.full((nblk + 1, S), NEG)
    a[0, 0] = 0.0
    for b in range(nblk):
        for sp in range(S):
            a[b + 1, sp] = _lse2(*[X[b, d, sp - d] + a[b, sp - d] for d in range(W) if sp - d >= 0])
    ll2 = _lse2(a[nblk, S - 1], a[nblk, S - 2] if S > 1 else NEG) if T > 0 else (0.0 if L == 0 else NEG)
    bh = np.full((nblk + 1, S), NEG)
    bh[nblk, S - 1] = 0.0
    if S > 1:
        bh[nblk, S - 2] = 0.0
    for b in range(nblk - 1, -1, -1):
        for s in range(S):
            bh[b, s] = _lse2(*[X[b, d, s] + bh[b + 1, s + d] for d in range(W) if s + d < S])
    alpha = np.full((T, S), NEG)
    beta = np.full((T, S), NEG)
    for b in range(nblk):
        t0, t1 = b * K, min((b + 1) * K, T) - 1
        prev = a[b].copy()
        for t in range(t0, t1 + 1):
            cur = np.full(S, NEG)
            for s in range(S):
                terms = [prev[s]]
                if s >= 1:
                    terms.append(prev[s - 1])
                if s >= 2 and skip[s]:
                    terms.append(prev[s - 2])
                cur[s] = E[t, s] + _lse2(*terms)
            alpha[t] = cur
            prev = cur
        nxt = E[t1] + bh[b + 1]
        beta[t1] = nxt
        for t in range(t1 - 1, t0 - 1, -1):
            cur = np.full(S, NEG)
            for s in range(S):
                terms = [nxt[s]]
                if s + 1 < S:
                    terms.append(nxt[s + 1])
                if s + 2 < S and skip[s + 2]:
                    terms.append(nxt[s + 2])
                cur[s] = E[t, s] + _lse2(*terms)
            beta[t] = cur
            nxt = cur
    return -ll2 * np.log(2.0), alpha, beta, E


def ctc_loss_grad_blocked(lp, labels, blank, K=8):
    """(nll, grad [T, C]) through the blocked decomposition; grad = exp(lp) - occupancy per class (gout = 1)."""
    lp = np.asarray(lp, dtype=np.float64)
    T, C = lp.shape
    nll, alpha, beta, E = blocked_alpha_beta(lp, labels, blank, K)
    grad = np.exp(lp)
    ll2 = -nll / np.log(2.0)
    S = 2 * len(labels) + 1
    for t in range(T):
        for s in range(S):
            c = blank if s % 2 == 0 else int(labels[s // 2])
            e = alpha[t, s] + beta[t, s] - E[t, s] - ll2
            if e > NEG:
                grad[t, c] -= 2.0 ** e
    return nll, grad
