"""fp64 numpy restatement of the CTC loss and gradient (test infrastructure only).

The arithmetic of this path lives in torch (un-vendored dependency of the reference, not
pinned there; torch 2.11.0 in this image): ``torch.nn.CTCLoss(blank, reduction='sum')`` at
lcasr/lib.py:492,575 and its autograd backward at :579.  This file restates the published
algorithm (Graves et al. 2006; formulas in SURVEY.md appendix A) in float64 and is PINNED by
tests/test_oracle_pins.py against torch's CPU implementation (fp64 and fp32) on ragged cases,
i.e. against the reference's own dependency executed in this container.
"""
import numpy as np

NEG = -np.inf


def _lse(*xs):
    m = np.maximum.reduce(xs)
    safe = np.where(np.isfinite(m), m, 0.0)
    s = sum(np.exp(x - safe) for x in xs)
    with np.errstate(divide="ignore"):
        return np.where(np.isfinite(m), safe + np.log(s), m)


def ctc_alpha_beta(x, labels, blank):
    """x [T,C] float64 log-probs, labels [L] ints -> (alpha [T,S], beta [T,S], nll) in natural log."""
    T, C = x.shape
    L = len(labels)
    S = 2 * L + 1
    ext = np.full(S, blank, dtype=np.int64)
    ext[1::2] = labels
    skip = np.zeros(S, dtype=bool)
    skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
    em = x[:, ext]                                       # [T,S]
    alpha = np.full((T, S), NEG)
    alpha[0, 0] = em[0, 0]
    if S > 1:
        alpha[0, 1] = em[0, 1]
    for t in range(1, T):
        p = alpha[t - 1]
        p1 = np.full(S, NEG)
        p1[1:] = p[:-1]
        p2 = np.full(S, NEG)
        p2[2:] = p[:-2]
        p2 = np.where(skip, p2, NEG)
        alpha[t] = em[t] + _lse(p, p1, p2)
    ll = _lse(alpha[T - 1, S - 1], alpha[T - 1, S - 2] if S > 1 else np.float64(NEG))
    beta = np.full((T, S), NEG)
    beta[T - 1, S - 1] = em[T - 1, S - 1]
    if S > 1:
        beta[T - 1, S - 2] = em[T - 1, S - 2]
    skipf = np.zeros(S, dtype=bool)                      # s -> s+2 allowed
    skipf[:-2] = (ext[:-2] != blank) & (ext[:-2] != ext[2:])
    for t in range(T - 2, -1, -1):
        p = beta[t + 1]
        p1 = np.full(S, NEG)
        p1[:-1] = p[1:]
        p2 = np.full(S, NEG)
        p2[:-2] = p[2:]
        p2 = np.where(skipf, p2, NEG)
        beta[t] = em[t] + _lse(p, p1, p2)
    return alpha, beta, -float(ll), ext


def ctc_loss_grad(lp, targets, input_lengths, target_lengths, blank, gout=None):
    """lp [T,N,C] -> (nll [N] float64, grad [T,N,C] float64) in torch's convention:
    grad = gout[n] * (exp(lp) - exp(ab + nll - lp)) for t < T_n, else 0."""
    lp = np.asarray(lp, dtype=np.float64)
    T, N, C = lp.shape
    nll = np.zeros(N)
    grad = np.zeros_like(lp)
    gout = np.ones(N) if gout is None else np.broadcast_to(np.asarray(gout, dtype=np.float64), (N,))
    for n in range(N):
        Tn, Ln = int(input_lengths[n]), int(target_lengths[n])
        x = lp[:Tn, n]
        alpha, beta, nl, ext = ctc_alpha_beta(x, np.asarray(targets[n][:Ln], dtype=np.int64), blank)
        nll[n] = nl
        ab = np.full((Tn, C), NEG)
        absum = alpha + beta
        for s in range(len(ext)):
            ab[:, ext[s]] = _lse(ab[:, ext[s]], absum[:, s])
        with np.errstate(over="ignore", invalid="ignore"):
            grad[:Tn, n] = gout[n] * (np.exp(x) - np.exp(ab + nl - x))
    return nll, grad
