"""numba-CUDA restatement of the reference's GPU soft-DTW path — BASELINE TIMING ONLY (bench.py's
`softdtw_vs_numba_cuda` table); test infrastructure, never imported by the product.

Follows the structure of lcasr_nemo/soft_dtw_cuda.py:33-111 and its host wrappers :114-174: one thread block per
sample, one thread per row (so max(N, M) <= 1024, the cap at :312-314), a block barrier after every anti-diagonal,
all operands in global memory, double-precision exp/log on fp32 storage, R and E padded to [B, N+2, M+2], the
backward pass preceded by the padded copy of D and the border patches of R (:158-166).  /root/reference does not
exist on the GPU box, hence a port ("kind": "port"); its outputs are checked against the fp64 oracle by
time_fwd_bwd() before any time is reported.
"""
import math

import torch
from numba import cuda


@cuda.jit
def _fwd(D, gamma, bandwidth, n_rows, n_cols, n_diag, R):
    s = cuda.blockIdx.x
    row = cuda.threadIdx.x
    ig = 1.0 / gamma
    for d in range(n_diag):
        col = d - row
        if row < n_rows and 0 <= col < n_cols:
            i, j = row + 1, col + 1
            if not (abs(i - j) > bandwidth > 0):
                x0 = -R[s, i - 1, j - 1] * ig
                x1 = -R[s, i - 1, j] * ig
                x2 = -R[s, i, j - 1] * ig
                top = max(max(x0, x1), x2)
                z = math.exp(x0 - top) + math.exp(x1 - top) + math.exp(x2 - top)
                R[s, i, j] = D[s, i - 1, j - 1] - gamma * (math.log(z) + top)
        cuda.syncthreads()


@cuda.jit
def _bwd(Dp, R, ig, bandwidth, n_rows, n_cols, n_diag, E):
    s = cuda.blockIdx.x
    row = cuda.threadIdx.x
    for step in range(n_diag):
        d = n_diag - 1 - step
        col = d - row
        if row < n_rows and 0 <= col < n_cols:
            i, j = row + 1, col + 1
            if math.isinf(R[s, i, j]):
                R[s, i, j] = -math.inf
            if not (abs(i - j) > bandwidth > 0):
                here = R[s, i, j]
                wa = math.exp((R[s, i + 1, j] - here - Dp[s, i + 1, j]) * ig)
                wb = math.exp((R[s, i, j + 1] - here - Dp[s, i, j + 1]) * ig)
                wc = math.exp((R[s, i + 1, j + 1] - here - Dp[s, i + 1, j + 1]) * ig)
                E[s, i, j] = E[s, i + 1, j] * wa + E[s, i, j + 1] * wb + E[s, i + 1, j + 1] * wc
        cuda.syncthreads()


def forward(D, gamma, bandwidth):
    B, N, M = D.shape
    threads = max(N, M)
    if threads > 1024:
        raise ValueError("the reference's CUDA path refuses sequences longer than 1024 (soft_dtw_cuda.py:312-314)")
    R = torch.full((B, N + 2, M + 2), math.inf, device=D.device, dtype=D.dtype)
    R[:, 0, 0] = 0
    _fwd[B, threads](cuda.as_cuda_array(D), float(gamma), float(bandwidth), N, M, 2 * threads - 1,
                     cuda.as_cuda_array(R))
    return R


def backward(D, R, gamma, bandwidth):
    B, N, M = D.shape
    threads = max(N, M)
    R = R.clone()
    Dp = torch.zeros((B, N + 2, M + 2), dtype=D.dtype, device=D.device)
    Dp[:, 1:N + 1, 1:M + 1] = D
    R[:, :, -1] = -math.inf
    R[:, -1, :] = -math.inf
    R[:, -1, -1] = R[:, -2, -2]
    E = torch.zeros((B, N + 2, M + 2), dtype=D.dtype, device=D.device)
    E[:, -1, -1] = 1
    _bwd[B, threads](cuda.as_cuda_array(Dp), cuda.as_cuda_array(R), 1.0 / float(gamma), float(bandwidth), N, M,
                     2 * threads - 1, cuda.as_cuda_array(E))
    return E[:, 1:N + 1, 1:M + 1]


def time_fwd_bwd(D, gamma, bandwidth, timer):
    """-> (forward seconds, backward seconds), host wrappers included as in _SoftDTWCUDA.forward/backward; the first
    sample's value and gradient are checked against the fp64 oracle first."""
    import numpy as np
    from . import softdtw_oracle as so
    R = forward(D, gamma, bandwidth)
    E = backward(D, R, gamma, bandwidth)
    torch.cuda.synchronize()
    Ro = so.forward(D[:1].cpu().numpy(), gamma, bandwidth)
    Eo = so.backward(D[:1].cpu().numpy(), Ro, gamma, bandwidth)
    assert abs(float(R[0, -2, -2]) - Ro[0, -2, -2]) <= 1e-4 * max(1.0, abs(Ro[0, -2, -2]))
    assert np.abs(E[0].cpu().numpy() - Eo[0]).max() <= 2e-3
    fw, _ = timer.time(lambda: forward(D, gamma, bandwidth), 5)
    bw, _ = timer.time(lambda: backward(D, R, gamma, bandwidth), 5)
    return fw, bw
