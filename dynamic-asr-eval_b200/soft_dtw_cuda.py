"""soft-DTW with the reference module's API (lcasr_nemo/soft_dtw_cuda.py), on the dae wavefront kernels.

``SoftDTW(use_cuda, gamma, normalize, bandwidth, dist_func)(X, Y) -> [B]`` (:273-352) and
``_SoftDTWCUDA.apply(D, gamma, bandwidth) -> [B]`` (:114-174) keep their signatures.  Unlike the
reference there is no 1024-frame cap and no silent CPU fallback (:312-314): every length runs on the
GPU, and CPU tensors are an error.  What the forward pass saves for the backward pass is not the fp32 R
matrix of the reference (:144) but the per-cell softmin weights ``W`` [B,N,M,2] — the coefficients a, b, c of
the reference's backward recurrence (:100-103) — so the backward pass is multiply-add only and the gradient is
good to ~1e-6 of its scale at 4096x4096 (the reference's fp32 R limits its own to ~1e-4 there).
"""
import torch

from . import _C, prof


def _scratch(B, N, M, dev):
    n = _C.lib().dae_softdtw_scratch_bytes(B, N, M)
    return torch.empty(n, dtype=torch.uint8, device=dev), n


def softdtw_forward(D, gamma, bandwidth=0.0, want_R=False):
    """D [B,N,M] fp32 CUDA -> (value [B], W [B,N,M,2] softmin weights, R [B,N,M] or None)."""
    _C.require_cuda(D, "D")
    D = D.detach()
    if D.dtype != torch.float32 or not D.is_contiguous():
        D = D.float().contiguous()
    B, N, M = D.shape
    dev = D.device
    W = torch.empty((B, N, M, 2), dtype=torch.float32, device=dev)
    R = torch.empty_like(D) if want_R else None
    out = torch.empty(B, dtype=torch.float32, device=dev)
    scratch, n = _scratch(B, N, M, dev)
    with torch.cuda.device(dev), prof.span("softdtw_fwd", 2 * B * N * M * 4):
        rc = _C.lib().dae_softdtw_fwd(D.data_ptr(), B, N, M, float(gamma), float(bandwidth), W.data_ptr(),
                                      _C.ptr(R), out.data_ptr(), scratch.data_ptr(), n, _C.stream_ptr(dev))
    _C.check(rc, "dae_softdtw_fwd")
    return out, W, R


def softdtw_backward(W, grad_out):
    """-> grad_out[b] * E [B,N,M] (soft_dtw_cuda.py:147-174) from the forward pass's weights."""
    _C.require_cuda(W, "W")
    B, N, M, _ = W.shape
    dev = W.device
    g = grad_out.detach().to(torch.float32).reshape(-1)
    if g.numel() == 1 and B > 1:
        g_stride = 0
    else:
        g = g.contiguous()
        g_stride = 1
    E = torch.empty((B, N, M), dtype=torch.float32, device=dev)
    scratch, n = _scratch(B, N, M, dev)
    with torch.cuda.device(dev), prof.span("softdtw_bwd", 3 * B * N * M * 4):
        rc = _C.lib().dae_softdtw_bwd(W.data_ptr(), g.data_ptr(), g_stride, B, N, M, E.data_ptr(),
                                      scratch.data_ptr(), n, _C.stream_ptr(dev))
    _C.check(rc, "dae_softdtw_bwd")
    return E


class _SoftDTWCUDA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, D, gamma, bandwidth):
        out, W, _ = softdtw_forward(D, gamma, bandwidth)
        ctx.save_for_backward(W)
        ctx.in_dtype = D.dtype
        return out.to(D.dtype)

    @staticmethod
    def backward(ctx, grad_output):
        W, = ctx.saved_tensors
        E = softdtw_backward(W, grad_output)
        return E.to(ctx.in_dtype), None, None


class SoftDTW(torch.nn.Module):
    def __init__(self, use_cuda=True, gamma=1.0, normalize=False, bandwidth=None, dist_func=None):
        super().__init__()
        self.normalize = normalize
        self.gamma = gamma
        self.bandwidth = 0 if bandwidth is None else float(bandwidth)
        self.use_cuda = use_cuda
        self.dist_func = dist_func if dist_func is not None else SoftDTW._euclidean_dist_func

    def _get_func_dtw(self, x, y):
        bx, lx, dx = x.shape
        by, ly, dy = y.shape
        assert bx == by
        assert dx == dy
        if not x.is_cuda:
            raise _C.DaeError("SoftDTW: inputs must be CUDA tensors; the dae build has no CPU path")
        return _SoftDTWCUDA.apply

    @staticmethod
    def _euclidean_dist_func(x, y):
        """Squared Euclidean distances [B,N,M] (soft_dtw_cuda.py:319-329) without the [B,N,M,d] expansion."""
        x2 = (x * x).sum(-1)[:, :, None]
        y2 = (y * y).sum(-1)[:, None, :]
        return (x2 + y2 - 2.0 * torch.bmm(x, y.transpose(1, 2))).clamp_min(0.0)

    def forward(self, X, Y):
        func_dtw = self._get_func_dtw(X, Y)
        if self.normalize:
            x = torch.cat([X, X, Y])
            y = torch.cat([Y, X, Y])
            D = self.dist_func(x, y)
            out = func_dtw(D, self.gamma, self.bandwidth)
            out_xy, out_xx, out_yy = torch.split(out, X.shape[0])
            return out_xy - 1 / 2 * (out_xx + out_yy)
        D_xy = self.dist_func(X, Y)
        return func_dtw(D_xy, self.gamma, self.bandwidth)
