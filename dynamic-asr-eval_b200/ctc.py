"""CTC loss with the reference's call signature, computed by libdae.so.

Drop-in for ``torch.nn.CTCLoss(blank, reduction)`` as called at lcasr/lib.py:492,575
(and :250,324-329 for AWMC; earnings_finetune/train.py:259,286 for ragged batches).
Forward launches the alpha/beta lattice kernel; backward launches the dense gradient
kernel with the real upstream gradient, so ``loss / (T*N)`` then ``.backward()`` costs one
read of ``log_probs`` and one write of its gradient.
"""
import torch

from . import _C, prof


class _CTCFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_probs, targets, input_lengths, target_lengths, blank, validate=True, grad_scale_hint=None):
        _C.require_cuda(log_probs, "log_probs")
        if log_probs.dtype != torch.float32:
            raise _C.DaeError("dae.CTCLoss computes in fp32; got " + str(log_probs.dtype))
        ctx.unbatched = log_probs.dim() == 2
        if ctx.unbatched:  # [T,C]
            log_probs = log_probs.unsqueeze(1)
            targets = targets.unsqueeze(0) if targets.dim() == 1 else targets
        T, N, C = log_probs.shape
        lp = log_probs.detach()
        if lp.stride(2) != 1:
            lp = lp.contiguous()
        dev = lp.device
        in_len = torch.as_tensor(input_lengths, dtype=torch.int64).reshape(-1).to(dev, non_blocking=True).contiguous()
        tg_len = torch.as_tensor(target_lengths, dtype=torch.int64).reshape(-1).to(dev, non_blocking=True).contiguous()
        tg = targets.to(device=dev, dtype=torch.int64)
        if tg.dim() == 1:
            # concatenated 1-D targets (torch allows this): repack to padded 2-D on the host side
            lens = tg_len.tolist()
            Lmax = max(lens) if lens else 0
            packed = tg.new_zeros((N, max(Lmax, 1)))
            o = 0
            for i, l in enumerate(lens):
                packed[i, :l] = tg[o:o + l]
                o += l
            tg = packed
        if tg.dim() != 2 or tg.shape[0] != N:
            raise _C.DaeError(f"targets must be [N, Lmax]; got {tuple(tg.shape)} for N={N}")
        if tg.stride(1) != 1:
            tg = tg.contiguous()
        Lmax = int(tg.shape[1])
        if validate and Lmax:
            # torch raises on labels outside [0, C); the kernels index class rows with the raw label, so check on
            # the device without a host sync (the error surfaces at the next synchronisation point)
            used = torch.arange(Lmax, device=dev)[None, :] < tg_len[:, None]
            torch._assert_async((((tg >= 0) & (tg < C)) | ~used).all(),
                                "dae.CTCLoss: a target label lies outside [0, num_classes)")
        lib = _C.lib()
        nbytes = lib.dae_ctc_scratch_bytes(T, N, Lmax)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        nll = torch.empty(N, dtype=torch.float32, device=dev)
        ctx.blank = int(blank)
        ctx.hint = None
        if grad_scale_hint is not None and log_probs.requires_grad:
            # the reference's `loss / (T*N); loss.backward()` (lcasr/lib.py:573-579): the upstream scale is known
            # now, so loss and gradient are one library call (the dense part of the gradient runs under the
            # lattice scan); backward() only checks the scale
            ctx.hint = float(grad_scale_hint)
            g = torch.full((1,), ctx.hint, dtype=torch.float32, device=dev)
            grad = torch.empty((T, N, C), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev), prof.span("ctc_loss_grad", 2 * T * N * C * 4):
                rc = lib.dae_ctc_loss_grad(lp.data_ptr(), lp.stride(0), lp.stride(1), T, N, C,
                                           tg.data_ptr() if Lmax else None, tg.stride(0), Lmax,
                                           in_len.data_ptr(), tg_len.data_ptr(), int(blank), nll.data_ptr(),
                                           g.data_ptr(), 0, grad.data_ptr(), scratch.data_ptr(), nbytes,
                                           _C.stream_ptr(dev))
            _C.check(rc, "dae_ctc_loss_grad")
            ctx.save_for_backward(grad)
            ctx.shape = (T, N, C)
            return nll
        with torch.cuda.device(dev), prof.span("ctc_lattice", T * N * C * 4):
            rc = lib.dae_ctc_lattice(lp.data_ptr(), lp.stride(0), lp.stride(1), T, N, C,
                                     tg.data_ptr() if Lmax else None, tg.stride(0), Lmax,
                                     in_len.data_ptr(), tg_len.data_ptr(), int(blank),
                                     nll.data_ptr(), scratch.data_ptr(), nbytes, _C.stream_ptr(dev))
        _C.check(rc, "dae_ctc_lattice")
        ctx.save_for_backward(lp, tg, in_len, tg_len, nll, scratch)
        return nll

    @staticmethod
    def backward(ctx, grad_nll):
        if ctx.hint is not None:
            grad, = ctx.saved_tensors
            T, N, C = ctx.shape
            g = grad_nll.to(torch.float32)
            if g.numel() == 1:
                g, g_stride = g.reshape(1), 0
            else:
                g, g_stride = g.contiguous(), 1
            with torch.cuda.device(grad.device):
                rc = _C.lib().dae_ctc_rescale(grad.data_ptr(), T, N, C, g.data_ptr(), g_stride, ctx.hint,
                                              _C.stream_ptr(grad.device))
            _C.check(rc, "dae_ctc_rescale")
            return (grad[:, 0] if ctx.unbatched else grad), None, None, None, None, None, None
        lp, tg, in_len, tg_len, nll, scratch = ctx.saved_tensors
        T, N, C = lp.shape
        unbatched = ctx.unbatched
        Lmax = int(tg.shape[1])
        g = grad_nll.to(torch.float32)
        if g.numel() == 1:
            g = g.reshape(1)
            g_stride = 0
        else:
            g = g.contiguous()
            g_stride = 1
        grad = torch.empty((T, N, C), dtype=torch.float32, device=lp.device)
        with torch.cuda.device(lp.device), prof.span("ctc_grad", 2 * T * N * C * 4):
            rc = _C.lib().dae_ctc_grad(lp.data_ptr(), lp.stride(0), lp.stride(1), T, N, C,
                                       tg.data_ptr() if Lmax else None, tg.stride(0), Lmax,
                                       in_len.data_ptr(), tg_len.data_ptr(), ctx.blank,
                                       nll.data_ptr(), g.data_ptr(), g_stride, grad.data_ptr(),
                                       scratch.data_ptr(), scratch.numel(), _C.stream_ptr(lp.device))
        _C.check(rc, "dae_ctc_grad")
        return (grad[:, 0] if unbatched else grad), None, None, None, None, None, None


def ctc_loss(log_probs, targets, input_lengths, target_lengths, blank=0, reduction="mean", zero_infinity=False,
             validate=True, grad_scale_hint=None):
    """Functional form, same argument meaning as ``torch.nn.functional.ctc_loss``.  ``validate=False`` skips the
    device-side label range check (the adapt loop's labels come straight from the tokenizer)."""
    if zero_infinity:
        raise _C.DaeError("zero_infinity=True is not used by the reference (SURVEY.md appendix A) and is not implemented")
    unbatched = log_probs.dim() == 2
    if grad_scale_hint is not None and reduction != "sum":
        raise _C.DaeError("grad_scale_hint assumes reduction='sum' (every sample sees the same upstream gradient)")
    nll = _CTCFunction.apply(log_probs, targets, input_lengths, target_lengths, blank, validate, grad_scale_hint)
    if reduction == "sum":
        return nll.sum()
    if reduction == "none":
        return nll[0] if unbatched else nll
    if reduction == "mean":
        tl = torch.as_tensor(target_lengths, dtype=torch.float32).reshape(-1).to(nll.device).clamp_min(1)
        return (nll / tl).mean()
    raise ValueError(f"unknown reduction {reduction!r}")


class CTCLoss(torch.nn.Module):
    """``torch.nn.CTCLoss`` call signature (lcasr/lib.py:492) on the dae CUDA kernels."""

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = False, validate: bool = True):
        super().__init__()
        self.blank, self.reduction, self.zero_infinity, self.validate = blank, reduction, zero_infinity, validate

    def with_scale(self, log_probs, targets, input_lengths, target_lengths, grad_scale_hint):
        """Same loss; the gradient is formed right away for the upstream scale the caller is about to apply
        (``loss * grad_scale_hint`` then ``.backward()``), see _CTCFunction.forward.  Any other upstream gradient
        still gives the right result (dae_ctc_rescale)."""
        return ctc_loss(log_probs, targets, input_lengths, target_lengths, self.blank, self.reduction,
                        self.zero_infinity, self.validate, grad_scale_hint)

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        return ctc_loss(log_probs, targets, input_lengths, target_lengths, self.blank, self.reduction,
                        self.zero_infinity, self.validate)
