"""Overlap-average stitch of window posteriors on the GPU (lcasr/lib.py:583-629).

The host computes the window positions with the reference's own arithmetic (Python float
ratio, ``int(overlap / ratio)``) and the kernel gathers, for every
covered output row, the windows that contain it: exp -> sum in window order -> / count -> log,
plus the row argmax so the whole-recording greedy decode needs no second pass.
"""
import torch

from . import _C, prof


def window_positions(starts, u_lens, ds_lens, overlap):
    """lib.py:586-588,615-621: first output row of each window, windows in sorted start order."""
    pos, out = 0, []
    for i, u_len, ds_len in zip(starts, u_lens, ds_lens):
        ratio = u_len / ds_len
        overlap_ds = int(overlap / ratio)
        pos -= overlap_ds if i != 0 else 0
        out.append(pos)
        pos += ds_len
    return out


def stitch_flat(flat, offs, pos, ds, want_path=True):
    """flat [rows, C] fp32 CUDA holding window w at rows offs[w] .. offs[w]+ds[w]; pos[w] = its first output
    row (non-decreasing).  Returns (log_probs [N_out, C], path [N_out] int32 or None) on the device."""
    _C.require_cuda(flat, "flat")
    dev = flat.device
    C = int(flat.shape[-1])
    if min(pos) < 0:
        raise _C.DaeError("negative window position: overlap exceeds the previous window")
    covered = torch.zeros(max(p + d for p, d in zip(pos, ds)), dtype=torch.bool)
    for p, d in zip(pos, ds):
        covered[p:p + d] = True
    row_map = torch.nonzero(covered).reshape(-1).to(torch.int64)
    n_out = int(row_map.numel())
    meta = torch.tensor([list(offs), list(pos), list(ds)], dtype=torch.int64).to(dev, non_blocking=True)
    row_map = row_map.to(dev, non_blocking=True)
    out = torch.empty((n_out, C), dtype=torch.float32, device=dev)
    path = torch.empty((n_out,), dtype=torch.int32, device=dev) if want_path else None
    if flat.dtype != torch.float32 or not flat.is_contiguous():
        flat = flat.float().contiguous()
    with torch.cuda.device(dev), prof.span("stitch", (sum(ds) + n_out) * C * 4):
        rc = _C.lib().dae_stitch(flat.data_ptr(), C, meta[0].data_ptr(), meta[1].data_ptr(), meta[2].data_ptr(),
                                 len(ds), row_map.data_ptr(), n_out, out.data_ptr(), _C.ptr(path),
                                 _C.stream_ptr(dev))
    _C.check(rc, "dae_stitch")
    return out, path


def stitch_windows(window_lps, starts, u_lens, overlap, want_path=True):
    """window_lps: list of [T'_w, C] fp32 CUDA log-prob tensors keyed by ``starts`` (any order).

    Returns (log_probs [N_out, C] fp32 CUDA, path [N_out] int32 CUDA or None).
    """
    order = sorted(range(len(starts)), key=lambda k: starts[k])
    wins = [window_lps[k] for k in order]
    starts = [starts[k] for k in order]
    u_lens = [u_lens[k] for k in order]
    C = int(wins[0].shape[-1])
    ds = [int(w.shape[0]) for w in wins]
    pos = window_positions(starts, u_lens, ds, overlap)
    flat = wins[0] if len(wins) == 1 else torch.cat([w.reshape(-1, C) for w in wins], 0)
    offs, o = [], 0
    for d in ds:
        offs.append(o)
        o += d
    return stitch_flat(flat, offs, pos, ds, want_path)
