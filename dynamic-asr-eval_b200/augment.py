"""SpecAugment with the call signature the reference uses, masks applied by libdae.so.

Drop-in for ``lcasr.utils.augmentation.SpecAugment`` (un-vendored; constructor keys from
lcasr/lib.py:102-112 and earnings_finetune/lcasr160rb1.yaml:73-81; called at lib.py:499,541,
:290 and earnings_finetune/train.py:230).  PARITY UNPINNED against lcasr's own class: the
semantics here are torchaudio's ``mask_along_axis_iid`` (mask value = mean of the input or 0;
width = floor(U*param), start = floor(U*(size-width)); per batch item iid; time bands first,
then frequency bands).  Band descriptors are drawn on the HOST with ``torch.rand`` in exactly
the order of :func:`draw_bands`, so an oracle can replay them (SURVEY.md §7 "RNG parity").
"""
import ctypes

import torch

from . import _C, prof

MAX_BANDS = 32


def draw_bands(n_items, n_masks, param, size, generator=None):
    """Draw ``n_masks`` (start, end) bands for each of ``n_items`` batch items.

    RNG order per mask (same as torchaudio.functional.mask_along_axis_iid): one
    ``torch.rand(n_items)`` for the widths, one for the starts.  Returns int32 [n_items, n_masks, 2].
    """
    out = torch.zeros((n_items, max(n_masks, 0), 2), dtype=torch.int32)
    for m in range(n_masks):
        value = torch.rand(n_items, generator=generator) * param
        min_value = torch.rand(n_items, generator=generator) * (size - value)
        start = min_value.long()
        end = start + value.long()
        out[:, m, 0] = start.to(torch.int32)
        out[:, m, 1] = end.to(torch.int32)
    return out


class SpecAugment(torch.nn.Module):
    def __init__(self, n_time_masks=0, n_freq_masks=0, freq_mask_param=42, time_mask_param=-1, min_p=0.05,
                 zero_masking=False, iid_masks=True, max_p=1.0):
        super().__init__()
        self.n_time_masks, self.n_freq_masks = int(n_time_masks), int(n_freq_masks)
        self.freq_mask_param, self.time_mask_param = freq_mask_param, time_mask_param
        self.min_p, self.max_p = min_p, max_p
        self.zero_masking, self.iid_masks = bool(zero_masking), iid_masks
        if self.n_time_masks > MAX_BANDS or self.n_freq_masks > MAX_BANDS:
            raise _C.DaeError(f"at most {MAX_BANDS} bands per axis")
        self.last_bands = None  # (fbands, tbands) of the most recent call, for tests/replay

    def _time_param(self, T):
        # time_mask_param <= 0 means "adaptive": a band may cover up to min_p of the frames
        # (lib.py:107-108 defaults time_mask_param=-1, min_p=0.05).  Unpinned: lcasr is absent.
        p = self.time_mask_param if self.time_mask_param > 0 else int(self.min_p * T)
        return min(p, int(self.max_p * T)) if self.max_p is not None else p

    def draw(self, B, F, T, generator=None):
        tb = draw_bands(B, self.n_time_masks, self._time_param(T), T, generator)
        fb = draw_bands(B, self.n_freq_masks, self.freq_mask_param, F, generator)
        return fb, tb

    @staticmethod
    def window_sums(spec: torch.Tensor, starts, lens):
        """Partial sums of every window ``spec[:, :, s:s+l]`` of a recording in ONE launch (dae_window_sums):
        -> float64 [n_win, dae_window_slices()] on the device.  Row w goes to ``forward(..., window_sums=sums[w])``,
        which then needs no reduction pass and no grid barrier of its own."""
        _C.require_cuda(spec, "spec")
        x = spec[0] if spec.dim() == 3 else spec
        if x.dtype != torch.float32 or x.stride(1) != 1:
            raise _C.DaeError("window_sums needs an fp32 spectrogram with contiguous time axis")
        lib, dev = _C.lib(), x.device
        n = len(starts)
        ws = torch.tensor(list(starts), dtype=torch.int64).to(dev, non_blocking=True)
        wl = torch.tensor(list(lens), dtype=torch.int64).to(dev, non_blocking=True)
        sums = torch.empty((n, lib.dae_window_slices()), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev), prof.span("window_sums", int(sum(lens)) * x.shape[0] * 4):
            rc = lib.dae_window_sums(x.data_ptr(), x.stride(0), x.shape[0], ws.data_ptr(), wl.data_ptr(), n,
                                     sums.data_ptr(), _C.stream_ptr(dev))
        _C.check(rc, "dae_window_sums")
        return sums

    def forward(self, specgram: torch.Tensor, lengths=None, n_clean: int = 0, bands=None, window_sums=None):
        """specgram [B,F,T] (or [F,T]) fp32 on the GPU -> masked copy, same shape.

        With ``n_clean=k`` the result is [B+k, F, T]: the B masked items followed by k verbatim
        copies of item 0 (the adapt step's ``repeat(2,1,1)`` batch, lib.py:539-541, in one pass).
        All B items must then be views of the same window (they are in the reference).
        """
        _C.require_cuda(specgram, "specgram")
        x = specgram
        squeeze = x.dim() == 2
        if squeeze:
            x = x.unsqueeze(0)
        if x.dtype != torch.float32:
            raise _C.DaeError("SpecAugment computes in fp32")
        B, F, T = x.shape
        fb, tb = bands if bands is not None else self.draw(B, F, T)
        self.last_bands = (fb, tb)
        lib = _C.lib()
        dev = x.device
        out = torch.empty((B + n_clean, F, T), dtype=torch.float32, device=dev)
        scratch = torch.empty(lib.dae_specaug_scratch_bytes(), dtype=torch.uint8, device=dev)
        nf, nt = int(fb.shape[1]), int(tb.shape[1])
        st = _C.stream_ptr(dev)
        with torch.cuda.device(dev), prof.span("specaug_repeat", (1 + B + n_clean) * F * T * 4):
            if window_sums is not None:
                # the window's partial sums are known (SpecAugment.window_sums): one plain single-pass launch
                src = x[0]
                if src.stride(1) != 1:
                    src = src.contiguous()
                fbc, tbc = fb.contiguous(), tb.contiguous()
                rc = lib.dae_specaug_repeat_premean(src.data_ptr(), src.stride(0), F, T, fbc.data_ptr() if nf else None,
                                                    nf, tbc.data_ptr() if nt else None, nt, int(self.zero_masking), B,
                                                    n_clean, out.data_ptr(), window_sums.data_ptr(), None, st)
                _C.check(rc, "dae_specaug_repeat_premean")
            elif n_clean:
                # one launch pair: B masked copies + n_clean clean copies of the shared window
                src = x[0]
                if src.stride(1) != 1:
                    src = src.contiguous()
                fbc, tbc = fb.contiguous(), tb.contiguous()
                rc = lib.dae_specaug_repeat(src.data_ptr(), src.stride(0), F, T, fbc.data_ptr() if nf else None, nf,
                                            tbc.data_ptr() if nt else None, nt, int(self.zero_masking), B, n_clean,
                                            out.data_ptr(), scratch.data_ptr(), None, st)
                _C.check(rc, "dae_specaug_repeat")
            else:
                # the mask value is the mean over the whole [B,F,T] input in the reference
                # (specgram.mean()); with B > 1 distinct items we use each call's own mean only
                # when B == 1, else the caller-visible mean of the batch computed on device.
                for b in range(B):
                    src = x[b]
                    if src.stride(1) != 1:
                        src = src.contiguous()
                    fbc, tbc = fb[b:b + 1].contiguous(), tb[b:b + 1].contiguous()
                    rc = lib.dae_specaug_repeat(src.data_ptr(), src.stride(0), F, T, fbc.data_ptr() if nf else None, nf,
                                                tbc.data_ptr() if nt else None, nt, int(self.zero_masking), 1, 0,
                                                out[b].data_ptr(), scratch.data_ptr(), None, st)
                    _C.check(rc, "dae_specaug_repeat")
        return out[0] if squeeze else out


CUTOUT_MODES = {"zero": 0, "mean": 1, "mean_recording": 2}


def draw_cutout_rects(spec_n, n_freq, seq_len, num_rectangles=5, max_width=100, max_height=10, generator=None):
    """Rectangle draw of lcasr/lib.py:391-401, same torch.randint order: widths, heights, start_x, start_y.
    Returns int32 [n, 4] = (start_x, end_x, start_y, end_y)."""
    ratio = spec_n / seq_len
    n = int(num_rectangles * ratio)
    widths = torch.randint(1, max_width, (n,), generator=generator)
    heights = torch.randint(1, max_height, (n,), generator=generator)
    sx = torch.randint(0, spec_n, (n,), generator=generator)
    ex = (sx + widths).clamp(max=spec_n)
    sy = torch.randint(0, n_freq, (n,), generator=generator)
    ey = (sy + heights).clamp(max=n_freq)
    return torch.stack([sx, ex, sy, ey], 1).to(torch.int32).contiguous()


def cutout(spec, seq_len, cutout_val="mean", num_rectangles=5, max_width=100, max_height=10, rects=None):
    """lcasr/lib.py:384-417 on the GPU, in place.  ``spec`` [1,F,T] (or [F,T]) fp32 CUDA.  ``rects`` overrides
    the random draw (tests / replay)."""
    if num_rectangles == 0:
        return spec
    _C.require_cuda(spec, "spec")
    if cutout_val not in CUTOUT_MODES:
        raise _C.DaeError(f"cutout_val must be one of {sorted(CUTOUT_MODES)}")
    x = spec[0] if spec.dim() == 3 else spec
    if spec.dim() == 3 and spec.shape[0] != 1:
        raise _C.DaeError("cutout assumes a batch size of 1 (lcasr/lib.py:387)")
    if x.dtype != torch.float32 or x.stride(1) != 1:
        raise _C.DaeError("cutout needs an fp32 tensor with contiguous time axis")
    F, T = x.shape
    if rects is None:
        rects = draw_cutout_rects(T, F, seq_len, num_rectangles, max_width, max_height)
    rects = rects.to(torch.int32).contiguous().cpu()
    n = int(rects.shape[0])
    if n == 0:
        return spec
    lib = _C.lib()
    nbytes = lib.dae_cutout_scratch_bytes(n)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device), prof.span("cutout", 0):
        rc = lib.dae_cutout(x.data_ptr(), x.stride(0), F, T, rects.data_ptr(), n, CUTOUT_MODES[cutout_val],
                            scratch.data_ptr(), nbytes, _C.stream_ptr(x.device))
    _C.check(rc, "dae_cutout")
    return spec


def draw_frame_shuffle(F, T, time_dimension=False, freq_dimension=False, generator=None):
    """Permutations of lcasr/lib.py:81-84 from the HOST generator in the reference's order: randperm(T) if
    time_dimension, then randperm(F) if freq_dimension.  -> (perm_t or None, perm_f or None), int64 CPU."""
    pt = torch.randperm(T, generator=generator) if time_dimension else None
    pf = torch.randperm(F, generator=generator) if freq_dimension else None
    return pt, pf


def frame_shuffle(spec, time_dimension=False, freq_dimension=False, perms=None):
    """lcasr/lib.py:81-84 on the GPU: ``spec[:, :, randperm(T)]`` then ``spec[:, randperm(F), :]`` as one gather
    kernel.  ``spec`` [B,F,T] fp32 CUDA; like the reference, ONE permutation per axis is shared by the batch.
    ``perms`` = (perm_t, perm_f) overrides the draw (tests / replay)."""
    if not (time_dimension or freq_dimension):
        return spec
    _C.require_cuda(spec, "spec")
    if spec.dim() != 3 or spec.dtype != torch.float32:
        raise _C.DaeError("frame_shuffle needs a [B,F,T] fp32 tensor")
    B, F, T = spec.shape
    pt, pf = perms if perms is not None else draw_frame_shuffle(F, T, time_dimension, freq_dimension)
    dev = spec.device
    dpt = pt.to(torch.int32).to(dev, non_blocking=True) if pt is not None else None
    dpf = pf.to(torch.int32).to(dev, non_blocking=True) if pf is not None else None
    out = torch.empty((B, F, T), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), prof.span("frame_shuffle", 2 * B * F * T * 4):
        for b in range(B):
            src = spec[b] if spec[b].stride(1) == 1 else spec[b].contiguous()
            rc = _C.lib().dae_frame_shuffle(src.data_ptr(), src.stride(0), F, T, _C.ptr(dpt), _C.ptr(dpf),
                                            out[b].data_ptr(), _C.stream_ptr(dev))
            _C.check(rc, "dae_frame_shuffle")
    return out


def add_random_noise(spec, noise_factor, z=None):
    """lcasr/lib.py:379-382 on the GPU, in place on ``spec`` ([B,F,T] or [F,T] fp32 CUDA):
    ``spec + torch.normal(0, spec.std(), size) * noise_factor``.  The reference draws on the CPU from the global
    generator; ``torch.normal(0, s, size)`` is bitwise ``torch.randn(size) * s`` with the same generator
    consumption, so the standard-normal field ``z`` is drawn on the host (or passed in) and the std / scale /
    add run in dae_add_noise."""
    if noise_factor == 0:
        return spec
    _C.require_cuda(spec, "spec")
    if spec.dtype != torch.float32 or not spec.is_contiguous():
        raise _C.DaeError("add_random_noise needs a contiguous fp32 tensor")
    if z is None:
        z = torch.randn(spec.shape)
    if not z.is_cuda:
        prof.count_h2d(z)
    zd = z.to(device=spec.device, dtype=torch.float32, non_blocking=True).contiguous()
    rows, T = spec.numel() // spec.shape[-1], spec.shape[-1]
    lib = _C.lib()
    n = lib.dae_noise_scratch_bytes()
    scratch = torch.empty(n, dtype=torch.uint8, device=spec.device)
    with torch.cuda.device(spec.device), prof.span("add_noise", 3 * spec.numel() * 4):
        rc = lib.dae_add_noise(spec.data_ptr(), T, rows, T, zd.data_ptr(), float(noise_factor), scratch.data_ptr(), n,
                               None, _C.stream_ptr(spec.device))
    _C.check(rc, "dae_add_noise")
    return spec
