"""numpy restatement of SpecAugment band masking + the [augmented, clean] batch build
(test infrastructure only).

Follows the call contract of lcasr.utils.augmentation.SpecAugment (un-vendored dependency,
version unpinned) at lcasr/lib.py:102-112,499,538-541 with torchaudio's
``mask_along_axis_iid`` semantics (torchaudio 2.11 in this image).  PARITY UNPINNED against
lcasr's own class; tests/test_oracle_pins.py pins the band draw + fill against
torchaudio.functional.mask_along_axis_iid run on the same RNG state.
"""
import numpy as np


def mask_value(x, zero_masking):
    """fp64 mean rounded once to fp32 (the kernel's definition), or 0."""
    return np.float32(0.0) if zero_masking else np.float32(np.asarray(x, dtype=np.float64).mean())


def apply_bands(x, fbands, tbands, fill):
    """x [F,T] fp32; fbands/tbands iterable of (start,end) half-open -> masked copy."""
    out = np.array(x, dtype=np.float32, copy=True)
    for s, e in tbands:
        out[:, int(s):int(e)] = fill
    for s, e in fbands:
        out[int(s):int(e), :] = fill
    return out


def specaug_repeat(x, fbands, tbands, zero_masking, n_clean, fill=None):
    """x [F,T]; fbands [n_aug][nf][2]; tbands [n_aug][nt][2] -> [n_aug+n_clean,F,T], fill used."""
    fill = mask_value(x, zero_masking) if fill is None else np.float32(fill)
    outs = [apply_bands(x, fbands[a], tbands[a], fill) for a in range(len(fbands))]
    outs += [np.array(x, dtype=np.float32, copy=True) for _ in range(n_clean)]
    return np.stack(outs), fill
