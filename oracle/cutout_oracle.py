"""numpy restatement of cutout (test infrastructure only): lcasr/lib.py:384-417 given the rectangle table.
Pinned by tests/golden/loop_toy.npz (keys cutout_*), produced by the reference's own cutout() under import
stubs with a fixed torch seed (tests/golden/make_loop_golden.py)."""
import numpy as np


def cutout(spec, rects, cutout_val):
    """spec [F,T] fp32; rects [n,4] = (sx, ex, sy, ey); returns a filled copy."""
    out = np.array(spec, dtype=np.float32, copy=True)
    if cutout_val == "mean_recording":
        whole = np.float32(out.mean())
    vals = None
    if cutout_val == "mean":
        vals = [np.float32(out[sy:ey, sx:ex].mean()) for sx, ex, sy, ey in rects]
    for i, (sx, ex, sy, ey) in enumerate(rects):
        if cutout_val == "mean":
            out[sy:ey, sx:ex] = vals[i]
        elif cutout_val == "mean_recording":
            out[sy:ey, sx:ex] = whole
        elif cutout_val == "zero":
            out[sy:ey, sx:ex] = 0
    return out
