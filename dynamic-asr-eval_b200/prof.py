"""Device-side timing of the dae kernels inside a running step (bench.py's live roofline numbers).

When enabled, every kernel wrapper brackets its launches with CUDA events on the launching stream;
``summary()`` synchronises once and returns per-kernel launch counts, total device time and
algorithmic bytes.  Disabled (the default) it costs one attribute check per call.
"""
import contextlib
from collections import defaultdict

import torch

_enabled = False
h2d_bytes = 0          # bytes the dae host code copied host->device / device->host since reset()
d2h_bytes = 0
_records = defaultdict(list)     # name -> [(start_event, end_event, nbytes)]


def enable(flag=True):
    global _enabled
    _enabled = bool(flag)


def reset():
    global h2d_bytes, d2h_bytes
    _records.clear()
    h2d_bytes = 0
    d2h_bytes = 0


def count_h2d(t):
    global h2d_bytes
    h2d_bytes += t.numel() * t.element_size()


def count_d2h(t):
    global d2h_bytes
    d2h_bytes += t.numel() * t.element_size()


@contextlib.contextmanager
def span(name, nbytes=0):
    if not _enabled:
        yield
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    try:
        yield
    finally:
        e.record()
        _records[name].append((s, e, int(nbytes)))


def summary():
    torch.cuda.synchronize()
    out = {}
    for name, recs in _records.items():
        ms = [s.elapsed_time(e) for s, e, _ in recs]
        out[name] = {"launches": len(recs), "total_ms": sum(ms), "avg_ms": sum(ms) / max(len(ms), 1),
                     "bytes_per_launch": sum(b for _, _, b in recs) / max(len(recs), 1)}
    return out
