"""Event-timed prefix beam search at the bench's cfg3 shape (1 h at 50 fps, V=32, beam 100, 360 segments)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dae.ctc_beam_search import _Search  # noqa: E402
from dae.ngram import NGramLM, read_arpa, write_synthetic_arpa  # noqa: E402
from dae.standin import peaky_log_probs  # noqa: E402

V, T, nseg = 31, 180000, 360
write_synthetic_arpa("/tmp/bt.arpa", V, order=4, counts=(None, 900, 20000, 80000), seed=4, fast=True)
order, grams = read_arpa("/tmp/bt.arpa")
lm = NGramLM(grams, order, V)
lp = torch.from_numpy(peaky_log_probs(T, V + 1, V, 3, sharp=5.0)).cuda()
sr = _Search(lp, [int(v) for v in np.linspace(0, T, nseg + 1)], lm, 100, 0.45, 1.53, V, 0.0, 0.0, -6, 3.17, n_best=1)
for rep in range(4):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out = sr.run_all()
    e.record()
    torch.cuda.synchronize()
    print(f"{nseg} segments x {T // nseg} frames: {s.elapsed_time(e):.2f} ms")
