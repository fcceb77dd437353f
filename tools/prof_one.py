"""Run ONE kernel family a few times at its BASELINE shape (for ncu captures under gpurun).

    python tools/prof_one.py ctc|greedy|specaug|stitch|softdtw|beam [--n N] [--reps R] [--noflush]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from kbench import peaky  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what")
    ap.add_argument("--n", type=int, default=1)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--noflush", action="store_true")
    ap.add_argument("--with-scale", action="store_true", help="ctc: loss and gradient as one call (the adapt step)")
    a = ap.parse_args()
    g = torch.Generator(device="cuda").manual_seed(0)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    if a.what == "ctc":
        from dae.ctc import CTCLoss
        from dae.greedy import greedy_ids_device
        T, C, N = 2048, 4096, a.n
        post = torch.stack([peaky(T, C, C - 1, g) for _ in range(N)], 1)
        labs = []
        for n in range(N):
            _, ids, k = greedy_ids_device(post[:, n], C - 1)
            labs.append(ids[0, :int(k[0])].long())
        Lmax = max(int(l.numel()) for l in labs)
        tg = torch.zeros(N, Lmax, dtype=torch.long, device="cuda")
        for n, l in enumerate(labs):
            tg[n, :l.numel()] = l
        tl = torch.tensor([int(l.numel()) for l in labs], device="cuda")
        il = torch.full((N,), T, device="cuda")
        x = post.clone().requires_grad_()
        f = CTCLoss(blank=C - 1, reduction="sum")
        for _ in range(a.reps):
            if not a.noflush:
                flush.add_(1.0)
            x.grad = None
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            loss = f.with_scale(x, tg, il, tl, 1.0 / T) if a.with_scale else f(x, tg, il, tl)
            e.record()
            scaled = loss / T
            s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(600000)           # keep the GPU busy (no memory traffic) while the CPU queues the backward
            s2.record()
            scaled.backward()
            e2.record()
            torch.cuda.synchronize()
            print("lattice ms", s.elapsed_time(e), "backward ms (incl. torch autograd glue)", s2.elapsed_time(e2),
                  "loss", loss.item())
    elif a.what == "greedy":
        from dae.greedy import greedy_ids_device
        lp = peaky(52000, 4096, 4095, g)
        for _ in range(a.reps):
            flush.add_(1.0)
            greedy_ids_device(lp, 4095)
        torch.cuda.synchronize()
    elif a.what == "specaug":
        from dae.augment import SpecAugment
        spec = torch.randn(1, 80, 120000, device="cuda")
        aug = SpecAugment(n_freq_masks=6, freq_mask_param=34)
        for _ in range(a.reps):
            flush.add_(1.0)
            aug(spec[:, :, 2048:2048 + 16384], n_clean=1)
        torch.cuda.synchronize()
    elif a.what == "stitch":
        from dae.stitch import stitch_flat, window_positions
        nwin, Tp, C = 52, 2048, 4096
        flat = torch.cat([peaky(Tp, C, C - 1, g) for _ in range(nwin)], 0)
        starts = [2048 * i for i in range(nwin)]
        pos = window_positions(starts, [16384] * nwin, [Tp] * nwin, 14336)
        for _ in range(a.reps):
            flush.add_(1.0)
            stitch_flat(flat, [Tp * i for i in range(nwin)], pos, [Tp] * nwin)
        torch.cuda.synchronize()
    elif a.what == "softdtw":
        from dae.soft_dtw_cuda import softdtw_backward, softdtw_forward
        B, N, M = 8, a.n if a.n > 1 else 4096, a.n if a.n > 1 else 4096
        x, y = torch.rand(B, N, 2, generator=g, device="cuda"), torch.rand(B, M, 2, generator=g, device="cuda")
        D = ((x[:, :, None, :] - y[:, None, :, :]) ** 2).sum(-1).contiguous()
        go = torch.ones(B, device="cuda")
        for _ in range(a.reps):
            flush.add_(1.0)
            _, W, _ = softdtw_forward(D, 1.0, 0.0)
            softdtw_backward(W, go)
        torch.cuda.synchronize()
    elif a.what == "beam":
        import numpy as np
        from dae.standin import peaky_log_probs
        from dae.ctc_beam_search import _Search
        from dae.ngram import NGramLM, read_arpa, write_synthetic_arpa
        V, T, nseg = 31, 18000, 36
        write_synthetic_arpa("/tmp/prof.arpa", V, order=4, counts=(None, 900, 20000, 80000), seed=4, fast=True)
        order, grams = read_arpa("/tmp/prof.arpa")
        lm = NGramLM(grams, order, V)
        lp = torch.from_numpy(peaky_log_probs(T, V + 1, V, 3, sharp=5.0)).cuda()
        sr = _Search(lp, [int(v) for v in np.linspace(0, T, nseg + 1)], lm, 100, 0.45, 1.53, V, 0.0, 0.0, -6, 3.17, n_best=1)
        for _ in range(a.reps):
            sr.run_all()
        torch.cuda.synchronize()
    else:
        raise SystemExit("unknown kernel family " + a.what)


if __name__ == "__main__":
    main()
