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
            loss = f(x, tg, il, tl)
            e.record()
            (loss / T).backward()
            torch.cuda.synchronize()
            print("lattice ms", s.elapsed_time(e), "loss", loss.item())
    else:
        raise SystemExit("unknown kernel family " + a.what)


if __name__ == "__main__":
    main()
