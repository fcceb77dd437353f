"""Repeatability of the blocked CTC lattice under load: the same call many times, with other kernels queued in front,
every overlap mode; every loss must equal torch's to 1e-4 and be bit-identical across repetitions."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from kbench import peaky  # noqa: E402
from dae import _C  # noqa: E402
from dae.ctc import CTCLoss  # noqa: E402
from dae.greedy import greedy_ids_device  # noqa: E402

bad = 0
for (T, C, N) in ((2048, 4096, 1), (2048, 4096, 2), (512, 64, 1), (40, 7, 2)):
    g = torch.Generator(device="cuda").manual_seed(7)
    post = torch.stack([peaky(T, C, C - 1, g) for _ in range(N)], 1)
    labs = []
    for n in range(N):
        _, ids, k = greedy_ids_device(post[:, n], C - 1)
        labs.append(ids[0, :max(int(k[0]), 1)].long())
    Lmax = max(int(l.numel()) for l in labs)
    tg = torch.zeros(N, Lmax, dtype=torch.long, device="cuda")
    for n, l in enumerate(labs):
        tg[n, :l.numel()] = l
    tl = torch.tensor([int(l.numel()) for l in labs], device="cuda")
    il = torch.full((N,), T, device="cuda")
    ref = torch.nn.functional.ctc_loss(post, tg, il, tl, blank=C - 1, reduction="sum")
    xr = post.clone().requires_grad_()
    torch.nn.functional.ctc_loss(xr, tg, il, tl, blank=C - 1, reduction="sum").backward()
    big = torch.randn(4096, 4096, device="cuda")
    small = torch.randn(1 << 20, device="cuda")
    f = CTCLoss(blank=C - 1, reduction="sum", validate=False)
    for overlap in (0, 1, 3):
        _C.ctc_configure(blocked=1, overlap=overlap)
        for front in ("nothing", "elementwise", "gemm"):
            vals, grads = [], []
            for rep in range(10):
                if front == "gemm":
                    big @ big
                if front == "elementwise":
                    small.mul_(1.0001)
                x = post.clone().requires_grad_()
                nll = f.with_scale(x, tg, il, tl, 1.0)
                nll.backward()
                vals.append(nll.detach().clone())
                grads.append(x.grad.clone())
            torch.cuda.synchronize()
            same = all(torch.equal(v, vals[0]) for v in vals) and all(torch.equal(gr, grads[0]) for gr in grads)
            err = max(((v - ref).abs() / ref.abs().clamp_min(1)).max().item() for v in vals)
            gerr = max((gr - xr.grad).abs().max().item() for gr in grads) / xr.grad.abs().max().item()
            ok = same and err < 1e-4 and gerr < 5e-3
            bad += not ok
            print(f"T={T} C={C} N={N} L={Lmax} overlap={overlap} front={front}: identical={same} loss rel err {err:.2e} "
                  f"grad err/scale vs torch {gerr:.2e} {'ok' if ok else 'FAIL'}", flush=True)
_C.ctc_configure()
print("FAILURES", bad)
sys.exit(1 if bad else 0)
