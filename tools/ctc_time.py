"""Event-timed CTC loss+grad at the adapt-step shape (T=2048, N, C=4096), both lattice implementations.
    python tools/ctc_time.py [N]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from kbench import Timer, peaky  # noqa: E402
from dae import _C  # noqa: E402
from dae.ctc import CTCLoss  # noqa: E402
from dae.greedy import greedy_ids_device  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T, C = 2048, 4096
g = torch.Generator(device="cuda").manual_seed(7)
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
x = post.requires_grad_()
f = CTCLoss(blank=C - 1, reduction="sum", validate=False)
tm = Timer()


def fb():
    x.grad = None
    (f.with_scale(x, tg, il, tl, 1.0 / (T * N)) / (T * N)).backward()


def fwd():
    with torch.no_grad():
        f(x, tg, il, tl)


for name, blocked in (("chain", 0), ("blocked", 1)):
    if blocked and N > 8:
        continue
    _C.ctc_configure(blocked=blocked)
    a, _ = tm.time(fb, 20)
    b, _ = tm.time(fwd, 20)
    print(f"N={N} L={Lmax} {name}: loss+grad {a * 1e6:.1f} us, lattice only {b * 1e6:.1f} us", flush=True)
_C.ctc_configure()
