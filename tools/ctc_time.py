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


def queued(fn, reps=30):
    """Device time of fn() when the host is ahead of the GPU (as inside the adapt step, behind the encoder's
    kernels): a ~1 ms GEMM is queued first, then event / fn / event."""
    big = torch.randn(8192, 8192, device="cuda")
    ts = []
    for _ in range(reps):
        big @ big
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def lossgrad_only():
    x.grad = None
    f.with_scale(x, tg, il, tl, 1.0 / (T * N))


for name, blocked, overlap in (("chain", 0, -1), ("blocked, serial (overlap=0)", 1, 0),
                               ("blocked, dense gradient under the scan (1)", 1, 1),
                               ("blocked, + label-class kernel resident early (3)", 1, 3)):
    if blocked and N > 8:
        continue
    if os.environ.get("DAE_ONLY_SERIAL") and overlap not in (0,):
        continue
    _C.ctc_configure(blocked=blocked, overlap=overlap)
    a, _ = tm.time(fb, 20)
    b, _ = tm.time(fwd, 20)
    qm, qb = queued(lossgrad_only)
    fm, fbst = queued(fwd)
    print(f"N={N} L={Lmax} {name}: loss+grad {a * 1e6:.1f} us, lattice only {b * 1e6:.1f} us | host ahead: "
          f"with_scale call {qm:.1f} us (best {qb:.1f}), lattice only {fm:.1f} us (best {fbst:.1f})", flush=True)
_C.ctc_configure()
