"""Event-timed soft-DTW forward/backward at cfg4 ([8,4096,4096]) or a given size."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dae.soft_dtw_cuda import softdtw_backward, softdtw_forward  # noqa: E402

B, N, M = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (8, 4096, 4096)))
g = torch.Generator(device="cuda").manual_seed(0)
x, y = torch.rand(B, N, 2, generator=g, device="cuda"), torch.rand(B, M, 2, generator=g, device="cuda")
D = ((x[:, :, None, :] - y[:, None, :, :]) ** 2).sum(-1).contiguous()
go = torch.ones(B, device="cuda")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for rep in range(5):
    flush.add_(1.0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    _, W, _ = softdtw_forward(D, 1.0, 0.0)
    ev[1].record()
    softdtw_backward(W, go)
    ev[2].record()
    torch.cuda.synchronize()
    print(f"[{B},{N},{M}] fwd {ev[0].elapsed_time(ev[1]):.3f} ms  bwd {ev[1].elapsed_time(ev[2]):.3f} ms")
