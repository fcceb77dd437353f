"""Per-window greedy and SpecAugment launches (for `ncu --metrics gpu__time_duration.sum`): 10 launches each with an
L2 flush in between.  Host-side event timing of single 5-15 us kernels measures the launch path, not the kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from kbench import peaky  # noqa: E402
from dae.augment import SpecAugment  # noqa: E402
from dae.greedy import greedy_ids_device  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
lp = peaky(2048, 4096, 4095, g)
spec = torch.randn(1, 80, 120000, device="cuda")
aug = SpecAugment(n_freq_masks=6, freq_mask_param=34)
starts = [2048 * k for k in range(10)]
sums = SpecAugment.window_sums(spec, starts, [16384] * 10)
for k in range(10):
    flush.add_(1.0)
    greedy_ids_device(lp, 4095)
    flush.add_(1.0)
    aug(spec[:, :, starts[k]:starts[k] + 16384], n_clean=1, window_sums=sums[k])
torch.cuda.synchronize()
print("ok")
