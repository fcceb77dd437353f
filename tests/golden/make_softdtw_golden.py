"""Generate tests/golden/softdtw_ref.npz from the REFERENCE's own soft-DTW file (imported read-only
from /root/reference/lcasr_nemo/soft_dtw_cuda.py, numba CPU kernels) at the shapes/seed its
profile() self-check uses (:382-428: seed 1234, (B,N,M,d) = (128,17,15,2), (512,64,64,2),
(512,256,256,2)) plus ragged, banded and gamma=1.5 cases.  Build container only.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "/root/reference/lcasr_nemo")
import soft_dtw_cuda as ref  # noqa: E402

CASES = [("p17x15", 128, 17, 15, 2, 1.0, 0.0, 8), ("p64", 512, 64, 64, 2, 1.0, 0.0, 4), ("p256", 512, 256, 256, 2, 1.0, 0.0, 2),
         ("rag", 3, 70, 45, 5, 1.5, 0.0, 3), ("band", 2, 96, 96, 3, 1.0, 12.0, 2), ("g01", 2, 50, 61, 4, 0.1, 0.0, 2)]


def main():
    out = {}
    torch.manual_seed(1234)
    for name, B, N, M, d, gamma, bw, keep in CASES:
        a = torch.rand((B, N, d))
        b = torch.rand((B, M, d))
        D = ref.SoftDTW._euclidean_dist_func(a, b).numpy().astype(np.float64)
        R = ref.compute_softdtw(D, gamma, bw)
        E = ref.compute_softdtw_backward(D, R.copy(), gamma, bw)
        out[f"{name}_x"] = a[:keep].numpy()
        out[f"{name}_y"] = b[:keep].numpy()
        out[f"{name}_val"] = R[:keep, -2, -2].astype(np.float64)
        out[f"{name}_E"] = E[:keep].astype(np.float32)
        out[f"{name}_meta"] = np.array([B, N, M, d, gamma, bw, keep], dtype=np.float64)
        # module-level API too (CPU autograd Function)
        x = a[:keep].clone().requires_grad_()
        v = ref.SoftDTW(False, gamma=gamma, bandwidth=bw if bw > 0 else None)(x, b[:keep])
        v.sum().backward()
        out[f"{name}_gx"] = x.grad.numpy()
        assert np.allclose(v.detach().numpy(), out[f"{name}_val"], rtol=1e-5)
        print(name, out[f"{name}_val"][:3])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "softdtw_ref.npz"), **out)


if __name__ == "__main__":
    main()
