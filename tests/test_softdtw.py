"""soft-DTW: oracle pinned to the reference's numba kernels (CPU), CUDA kernel vs oracle (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import softdtw_oracle as so

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "softdtw_ref.npz"))
CASES = ["p17x15", "p64", "p256", "rag", "band", "g01"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_numba(name):
    B, N, M, d, gamma, bw, keep = GOLD[f"{name}_meta"]
    x, y = GOLD[f"{name}_x"].astype(np.float64), GOLD[f"{name}_y"].astype(np.float64)
    D = so.sqeuclidean(x, y)
    R = so.forward(D, gamma, bw)
    np.testing.assert_allclose(R[:, -2, -2], GOLD[f"{name}_val"], rtol=1e-6)      # fp32 inputs, fp64 math
    E = so.backward(D, R, gamma, bw)
    np.testing.assert_allclose(E, GOLD[f"{name}_E"], rtol=2e-5, atol=1e-7)


def test_oracle_borders_and_band():
    D = np.random.default_rng(0).random((1, 5, 7))
    R = so.forward(D, 1.0, 2.0)
    assert R[0, 0, 0] == 0 and np.isinf(R[0, 0, 1:]).all() and np.isinf(R[0, 1:, 0]).all()
    assert np.isinf(R[0, 1, 4]) and np.isfinite(R[0, 3, 5])                        # |i-j| > 2 is pruned
    E = so.backward(D, R, 1.0, 2.0)
    assert E[0, 0, 3] == 0 and E[0, 4, 1] == 0 and abs(E[0, 4, 6] - 1) < 1e-12    # pruned cells get no gradient


# ------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference_golden(cuda, name):
    from dae.soft_dtw_cuda import SoftDTW
    B, N, M, d, gamma, bw, keep = GOLD[f"{name}_meta"]
    x = torch.from_numpy(GOLD[f"{name}_x"]).to(cuda).requires_grad_()
    y = torch.from_numpy(GOLD[f"{name}_y"]).to(cuda)
    sd = SoftDTW(True, gamma=float(gamma), bandwidth=float(bw) if bw > 0 else None)
    v = sd(x, y)
    v.sum().backward()
    # north_star: within 1e-4 of the numba reference
    np.testing.assert_allclose(v.detach().cpu().numpy(), GOLD[f"{name}_val"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(x.grad.cpu().numpy(), GOLD[f"{name}_gx"], rtol=1e-3, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,M,gamma,bw", [(2, 100, 37, 1.0, 0.0), (3, 33, 129, 1.5, 0.0), (1, 1, 1, 1.0, 0.0),
                                            (2, 31, 64, 0.5, 0.0), (2, 257, 250, 1.0, 20.0), (1, 1024, 1024, 1.0, 0.0),
                                            (2, 65, 1, 1.0, 0.0), (1, 5, 300, 2.0, 0.0)])
def test_cuda_matches_oracle_matrices(cuda, B, N, M, gamma, bw):
    from dae.soft_dtw_cuda import _SoftDTWCUDA
    g = torch.Generator().manual_seed(N * 1000 + M)
    D = torch.rand(B, N, M, generator=g) * 2
    Dg = D.to(cuda).requires_grad_()
    gout = torch.rand(B, generator=g) + 0.5
    v = _SoftDTWCUDA.apply(Dg, gamma, bw)
    (v * gout.to(cuda)).sum().backward()
    R = so.forward(D.numpy(), gamma, bw)
    E = so.backward(D.numpy(), R, gamma, bw) * gout.numpy()[:, None, None]
    np.testing.assert_allclose(v.detach().cpu().numpy(), R[:, -2, -2], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(Dg.grad.cpu().numpy(), E, rtol=1e-3, atol=1e-5)


@pytest.mark.gpu
def test_cuda_full_matrix_R(cuda):
    from dae.soft_dtw_cuda import softdtw_forward
    D = torch.rand(2, 70, 90, generator=torch.Generator().manual_seed(1))
    _, R, _ = softdtw_forward(D.to(cuda), 1.0, 0.0)
    ref = so.forward(D.numpy(), 1.0, 0.0)[:, 1:-1, 1:-1]
    np.testing.assert_allclose(R.cpu().numpy(), ref, rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_cuda_large_properties(cuda):
    """BASELINE size (cfg4 is [8,4096,4096]; one sample here keeps the CPU oracle to seconds):
    value vs the fp64 oracle, and sum(E * D-perturbation) consistency via the normalize identity."""
    from dae.soft_dtw_cuda import SoftDTW, _SoftDTWCUDA
    g = torch.Generator().manual_seed(1234)
    a, b = torch.rand(1, 4096, 2, generator=g), torch.rand(1, 4096, 2, generator=g)
    D = SoftDTW._euclidean_dist_func(a, b)
    Dg = D.to(cuda).requires_grad_()
    v = _SoftDTWCUDA.apply(Dg, 1.0, 0.0)
    v.sum().backward()
    R = so.forward(D.numpy(), 1.0, 0.0)
    assert abs(v.item() - R[0, -2, -2]) <= 1e-4 * abs(R[0, -2, -2])
    E = so.backward(D.numpy(), R, 1.0, 0.0)
    # R ~ -5000 is stored in fp32 (ulp 5e-4), as in the reference's own CUDA path (dtype=D.dtype, :133), so
    # the transition weights exp((R' - R - D)/gamma) carry ~5e-4 relative noise; the reference accepts
    # atol 1e-3 between its CPU and CUDA paths already at 256x256 (:405-406,426-428).
    np.testing.assert_allclose(Dg.grad.cpu().numpy(), E, rtol=1e-2, atol=1e-3)
    assert np.abs(Dg.grad.cpu().numpy() - E).mean() < 2e-5
    # every alignment passes through exactly one cell of the first row and of the first column pair:
    # the gradient mass entering the last cell is 1
    assert abs(Dg.grad[0, -1, -1].item() - 1.0) < 1e-5


@pytest.mark.gpu
def test_softdtw_module_normalize_and_errors(cuda):
    import dae._C as C
    from dae.soft_dtw_cuda import SoftDTW
    g = torch.Generator().manual_seed(3)
    X, Y = torch.rand(4, 48, 3, generator=g), torch.rand(4, 48, 3, generator=g)   # normalize needs equal lengths (:342)
    v = SoftDTW(True, gamma=1.5, normalize=True)(X.to(cuda), Y.to(cuda)).cpu().numpy()
    def val(p, q):
        return so.forward(so.sqeuclidean(p.numpy(), q.numpy()), 1.5)[:, -2, -2]
    ref = val(X, Y) - 0.5 * (val(X, X) + val(Y, Y))
    np.testing.assert_allclose(v, ref, rtol=1e-3, atol=1e-3)
    with pytest.raises(C.DaeError):
        SoftDTW(True)(X, Y)
