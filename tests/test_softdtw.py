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
    # gx = sum_j E * dD/dx: 1e-4 of the gradient's scale (the golden E itself carries the reference's fp32 R)
    gx = GOLD[f"{name}_gx"]
    np.testing.assert_allclose(x.grad.cpu().numpy(), gx, rtol=1e-4, atol=1e-4 * np.abs(gx).max())


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
    np.testing.assert_allclose(v.detach().cpu().numpy(), R[:, -2, -2], rtol=1e-5, atol=1e-5)
    # north_star: 1e-4; measured ~1e-6 of the gradient scale (E <= gout) thanks to the centred forward pass
    np.testing.assert_allclose(Dg.grad.cpu().numpy(), E, rtol=1e-4, atol=1e-5 * gout.numpy().max())


@pytest.mark.gpu
def test_cuda_full_matrix_R(cuda):
    """The optional R output (interior of the reference's padded R) and the saved softmin weights: they are the
    reference's backward coefficients a, b, c (:100-103) evaluated from the fp64 R."""
    from dae.soft_dtw_cuda import softdtw_forward
    D = torch.rand(2, 70, 90, generator=torch.Generator().manual_seed(1))
    for gamma, bw in ((1.0, 0.0), (0.3, 0.0), (1.0, 12.0)):
        val, W, R = softdtw_forward(D.to(cuda), gamma, bw, want_R=True)
        Rp = so.forward(D.numpy(), gamma, bw)
        ref = Rp[:, 1:-1, 1:-1]
        got = R.cpu().numpy()
        assert (np.isinf(got) == np.isinf(ref)).all()
        fin = np.isfinite(ref)
        np.testing.assert_allclose(got[fin], ref[fin], rtol=1e-6, atol=1e-5)
        np.testing.assert_allclose(val.cpu().numpy(), ref[:, -1, -1], rtol=1e-6, atol=1e-5)
        with np.errstate(invalid="ignore", over="ignore"):
            w_up = np.exp((Rp[:, 1:-1, 1:-1] - D.numpy() - Rp[:, :-2, 1:-1]) / gamma)      # weight of (i-1, j)
            w_left = np.exp((Rp[:, 1:-1, 1:-1] - D.numpy() - Rp[:, 1:-1, :-2]) / gamma)   # weight of (i, j-1)
        Wc = W.cpu().numpy()
        np.testing.assert_allclose(Wc[..., 0][fin], np.nan_to_num(w_up)[fin], rtol=0, atol=1e-5)
        np.testing.assert_allclose(Wc[..., 1][fin], np.nan_to_num(w_left)[fin], rtol=0, atol=1e-5)
        assert (Wc[~fin] == 0).all()


@pytest.mark.gpu
def test_cuda_large_properties(cuda):
    """One BASELINE-sized sample (all 8 of cfg4 are checked in test_fullsize_gpu.py): value and gradient vs the
    fp64 oracle at the north-star tolerance."""
    from dae.soft_dtw_cuda import SoftDTW, _SoftDTWCUDA
    g = torch.Generator().manual_seed(1234)
    a, b = torch.rand(1, 4096, 2, generator=g), torch.rand(1, 4096, 2, generator=g)
    D = SoftDTW._euclidean_dist_func(a, b)
    Dg = D.to(cuda).requires_grad_()
    v = _SoftDTWCUDA.apply(Dg, 1.0, 0.0)
    v.sum().backward()
    R = so.forward(D.numpy(), 1.0, 0.0)
    assert abs(v.item() - R[0, -2, -2]) <= 1e-6 * abs(R[0, -2, -2])
    E = so.backward(D.numpy(), R, 1.0, 0.0)
    # north_star: within 1e-4 (of the gradient's scale, E <= 1).  The reference itself stores R ~ -5000 in fp32
    # (:144,256-258) and sits 1.2e-4 from this fp64 oracle; the centred forward + saved weights are ~1e-6.
    got = Dg.grad.cpu().numpy()
    assert np.abs(got - E).max() <= 1e-5
    big = E > 1e-3
    assert (np.abs(got - E)[big] / E[big]).max() <= 1e-4
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


@pytest.mark.gpu
def test_teacher_student_softdtw_loss_matches_oracle(cuda):
    """The soft-DTW consumer sketched at wav2vec2/lib.py:184-191: loss value and the gradient reaching the student
    posteriors against the fp64 oracle (E from softdtw_oracle.c, chained through the distance in torch fp64)."""
    from dae.softdtw_loss import teacher_student_softdtw_loss
    g = torch.Generator().manual_seed(17)
    post = (torch.randn(3, 90, 24, generator=g) * 1.5).log_softmax(-1)            # 2 students + teacher
    x = post.to(cuda).requires_grad_()
    loss = teacher_student_softdtw_loss(x, gamma=1.5)
    loss.backward()
    xd = post.double().requires_grad_()
    teacher, students = xd[-1].detach().unsqueeze(0).repeat(2, 1, 1), xd[:2]
    D = ((teacher[:, :, None, :] - students[:, None, :, :]) ** 2).sum(-1)          # soft_dtw_cuda.py:319-329
    R = so.forward(D.detach().numpy(), 1.5, 0.0)
    E = so.backward(D.detach().numpy(), R, 1.5, 0.0)
    ref = R[:, -2, -2].mean()
    (torch.from_numpy(E) * D).sum().div(2).backward()                              # d loss / d students via E
    assert abs(loss.item() - ref) <= 1e-5 * abs(ref)
    assert x.grad[-1].abs().max().item() == 0                                      # the teacher row is detached
    np.testing.assert_allclose(x.grad[:2].cpu().numpy(), xd.grad[:2].numpy(), rtol=1e-4,
                               atol=1e-5 * float(xd.grad.abs().max()))
