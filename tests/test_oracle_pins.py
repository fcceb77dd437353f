"""Pin the CPU oracles against the reference's own dependencies / golden vectors (no GPU)."""
import numpy as np
import pytest
import torch

from oracle import ctc_oracle, greedy_oracle, specaug_oracle, stitch_oracle


@pytest.mark.parametrize("seed,T,N,C,Lmax", [(0, 40, 2, 7, 9), (1, 25, 3, 5, 6), (2, 60, 1, 33, 20)])
def test_ctc_oracle_vs_torch_cpu(seed, T, N, C, Lmax):
    g = torch.Generator().manual_seed(seed)
    blank = C - 1
    lp = (torch.randn(T, N, C, generator=g, dtype=torch.float64) * 2).log_softmax(-1).requires_grad_()
    tg = torch.randint(0, blank, (N, Lmax), generator=g)
    tg[0, 1] = tg[0, 0]                                  # force a repeated label
    il = torch.tensor([T] + [int(x) for x in torch.randint(T // 2 + Lmax, T + 1, (N - 1,), generator=g)])
    tl = torch.tensor([Lmax] + [int(x) for x in torch.randint(0, Lmax + 1, (N - 1,), generator=g)])
    loss = torch.nn.CTCLoss(blank=blank, reduction="sum")(lp, tg, il, tl)
    (loss * 0.5).backward()
    nll, grad = ctc_oracle.ctc_loss_grad(lp.detach().numpy(), tg.numpy(), il.numpy(), tl.numpy(), blank, gout=0.5)
    assert abs(nll.sum() - loss.item()) < 1e-10 * max(1.0, abs(loss.item()))
    np.testing.assert_allclose(grad, lp.grad.numpy(), atol=1e-12)
    # rows of torch's "gradient" sum to zero (it is the gradient w.r.t. the logits)
    assert np.abs(grad.sum(-1)).max() < 1e-12
    # frames past input_lengths get exactly zero
    for n in range(N):
        assert np.all(grad[int(il[n]):, n] == 0)


def test_ctc_oracle_vs_torch_fp32():
    g = torch.Generator().manual_seed(3)
    T, N, C, L = 120, 2, 50, 30
    lp = (torch.randn(T, N, C, generator=g) * 3).log_softmax(-1).requires_grad_()
    tg = torch.randint(0, C - 1, (N, L), generator=g)
    il, tl = torch.tensor([T, T - 7]), torch.tensor([L, L - 11])
    loss = torch.nn.CTCLoss(blank=C - 1, reduction="sum")(lp, tg, il, tl)
    loss.backward()
    nll, grad = ctc_oracle.ctc_loss_grad(lp.detach().numpy(), tg.numpy(), il.numpy(), tl.numpy(), C - 1)
    assert abs(nll.sum() - loss.item()) < 1e-4 * abs(loss.item())
    # torch's fp32 log-space lattice is itself only good to ~1e-3 here (|log-lik| ~ 500 -> ulp 6e-5 per
    # step, accumulated over T steps); the fp64 oracle is the truth, torch fp32 a sanity bound.
    np.testing.assert_allclose(grad, lp.grad.numpy(), atol=1e-3)


def test_ctc_oracle_infeasible_is_inf():
    lp = torch.randn(3, 1, 5).log_softmax(-1).numpy()
    nll, _ = ctc_oracle.ctc_loss_grad(lp, np.array([[1, 1, 1, 1]]), [3], [4], 4)
    assert np.isinf(nll[0]) and nll[0] > 0


def test_greedy_oracle_vs_torch():
    g = torch.Generator().manual_seed(0)
    lp = torch.randn(300, 17, generator=g)
    lp[10:20] = lp[10]                                   # a run of identical frames
    lp[30, 3] = lp[30, 9] = lp[30].max() + 1             # exact tie -> lower index
    lp[40, 5] = float("nan")                             # NaN counts as max in torch.argmax
    lp[50] = float("-inf")                               # all -inf -> index 0
    path = greedy_oracle.argmax_rows(lp.numpy())
    np.testing.assert_array_equal(path, lp.argmax(-1).numpy())
    blank = 16
    ref = [int(i) for i in torch.unique_consecutive(lp.argmax(-1)) if i != blank]
    assert greedy_oracle.collapse(path, blank) == ref
    assert greedy_oracle.collapse(np.array([], dtype=np.int32), blank) == []


def test_specaug_oracle_vs_torchaudio():
    ta = pytest.importorskip("torchaudio")
    from dae.augment import draw_bands
    F, T, param, n = 80, 400, 34, 3
    x = torch.randn(1, 1, F, T)
    fill = float(x.mean())
    torch.manual_seed(7)
    y = x.clone()
    for _ in range(n):                                   # torchaudio's own band draw + fill, iid per item
        y = ta.functional.mask_along_axis_iid(y, param, fill, 2)
    torch.manual_seed(7)
    bands = draw_bands(1, n, param, F)
    out = specaug_oracle.apply_bands(x[0, 0].numpy(), bands[0].tolist(), [], np.float32(fill))
    np.testing.assert_array_equal(out, y[0, 0].numpy())


def test_specaug_oracle_mean_and_repeat():
    x = np.random.default_rng(0).standard_normal((8, 64)).astype(np.float32)
    out, fill = specaug_oracle.specaug_repeat(x, [[(1, 3)]], [[(10, 20)]], False, 1)
    assert out.shape == (2, 8, 64)
    np.testing.assert_array_equal(out[1], x)
    assert np.all(out[0, 1:3] == fill) and np.all(out[0, :, 10:20] == fill)
    assert abs(float(fill) - float(torch.from_numpy(x).mean())) <= 2e-7 * max(1.0, abs(float(fill)))
    np.testing.assert_array_equal(out[0, 0, :10], x[0, :10])


def _ds(n):  # three k3/s2/p1 convs: T' = floor((L-1)/2)+1 thrice
    for _ in range(3):
        n = (n - 1) // 2 + 1
    return n


@pytest.mark.parametrize("spec_n,n_win,last,last_ds", [
    (6000, 1, (0, 6000), 750), (120000, 52, (104448, 15552), 1944),
    (360000, 169, (344064, 15936), 1992), (415990, 197, (401408, 14582), 1823)])
def test_chunking_golden_vectors(spec_n, n_win, last, last_ds):
    """Golden index vectors computed from the reference's prepare_chunks."""
    ch = stitch_oracle.prepare_chunks(spec_n, 16384, 14336)
    assert len(ch) == n_win and ch[-1] == last and _ds(ch[-1][1]) == last_ds
    if n_win > 1:
        assert _ds(16384) == 2048 and int(14336 / (16384 / 2048)) == 1792


def test_prepare_chunks_matches_reference_semantics():
    # restated reference loop on a real tensor (views), incl. the "one window after the first shorter one" rule
    def ref(spec, seq_len, overlap):
        spec_n = spec.shape[-1]
        last_ulen, kill_next = None, False
        if spec_n <= seq_len:
            return [(0, spec_n)]
        out = []
        for i in range(0, spec_n, seq_len - overlap):
            chunk = spec[:, :, i:i + seq_len]
            u_len = chunk.shape[-1]
            if kill_next:
                break
            elif last_ulen is not None and u_len < last_ulen:
                kill_next = True
            last_ulen = u_len
            out.append((i, u_len))
        return out
    for spec_n, seq, ov in [(1000, 256, 192), (1000, 256, 0), (255, 256, 128), (777, 128, 64), (4096, 1024, 896)]:
        assert stitch_oracle.prepare_chunks(spec_n, seq, ov) == ref(torch.zeros(1, 2, spec_n), seq, ov)


def test_stitch_last_window_shift_quirk():
    # overlap_ds becomes 1793 for a short last window with r/k >= 0.00446
    u_len, k = 8 * 1500 - 7, 1500
    assert int(14336 / (u_len / k)) == 1793
    pos = stitch_oracle.window_positions([0, 2048], [16384, u_len], [2048, k], 14336)
    assert pos == [0, 2048 - 1793]


def test_stitch_oracle_small():
    rng = np.random.default_rng(0)
    C = 5
    wins = [np.log(rng.dirichlet(np.ones(C), size=8)).astype(np.float32) for _ in range(3)]
    out = stitch_oracle.stitch(wins, [0, 16, 32], [64, 64, 64], 48, buf_rows=64)
    # ratio 8, overlap_ds 6: positions 0, 2, 4 -> 12 covered rows
    assert out.shape == (12, C)
    np.testing.assert_allclose(out[0], wins[0][0], rtol=1e-6)
    exp_row5 = np.log((np.exp(wins[0][5]) + np.exp(wins[1][3]) + np.exp(wins[2][1])) / 3)
    np.testing.assert_allclose(out[5], exp_row5, rtol=1e-6)


@pytest.mark.parametrize("mode", ["mean", "mean_recording", "zero"])
def test_cutout_oracle_matches_reference(mode):
    """Rectangle draw order + fill semantics vs the reference's own cutout() (golden, lib.py:384-417)."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from toy import toy_spec
    from dae.augment import draw_cutout_rects
    from oracle import cutout_oracle
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loop_toy.npz"))[f"cutout_{mode}"]
    spec = toy_spec(3, 700)[0].numpy()
    torch.manual_seed(77)
    rects = draw_cutout_rects(700, 80, 512, num_rectangles=40, max_width=60, max_height=12).tolist()
    assert len(rects) == int(40 * 700 / 512)
    out = cutout_oracle.cutout(spec, rects, mode)
    np.testing.assert_allclose(out, gold, rtol=0, atol=1e-6)        # means: torch fp32 sum vs numpy pairwise
    assert (out != spec).any()


def _toy_gold():
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from toy import toy_spec
    return toy_spec, np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loop_toy.npz"))


@pytest.mark.parametrize("td,fd", [(True, False), (False, True), (True, True)])
def test_frame_shuffle_oracle_matches_reference(td, fd):
    """Permutation draw order (time, then frequency) + gather vs the reference's own frame_shuffle() (lib.py:81-84)."""
    from dae.augment import draw_frame_shuffle
    from oracle import augment_extra_oracle as ax
    toy_spec, gold = _toy_gold()
    spec = toy_spec(4, 200).numpy()
    torch.manual_seed(78)
    pt, pf = draw_frame_shuffle(80, 200, td, fd)
    out = ax.frame_shuffle(spec, None if pt is None else pt.numpy(), None if pf is None else pf.numpy())
    np.testing.assert_array_equal(out[0], gold[f"frame_shuffle_t{int(td)}f{int(fd)}"])


def test_add_random_noise_oracle_matches_reference():
    """normal(0, std, size) == randn(size) * std for the same generator state, and the restated std/scale/add vs
    the reference's own add_random_noise() (lib.py:379-382)."""
    from oracle import augment_extra_oracle as ax
    toy_spec, gold = _toy_gold()
    spec = toy_spec(4, 200)
    torch.manual_seed(79)
    a = torch.normal(0, std=spec.std(), size=spec.shape)
    torch.manual_seed(79)
    z = torch.randn(spec.shape)
    assert torch.equal(a, z * spec.std())
    out = ax.add_random_noise(spec.numpy(), z.numpy(), 0.3)
    np.testing.assert_allclose(out[0], gold["add_random_noise_0p3"], rtol=0, atol=5e-7)
    assert np.abs(out - spec.numpy()).max() > 0.1


@pytest.mark.parametrize("T,C,L,K", [(37, 11, 9, 8), (16, 6, 7, 8), (8, 5, 0, 8), (5, 5, 2, 8), (20, 4, 9, 4), (19, 7, 5, 16)])
def test_blocked_ctc_decomposition_equals_per_frame_recursion(T, C, L, K):
    """The time-blocked factorisation the GPU path uses (transfer bands, boundary scan in both directions, block
    fill) is the same function as the per-frame recursion: loss and gradient agree with ctc_oracle in fp64."""
    from oracle import ctc_blocked_oracle, ctc_oracle
    rng = np.random.default_rng(T * 100 + L)
    lp = np.log(rng.dirichlet(np.ones(C), size=T))
    labels = list(rng.integers(0, C - 1, size=L))
    if L > 3:
        labels[2] = labels[1]                                  # a repeated label: no skip transition there
    nll_b, grad_b = ctc_blocked_oracle.ctc_loss_grad_blocked(lp, labels, C - 1, K)
    nll, grad = ctc_oracle.ctc_loss_grad(lp[:, None, :], [labels], [T], [L], C - 1)
    assert abs(nll_b - nll[0]) <= 1e-10 * max(1.0, abs(nll[0]))
    np.testing.assert_allclose(grad_b, grad[:, 0], rtol=0, atol=1e-10)
