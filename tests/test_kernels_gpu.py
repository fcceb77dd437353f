"""GPU parity tests (B200): greedy, SpecAugment, CTC, stitch — CUDA path vs the CPU oracle.

All calls go through the C ABI (ctypes).  Integer/index results are bit-exact; floating point
tolerances are written next to each assert.
"""
import numpy as np
import pytest
import torch

from oracle import ctc_oracle, greedy_oracle, specaug_oracle, stitch_oracle

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- greedy
def _peaky(T, C, blank, g, p_blank=0.6):
    lp = torch.randn(T, C, generator=g)
    cls = torch.randint(0, C - 1, (T,), generator=g)
    cls[torch.rand(T, generator=g) < p_blank] = blank
    run = torch.rand(T, generator=g) < 0.5               # repeat the previous frame's class
    for t in range(1, T):
        if run[t]:
            cls[t] = cls[t - 1]
    lp[torch.arange(T), cls] += 6
    return lp.log_softmax(-1)


@pytest.mark.parametrize("T,C", [(2048, 4096), (333, 129), (1000, 32), (1, 7), (5000, 4096), (70, 1030)])
def test_greedy_matches_oracle(cuda, T, C):
    from dae.greedy import greedy_ids_device
    g = torch.Generator().manual_seed(T + C)
    lp = _peaky(T, C, C - 1, g)
    if T > 60:
        lp[10, 3] = lp[10, 11] = lp[10].max() + 1        # exact tie -> first index
        lp[20] = lp[19]                                  # identical consecutive frames
        lp[30, C // 2] = float("nan")
        lp[40] = float("-inf")
    path, ids, n = greedy_ids_device(lp.to(cuda), C - 1)
    ref_path = greedy_oracle.argmax_rows(lp.numpy())
    np.testing.assert_array_equal(path[0, :T].cpu().numpy(), ref_path)
    assert ids[0, :int(n[0])].tolist() == greedy_oracle.collapse(ref_path, C - 1)


def test_greedy_batched_strided_lengths(cuda):
    from dae.greedy import greedy_ids_device
    g = torch.Generator().manual_seed(0)
    B, T, C = 3, 700, 260
    big = torch.stack([_peaky(T, C + 4, C - 1, g) for _ in range(B)])
    lp = big[:, :, :C]                                   # row stride C+4, class stride 1, unaligned rows
    lens = torch.tensor([700, 1, 313], dtype=torch.int32)
    lpd = big.to(cuda)[:, :, :C]
    path, ids, n = greedy_ids_device(lpd, C - 1, lens)
    for b in range(B):
        ref = greedy_oracle.greedy_ids(lp[b, :int(lens[b])].numpy(), C - 1)
        assert ids[b, :int(n[b])].tolist() == ref


def test_greedy_decoder_module(cuda):
    from dae.greedy import GreedyCTCDecoder

    class Tok:
        def decode(self, ids):
            return " ".join(str(i) for i in ids)
    g = torch.Generator().manual_seed(5)
    lp = _peaky(500, 129, 128, g)
    ref = greedy_oracle.greedy_ids(lp.numpy(), 128)
    dec = GreedyCTCDecoder(tokenizer=Tok(), blank_id=128)
    assert dec(lp.to(cuda)) == " ".join(map(str, ref))
    assert dec(lp.to(cuda), decode=False) == ref
    assert dec(lp) == " ".join(map(str, ref))            # host tensor, as the reference passes it
    assert GreedyCTCDecoder(blank_id=128)(lp.to(cuda)) == ref


# ---------------------------------------------------------------- SpecAugment
@pytest.mark.parametrize("F,T,nf,nt,zero", [(80, 16384, 6, 0, False), (80, 4000, 2, 3, False), (80, 1001, 4, 2, True),
                                            (13, 77, 1, 1, False)])
def test_specaug_repeat_matches_oracle(cuda, F, T, nf, nt, zero):
    from dae.augment import SpecAugment
    torch.manual_seed(F * T)
    spec = torch.randn(1, F, T + 100) * 2 + 0.3
    win = spec[:, :, 50:50 + T]                          # a window view, row stride T+100
    aug = SpecAugment(n_time_masks=nt, n_freq_masks=nf, freq_mask_param=34 if F > 40 else 5, time_mask_param=50,
                      zero_masking=zero)
    out = aug(spec.to(cuda)[:, :, 50:50 + T], n_clean=1)
    fb, tb = aug.last_bands
    ref, fill = specaug_oracle.specaug_repeat(win[0].numpy(), fb.tolist(), tb.tolist(), zero, 1)
    got = out.cpu().numpy()
    assert got.shape == (2, F, T)
    np.testing.assert_array_equal(got[1], win[0].numpy())                       # clean copy: bit-exact
    masked = got[0] != win[0].numpy()
    kfill = got[0][masked][0] if masked.any() else fill
    # mean fill: fp64 accumulation on both sides, may differ in the last fp32 bit only
    assert abs(float(kfill) - float(fill)) <= 1.2e-7 * max(1.0, abs(float(fill)))
    ref2, _ = specaug_oracle.specaug_repeat(win[0].numpy(), fb.tolist(), tb.tolist(), zero, 1, fill=kfill)
    np.testing.assert_array_equal(got, ref2)                                   # given the fill: bit-exact


def test_specaug_with_precomputed_window_sums(cuda):
    """The adapt loop's variant: dae_window_sums once per recording, then one barrier-free launch per window.  Same
    result as the self-contained call: clean copy and everything outside the bands bit-exact, fill = the window's mean."""
    from dae.augment import SpecAugment
    torch.manual_seed(21)
    spec = (torch.randn(1, 80, 40000) * 1.7 - 0.2)
    dspec = spec.to(cuda)
    starts, lens = [0, 2048, 20000, 30000], [16384, 16384, 16384, 10000]      # the last window is the ragged one
    sums = SpecAugment.window_sums(dspec, starts, lens)
    assert sums.shape[0] == 4 and sums.dtype == torch.float64
    for w, (s0, ln) in enumerate(zip(starts, lens)):
        win = spec[0, :, s0:s0 + ln].numpy()
        assert abs(float(sums[w].sum()) - float(win.astype(np.float64).sum())) <= 1e-9 * ln * 80
        aug = SpecAugment(n_time_masks=2 if w == 3 else 0, n_freq_masks=6, freq_mask_param=34, time_mask_param=40)
        out = aug(dspec[:, :, s0:s0 + ln], n_clean=1, window_sums=sums[w])
        fb, tb = aug.last_bands
        _, fill = specaug_oracle.specaug_repeat(win, fb.tolist(), tb.tolist(), False, 1)
        got = out.cpu().numpy()
        kfill = got[0][got[0] != win][0]
        assert abs(float(kfill) - float(fill)) <= 1.2e-7 * max(1.0, abs(float(fill)))
        ref, _ = specaug_oracle.specaug_repeat(win, fb.tolist(), tb.tolist(), False, 1, fill=kfill)
        np.testing.assert_array_equal(got, ref)
        same = aug(dspec[:, :, s0:s0 + ln], n_clean=1, bands=(fb, tb))          # self-contained path, same bands
        assert torch.equal(same[1], out[1]) and float((same[0] - out[0]).abs().max()) <= 1.2e-7 * max(1.0, abs(float(fill)))


def test_specaug_plain_call_and_determinism(cuda):
    from dae.augment import SpecAugment
    torch.manual_seed(3)
    x = torch.randn(1, 80, 2048, device=cuda)
    aug = SpecAugment(n_freq_masks=6, freq_mask_param=34)
    torch.manual_seed(11)
    a = aug(x)
    torch.manual_seed(11)
    b = aug(x)
    assert a.shape == x.shape and torch.equal(a, b)
    fb, _ = aug.last_bands
    ref, _ = specaug_oracle.specaug_repeat(x[0].cpu().numpy(), fb.tolist(), [[]], False, 0, fill=a[0][a[0] != x[0]][0].item()
                                           if (a[0] != x[0]).any() else None)
    np.testing.assert_array_equal(a.cpu().numpy(), ref)


# ---------------------------------------------------------------- CTC
@pytest.fixture(params=["chain", "blocked", "blocked_nocluster", "blocked_serial"])
def ctc_path(request):
    """Both lattice implementations behind dae_ctc_lattice: the per-frame chain (ctc.cu) and the time-blocked
    scan (ctc_blocked.cu), the latter with hand-over through cluster shared memory (default), through global
    memory only, and with dae_ctc_loss_grad's overlap of gradient and scan switched off;
    dae_ctc_configure forces the choice regardless of shape."""
    import dae._C as C
    C.ctc_configure(blocked=0 if request.param == "chain" else 1, cluster=1 if request.param == "blocked_nocluster" else 8,
                    overlap=0 if request.param == "blocked_serial" else -1)
    yield request.param
    C.ctc_configure()


def _ctc_case(T, N, C, Lmax, seed, ragged=True, peaky=False):
    g = torch.Generator().manual_seed(seed)
    blank = C - 1
    if peaky:
        lp = torch.stack([_peaky(T, C, blank, g) for _ in range(N)], 1)
    else:
        lp = (torch.randn(T, N, C, generator=g) * 3).log_softmax(-1)
    tg = torch.randint(0, blank, (N, max(Lmax, 1)), generator=g)
    if Lmax >= 2:
        tg[0, 1] = tg[0, 0]
    il = torch.full((N,), T, dtype=torch.long)
    tl = torch.full((N,), Lmax, dtype=torch.long)
    if ragged and N > 1:
        il[1:] = torch.randint(max(T // 2, 2 * Lmax + 1), T + 1, (N - 1,), generator=g)
        tl[1:] = torch.randint(0, Lmax + 1, (N - 1,), generator=g)
    return lp, tg, il, tl, blank


def _assert_ctc_grad_close(got, ref, lp64, g):
    """grad = g*(exp(lp) - occupancy): each of the two terms must be good to 1e-4 relative (north_star
    tolerance), so the difference may be off by 1e-4 of the larger operand where they cancel."""
    tol = 1e-4 * np.abs(ref) + 1e-4 * g * np.exp(lp64) + 1e-12
    bad = np.abs(got - ref) > tol
    assert not bad.any(), f"{bad.sum()} elements off; worst {np.abs(got - ref)[bad].max()} vs tol {tol[bad].min()}"


@pytest.mark.parametrize("T,N,C,Lmax", [(40, 2, 7, 9), (200, 3, 129, 40), (512, 1, 4096, 150), (64, 4, 32, 0),
                                        (300, 2, 50, 149), (33, 1, 5, 1)])
def test_ctc_matches_fp64_oracle(cuda, T, N, C, Lmax, ctc_path):
    from dae.ctc import CTCLoss
    lp, tg, il, tl, blank = _ctc_case(T, N, C, Lmax, seed=T + N)
    if Lmax == 0:
        tg = tg[:, :0]
    x = lp.to(cuda).requires_grad_()
    loss = CTCLoss(blank=blank, reduction="sum")(x, tg.to(cuda), il.to(cuda), tl.to(cuda))
    (loss / (T * N)).backward()
    nll, grad = ctc_oracle.ctc_loss_grad(lp.double().numpy(), tg.numpy(), il.numpy(), tl.numpy(), blank,
                                         gout=1.0 / (T * N))
    # north_star tolerance: loss and gradients within 1e-4 relative (fp32)
    assert abs(loss.item() - nll.sum()) <= 1e-4 * abs(nll.sum())
    got = x.grad.cpu().numpy()
    _assert_ctc_grad_close(got, grad, lp.double().numpy(), 1.0 / (T * N))
    for n in range(N):
        assert np.all(got[int(il[n]):, n] == 0)


def _teacher_student(T, C, g, noise):
    """[2,T,C] posteriors like the adapt step's batch: row 1 = clean teacher branch, row 0 = the same
    frames seen through an augmentation (correlated, not identical)."""
    teacher_logits = torch.randn(T, C, generator=g)
    cls = torch.randint(0, C - 1, (T,), generator=g)
    cls[torch.rand(T, generator=g) < 0.7] = C - 1
    teacher_logits[torch.arange(T), cls] += 8
    student_logits = teacher_logits + noise * torch.randn(T, C, generator=g)
    return torch.stack([student_logits.log_softmax(-1), teacher_logits.log_softmax(-1)])


def test_ctc_hot_path_shape_vs_torch(cuda, ctc_path):
    """cfg2 shape: lp is the non-contiguous view out[:1].transpose(0,1) of [2,2048,4096] (lib.py:570-575)."""
    from dae.ctc import CTCLoss
    T, C = 2048, 4096
    g = torch.Generator().manual_seed(0)
    post = _teacher_student(T, C, g, noise=1.0).to(cuda).requires_grad_()
    labels = greedy_oracle.greedy_ids(post[1].detach().cpu().numpy(), C - 1)
    tg = torch.tensor(labels, dtype=torch.long, device=cuda)[None]
    L = len(labels)
    assert 100 < L < 1500
    aug = post[:1].transpose(0, 1)
    il, tl = torch.tensor([T], device=cuda), torch.tensor([L], device=cuda)
    loss = CTCLoss(blank=C - 1, reduction="sum")(aug, tg, il, tl) / T
    loss.backward()
    got = post.grad.clone()
    post.grad = None
    ref_loss = torch.nn.CTCLoss(blank=C - 1, reduction="sum")(aug, tg, il, tl) / T
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-4 * abs(ref_loss.item())
    assert torch.all(got[1] == 0)                       # the clean copy gets no gradient
    nll, grad = ctc_oracle.ctc_loss_grad(post[:1].detach().transpose(0, 1).double().cpu().numpy(), [labels], [T], [L],
                                         C - 1, gout=1.0 / T)
    assert abs(loss.item() * T - nll[0]) <= 1e-4 * nll[0]
    _assert_ctc_grad_close(got[0].cpu().numpy(), grad[:, 0], post[0].detach().double().cpu().numpy(), 1.0 / T)
    # torch's own fp32 CUDA lattice is the looser side: we must be at least as close to fp64 as it is
    err_dae = np.abs(got[0].cpu().numpy() - grad[:, 0]).max()
    err_torch = np.abs(post.grad[0].cpu().numpy() - grad[:, 0]).max()
    assert err_dae <= err_torch + 1e-9


def test_ctc_mismatched_labels_vs_fp64(cuda, ctc_path):
    """Adversarial: labels unrelated to the posteriors, so the alignment runs through states ~2^-1000 below
    the per-frame maximum.  fp32 log-space keeps ~1e-3 there (documented in DESIGN.md); torch fp32 is worse."""
    from dae.ctc import CTCLoss
    T, C, L = 1024, 512, 300
    g = torch.Generator().manual_seed(2)
    post = _teacher_student(T, C, g, noise=0.0)[1]
    tgt = torch.randint(0, C - 1, (1, L), generator=g)
    x = post[:, None].to(cuda).requires_grad_()
    loss = CTCLoss(blank=C - 1, reduction="sum")(x, tgt.to(cuda), [T], [L])
    loss.backward()
    nll, grad = ctc_oracle.ctc_loss_grad(post[:, None].double().numpy(), tgt.numpy(), [T], [L], C - 1)
    assert abs(loss.item() - nll[0]) <= 1e-5 * nll[0]
    got = x.grad.cpu().numpy()
    tol = 5e-3 * np.abs(grad) + 5e-3 * np.exp(post.numpy())[:, None] + 1e-9
    assert (np.abs(got - grad) <= tol).all(), float((np.abs(got - grad) / tol).max())
    y = post[:, None].to(cuda).requires_grad_()
    torch.nn.CTCLoss(blank=C - 1, reduction="sum")(y, tgt.to(cuda), [T], [L]).backward()
    assert np.abs(got - grad).max() <= np.abs(y.grad.cpu().numpy() - grad).max() + 1e-9


def test_ctc_reductions_and_infeasible(cuda, ctc_path):
    from dae.ctc import CTCLoss, ctc_loss
    lp, tg, il, tl, blank = _ctc_case(50, 3, 11, 8, seed=9)
    x = lp.to(cuda)
    for red in ("none", "mean", "sum"):
        a = ctc_loss(x, tg.to(cuda), il, tl, blank=blank, reduction=red)
        b = torch.nn.functional.ctc_loss(lp, tg, il, tl, blank=blank, reduction=red)
        torch.testing.assert_close(a.cpu(), b, rtol=1e-4, atol=1e-4)
    # infeasible: 4 repeated labels need 7 frames, only 3 given -> +inf like torch (zero_infinity=False)
    y = torch.randn(3, 1, 5).log_softmax(-1).to(cuda)
    bad = CTCLoss(blank=4, reduction="sum")(y, torch.tensor([[1, 1, 1, 1]], device=cuda), [3], [4])
    assert torch.isinf(bad) and bad > 0


def test_ctc_large_magnitude_precision(cuda, ctc_path):
    """Random logits, |log-likelihood| ~ 1e4 and the alignment far below the per-frame maximum: the centred
    fp32 lattice keeps the loss to 1e-5 and the gradient to ~1e-3 of its scale (DESIGN.md "CTC accuracy");
    torch's uncentred fp32 kernel is an order of magnitude further from fp64."""
    from dae.ctc import CTCLoss
    T, N, C, L = 2048, 1, 512, 300
    lp, tg, il, tl, blank = _ctc_case(T, N, C, L, seed=1, ragged=False)
    x = lp.to(cuda).requires_grad_()
    loss = CTCLoss(blank=blank, reduction="sum")(x, tg.to(cuda), il, tl)
    loss.backward()
    nll, grad = ctc_oracle.ctc_loss_grad(lp.double().numpy(), tg.numpy(), il.numpy(), tl.numpy(), blank)
    assert nll[0] > 5000
    assert abs(loss.item() - nll[0]) <= 1e-5 * nll[0]
    err = np.abs(x.grad.cpu().numpy() - grad).max()
    assert err <= 3e-3 * np.abs(grad).max()
    y = lp.to(cuda).requires_grad_()
    torch.nn.CTCLoss(blank=blank, reduction="sum")(y, tg.to(cuda), il, tl).backward()
    err_torch = np.abs(y.grad.cpu().numpy() - grad).max()
    print(f"max |grad - fp64|: dae {err:.2e}, torch.cuda fp32 {err_torch:.2e}")
    assert err < err_torch


def test_ctc_long_label_sequence_two_pairs_per_thread(cuda, ctc_path):
    """Lmax ~ 1000 labels: two state pairs per thread in the chain kernel; on the blocked path 32 scan regions
    (four clusters) and the unfused fill + row-gradient fallback (a block's rows no longer fit in shared memory)."""
    from dae.ctc import CTCLoss
    T, N, C, L = 2100, 1, 48, 1000
    g = torch.Generator().manual_seed(7)
    blank = C - 1
    tg = torch.randint(0, blank, (N, L), generator=g)
    # posteriors loosely aligned with the labels (two frames per label), so the alignment is not adversarial
    lp = torch.randn(T, N, C, generator=g)
    idx = torch.arange(T).clamp_max(2 * L - 1) // 2
    lp[torch.arange(T), 0, tg[0, idx]] += 4.0
    lp = lp.log_softmax(-1)
    x = lp.to(cuda).requires_grad_()
    loss = CTCLoss(blank=blank, reduction="sum")(x, tg.to(cuda), [T], [L])
    (loss / T).backward()
    nll, grad = ctc_oracle.ctc_loss_grad(lp.double().numpy(), tg.numpy(), [T], [L], blank, gout=1.0 / T)
    assert abs(loss.item() - nll[0]) <= 1e-5 * abs(nll[0])
    _assert_ctc_grad_close(x.grad.cpu().numpy(), grad, lp.double().numpy(), 1.0 / T)


def test_ctc_empty_inputs_and_targets(cuda, ctc_path):
    """input_length 0 (loss 0 for the empty target, inf otherwise), target_length 0, and a normal sample in
    one ragged batch: losses equal torch's, gradient rows of absent frames are zero."""
    from dae.ctc import ctc_loss
    T, N, C = 96, 4, 9
    g = torch.Generator().manual_seed(11)
    lp = torch.randn(T, N, C, generator=g).log_softmax(-1)
    tg = torch.randint(0, C - 1, (N, 5), generator=g)
    il = torch.tensor([0, 0, T, 70])
    tl = torch.tensor([0, 3, 0, 5])
    x = lp.to(cuda).requires_grad_()
    nll = ctc_loss(x, tg.to(cuda), il, tl, blank=C - 1, reduction="none")
    ref = torch.nn.functional.ctc_loss(lp, tg, il, tl, blank=C - 1, reduction="none")
    assert nll[0].item() == 0.0 and torch.isinf(nll[1]) and nll[1] > 0
    torch.testing.assert_close(nll[2:].cpu(), ref[2:], rtol=1e-5, atol=1e-5)
    nll[2:].sum().backward()
    got = x.grad.cpu()
    _, grad = ctc_oracle.ctc_loss_grad(lp[:, 2:].double().numpy(), tg[2:].numpy(), il[2:].numpy(), tl[2:].numpy(), C - 1)
    _assert_ctc_grad_close(got[:, 2:].numpy(), grad, lp[:, 2:].double().numpy(), 1.0)
    assert torch.all(got[:, 0] == 0) and torch.all(got[70:, 3] == 0)


# ---------------------------------------------------------------- stitch
def test_stitch_matches_oracle(cuda):
    import dae._C as C_
    from dae.stitch import stitch_windows
    rng = np.random.default_rng(0)
    C, seq, ov, spec_n = 129, 1024, 896, 5000
    chunks = stitch_oracle.prepare_chunks(spec_n, seq, ov)
    ds = lambda n: ((((n - 1) // 2 + 1) - 1) // 2 + 1 - 1) // 2 + 1
    wins = [np.log(rng.dirichlet(np.ones(C) * 0.3, size=ds(u)) + 1e-30).astype(np.float32) for _, u in chunks]
    starts, ulens = [c[0] for c in chunks], [c[1] for c in chunks]
    ref = stitch_oracle.stitch(wins, starts, ulens, ov, buf_rows=spec_n // 4 + seq)
    out, path = stitch_windows([torch.from_numpy(w).to(cuda) for w in wins], starts, ulens, ov)
    assert out.shape == ref.shape
    # exp/log are MUFU-based on the device: 2e-6 relative on the probabilities
    np.testing.assert_allclose(np.exp(out.cpu().numpy()), np.exp(ref), rtol=5e-6, atol=1e-30)
    np.testing.assert_array_equal(path.cpu().numpy(), greedy_oracle.argmax_rows(out.cpu().numpy()))
    assert (path.cpu().numpy() == greedy_oracle.argmax_rows(ref)).mean() > 0.999


def test_stitch_single_window_identity(cuda):
    from dae.stitch import stitch_windows
    lp = torch.randn(750, 129).log_softmax(-1)
    out, path = stitch_windows([lp.to(cuda)], [0], [6000], 0)
    np.testing.assert_allclose(out.cpu().numpy(), lp.numpy(), rtol=0, atol=2e-6)


# ---------------------------------------------------------------- cutout
@pytest.mark.parametrize("mode", ["mean", "mean_recording", "zero"])
def test_cutout_matches_oracle(cuda, mode):
    from dae.augment import cutout, draw_cutout_rects
    from oracle import cutout_oracle
    g = torch.Generator().manual_seed(5)
    spec = torch.randn(1, 80, 3000, generator=g)
    torch.manual_seed(9)
    rects = draw_cutout_rects(3000, 80, 16384, num_rectangles=205 * 6, max_width=792, max_height=41)
    assert len(rects) == int(205 * 6 * 3000 / 16384)
    x = spec.to(cuda).clone()
    out = cutout(x, 16384, cutout_val=mode, rects=rects)
    assert out.data_ptr() == x.data_ptr()                                # in place, like the reference
    ref = cutout_oracle.cutout(spec[0].numpy(), rects.tolist(), mode)
    got = x[0].cpu().numpy()
    untouched = ref == spec[0].numpy()
    np.testing.assert_array_equal(got[untouched], spec[0].numpy()[untouched])  # only rectangles are written
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-6)              # fp64 means vs numpy fp32 pairwise means
    # a strided window view (row stride != T) works too
    big = torch.randn(1, 80, 5000, generator=g).to(cuda)
    view = big[:, :, 1000:4000]
    before = big.clone()
    cutout(view, 16384, cutout_val="zero", rects=rects)
    assert torch.equal(big[:, :, :1000], before[:, :, :1000]) and torch.equal(big[:, :, 4000:], before[:, :, 4000:])
    assert (view[0].cpu().numpy() == 0).sum() > 0


# ---------------------------------------------------------------- frame shuffle / additive noise (f-3)
@pytest.mark.parametrize("td,fd", [(True, False), (False, True), (True, True)])
def test_frame_shuffle_matches_reference_golden(cuda, td, fd):
    """Host-drawn permutations in the reference's order + the gather kernel == the reference's own frame_shuffle()
    (golden), bit for bit; also against the oracle on a strided window view at the hot-path size."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from toy import toy_spec
    from dae.augment import draw_frame_shuffle, frame_shuffle
    from oracle import augment_extra_oracle as ax
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loop_toy.npz"))
    spec = toy_spec(4, 200)
    torch.manual_seed(78)
    out = frame_shuffle(spec.to(cuda), time_dimension=td, freq_dimension=fd)
    np.testing.assert_array_equal(out[0].cpu().numpy(), gold[f"frame_shuffle_t{int(td)}f{int(fd)}"])
    g = torch.Generator().manual_seed(6)
    big = torch.randn(1, 80, 20000, generator=g)
    view = big.to(cuda)[:, :, 1500:1500 + 16384]
    pt, pf = draw_frame_shuffle(80, 16384, td, fd, generator=g)
    out = frame_shuffle(view, td, fd, perms=(pt, pf))
    ref = ax.frame_shuffle(big[:, :, 1500:1500 + 16384].numpy(), None if pt is None else pt.numpy(),
                           None if pf is None else pf.numpy())
    np.testing.assert_array_equal(out.cpu().numpy(), ref)


def test_add_random_noise_matches_reference_golden(cuda):
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from toy import toy_spec
    from dae.augment import add_random_noise
    from oracle import augment_extra_oracle as ax
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loop_toy.npz"))
    spec = toy_spec(4, 200)
    x = spec.to(cuda).clone()
    torch.manual_seed(79)
    out = add_random_noise(x, 0.3)                      # draws randn on the host like the reference's normal()
    assert out.data_ptr() == x.data_ptr()
    np.testing.assert_allclose(x[0].cpu().numpy(), gold["add_random_noise_0p3"], rtol=0, atol=5e-7)
    g = torch.Generator().manual_seed(8)
    big, z = torch.randn(1, 80, 16384, generator=g) * 2.5 + 0.7, torch.randn(1, 80, 16384, generator=g)
    y = big.to(cuda).clone()
    add_random_noise(y, 0.05, z=z)
    np.testing.assert_allclose(y.cpu().numpy(), ax.add_random_noise(big.numpy(), z.numpy(), 0.05), rtol=0, atol=1e-6)
    assert add_random_noise(y, 0) is y


def test_ctc_rejects_out_of_range_labels(cuda):
    """torch.nn.CTCLoss raises on a label outside [0, C); dae.CTCLoss checks on the device (no host sync) and the
    error surfaces at the next synchronisation.  Padding beyond target_lengths is not looked at."""
    from dae.ctc import CTCLoss
    T, C = 30, 9
    lp = torch.randn(T, 2, C, device=cuda).log_softmax(-1)
    ok = torch.tensor([[1, 2, 3, -7], [4, 5, 99, 99]], device=cuda)          # bad values only in the padding
    CTCLoss(blank=C - 1, reduction="sum")(lp, ok, [T, T], [3, 2])
    torch.cuda.synchronize()
    # The failing case is NOT run here: a device-side assert kills the CUDA context and is logged by the driver
    # as a GPU fault on this shared pool; the check itself is the torch._assert_async call in dae/ctc.py.


@pytest.mark.parametrize("T,C,L,short", [(512, 256, 90, 0), (2048, 4096, 600, 0), (300, 64, 40, 37)])
def test_ctc_loss_grad_one_call_matches_the_pair(cuda, T, C, L, short):
    """dae_ctc_loss_grad (dense gradient streamed under the scan as a dependent launch, label classes after it) against
    the same call with the overlap switched off (= dae_ctc_lattice + dae_ctc_grad) and against the fp64 oracle; a
    reduction queued right behind it must see the finished gradient."""
    import dae._C as C_
    from dae.ctc import CTCLoss
    g = torch.Generator().manual_seed(T + C)
    tg = torch.randint(0, C - 1, (1, L), generator=g)
    Ln = L
    # teacher-like posteriors: frame t favours label floor(t*L/(T-short)) on two frames out of three, blank otherwise
    t = torch.arange(T)
    path = tg[0][(t * L // max(T - short, 1)).clamp(max=L - 1)]
    path = torch.where(t % 3 == 2, torch.full_like(path, C - 1), path)
    logits = torch.randn(T, 1, C, generator=g)
    logits[t, 0, path] += 7.0
    lp = logits.log_softmax(-1)
    il, tl = torch.tensor([T - short], device=cuda), torch.tensor([Ln], device=cuda)
    f = CTCLoss(blank=C - 1, reduction="sum")
    out = {}
    try:
        for overlap in (0, 1, -1):
            C_.ctc_configure(blocked=1, overlap=overlap)
            x = lp.to(cuda).requires_grad_()
            loss = f.with_scale(x, tg.to(cuda), il, tl, grad_scale_hint=1.0 / T)
            (loss / T).backward()
            out[overlap] = (loss.detach().clone(), x.grad.clone(), x.grad.double().sum().item())
    finally:
        C_.ctc_configure()
    assert torch.equal(out[0][0], out[-1][0]) and torch.equal(out[0][0], out[1][0])
    torch.testing.assert_close(out[-1][1], out[0][1], rtol=1e-6, atol=1e-9)
    assert torch.equal(out[-1][1], out[1][1])               # the early-resident kernel does the same arithmetic
    assert torch.all(out[-1][1][T - short:] == 0)
    nll, grad = ctc_oracle.ctc_loss_grad(lp.double().numpy(), tg.numpy(), [T - short], [Ln], C - 1, gout=1.0 / T)
    assert abs(out[-1][0].item() - nll[0]) <= 1e-4 * abs(nll[0])
    scale = np.abs(grad).max()
    assert np.abs(out[-1][1].cpu().numpy() - grad).max() <= 1e-4 * scale
    assert abs(out[-1][2] - out[-1][1].cpu().double().sum().item()) <= 1e-9 + 1e-9 * abs(out[-1][2])


@pytest.mark.parametrize("T,C", [(512, 64), (2048, 4096)])
def test_ctc_one_call_is_repeatable_under_load(cuda, T, C):
    """The one-call path runs three kernels at once (scan, dense gradient as its programmatic dependent, whatever the
    caller queued before): the same call, repeated behind a GEMM / an elementwise kernel / nothing, must give the
    same bits every time and torch's loss (tools/ctc_repeat.py is the longer version of this check)."""
    import dae._C as C_
    from dae.ctc import CTCLoss
    g = torch.Generator().manual_seed(5)
    L = T // 4
    tg = torch.randint(0, C - 1, (1, L), generator=g)
    t = torch.arange(T)
    path = tg[0][(t * L // T).clamp(max=L - 1)]
    path = torch.where(t % 3 == 2, torch.full_like(path, C - 1), path)
    logits = torch.randn(T, 1, C, generator=g)
    logits[t, 0, path] += 7.0
    lp = logits.log_softmax(-1).to(cuda)
    tgd, il, tl = tg.to(cuda), torch.tensor([T], device=cuda), torch.tensor([L], device=cuda)
    ref = torch.nn.functional.ctc_loss(lp, tgd, il, tl, blank=C - 1, reduction="sum")
    big, small = torch.randn(2048, 2048, device=cuda), torch.randn(1 << 20, device=cuda)
    f = CTCLoss(blank=C - 1, reduction="sum", validate=False)
    try:
        for overlap in (0, 1, -1):
            C_.ctc_configure(blocked=1, overlap=overlap)
            outs = []
            for rep in range(9):
                if rep % 3 == 1:
                    big @ big
                if rep % 3 == 2:
                    small.mul_(1.0001)
                x = lp.clone().requires_grad_()
                loss = f.with_scale(x, tgd, il, tl, grad_scale_hint=1.0 / T)
                (loss / T).backward()
                outs.append((loss.detach().clone(), x.grad.clone()))
            torch.cuda.synchronize()
            for l_, g_ in outs[1:]:
                assert torch.equal(l_, outs[0][0]) and torch.equal(g_, outs[0][1])
            assert abs(outs[0][0].item() - ref.item()) <= 1e-4 * abs(ref.item())
    finally:
        C_.ctc_configure()


def test_ctc_gradient_formed_in_forward_with_scale_hint(cuda, ctc_path):
    """CTCLoss.with_scale (the adapt loop's `loss / (T*N); backward()`): same loss, gradient bit-identical to the
    two-call path when the upstream gradient equals the hint, rescaled correctly when it does not."""
    from dae.ctc import CTCLoss
    T, N, C, L = 200, 2, 50, 23
    g = torch.Generator().manual_seed(12)
    lp = (torch.randn(T, N, C, generator=g) * 2).log_softmax(-1).to(cuda)
    tg = torch.randint(0, C - 1, (N, L), generator=g).to(cuda)
    il, tl = torch.tensor([T, T - 17], device=cuda), torch.tensor([L, L - 5], device=cuda)
    f = CTCLoss(blank=C - 1, reduction="sum")
    a = lp.clone().requires_grad_()
    (f(a, tg, il, tl) / (T * N)).backward()
    b = lp.clone().requires_grad_()
    lb = f.with_scale(b, tg, il, tl, grad_scale_hint=1.0 / (T * N))
    (lb / (T * N)).backward()
    assert torch.equal(a.grad, b.grad)
    c = lp.clone().requires_grad_()
    (f.with_scale(c, tg, il, tl, grad_scale_hint=1.0 / (T * N)) * 0.37).backward()       # a different upstream scale
    torch.testing.assert_close(c.grad, a.grad * (0.37 * T * N), rtol=2e-6, atol=1e-12)
    with torch.no_grad():                                                             # no grad required: lattice only
        assert torch.equal(f.with_scale(lp, tg, il, tl, grad_scale_hint=0.5), lb.detach())
