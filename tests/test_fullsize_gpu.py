"""BASELINE.json-sized inputs, checked through size-independent properties (the CPU oracles would take
minutes to hours here): batch independence, conservation laws, idempotence, determinism."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _peaky_gpu(T, C, blank, seed, p_blank=0.7):
    g = torch.Generator(device="cuda").manual_seed(seed)
    lp = torch.randn(T, C, generator=g, device="cuda")
    cls = torch.randint(0, C - 1, (T,), generator=g, device="cuda")
    cls[torch.rand(T, generator=g, device="cuda") < p_blank] = blank
    lp[torch.arange(T, device="cuda"), cls] += 8
    return lp.log_softmax(-1)


def test_greedy_whole_recording(cuda):
    """cfg2 whole-recording decode [52000, 4096]: path == torch.argmax on the device, collapse is idempotent."""
    from dae.greedy import collapse_path_device, greedy_ids_device
    lp = _peaky_gpu(52000, 4096, 4095, 1)
    path, ids, n = greedy_ids_device(lp, 4095)
    assert torch.equal(path[0].long(), lp.argmax(-1))
    k = int(n[0])
    got = ids[0, :k].tolist()
    ref = torch.unique_consecutive(lp.argmax(-1))
    assert got == ref[ref != 4095].tolist()
    assert collapse_path_device(path[0], 4095) == got


def test_ctc_batch_of_64_properties(cuda):
    """[2048, 64, 4096] (4.3 GB of gradient): per-sample losses equal single-sample runs; every gradient row
    sums to zero (sum_c exp(lp) = 1 = sum_c occupancy); frames past input_length are exactly zero."""
    from dae.ctc import ctc_loss
    from dae.greedy import greedy_ids_device
    T, N, C = 2048, 64, 4096
    post = torch.stack([_peaky_gpu(T, C, C - 1, 100 + n) for n in range(N)], 1)
    labs = []
    for n in range(N):
        _, ids, k = greedy_ids_device(post[:, n], C - 1)
        labs.append(ids[0, :int(k[0])].long())
    Lmax = max(int(l.numel()) for l in labs)
    tg = torch.zeros(N, Lmax, dtype=torch.long, device=cuda)
    for n, l in enumerate(labs):
        tg[n, :l.numel()] = l
    tl = torch.tensor([int(l.numel()) for l in labs], device=cuda)
    il = torch.full((N,), T, device=cuda)
    il[3] = 1500
    il[40] = 1900
    tl[3] = min(int(tl[3]), 300)
    x = post.clone().requires_grad_()
    nll = ctc_loss(x, tg, il, tl, blank=C - 1, reduction="none")
    nll.sum().backward()
    assert torch.isfinite(nll).all()
    for n in (0, 3, 17, 40, 63):
        single = ctc_loss(post[:, n:n + 1].contiguous(), tg[n:n + 1], il[n:n + 1], tl[n:n + 1], blank=C - 1,
                          reduction="none")
        assert abs(single.item() - nll[n].item()) <= 1e-5 * abs(nll[n].item())
    rowsum = x.grad.sum(-1)                                     # [T, N]
    assert rowsum.abs().max().item() < 2e-4
    assert torch.all(x.grad[1500:, 3] == 0) and torch.all(x.grad[1900:, 40] == 0)
    assert x.grad[:1500, 3].abs().sum() > 0


@pytest.fixture
def C_():
    import dae._C as C
    yield C
    C.ctc_configure()


def test_ctc_adapt_step_blocked_equals_chain(cuda, C_):
    """The adapt step's shape ([2048, N, 4096], N = 1 and a ragged N = 2 as in AWMC): the time-blocked lattice
    (default for few samples) and the per-frame chain are two implementations of the same function: losses agree
    to 1e-6 relative, gradients within the 1e-4 elementwise tolerance, gradient rows sum to zero, padding frames are zero."""
    from dae.ctc import ctc_loss
    from dae.greedy import greedy_ids_device
    T, C = 2048, 4096
    for N, in_lens in ((1, [T]), (2, [T, T - 333])):
        post = torch.stack([_peaky_gpu(T, C, C - 1, 300 + n) for n in range(N)], 1)
        labs = []
        for n in range(N):
            _, ids, k = greedy_ids_device(post[:in_lens[n], n], C - 1)
            labs.append(ids[0, :int(k[0])].long())
        Lmax = max(int(l.numel()) for l in labs)
        tg = torch.zeros(N, Lmax, dtype=torch.long, device="cuda")
        for n, l in enumerate(labs):
            tg[n, :l.numel()] = l
        il = torch.tensor(in_lens, device="cuda")
        tl = torch.tensor([int(l.numel()) for l in labs], device="cuda")
        out = {}
        for path in ("0", "1"):
            C_.ctc_configure(blocked=int(path))
            x = post.clone().requires_grad_()
            nll = ctc_loss(x, tg, il, tl, blank=C - 1, reduction="none")
            (nll.sum() / T).backward()
            out[path] = (nll.detach().reshape(-1), x.grad)
        (nll_c, g_c), (nll_b, g_b) = out["0"], out["1"]
        # the blocked path is bit-reproducible: frames of reference and hand-over depend on values, not on timing
        x2 = post.clone().requires_grad_()
        nll2 = ctc_loss(x2, tg, il, tl, blank=C - 1, reduction="none")
        (nll2.sum() / T).backward()
        assert torch.equal(nll2.detach().reshape(-1), nll_b) and torch.equal(x2.grad, g_b)
        assert torch.allclose(nll_b, nll_c, rtol=1e-6, atol=0)
        # north_star tolerance (each term of g*(exp(lp) - occupancy) good to 1e-4 relative), as in test_kernels_gpu
        tol = 1e-4 * g_c.abs() + 1e-4 * (1.0 / T) * post.exp() + 1e-12
        assert float(((g_b - g_c).abs() / tol).max()) <= 1.0
        assert float(g_b.sum(-1).abs().max()) <= 1e-6          # sum_c exp(lp) = 1 = sum_c occupancy
        for n in range(N):
            assert torch.all(g_b[in_lens[n]:, n] == 0)


def test_stitch_recording_sized(cuda):
    """52 windows x [2048, 4096] (1.7 GB): rows are probability distributions, rows covered by one window are
    that window's rows, fused argmax == argmax of the output."""
    from dae.stitch import stitch_flat, window_positions
    nwin, Tp, C = 52, 2048, 4096
    flat = torch.cat([_peaky_gpu(Tp, C, C - 1, 200 + w) for w in range(nwin)], 0)
    starts = [2048 * i for i in range(nwin)]
    pos = window_positions(starts, [16384] * nwin, [Tp] * nwin, 14336)
    assert pos[1] == 256 and pos[-1] == 256 * (nwin - 1)
    out, path = stitch_flat(flat, [Tp * i for i in range(nwin)], pos, [Tp] * nwin)
    assert out.shape == (256 * (nwin - 1) + Tp, C)
    assert (out.exp().sum(-1) - 1).abs().max().item() < 1e-4
    torch.testing.assert_close(out[:256], flat[:256], rtol=0, atol=2e-6)         # first 256 rows: window 0 only
    torch.testing.assert_close(out[-256:], flat[-256:], rtol=0, atol=2e-6)       # last 256 rows: last window only
    assert torch.equal(path.long(), out.argmax(-1))


def test_softdtw_cfg4_properties(cuda):
    """[8, 4096, 4096] (BASELINE.json configs[3]): ALL 8 samples against the fp64 oracle (value 1e-6 relative,
    gradient 1e-5 of its scale and 1e-4 relative where E > 1e-3: inside the north star's 1e-4), plus batch
    independence, E >= 0, and conservation: every alignment starts at (0,0) and ends at (N-1,M-1), so E there
    equals the upstream gradient."""
    from oracle import softdtw_oracle as so
    from dae.soft_dtw_cuda import softdtw_backward, softdtw_forward
    g = torch.Generator(device="cuda").manual_seed(1234)
    B, N, M = 8, 4096, 4096
    a, b = torch.rand(B, N, 2, generator=g, device="cuda"), torch.rand(B, M, 2, generator=g, device="cuda")
    D = ((a[:, :, None, :] - b[:, None, :, :]) ** 2).sum(-1).contiguous()
    val, W, _ = softdtw_forward(D, 1.0, 0.0)
    v1, _, _ = softdtw_forward(D[5:6].contiguous(), 1.0, 0.0)
    assert v1.item() == val[5].item()                                            # bit-identical regardless of batch
    go = torch.arange(1, B + 1, device="cuda", dtype=torch.float32)
    E = softdtw_backward(W, go)
    assert torch.isfinite(E).all() and (E >= 0).all()
    torch.testing.assert_close(E[:, -1, -1], go, rtol=1e-6, atol=0)
    torch.testing.assert_close(E[:, 0, 0], go, rtol=2e-5, atol=0)                # 8191 multiply-add steps
    worst_abs = worst_rel = 0.0
    for s in range(B):                                                           # ~8 s of fp64 C per sample
        Dn = D[s:s + 1].cpu().numpy()
        Rr = so.forward(Dn, 1.0, 0.0)
        assert abs(val[s].item() - Rr[0, -2, -2]) <= 1e-6 * abs(Rr[0, -2, -2])
        Er = so.backward(Dn, Rr, 1.0, 0.0)[0] * float(go[s])
        got = E[s].cpu().numpy().astype(np.float64)
        err = np.abs(got - Er)
        worst_abs = max(worst_abs, err.max() / float(go[s]))
        big = Er > 1e-3 * float(go[s])
        worst_rel = max(worst_rel, (err[big] / Er[big]).max())
    print(f"softdtw cfg4 vs fp64 oracle: max |dE| / gout = {worst_abs:.3g}, max relative (E > 1e-3) = {worst_rel:.3g}")
    assert worst_abs <= 1e-5 and worst_rel <= 1e-4


def test_beam_cfg3_properties(cuda, tmp_path):
    """1 h of 50 fps posteriors, V=32, beam 100, 360 segments in one launch: equal to searching sampled segments
    alone (bit-exact), deterministic, scores ordered, start frames increasing and inside the segment."""
    from dae.ctc_beam_search import beam_search_batch
    from dae.ngram import NGramLM, read_arpa, write_synthetic_arpa
    from dae.standin import peaky_log_probs
    V, T, nseg, W = 31, 180000, 360, 100
    arpa = str(tmp_path / "lm.arpa")
    write_synthetic_arpa(arpa, V, order=4, counts=(None, 900, 20000, 60000), seed=4, fast=True)
    order, grams = read_arpa(arpa)
    lm = NGramLM(grams, order, V)
    lp = torch.from_numpy(peaky_log_probs(T, V + 1, V, 3, sharp=5.0)).to(cuda)
    offs = [int(v) for v in np.linspace(0, T, nseg + 1)]
    kw = dict(alpha=0.45, beta=1.53, blank_id=V, top_am_threshold=-6, prune_less_than_val=3.17)
    res = beam_search_batch(lp, offs, lm, W, n_best=5, **kw)
    res2 = beam_search_batch(lp, offs, lm, W, n_best=5, **kw)
    assert len(res) == nseg
    for g in range(nseg):
        beams = res[g]
        assert 1 <= len(beams) <= 5
        scores = [float(s) for s, _, _, _ in beams]
        assert all(np.isfinite(scores)) and scores == sorted(scores, reverse=True)
        for s, toks, times, _ in beams:
            assert len(toks) == len(times) and all(1 <= t_ <= V - 1 for t_ in toks)
            assert all(t2 > t1 for t1, t2 in zip(times, times[1:])) and (not times or times[-1] < offs[g + 1] - offs[g])
        assert [(np.float32(s).tobytes(), t_, tm) for s, t_, tm, _ in beams] == \
               [(np.float32(s).tobytes(), t_, tm) for s, t_, tm, _ in res2[g]]
    for g in (0, 123, 359):                                                       # batch == alone
        alone = beam_search_batch(lp[offs[g]:offs[g + 1]].contiguous(), None, lm, W, n_best=5, **kw)[0]
        assert [(np.float32(s).tobytes(), t_, tm) for s, t_, tm, _ in alone] == \
               [(np.float32(s).tobytes(), t_, tm) for s, t_, tm, _ in res[g]]
