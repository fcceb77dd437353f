"""Kernel micro-benchmarks at the BASELINE shapes: CUDA events on the launching
stream, L2 flushed between iterations, GB/s of ALGORITHMIC bytes against MEASURED_PEAKS.json.

    python tools/kbench.py [--only ctc,greedy,...] [--iters 20] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


class Timer:
    def __init__(self, flush_mb=256):
        self.flush = torch.empty(flush_mb * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")

    def time(self, fn, iters=20, warmup=3, flush=True):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            if flush:
                self.flush.add_(1.0)              # 256 MB read+write > 126 MB L2
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e-3)
        ts.sort()
        return ts[len(ts) // 2], ts[0]


def peaky(T, C, blank, g, p_blank=0.7, dev="cuda"):
    lp = torch.randn(T, C, generator=g, device=dev)
    cls = torch.randint(0, C - 1, (T,), generator=g, device=dev)
    cls[torch.rand(T, generator=g, device=dev) < p_blank] = blank
    lp[torch.arange(T, device=dev), cls] += 8
    return lp.log_softmax(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    want = lambda k: not only or k in only
    import dae._C as C_
    from dae.augment import SpecAugment
    from dae.ctc import CTCLoss
    from dae.greedy import greedy_ids_device
    peak, how = peak_gbs()
    tm = Timer()
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = []

    def report(name, shape, nbytes, med, best, extra=None):
        r = {"kernel": name, "shape": shape, "bytes": nbytes, "ms_median": med * 1e3, "ms_best": best * 1e3,
             "gbs": nbytes / med / 1e9, "frac_of_%s_peak" % how: nbytes / med / 1e9 / peak}
        if extra:
            r.update(extra)
        rows.append(r)
        print(json.dumps(r), flush=True)

    if want("greedy"):
        for T in (2048, 52000):
            lp = peaky(T, 4096, 4095, g)
            med, best = tm.time(lambda: greedy_ids_device(lp, 4095), args.iters)
            report("greedy_collapse", [T, 4096], T * 4096 * 4, med, best)
            med, best = tm.time(lambda: lp.argmax(-1), args.iters)
            report("torch.argmax(cuda)", [T, 4096], T * 4096 * 4, med, best)

    if want("specaug"):
        spec = torch.randn(1, 80, 120000, device="cuda")
        win = spec[:, :, 2048:2048 + 16384]
        aug = SpecAugment(n_freq_masks=6, freq_mask_param=34)
        bands = aug.draw(1, 80, 16384)
        med, best = tm.time(lambda: aug(win, n_clean=1, bands=bands), args.iters)
        report("specaug_repeat", [80, 16384], 3 * 80 * 16384 * 4, med, best)
        med, best = tm.time(lambda: aug(win, n_clean=1), args.iters)
        report("specaug_repeat(+host band draw)", [80, 16384], 3 * 80 * 16384 * 4, med, best)

    if want("ctc"):
        for N in (1, 8, 64):
            T, Cc = 2048, 4096
            post = torch.stack([peaky(T, Cc, Cc - 1, g) for _ in range(N)], 1)      # [T,N,C]
            labs = []
            for n in range(N):
                _, ids, k = greedy_ids_device(post[:, n], Cc - 1)
                labs.append(ids[0, :int(k[0])].long())
            Lmax = max(int(l.numel()) for l in labs)
            tg = torch.zeros(N, Lmax, dtype=torch.long, device="cuda")
            for n, l in enumerate(labs):
                tg[n, :l.numel()] = l
            tl = torch.tensor([int(l.numel()) for l in labs], device="cuda")
            il = torch.full((N,), T, device="cuda")
            x = post.clone().requires_grad_()
            lossf = CTCLoss(blank=Cc - 1, reduction="sum")

            def fb():
                x.grad = None
                (lossf(x, tg, il, tl) / (T * N)).backward()
            def fwd_only():
                with torch.no_grad():
                    lossf(x, tg, il, tl)
            nbytes = 2 * T * N * Cc * 4
            med, best = tm.time(fb, args.iters)
            report("ctc_fwd_bwd(dae)", [T, N, Cc, Lmax], nbytes, med, best)
            med, best = tm.time(fwd_only, args.iters)
            report("ctc_lattice_only(dae)", [T, N, Cc, Lmax], nbytes, med, best)
            tl_f = torch.nn.CTCLoss(blank=Cc - 1, reduction="sum")

            def fb_t():
                x.grad = None
                (tl_f(x, tg, il, tl) / (T * N)).backward()
            med, best = tm.time(fb_t, max(5, args.iters // 2))
            report("ctc_fwd_bwd(torch.cuda)", [T, N, Cc, Lmax], nbytes, med, best)
            del post, x

    if want("stitch"):
        from dae.stitch import stitch_windows
        nwin, Tp, Cc = 52, 2048, 4096
        wins = [peaky(Tp, Cc, Cc - 1, g) for _ in range(nwin)]
        starts = [2048 * i for i in range(nwin)]
        ul = [16384] * nwin
        med, best = tm.time(lambda: stitch_windows(wins, starts, ul, 14336), max(3, args.iters // 4))
        n_out = 2048 + 256 * (nwin - 1)
        report("stitch(+cat)", [nwin, Tp, Cc], nwin * Tp * Cc * 4 + n_out * Cc * 4, med, best)

    if want("softdtw"):
        from dae.soft_dtw_cuda import softdtw_backward, softdtw_forward
        for (B, N, M) in ((8, 4096, 4096), (8, 1024, 1024), (512, 256, 256)):
            a = torch.rand(B, N, 2, generator=g, device="cuda")
            b = torch.rand(B, M, 2, generator=g, device="cuda")
            D = ((a[:, :, None, :] - b[:, None, :, :]) ** 2).sum(-1).contiguous()
            med, best = tm.time(lambda: softdtw_forward(D, 1.0, 0.0), max(5, args.iters // 2))
            report("softdtw_fwd", [B, N, M], 2 * B * N * M * 4, med, best)
            _, W, _ = softdtw_forward(D, 1.0, 0.0)
            go = torch.ones(B, device="cuda")
            med, best = tm.time(lambda: softdtw_backward(W, go), max(5, args.iters // 2))
            report("softdtw_bwd", [B, N, M], 3 * B * N * M * 4, med, best)
            del D, W

    if want("beam"):
        import numpy as np
        from dae.standin import peaky_log_probs
        from dae.ctc_beam_search import _Search
        from dae.ngram import NGramLM, read_arpa, write_synthetic_arpa
        V = 31
        write_synthetic_arpa("/tmp/kbench.arpa", V, order=4, counts=(None, 900, 200000, 800000), seed=4, fast=True)
        order, grams = read_arpa("/tmp/kbench.arpa")
        lm = NGramLM(grams, order, V)
        T = 180000
        lp = torch.from_numpy(peaky_log_probs(T, V + 1, V, 3, sharp=5.0)).cuda()
        for nseg in (360, 1):
            offs = [int(x) for x in np.linspace(0, T, nseg + 1)]
            sr = _Search(lp, offs, lm, 100, 0.45, 1.53, V, 0.0, 0.0, -6, 3.17, n_best=1)
            med, best = tm.time(lambda: sr.run_all(), 3, warmup=1)
            report("beam_search(beam=100,4-gram %d nodes %.0f MB)" % (lm.n_nodes, lm.nbytes() / 1e6), [T, V + 1, nseg],
                   T * (V + 1) * 4, med, best, {"frames_per_s": T / med, "audio_hours_per_s_at_50fps": T / 50 / 3600 / med})

    print("launches:", C_.launch_count())
    if args.json:
        json.dump({"peak_gbs": peak, "peak_source": how, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
