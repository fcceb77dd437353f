"""bench.py — headline benchmark: audio-hours/sec of dynamic evaluation (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl dae|reference] [--frames F]

Workload (BASELINE.json configs[1]): stand-in lcasr160rb1-sized CTC encoder (6 layers, d=768, fp32,
random-init, blank prior calibrated so teacher labels are speech-like) doing dynamic evaluation of
one Earnings22-shaped synthetic recording per step: seq 16384 / overlap 14336, 1 epoch, 6 frequency
masks x 34, MADGRAD lr 9e-5, then the no-grad final pass and overlap stitch, then greedy decode.
One "step" = one whole recording.  Under torchrun every rank processes its own recording per step
(weak scaling, no data-path collective; one int64[5] all-reduce of WER counts per step).

The JSON line carries: value (inputs resident in HBM), e2e (host spectrogram -> host hypothesis ids,
copies inside the timed region), roofline (dominant dae kernel, CUDA events live in the timed
region), cpu_baseline (the oracle loop on host cores, bounded sample), clocks, gpu_launches.
`--impl reference` times the CPU restatement of the reference loop (oracle/ref_loop.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEQ_LEN, OVERLAP, FPS = 16384, 14336, 100
KW = dict(epochs=1, shuffle=True, spec_augment_n_freq_masks=6, spec_augment_freq_mask_param=34,
          spec_augment_n_time_masks=0, optim_lr=9e-5)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


ONLINE = False      # --online: the configuration of the reference's only published timing (timeit_earnings22.sh:1)


def make_args(config):
    from types import SimpleNamespace
    a = SimpleNamespace(config=config)
    a.__dict__.update(KW)
    if ONLINE:
        a.__dict__.update(online=True, spec_augment_freq_mask_param=10)
    a.__dict__["_record_steps"] = True        # per-window pseudo-label lengths (host ints, already on the host)
    return a


def n_windows(frames):
    n, i, last, kill = 0, 0, None, False
    if frames <= SEQ_LEN:
        return 1
    for i in range(0, frames, SEQ_LEN - OVERLAP):
        u = min(SEQ_LEN, frames - i)
        if kill:
            break
        if last is not None and u < last:
            kill = True
        last = u
        n += 1
    return n


def run_reference(a, rank, world):
    """The reference's CPU path (oracle restatement of lcasr/lib.py:450-640), model on the host too."""
    if rank != 0:
        return
    import random
    from oracle.ref_loop import dynamic_eval_reference
    from dae import standin
    from dae.optim import MADGRAD
    torch.set_num_threads(os.cpu_count() or 1)
    tok = standin.SyntheticTokenizer()
    on_gpu = a.ref_device == "cuda" and torch.cuda.is_available()
    model = standin.build_model(tok.vocab_size(), device="cuda:0" if on_gpu else "cpu", seed=0)
    spec = torch.randn(1, 80, a.frames, generator=torch.Generator().manual_seed(1))
    model.set_spike_prior(2048, nonblank_frac=0.3, seed=0)
    args = make_args(standin.default_config())
    sample_windows = n_windows(a.frames) if on_gpu else 1     # the CPU arm times a bounded sample and extrapolates
    audio_h = a.frames / FPS / 3600.0 / n_windows(a.frames) * sample_windows

    def step():
        random.seed(0)
        torch.manual_seed(0)
        dynamic_eval_reference(args, model, spec, SEQ_LEN, OVERLAP, tok, MADGRAD,
                               max_windows=None if on_gpu else sample_windows)
        if on_gpu:
            torch.cuda.synchronize()
    for _ in range(a.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    val = audio_h * a.steps / dt
    cores = torch.get_num_threads()
    sample = (f"{sample_windows} of {n_windows(a.frames)} windows of a {a.frames}-frame recording per step "
              f"(adapt step + final pass + stitch), " +
              ("stand-in encoder and torch CTC on cuda:0, augmentation / greedy / stitch on the host (the reference's "
               "own placement)" if on_gpu else "model and CTC on the CPU; value extrapolated linearly in windows"))
    print(json.dumps({
        "impl": "reference", "metric": "audio-hours/sec dynamic-eval", "value": val, "unit": "audio-hours/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a),
        "cpu_baseline": {"value": val, "unit": "audio-hours/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "audio-hours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "ONE host process on rank 0 whatever --gpus says (the contract of this arm): at N > 1 compare it with the "
                "dae arm's value / N; the same loop with the encoder on the GPU (the reference's real placement) is the "
                "dae arm's `gpu_reference_loop` key, or `--ref-device cuda` here",
    }), flush=True)


def workload_config(a):
    mode = ("online=True: 1 epoch in window order, teacher posteriors stitched, no final pass, 6 freq masks x 10 "
            "(lcasr/launch_scripts/timeit_earnings22.sh:1)") if ONLINE else "1 epoch + final pass + stitch + greedy"
    return {"workload": f"dynamic eval of one synthetic Earnings22-shaped recording per step ({a.frames} frames = "
                        f"{a.frames / FPS / 60:.1f} min, 80-mel, seq {SEQ_LEN} overlap {OVERLAP}, "
                        f"{n_windows(a.frames)} windows, {mode})",
            "standin_encoder": "lcasr160rb1-shaped (6 layers, d=768, 6x128 heads, conv k=9, x8 subsampling, C=4096), "
                               "fp32, random-init, PyTorch (not the product)",
            "spec_augment": ("6 freq masks x 10" if ONLINE else "6 freq masks x 34") + ", 0 time masks", "optimizer": "MADGRAD lr 9e-5",
            "l2": "inputs larger than L2 (model weights 0.36 GB + activations stream through every step)",
            "parallelism": f"recordings sharded over {a.gpus} rank(s), one int64[5] all-reduce per step"}


def cpu_baseline_sample(frames):
    """Bounded CPU sample of the same workload through the oracle loop (reported baseline only)."""
    import random
    from oracle.ref_loop import dynamic_eval_reference
    from dae import standin
    from dae.optim import MADGRAD
    torch.set_num_threads(os.cpu_count() or 1)
    tok = standin.SyntheticTokenizer()
    model = standin.build_model(tok.vocab_size(), device="cpu", seed=0)
    spec = torch.randn(1, 80, frames, generator=torch.Generator().manual_seed(1))
    model.set_spike_prior(2048, nonblank_frac=0.3, seed=0)
    args = make_args(standin.default_config())
    k = 2
    random.seed(0)
    torch.manual_seed(0)
    t0 = time.perf_counter()
    dynamic_eval_reference(args, model, spec, SEQ_LEN, OVERLAP, tok, MADGRAD, max_windows=k)
    dt = time.perf_counter() - t0
    audio_h = frames / FPS / 3600.0 / n_windows(frames) * k
    return {"value": audio_h / dt, "unit": "audio-hours/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{k} of {n_windows(frames)} windows of the {frames}-frame recording (adapt + final pass + "
                      f"stitch) through oracle/ref_loop.py with model and CTC on the CPU, {dt:.1f} s"}


def aux_kernels(peak):
    """The north star's other kernels at their BASELINE.json shapes (configs 3 and 4 + the whole-recording
    greedy), timed alone with CUDA events and an L2 flush between iterations.  Reported next to the step's own
    kernels; not part of `value`."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from kbench import Timer, peaky
    from dae.ctc_beam_search import _Search
    from dae.greedy import greedy_ids_device
    from dae.ngram import NGramLM, read_arpa, write_synthetic_arpa
    from dae.soft_dtw_cuda import softdtw_backward, softdtw_forward
    from dae.standin import peaky_log_probs
    tm = Timer()
    g = torch.Generator(device="cuda").manual_seed(0)
    out = {}

    def rec(name, shape, nbytes, med, extra=None):
        out[name] = {"shape": shape, "algorithmic_bytes": nbytes, "ms": med * 1e3, "gbs": nbytes / med / 1e9,
                     "frac_of_hbm_peak": nbytes / med / 1e9 / peak}
        if extra:
            out[name].update(extra)
    lp = peaky(52000, 4096, 4095, g)
    med, _ = tm.time(lambda: greedy_ids_device(lp, 4095), 10)
    rec("greedy_collapse[52000x4096]", [52000, 4096], 52000 * 4096 * 4, med)
    del lp
    B, N, M = 8, 4096, 4096
    x, y = torch.rand(B, N, 2, generator=g, device="cuda"), torch.rand(B, M, 2, generator=g, device="cuda")
    D = ((x[:, :, None, :] - y[:, None, :, :]) ** 2).sum(-1).contiguous()
    med, _ = tm.time(lambda: softdtw_forward(D, 1.0, 0.0), 6)
    rec("softdtw_fwd[8x4096x4096]", [B, N, M], 2 * B * N * M * 4, med)
    _, W, _ = softdtw_forward(D, 1.0, 0.0)
    go = torch.ones(B, device="cuda")
    med, _ = tm.time(lambda: softdtw_backward(W, go), 6)
    rec("softdtw_bwd[8x4096x4096]", [B, N, M], 3 * B * N * M * 4, med)
    del D, W
    V, T, nseg = 31, 180000, 360
    arpa = "/tmp/dae_bench_4gram.arpa"
    write_synthetic_arpa(arpa, V, order=4, counts=(None, 900, 200000, 800000), seed=4, fast=True)
    order, grams = read_arpa(arpa)
    lm = NGramLM(grams, order, V)
    offs = [int(v) for v in np.linspace(0, T, nseg + 1)]

    def beam_case(name, lpb, lm_, dense, segs, note):
        sr = _Search(lpb, segs, lm_, 100, 0.45, 1.53, V, 0.0, 0.0, -6, 3.17, n_best=1, dense_lm=dense)
        med, _ = tm.time(lambda: sr.run_all(), 3, warmup=1)
        st = sr.stats()
        n_ctx = int((lm_.depth < lm_.order).sum())
        rec(name, [T, V + 1, len(segs) - 1], T * (V + 1) * 4, med,
            {"frames_per_s": T / med, "audio_hours_per_s": T / 50 / 3600 / med, "lm_nodes": lm_.n_nodes,
             "lm_hbm_bytes": lm_.nbytes() + (8 * n_ctx * V if sr.row is not None else 0),
             "lm_path": "dense row/next tables" if sr.row is not None else "trie walk (fail links + binary search)",
             "candidates_per_frame": st["candidates"] / T, "ns_per_candidate": med * 1e9 / max(st["candidates"], 1),
             "us_per_frame_per_segment": med * 1e6 / (T / (len(segs) - 1)),
             "lm_loads": st["lm_loads"], "lm_probe_bytes_32B_sectors": 32 * st["lm_loads"],
             "algorithmic_gbs_incl_lm_probes": (T * (V + 1) * 4 + 32 * st["lm_loads"]) / med / 1e9, "note": note})
        del sr
    lpb = torch.from_numpy(peaky_log_probs(T, V + 1, V, 3, sharp=5.0)).cuda()
    beam_case("beam_search[1h@50fps,V=32,beam=100,360 segments]", lpb, lm, True, offs,
              "flat synthetic posteriors (sharp=5): ~25 of 32 classes pass the -6 AM threshold every frame")
    beam_case("beam_search[1h,one sequence]", lpb, lm, True, [0, T], "the same hour as ONE sequence: per-frame latency")
    big = NGramLM.synthetic(V, 5, 8_000_000, seed=4)
    beam_case("beam_search[360 segments, 198 MB 5-gram trie walked in HBM]", lpb, big, False, offs,
              "LM larger than L2 (8.2 M nodes), no dense tables: every LM query walks fail links in HBM")
    del big
    lps = torch.from_numpy(peaky_log_probs(T, V + 1, V, 3, sharp=12.0)).cuda()
    beam_case("beam_search[360 segments, speech-like sharp posteriors]", lps, lm, True, offs,
              "top class ~0.99 per frame (sharp=12), as trained CTC models emit: 1-3 candidate classes per frame")
    beam_case("beam_search[1h, one sequence, speech-like sharp posteriors]", lps, lm, True, [0, T],
              "per-frame latency of ONE search on speech-like posteriors")
    return out


def ctc_table(peak, tm):
    """CTC loss+grad at the adapt step's frame/class counts for N in {1, 8, 64, 256}: torch's CUDA
    ctc_loss (what the reference calls on a GPU, lcasr/lib.py:492,575-579) beside dae's two implementations.
    Labels = greedy of the posteriors themselves (teacher = student), L ~ 600 per sample."""
    from dae import _C
    from dae.ctc import CTCLoss
    from dae.greedy import greedy_ids_device
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from kbench import peaky
    g = torch.Generator(device="cuda").manual_seed(7)
    T, Cc = 2048, 4096
    out = {}
    for N in (1, 8, 64, 256):
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
        x = post.requires_grad_()
        nbytes = 2 * T * N * Cc * 4
        row = {"shape": [T, N, Cc], "Lmax": Lmax, "algorithmic_bytes": nbytes}

        def run(lossf):
            def fb():
                x.grad = None
                (lossf(x, tg, il, tl) / (T * N)).backward()
            med, _ = tm.time(fb, 5 if N >= 64 else 10)
            return {"ms": med * 1e3, "gbs": nbytes / med / 1e9, "frac_of_hbm_peak": nbytes / med / 1e9 / peak}
        row["torch_cuda_ctc_loss"] = run(torch.nn.CTCLoss(blank=Cc - 1, reduction="sum"))
        dae_f = CTCLoss(blank=Cc - 1, reduction="sum", validate=False)
        try:
            _C.ctc_configure(blocked=0)
            row["dae_chain"] = run(dae_f)
            if N <= 8:
                _C.ctc_configure(blocked=1)
                row["dae_time_blocked"] = run(dae_f)
        finally:
            _C.ctc_configure()
        row["dae_default"] = "time_blocked" if N <= 8 else "chain"
        best = min(row[k]["ms"] for k in ("dae_chain", "dae_time_blocked") if k in row)
        row["speedup_vs_torch_cuda"] = row["torch_cuda_ctc_loss"]["ms"] / best
        out[f"N={N}"] = row
        del post, x, tg
        torch.cuda.empty_cache()
    return out


def numba_softdtw_table(tm):
    """The reference's numba-CUDA soft-DTW algorithm (one block per sample, max(N,M) <= 1024 threads, global-memory
    operands, fp64 math: lcasr_nemo/soft_dtw_cuda.py:33-111) restated in oracle/softdtw_numba_port.py and timed on
    this GPU beside dae at the reference's profile() shapes (:382-428) and at its 1024 cap."""
    from dae.soft_dtw_cuda import softdtw_backward, softdtw_forward
    out = {}
    try:
        from oracle import softdtw_numba_port as nb
        nb_err = None
    except Exception as e:                                      # numba missing / cannot target this GPU
        nb, nb_err = None, f"{type(e).__name__}: {e}"
    g = torch.Generator(device="cuda").manual_seed(1234)
    for (B, N, M) in ((128, 17, 15), (512, 64, 64), (512, 256, 256), (8, 1024, 1024)):
        a_, b_ = torch.rand(B, N, 2, generator=g, device="cuda"), torch.rand(B, M, 2, generator=g, device="cuda")
        D = ((a_[:, :, None, :] - b_[:, None, :, :]) ** 2).sum(-1).contiguous()
        row = {}
        fw, _ = tm.time(lambda: softdtw_forward(D, 1.0, 0.0), 10)
        _, W, _ = softdtw_forward(D, 1.0, 0.0)
        go = torch.ones(B, device="cuda")
        bw, _ = tm.time(lambda: softdtw_backward(W, go), 10)
        row["dae_fwd_ms"], row["dae_bwd_ms"] = fw * 1e3, bw * 1e3
        if nb is not None:
            try:
                f2, b2 = nb.time_fwd_bwd(D, 1.0, 0.0, tm)
                row["numba_cuda_fwd_ms"], row["numba_cuda_bwd_ms"] = f2 * 1e3, b2 * 1e3
                row["speedup_fwd_bwd"] = (f2 + b2) / (fw + bw)
            except Exception as e:
                row["numba_cuda_error"] = f"{type(e).__name__}: {str(e)[:200]}"
        else:
            row["numba_cuda_error"] = nb_err
        out[f"[{B},{N},{M}]"] = row
    return out


def gpu_reference_loop(frames, dev, windows=None):
    """The reference's real device placement (lcasr/lib.py:538-629): encoder + torch CTC on the GPU, SpecAugment on
    the CPU, H2D of every [2,80,16384] batch, D2H of the posteriors for greedy decoding, CPU overlap stitch —
    through oracle/ref_loop.py on the same stand-in model, one recording, wall clock."""
    import random
    from oracle.ref_loop import dynamic_eval_reference
    from dae import standin
    from dae.optim import MADGRAD
    tok = standin.SyntheticTokenizer()
    model = standin.build_model(tok.vocab_size(), device=dev, seed=0)
    model.set_spike_prior(2048, nonblank_frac=0.3, seed=0)
    spec = torch.randn(1, 80, frames, generator=torch.Generator().manual_seed(100))
    args = make_args(standin.default_config())
    nw = n_windows(frames)
    k = nw if windows is None else min(windows, nw)

    def once():
        random.seed(0)
        torch.manual_seed(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dynamic_eval_reference(args, model, spec, SEQ_LEN, OVERLAP, tok, MADGRAD, max_windows=None if k == nw else k)
        torch.cuda.synchronize()
        return time.perf_counter() - t0
    once() if k < nw else None                                 # warm-up (cuDNN/cuBLAS plans) on the bounded form only
    dt = once()
    audio_h = frames / FPS / 3600.0 * k / nw
    return {"value": audio_h / dt, "unit": "audio-hours/s", "seconds": dt, "windows": k, "of_windows": nw,
            "cores": torch.get_num_threads(),
            "what": "reference loop restated (oracle/ref_loop.py): stand-in encoder + torch.nn.CTCLoss on the GPU, "
                    "CPU SpecAugment, per-step H2D of the batch and D2H of [2048,4096] posteriors for greedy, CPU stitch"}


def awmc_throughput(frames, dev, tok, model):
    """lib.AWMC (lcasr/lib.py:206-376; published RTF 0.097, timeit_earnings22.sh:10-12) on one recording."""
    import random
    from dae import lib
    from dae.optim import MADGRAD
    spec = torch.randn(1, 80, frames, generator=torch.Generator().manual_seed(100)).to(dev)
    args = make_args(__import__("dae.standin", fromlist=["x"]).default_config())

    def once():
        random.seed(0)
        torch.manual_seed(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ids = lib.AWMC(args, model, spec, SEQ_LEN, OVERLAP, tok, use_tqdm=False, optim=MADGRAD, output="greedy")
        torch.cuda.synchronize()
        return time.perf_counter() - t0, len(ids)
    dt, n = once()
    dt, n = once()
    return {"value": frames / FPS / 3600.0 / dt, "unit": "audio-hours/s", "seconds": dt, "rtf": dt / (frames / FPS),
            "hyp_ids": n, "what": "dae.lib.AWMC, 1 epoch, same recording/model/SpecAugment as the headline"}


def run_dae(a, rank, world, local):
    import random
    import torch.distributed as dist
    from dae import _C, lib, prof, standin
    from dae.optim import MADGRAD
    from dae.shard import all_reduce_counts
    from dae.wer import word_error_counts
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the dae kernels have no CPU path")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    _C.lib()
    tok = standin.SyntheticTokenizer()
    model = standin.build_model(tok.vocab_size(), device=dev, seed=0)
    args = make_args(standin.default_config())
    # two synthetic recordings per rank, alternated so no step re-reads the previous step's input
    specs_host = [torch.randn(1, 80, a.frames, generator=torch.Generator().manual_seed(100 * rank + k)).pin_memory()
                  for k in range(2)]
    specs_dev = [s.to(dev) for s in specs_host]
    model.set_spike_prior(2048, nonblank_frac=0.3, seed=0)     # speech-like pseudo-label density, see standin.py
    gold = [" ".join(tok.decode([i]) for i in random.Random(k).choices(range(1, 4095), k=a.frames // 40))
            for k in range(2)]
    audio_h = a.frames / FPS / 3600.0

    label_lens = []

    def step(k, host):
        random.seed(k)
        torch.manual_seed(k)
        spec = specs_host[k % 2] if host else specs_dev[k % 2]
        ids = lib.dynamic_eval(args, model, spec, SEQ_LEN, OVERLAP, tok, use_tqdm=False, optim=MADGRAD,
                               output="greedy")
        label_lens.extend(len(r["ids"]) for r in args.__dict__.get("_step_log", []))
        counts = word_error_counts([tok.decode(ids)], [gold[k % 2]])
        all_reduce_counts(counts, dev)
        return ids

    def timed(n_steps, host, first):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        n_ids = 0
        for k in range(n_steps):
            n_ids += len(step(first + k, host))
        e.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), n_ids

    for k in range(a.warmup):
        step(k, False)
    sampler = ClockSampler(local)
    sampler.start()
    prof.reset()
    prof.enable(True)
    l0 = _C.launch_count()
    t_dev, _ = timed(a.steps, False, a.warmup)
    launches = _C.launch_count() - l0
    kern = prof.summary()
    prof.enable(False)
    prof.reset()
    t_e2e, n_ids = timed(a.steps, True, a.warmup + a.steps)
    h2d_step, d2h_step = prof.h2d_bytes // max(a.steps, 1), prof.d2h_bytes // max(a.steps, 1)
    clocks = sampler.stop()

    value = audio_h * a.steps * world / t_dev
    e2e = audio_h * a.steps * world / t_e2e
    if rank != 0:
        return
    peak, peak_src = measured_peak()
    # dominant dae op by device time: the CTC loss+grad pair (lattice launch + dense gradient launch)
    nwin = n_windows(a.frames)
    table = {}
    for name, r in kern.items():
        table[name] = {"launches_per_step": r["launches"] / a.steps, "avg_us": r["avg_ms"] * 1e3,
                       "algorithmic_bytes_per_launch": r["bytes_per_launch"],
                       "gbs": (r["bytes_per_launch"] / (r["avg_ms"] * 1e-3) / 1e9) if r["avg_ms"] > 0 else None}
    roof = None
    if "ctc_loss_grad" in kern or ("ctc_lattice" in kern and "ctc_grad" in kern):
        if "ctc_loss_grad" in kern:                               # one library call: bands, scan, dense + sparse gradient
            t_pair = kern["ctc_loss_grad"]["avg_ms"]
            by = kern["ctc_loss_grad"]["bytes_per_launch"]
        else:
            t_pair = kern["ctc_lattice"]["avg_ms"] + kern["ctc_grad"]["avg_ms"]
            by = kern["ctc_grad"]["bytes_per_launch"]             # 2*T*N*C*4: read lp once, write grad once
        ach = by / (t_pair * 1e-3) / 1e9
        traffic, traffic_warm = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get("ctc_lattice+ctc_grad")
            traffic_warm = (tj.get("ctc_lattice+ctc_grad_warm_l2") or {}).get("bytes")
        roof = {"kernel": "ctc_lattice+ctc_grad (CTC loss+grad of one adapt step, T=2048 N=1 C=4096)", "bound": "hbm",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_note": "dram read+write of the call's launches (bands, scan, dense gradient, label-class gradient) "
                                "from one ncu --set full capture (caches flushed between kernels, so the gradient "
                                "rows the last launch patches are fetched again); with ncu --cache-control none "
                                f"the same launches move {traffic_warm} bytes (bands, emissions and the fresh "
                                "gradient rows stay in the 126 MB L2)",
                "algorithmic_bytes": by, "peak_source": peak_src, "avg_launch_us": t_pair * 1e3,
                "note": "N=1: time-blocked lattice (transfer bands + 256-step boundary scan; the dense part of the gradient streams under the scan, the label classes follow it); the scan is a dependent chain (latency-bound); see DESIGN.md"}
    cpu = cpu_baseline_sample(a.frames) if world == 1 else None
    aux = aux_kernels(peak) if (world == 1 and not a.no_aux) else None
    extra = {}
    if world == 1 and not a.no_aux:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from kbench import Timer
        tm = Timer()
        extra["ctc_vs_torch_cuda"] = ctc_table(peak, tm)
        extra["softdtw_vs_numba_cuda"] = numba_softdtw_table(tm)
        del tm
        torch.cuda.empty_cache()
        extra["awmc"] = awmc_throughput(a.frames, dev, tok, model)
        extra["gpu_reference_loop"] = gpu_reference_loop(a.frames, dev)
        extra["gpu_reference_loop"]["dae_speedup_e2e"] = e2e / extra["gpu_reference_loop"]["value"]
    line = {
        "metric": "audio-hours/sec dynamic-eval", "value": value, "unit": "audio-hours/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": t_dev / a.steps * 1e3, "higher_is_better": True,
        "scaling": "weak",
        # BASELINE.md's only published number for this metric: 95.77 s for one 415 990-frame recording, online=True
        # (lcasr/launch_scripts/timeit_earnings22.sh:1,6-8; other GPU, the real model) = 0.012066 audio-hours/s
        "vs_baseline": (value / world / (4159.90 / 95.77 / 3600.0)) if (ONLINE and a.frames == 415990) else None,
        "dtype": "f32",
        "data": "synthetic (N(0,1) log-mel stand-in, random-init weights + a fixed logit spike pattern so that pseudo-labels are speech-like, ~600 labels per window)",
        "config": workload_config(a),
        "e2e": {"value": e2e, "unit": "audio-hours/s", "h2d_bytes_per_step": int(h2d_step),
                "d2h_bytes_per_step": int(d2h_step), "ms_per_step": t_e2e / a.steps * 1e3},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": table, "cpu_baseline": cpu,
        "kernels_at_baseline_shapes": aux, **extra,
        "avg_pseudo_label_len": (sum(label_lens) / len(label_lens)) if label_lens else None,
    }
    print(json.dumps(line), flush=True)


def run_sweep(a, rank, world, local):
    """BASELINE.json configs[4]: a FIXED ragged set of synthetic recordings (TED-LIUM-, Rev16- or Earnings22-shaped,
    SURVEY.md §8d cfg5) LPT-sharded over the ranks by frame count: strong scaling.  One step = the whole sweep.
    Prints the all-reduced (S, D, I, words, n) and a hash of all hypotheses so lines at different N can be compared
    for bit-equal WER; the limiter is named from the per-rank busy times."""
    import hashlib
    import random
    import zlib
    import torch.distributed as dist
    from dae import _C, lib, standin
    from dae.optim import MADGRAD
    from dae.shard import all_reduce_counts, gather_objects, lpt_assign
    from dae.wer import rates_from_counts, word_error_counts
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    _C.lib()
    tok = standin.SyntheticTokenizer()
    model = standin.build_model(tok.vocab_size(), device=dev, seed=0)
    model.set_spike_prior(2048, nonblank_frac=0.3, seed=0)
    args = make_args(standin.default_config())
    recs = standin.synthetic_recordings(a.sweep, tokenizer=tok, scale=a.sweep_scale)
    frames = [r["frames"] for r in recs]
    mine = lpt_assign(frames, world)[rank]
    specs = {i: recs[i]["process_fn"](recs[i])[0].pin_memory() for i in mine}        # host spectrograms (pinned)
    audio_h = sum(frames) / FPS / 3600.0

    def sweep():
        hyp, t_busy = {}, 0.0
        for i in mine:
            key = zlib.crc32(f"0|0|{recs[i]['id']}".encode())                        # seeded per recording, not per rank
            random.seed(key)
            torch.manual_seed(key ^ 0x5bd1e995)
            t0 = time.perf_counter()
            hyp[i] = lib.dynamic_eval(args, model, specs[i], SEQ_LEN, OVERLAP, tok, use_tqdm=False, optim=MADGRAD,
                                      output="greedy")
            t_busy += time.perf_counter() - t0
        counts = word_error_counts([tok.decode(hyp[i]) for i in mine], [recs[i]["text"] for i in mine])
        total = all_reduce_counts(counts, dev)
        return hyp, total, t_busy

    def timed(n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            out = sweep()
        e.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out
    # warm-up: W passes over this rank's SHORTEST recording (plans, allocator), not the whole sweep
    if mine:
        j = min(mine, key=lambda i: frames[i])
        for _ in range(a.warmup):
            lib.dynamic_eval(args, model, specs[j], SEQ_LEN, OVERLAP, tok, use_tqdm=False, optim=MADGRAD, output="greedy")
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _C.launch_count()
    t, (hyp, total, busy) = timed(a.steps)
    launches = _C.launch_count() - l0
    clocks = sampler.stop()
    parts = gather_objects((hyp, busy / 1.0, [frames[i] for i in mine]))
    if rank != 0:
        return
    all_h = {}
    for h, _, _ in parts:
        all_h.update(h)
    digest = hashlib.sha256(repr([(i, all_h[i]) for i in sorted(all_h)]).encode()).hexdigest()[:16]
    wer, words, ir, dr, sr = rates_from_counts(total)
    busy_s = [b for _, b, _ in parts]
    load = [sum(f) for _, _, f in parts]
    longest = max(frames)
    lim = ("LPT imbalance: the busiest rank holds %.1f%% of the frames vs %.1f%% ideal; the longest recording alone is %.1f%%"
           % (100 * max(load) / sum(frames), 100 / world, 100 * longest / sum(frames)))
    print(json.dumps({
        "metric": "audio-hours/sec dynamic-eval", "value": audio_h * a.steps / t, "unit": "audio-hours/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": t / a.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"sweep over {len(recs)} {a.sweep}-shaped synthetic recordings "
                               f"({audio_h:.2f} audio hours, {min(frames)}-{max(frames)} frames), LPT-sharded by frames",
                   "sweep_scale": a.sweep_scale, "parallelism": f"{world} rank(s), one int64[5] all-reduce per sweep",
                   "l2": "inputs larger than L2"},
        "wer_counts_SDIWn": [int(x) for x in total.tolist()], "wer": wer, "hypotheses_sha256_16": digest,
        "per_rank_busy_s": busy_s, "per_rank_frames": load, "limiter": lim, "gpu_launches": int(launches),
        "e2e": {"value": audio_h * a.steps / t, "unit": "audio-hours/s",
                "h2d_bytes_per_step": int(sum(frames) * 80 * 4), "d2h_bytes_per_step": int(sum(4 * len(v) + 4 for v in all_h.values())),
                "note": "sweep inputs start in pinned host memory: value == e2e"},
        "clocks": clocks}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dae", choices=["dae", "reference"])
    ap.add_argument("--frames", type=int, default=120000, help="frames per synthetic recording (100 fps)")
    ap.add_argument("--no-aux", dest="no_aux", action="store_true", help="skip the BASELINE-shape kernel table")
    ap.add_argument("--online", action="store_true",
                    help="online=True + freq_mask_param 10: the reference's published-timing configuration")
    ap.add_argument("--sweep", default="", choices=["", "tedlium", "rev16", "earnings22"],
                    help="strong-scaling sweep over a fixed ragged recording set (BASELINE.json configs[4])")
    ap.add_argument("--sweep-scale", dest="sweep_scale", type=float, default=1.0, help="shrink the sweep's durations")
    ap.add_argument("--ref-device", dest="ref_device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference: where the stand-in encoder + torch CTC run (cpu = host cores, the "
                         "contract's arm; cuda = the reference's real placement, lcasr/lib.py:549)")
    a = ap.parse_args()
    global ONLINE
    ONLINE = bool(a.online)
    if a.impl == "reference":
        run_reference(a, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return
    from dae.shard import init_distributed
    rank, world, local = init_distributed()
    if a.sweep:
        run_sweep(a, rank, world, local)
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    if world != a.gpus and rank == 0:
        print(f"[bench] note: --gpus {a.gpus} but WORLD_SIZE={world}; launch N>1 with torch.distributed.run",
              file=sys.stderr)
    run_dae(a, rank, world, local)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
