"""Where does one bench step go?  torch.profiler kernel table for one dynamic_eval call of the bench workload.

    python tools/step_profile.py [--frames 30000]
Diagnostic only (never a bench number): the profiler serialises launches.
"""
import argparse
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=30000)
    ap.add_argument("--rows", type=int, default=30)
    a = ap.parse_args()
    from dae import _C, lib, standin
    from dae.optim import MADGRAD
    dev = torch.device("cuda", 0)
    _C.lib()
    tok = standin.SyntheticTokenizer()
    model = standin.build_model(tok.vocab_size(), device=dev, seed=0)
    args = bench.make_args(standin.default_config())
    model.set_spike_prior(2048, nonblank_frac=0.3, seed=0)
    spec = torch.randn(1, 80, a.frames, generator=torch.Generator().manual_seed(1)).to(dev)

    def step(k):
        random.seed(k)
        torch.manual_seed(k)
        return lib.dynamic_eval(args, model, spec, bench.SEQ_LEN, bench.OVERLAP, tok, use_tqdm=False, optim=MADGRAD,
                                output="greedy")

    step(0)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    step(1)
    e.record()
    torch.cuda.synchronize()
    print(f"unprofiled step: {s.elapsed_time(e):.1f} ms for {bench.n_windows(a.frames)} windows")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
        step(2)
        torch.cuda.synchronize()
    print(p.key_averages().table(sort_by="cuda_time_total", row_limit=a.rows, max_name_column_width=70))


if __name__ == "__main__":
    main()
