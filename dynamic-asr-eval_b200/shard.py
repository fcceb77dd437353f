"""Recording-level data parallelism (new; the reference has no distributed code).

One process per GPU, each with a full model replica.  Recordings are independent units
(lcasr/run_dynamic_eval_full.py:80-110; weights are restored after each one, lib.py:636-637), so they
are assigned longest-processing-time-first by frame count and the only exchange is ONE
all-reduce(SUM) of an int64[5] = (S, D, I, ref_words, n_recordings) vector per repeat — over NCCL on
GPUs (gloo in the CPU tests).  Integer sums are order independent, so WER is bit-identical to a
single-process run.
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment; no-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, 0
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend)
    return rank, world, local


def lpt_assign(costs, world):
    """Longest-processing-time-first: returns, per rank, the list of item indices (deterministic)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(x) for x in out]


def all_reduce_counts(counts, device=None):
    """Sum an int64 count vector over all ranks (one collective); identity for a single process."""
    t = torch.as_tensor(counts, dtype=torch.int64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu()


def gather_objects(obj):
    """all_gather_object of per-rank (index, hypothesis) lists so rank 0 can rebuild the pickle."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, obj)
        return out
    return [obj]
