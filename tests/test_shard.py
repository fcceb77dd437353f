"""Multi-process path on CPU (gloo, world_size 2): LPT sharding + the single all-reduce of WER counts give
exactly the single-process numbers (integer sums are order independent)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HYPS = ["a b c d", "x y", "the cat sat", "", "one two three four five", "q"]
REFS = ["a c c d e", "x y z", "the cat sat", "hello", "one three four five", "q r"]
COSTS = [500, 200, 900, 100, 700, 300]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from dae.shard import all_reduce_counts, gather_objects, init_distributed, lpt_assign
    from dae.wer import rates_from_counts, word_error_counts
    r, w, _ = init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    mine = lpt_assign(COSTS, world)[rank]
    counts = word_error_counts([HYPS[i] for i in mine], [REFS[i] for i in mine])
    total = all_reduce_counts(counts)
    parts = gather_objects({i: HYPS[i] for i in mine})
    np.save(os.path.join(out_dir, f"r{rank}.npy"), total.numpy())
    if rank == 0:
        merged = {}
        for p in parts:
            merged.update(p)
        assert sorted(merged) == list(range(len(HYPS)))
        assert rates_from_counts(total)[1] == sum(len(x.split()) for x in REFS)
    dist.destroy_process_group()


def test_gloo_world2_counts_equal_single_process(tmp_path):
    from dae.shard import lpt_assign
    from dae.wer import word_error_counts
    single = word_error_counts(HYPS, REFS)
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        np.testing.assert_array_equal(np.load(tmp_path / f"r{rank}.npy"), single)
    shards = lpt_assign(COSTS, 2)
    assert sorted(shards[0] + shards[1]) == list(range(6))
    loads = [sum(COSTS[i] for i in s) for s in shards]
    assert abs(loads[0] - loads[1]) <= max(COSTS)


def test_lpt_is_deterministic_and_balanced():
    from dae.shard import lpt_assign
    costs = [415990, 120000, 360000, 240000, 180000, 300000, 90000, 60000]
    a = lpt_assign(costs, 4)
    assert a == lpt_assign(costs, 4)
    loads = [sum(costs[i] for i in s) for s in a]
    assert max(loads) <= 1.34 * (sum(costs) / 4)
    assert lpt_assign(costs, 1) == [list(range(8))]


def test_wer_counts_properties():
    from dae.wer import edit_counts, word_error_rate_detail
    assert edit_counts([], ["a", "b"]) == (0, 2, 0)
    assert edit_counts(["a", "b"], []) == (0, 0, 2)                      # empty reference: all insertions
    wer, words, ins, dele, sub = word_error_rate_detail(["a b c", "x"], ["a c c d", "x y"])
    assert words == 6 and abs(wer - (ins + dele + sub)) < 1e-15 and wer == 0.5
    assert word_error_rate_detail(["same words"], ["same words"])[0] == 0.0
    assert word_error_rate_detail(["ab"], ["abc"], use_cer=True)[1] == 3
