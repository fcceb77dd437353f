"""numpy restatement of greedy CTC decoding (test infrastructure only).

Follows the usage of lcasr.decoding.greedy.GreedyCTCDecoder (un-vendored dependency of the
reference, version unpinned) at lcasr/lib.py:559 and lcasr/run_dynamic_eval_full.py:100:
argmax over classes -> collapse consecutive repeats -> drop blank.  PARITY UNPINNED against
lcasr's own class (absent); tie-break is torch.argmax's (first maximal index, NaN counts as
maximal), which tests/test_oracle_pins.py checks against torch itself.
"""
import numpy as np


def argmax_rows(lp):
    lp = np.asarray(lp)
    nan = np.isnan(lp)
    has_nan = nan.any(axis=-1)
    return np.where(has_nan, nan.argmax(axis=-1), np.where(nan, -np.inf, lp).argmax(axis=-1)).astype(np.int32)


def collapse(path, blank):
    path = np.asarray(path)
    if path.size == 0:
        return []
    keep = np.ones(len(path), dtype=bool)
    keep[1:] = path[1:] != path[:-1]
    keep &= path != blank
    return path[keep].astype(np.int64).tolist()


def greedy_ids(lp, blank):
    return collapse(argmax_rows(lp), blank)
