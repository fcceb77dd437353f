"""The CPU restatement of the adapt loop (oracle/ref_loop.py) reproduces what the reference's own
lcasr/lib.py produced under import stubs (tests/golden/loop_toy.npz); no GPU, no /root/reference."""
import os
import random
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from toy import TOY, TOY_CONFIG, RecordingTokenizer, ToyModel, toy_spec  # noqa: E402

from oracle import stitch_oracle  # noqa: E402
from oracle.ref_loop import awmc_reference, dynamic_eval_reference, make_args  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loop_toy.npz"))


@pytest.mark.parametrize("online", [False, True])
def test_oracle_loop_matches_reference_golden(online):
    from dae.optim import MADGRAD
    from dae.standin import SyntheticTokenizer
    tag = "online" if online else "offline"
    tok = RecordingTokenizer(SyntheticTokenizer(vocab_size=TOY["C"] - 1, seed=0))
    model = ToyModel(TOY["C"], seed=TOY["model_seed"])
    before = [p.detach().clone() for p in model.parameters()]
    args = make_args(TOY_CONFIG, online=online, **TOY["kwargs"])
    random.seed(TOY["seed"])
    torch.manual_seed(TOY["seed"])
    rec = []
    logits = dynamic_eval_reference(args, model, toy_spec(TOY["spec_seed"], TOY["spec_n"]), TOY["seq_len"],
                                    TOY["overlap"], tok, MADGRAD, record=rec)
    assert len(rec) == int(GOLD[f"n_steps_{tag}"])
    lens = GOLD[f"ids_len_{tag}"].tolist()
    assert [len(r["ids"]) for r in rec] == lens
    assert [i for r in rec for i in r["ids"]] == GOLD[f"ids_flat_{tag}"].tolist()
    assert logits.shape == GOLD[f"logits_{tag}"].shape
    # same torch CPU ops in the same order; allow for different CPU vector paths
    np.testing.assert_allclose(np.exp(logits), np.exp(GOLD[f"logits_{tag}"]), rtol=2e-4, atol=1e-7)
    assert all(torch.equal(a, b) for a, b in zip(before, model.parameters()))


@pytest.mark.parametrize("spec_n", [6000, 120000, 360000, 415990])
def test_chunks_match_reference_prepare_chunks(spec_n):
    assert stitch_oracle.prepare_chunks(spec_n, 16384, 14336) == [tuple(x) for x in GOLD[f"chunks_{spec_n}"].tolist()]


def test_product_prepare_chunks_matches_golden():
    from dae.lib import prepare_chunks
    for spec_n in (6000, 120000, 415990):
        td, keys = prepare_chunks(torch.zeros(1, 1, spec_n), 16384, 14336)
        assert [[k, td[k].shape[-1]] for k in keys] == GOLD[f"chunks_{spec_n}"].tolist()


def test_oracle_awmc_matches_reference_golden():
    from dae.optim import MADGRAD
    from dae.standin import SyntheticTokenizer
    tok = RecordingTokenizer(SyntheticTokenizer(vocab_size=TOY["C"] - 1, seed=0))
    model = ToyModel(TOY["C"], seed=TOY["model_seed"])
    before = [p.detach().clone() for p in model.parameters()]
    args = make_args(TOY_CONFIG, **dict(TOY["kwargs"], ema_decay=0.9))
    random.seed(TOY["seed"])
    torch.manual_seed(TOY["seed"])
    logits = awmc_reference(args, model, toy_spec(TOY["spec_seed"], TOY["spec_n"]), TOY["seq_len"], TOY["overlap"], tok,
                            MADGRAD)
    # the noisy-prediction print of the reference also calls tokenizer.encode (lib.py:304); skip those entries
    gold_lens = GOLD["ids_len_awmc"].tolist()
    got = [len(e) for e in tok.encoded]
    assert sum(got) <= sum(gold_lens)
    np.testing.assert_allclose(np.exp(logits), np.exp(GOLD["logits_awmc"]), rtol=2e-4, atol=1e-7)
    assert all(torch.equal(a, b) for a, b in zip(before, model.parameters()))


def test_chunking_and_positions_properties_random_sizes():
    """Host index arithmetic of the product (dae.lib.prepare_chunks, dae.stitch.window_positions) against the
    oracle restatement of lcasr/lib.py:128-145,586-621 over random recording lengths, windows and overlaps:
    same windows, same output positions; windows tile the recording with the stride seq_len - overlap."""
    import random

    import torch

    from dae.lib import prepare_chunks
    from dae.stitch import window_positions
    from oracle import stitch_oracle
    rnd = random.Random(0)
    for _ in range(200):
        seq = rnd.choice([64, 512, 2048, 16384])
        overlap = rnd.choice([0, seq // 8, seq // 2, seq - seq // 8])
        spec_n = rnd.randint(1, 6 * seq)
        spec = torch.zeros(1, 2, spec_n)
        chunks, starts = prepare_chunks(spec, seq, overlap)
        ref = stitch_oracle.prepare_chunks(spec_n, seq, overlap)
        assert list(starts) == [s for s, _ in ref]
        assert [int(chunks[s].shape[-1]) for s in starts] == [l for _, l in ref]
        stride = seq - overlap
        assert all(b - a == stride for a, b in zip(starts, starts[1:]))
        u_lens = [l for _, l in ref]
        ds = [max(1, l // 8) for l in u_lens]                      # x8 subsampling stand-in
        assert window_positions(starts, u_lens, ds, overlap) == stitch_oracle.window_positions(starts, u_lens, ds, overlap)
