"""GPU: dae.lib.dynamic_eval (CUDA kernels) vs the CPU-style oracle loop, same toy model on the same
device, same seeds and band draws.  Pseudo-label ids per step are bit-exact; losses and the stitched
posteriors agree to fp32 tolerances (the two loops share the model's cuBLAS kernels)."""
import os
import random
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from toy import TOY, TOY_CONFIG, RecordingTokenizer, ToyModel, toy_spec  # noqa: E402

from oracle import greedy_oracle  # noqa: E402
from oracle.ref_loop import dynamic_eval_reference, make_args  # noqa: E402

pytestmark = pytest.mark.gpu


def _run_pair(cuda, online, **over):
    from dae import lib
    from dae.optim import MADGRAD
    from dae.standin import SyntheticTokenizer
    kw = dict(TOY["kwargs"], **over)
    spec = toy_spec(TOY["spec_seed"], TOY["spec_n"])
    outs = []
    for which in ("oracle", "dae"):
        tok = RecordingTokenizer(SyntheticTokenizer(vocab_size=TOY["C"] - 1, seed=0))
        model = ToyModel(TOY["C"], seed=TOY["model_seed"]).to(cuda)
        model.device = cuda
        before = [p.detach().clone() for p in model.parameters()]
        args = make_args(TOY_CONFIG, online=online, _record_steps=True, **kw)
        random.seed(TOY["seed"])
        torch.manual_seed(TOY["seed"])
        if which == "oracle":
            rec = []
            logits = dynamic_eval_reference(args, model, spec, TOY["seq_len"], TOY["overlap"], tok, MADGRAD, record=rec)
        else:
            logits = lib.dynamic_eval(args, model, spec, TOY["seq_len"], TOY["overlap"], tok, use_tqdm=False,
                                      optim=MADGRAD)
            rec = args._step_log
        assert all(torch.equal(a, b) for a, b in zip(before, model.parameters())), "parameters must be restored"
        outs.append((logits, rec, tok.encoded))
    return outs


@pytest.mark.parametrize("online", [False, True])
def test_dynamic_eval_matches_oracle_loop(cuda, online):
    (lo, ro, eo), (ld, rd, ed) = _run_pair(cuda, online)
    assert [r["key"] for r in ro] == [r["key"] for r in rd]           # same window order (shuffle RNG)
    assert eo == ed                                                    # pseudo-label ids: bit-exact, every step
    for a, b in zip(ro, rd):
        assert abs(a["loss"] - b["loss"]) <= 1e-4 * abs(a["loss"])     # north_star CTC tolerance
    assert lo.shape == ld.shape and ld.dtype == np.float32
    np.testing.assert_allclose(np.exp(ld), np.exp(lo), rtol=2e-3, atol=1e-6)
    assert (greedy_oracle.argmax_rows(ld) == greedy_oracle.argmax_rows(lo)).mean() > 0.995


def test_dynamic_eval_with_cutout_matches_oracle_loop(cuda):
    """Secondary augmentation: cutout rectangles on the augmented copy, same RNG stream."""
    (lo, ro, eo), (ld, rd, ed) = _run_pair(cuda, False, cutout_num_rectangles=30, cutout_max_width=60,
                                          cutout_max_height=12, cutout_value='mean')
    assert eo == ed
    for a, b in zip(ro, rd):
        assert abs(a["loss"] - b["loss"]) <= 1e-4 * abs(a["loss"])
    np.testing.assert_allclose(np.exp(ld), np.exp(lo), rtol=2e-3, atol=1e-6)


def test_dynamic_eval_epochs0_is_plain_inference(cuda):
    (lo, ro, _), (ld, rd, _) = _run_pair(cuda, False, epochs=0)
    assert ro == [] and rd == []
    np.testing.assert_allclose(np.exp(ld), np.exp(lo), rtol=1e-4, atol=1e-7)


def test_dynamic_eval_output_modes(cuda):
    from dae import lib
    from dae.optim import MADGRAD
    from dae.standin import SyntheticTokenizer
    tok = SyntheticTokenizer(vocab_size=TOY["C"] - 1, seed=0)
    spec = toy_spec(3, 2000)
    res = {}
    for mode in ("numpy", "device", "greedy"):
        model = ToyModel(TOY["C"], seed=1).to(cuda)
        model.device = cuda
        args = make_args(TOY_CONFIG, **dict(TOY["kwargs"], shuffle=False, epochs=1))
        random.seed(0)
        torch.manual_seed(0)
        res[mode] = lib.dynamic_eval(args, model, spec, 1024, 512, tok, use_tqdm=False, optim=MADGRAD, output=mode)
    assert isinstance(res["numpy"], np.ndarray) and res["device"].is_cuda
    np.testing.assert_array_equal(res["numpy"], res["device"].cpu().numpy())
    assert res["greedy"] == greedy_oracle.greedy_ids(res["numpy"], TOY["C"] - 1)
    logits, params = lib.dynamic_eval(args, model, spec, 1024, 512, tok, use_tqdm=False, optim=MADGRAD, return_params=True)
    assert len(params) == len(list(model.parameters())) and not params[0].is_cuda
    assert any(not torch.equal(p.cpu(), q) for p, q in zip(model.parameters(), params))   # adapted copy differs


def test_short_recording_single_window(cuda):
    from dae import lib
    from dae.optim import MADGRAD
    from dae.standin import SyntheticTokenizer
    tok = SyntheticTokenizer(vocab_size=TOY["C"] - 1, seed=0)
    model = ToyModel(TOY["C"], seed=2).to(cuda)
    model.device = cuda
    args = make_args(TOY_CONFIG, **TOY["kwargs"])
    out = lib.dynamic_eval(args, model, toy_spec(1, 600), 1024, 512, tok, use_tqdm=False, optim=MADGRAD)
    assert out.shape == (75, TOY["C"])
    np.testing.assert_allclose(np.exp(out).sum(-1), 1.0, rtol=1e-4)


def test_awmc_matches_oracle_loop(cuda):
    from dae import lib
    from dae.optim import MADGRAD
    from dae.standin import SyntheticTokenizer
    from oracle.ref_loop import awmc_reference
    spec = toy_spec(TOY["spec_seed"], TOY["spec_n"])
    outs = []
    for which in ("oracle", "dae"):
        tok = RecordingTokenizer(SyntheticTokenizer(vocab_size=TOY["C"] - 1, seed=0))
        model = ToyModel(TOY["C"], seed=TOY["model_seed"]).to(cuda)
        model.device = cuda
        before = [p.detach().clone() for p in model.parameters()]
        args = make_args(TOY_CONFIG, **dict(TOY["kwargs"], ema_decay=0.9))
        random.seed(TOY["seed"])
        torch.manual_seed(TOY["seed"])
        if which == "oracle":
            logits = awmc_reference(args, model, spec, TOY["seq_len"], TOY["overlap"], tok, MADGRAD)
        else:
            logits = lib.AWMC(args, model, spec, TOY["seq_len"], TOY["overlap"], tok, use_tqdm=False, optim=MADGRAD)
        assert all(torch.equal(a, b) for a, b in zip(before, model.parameters()))
        outs.append((logits, tok.encoded))
    (lo, eo), (ld, ed) = outs
    assert eo == ed                                                    # anchor/leader label banks: bit-exact
    np.testing.assert_allclose(np.exp(ld), np.exp(lo), rtol=2e-3, atol=1e-6)


def test_run_dynamic_eval_full_main_and_beamsearch(cuda, tmp_path):
    """Entry point end to end on synthetic recordings: greedy path, pickle schema, then the LM beam-search path."""
    import pickle
    from types import SimpleNamespace
    from dae import lib, run_dynamic_eval_full as r
    from dae.ngram import write_synthetic_arpa
    from dae.standin import SyntheticTokenizer, synthetic_recordings
    V = TOY["C"] - 1
    tok = SyntheticTokenizer(vocab_size=V, seed=0)
    data = synthetic_recordings("tedlium", tokenizer=tok, scale=0.004)[:3]           # a few seconds each
    model = ToyModel(TOY["C"], seed=TOY["model_seed"])
    save = str(tmp_path / "res.pkl")
    args = SimpleNamespace(split="test", dataset="tedlium", repeats=1, save_path=save, log="", checkpoint="toy",
                           seq_len=1024, overlap=512, awmc=False, consistency=False, config=TOY_CONFIG,
                           **dict(TOY["kwargs"], epochs=1))
    wer = r.main(args, model=model, tokenizer=tok, data=data, normalize=lambda s: s)
    assert wer > 0 and np.isfinite(wer)
    saved = pickle.load(open(save.replace(".pkl", "_1.pkl"), "rb"))
    assert set(saved) >= {"wer", "words", "ins_rate", "del_rate", "sub_rate", "model_output", "gold", "elapsed_times",
                          "args_dict", "repeat"}
    assert saved["wer"] == wer and len(saved["model_output"]) == 3
    assert abs(saved["ins_rate"] + saved["del_rate"] + saved["sub_rate"] - wer) < 1e-12
    # beam-search path (run_dynamic_eval_full.py:56-65,101-104): n-gram LM from an ARPA file
    arpa = str(tmp_path / "lm.arpa")
    write_synthetic_arpa(arpa, V, order=3, counts=(None, 400, 600), seed=2)
    bs = lib.load_beamsearch(arpa, tokenizer=tok)
    args.lm_eval_beams, args.lm_tta_beams, args.save_path = 5, 3, ""
    wer2 = r.main(args, model=model, tokenizer=tok, data=data[:1], normalize=lambda s: s, beamsearch=bs)
    assert np.isfinite(wer2)


def test_cli_beamsearch_flag_builds_the_lm_decoder(cuda, tmp_path, monkeypatch):
    """`-beamsearch -kwargs lm_path=... lm_alpha=...` (run_dynamic_eval_full.py:56-65): the CLI builds the n-gram
    BeamSearch itself, uses it for the in-loop pseudo-labels (lm_tta_beams) and the final decode (lm_eval_beams),
    and the result equals passing the same decoder programmatically."""
    from dae import ctc_beam_search, lib, run_dynamic_eval_full as r
    from dae.ngram import write_synthetic_arpa
    from dae.standin import SyntheticTokenizer, synthetic_recordings
    V = TOY["C"] - 1
    tok = SyntheticTokenizer(vocab_size=V, seed=0)
    data = synthetic_recordings("tedlium", tokenizer=tok, scale=0.004)[:1]
    arpa = str(tmp_path / "lm.arpa")
    write_synthetic_arpa(arpa, V, order=3, counts=(None, 400, 600), seed=2)
    argv = ['-beamsearch', '-seq', '1024', '-o', '512', '-d', 'tedlium', '-kwargs', f'lm_path={arpa}', 'lm_alpha=0.4016',
            'lm_beta=1.625', 'lm_prune_less_than_val=3.221', 'lm_eval_beams=5', 'lm_tta_beams=3', 'optim_lr=1e-3',
            'spec_augment_n_freq_masks=2', 'spec_augment_freq_mask_param=10']
    widths = []
    orig = ctc_beam_search.BeamSearch.__init__

    def spy(self, *a, **k):
        widths.append((k.get('beam_width'), k.get('alpha'), k.get('beta'), k.get('prune_less_than_val')))
        return orig(self, *a, **k)
    monkeypatch.setattr(ctc_beam_search.BeamSearch, "__init__", spy)
    texts = {}
    for how in ("cli", "programmatic"):
        args = lib.apply_args(r.build_parser(), argv if how == "cli" else [a for a in argv if a != '-beamsearch'])
        args.config = TOY_CONFIG
        model = ToyModel(TOY["C"], seed=TOY["model_seed"])
        captured = []
        bs = None if how == "cli" else lib.load_beamsearch(arpa, alpha=0.4016, beta=1.625, prune_less_than_val=3.221,
                                                           tokenizer=tok)
        wer = r.main(args, model=model, tokenizer=tok, data=data, normalize=lambda s: captured.append(s) or s,
                     beamsearch=bs)
        assert np.isfinite(wer)
        texts[how] = (captured, wer)
    assert texts["cli"] == texts["programmatic"]
    n_windows = sum(1 for w in widths if w[0] == 3)
    assert n_windows >= 2 and sum(1 for w in widths if w[0] == 5) == 2          # in-loop beams 3, final decode beams 5
    assert all(w[1:] == (0.4016, 1.625, 3.221) for w in widths)


def test_adapt_on_concat_only_equals_return_params(cuda):
    """run_half_concat_eval.py:64-160: the adapt-only pass yields the same parameters as return_params=True."""
    from dae import lib
    from dae.optim import MADGRAD
    from dae.standin import SyntheticTokenizer
    tok = SyntheticTokenizer(vocab_size=TOY["C"] - 1, seed=0)
    spec = toy_spec(4, 2600)
    outs = []
    for mode in ("full", "adapt_only"):
        model = ToyModel(TOY["C"], seed=3).to(cuda)
        model.device = cuda
        before = [p.detach().clone() for p in model.parameters()]
        args = make_args(TOY_CONFIG, seq_len=1024, awmc=False, **dict(TOY["kwargs"], epochs=1))
        random.seed(5)
        torch.manual_seed(5)
        if mode == "full":
            _, params = lib.dynamic_eval(args, model, spec, 1024, 512, tok, use_tqdm=False, optim=MADGRAD,
                                         return_params=True)
        else:
            params = lib.adapt_on_concat_only(args, model, spec, tok, adapt_overlap=512)
        assert all(torch.equal(a, b) for a, b in zip(before, model.parameters()))
        outs.append(params)
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_cross_dataset_driver_matches_sequential_reference_flow(cuda, tmp_path):
    """run_cross_dataset_eval (lcasr/run_cross_dataset_eval.py:96-212): baselines by `epochs=0`, adapt on A[i] with
    return_params, evaluate on B and A leave-one-out, parameters restored; result schema as the reference pickles."""
    import pickle
    from types import SimpleNamespace
    from dae import lib, run_cross_dataset_eval as x
    from dae.optim import MADGRAD
    from dae.standin import SyntheticTokenizer, synthetic_recordings
    from dae.wer import rates_from_counts, word_error_counts
    tok = SyntheticTokenizer(vocab_size=TOY["C"] - 1, seed=0)
    data_a = synthetic_recordings("earnings22", tokenizer=tok, scale=0.0015)[:2]
    data_b = synthetic_recordings("tedlium", tokenizer=tok, scale=0.003)[:2]
    model = ToyModel(TOY["C"], seed=TOY["model_seed"])
    save = str(tmp_path / "x.pkl")
    args = SimpleNamespace(dataset="earnings22", dataset2="tedlium", repeats=1, save_path=save, seq_len=1024, overlap=512,
                           adapt_overlap=None, awmc=False, config=TOY_CONFIG, **dict(TOY["kwargs"], epochs=1))
    before = [p.detach().clone() for p in model.parameters()]
    res = x.main(args, model, tok, data_a, data_b)[0]
    assert all(torch.equal(a, b.to(a.device)) for a, b in zip(model.parameters(), before))
    assert set(res) >= {"a_baseline", "b_baseline", "a_to_b", "a_to_a_loo", "dataset_a", "dataset_b", "args_dict", "repeat"}
    assert len(res["a_to_b"]) == 2 and len(res["a_to_a_loo"]) == 2
    assert pickle.load(open(save.replace(".pkl", "_1.pkl"), "rb"))["a_baseline"] == res["a_baseline"]
    # the b baseline equals plain epochs=0 inference on every B recording
    base = SimpleNamespace(**dict(vars(args), epochs=0))
    hyp, gold = [], []
    for r in data_b:
        spec, g_ = r["process_fn"](r)
        hyp.append(tok.decode(lib.dynamic_eval(base, model, spec, 1024, 512, tok, use_tqdm=False, optim=MADGRAD,
                                               output="greedy")).lower())
        gold.append(g_)
    assert rates_from_counts(word_error_counts(hyp, gold))[0] == res["b_baseline"]["wer"]
    assert res["a_to_b"][0]["words"] == res["b_baseline"]["words"]


def test_within_recording_loo_driver(cuda, tmp_path):
    """run_within_recording_loo_eval (lcasr/run_within_recording_loo_eval.py:103-237): outer chunks, adapt on one,
    no-adapt inference on the audio-disjoint ones, probability-space average at the downsampled positions.  Checked
    against the same flow spelled out with plain lib.dynamic_eval calls and a host-side accumulation."""
    from types import SimpleNamespace
    from dae import lib, run_within_recording_loo_eval as w
    from dae.optim import MADGRAD
    from dae.standin import SyntheticTokenizer
    tok = SyntheticTokenizer(vocab_size=TOY["C"] - 1, seed=0)
    args = SimpleNamespace(dataset="toy", repeats=1, save_path="", seq_len=1024, overlap=512, loo_seq_len=2048,
                           loo_overlap=1024, awmc=False, config=TOY_CONFIG, **dict(TOY["kwargs"], epochs=1))
    spec = toy_spec(6, 5200)
    model = ToyModel(TOY["C"], seed=TOY["model_seed"]).to(cuda)
    model.device = cuda
    before = [p.detach().clone() for p in model.parameters()]
    random.seed(3)
    torch.manual_seed(3)
    got, info = w.loo_eval(args, model, spec, tok)
    assert info["mode"] == "loo" and info["n_chunks"] >= 3
    assert all(torch.equal(a, b) for a, b in zip(model.parameters(), before))
    # the same flow by hand
    random.seed(3)
    torch.manual_seed(3)
    base = SimpleNamespace(**dict(vars(args), epochs=0))
    chunks, keys = lib.prepare_chunks(spec, 2048, 1024)
    keys = sorted(keys)
    clen = {k: chunks[k].shape[-1] for k in keys}
    acc = np.zeros((5200 // 8 + 2048, TOY["C"]), np.float64)
    cnt = np.zeros(5200 // 8 + 2048)
    for a in keys:
        ev = [e for e in keys if e >= a + clen[a] or a >= e + clen[e]]
        if not ev:
            continue
        for p, u in zip(model.parameters(), before):
            p.data = u.data.clone()
        _, upd = lib.dynamic_eval(args, model, chunks[a], 1024, 512, tok, use_tqdm=False, optim=MADGRAD, return_params=True)
        for p, u in zip(model.parameters(), upd):
            p.data = u.data.to(p.device)
        for e in ev:
            lp = lib.dynamic_eval(base, model, chunks[e], 1024, 512, tok, use_tqdm=False, optim=MADGRAD)
            acc[e // 8:e // 8 + lp.shape[0]] += np.exp(lp.astype(np.float64))
            cnt[e // 8:e // 8 + lp.shape[0]] += 1
    for p, u in zip(model.parameters(), before):
        p.data = u.data.clone()
    ref = np.log(acc[cnt != 0] / cnt[cnt != 0][:, None])
    assert got.shape == ref.shape
    np.testing.assert_allclose(np.exp(got.cpu().numpy()), np.exp(ref), rtol=2e-4, atol=1e-7)
    # entry point: schema + sharding plumbing on two tiny recordings
    from dae.standin import synthetic_recordings
    data = synthetic_recordings("tedlium", tokenizer=tok, scale=0.004)[:2]
    res = w.main(args, ToyModel(TOY["C"], seed=TOY["model_seed"]), tok, data)[0]
    assert set(res) >= {"loo", "baseline", "model_output", "baseline_model_output", "gold", "per_recording_meta", "repeat"}
    assert len(res["model_output"]) == 2 and np.isfinite(res["loo"]["wer"])


def test_split_branches_equals_batched_call(cuda):
    """The teacher branch without a graph + the augmented branch alone (default) against the reference's single
    [augmented, clean] batch (`split_branches=False`): same pseudo-labels at every step (bit-exact ids), same
    stitched posteriors and adapted parameters up to GEMM summation order — the clean row's upstream gradient is
    exactly zero, so dropping its backward changes nothing mathematically."""
    from dae import lib
    from dae.optim import MADGRAD
    from dae.standin import SyntheticTokenizer
    outs = []
    for split in (True, False):
        tok = RecordingTokenizer(SyntheticTokenizer(vocab_size=TOY["C"] - 1, seed=0))
        model = ToyModel(TOY["C"], seed=TOY["model_seed"]).to(cuda)
        model.device = cuda
        args = make_args(TOY_CONFIG, split_branches=split, **TOY["kwargs"])
        random.seed(TOY["seed"])
        torch.manual_seed(TOY["seed"])
        logits, params = lib.dynamic_eval(args, model, toy_spec(TOY["spec_seed"], TOY["spec_n"]), TOY["seq_len"],
                                          TOY["overlap"], tok, use_tqdm=False, optim=MADGRAD, return_params=True)
        outs.append((logits, tok.encoded, params))
    (la, ea, pa), (lb, eb, pb) = outs
    assert ea == eb
    # fp32 GEMM summation order differs between a [1,..] and a [2,..] batch; ten MADGRAD steps of the sharp toy model
    # amplify that to ~4e-4 of a probability (the oracle-loop comparisons above allow 2e-3 for the same reason)
    np.testing.assert_allclose(np.exp(la), np.exp(lb), rtol=2e-3, atol=1e-6)
    for a, b in zip(pa, pb):
        torch.testing.assert_close(a, b, rtol=2e-3, atol=1e-5)
