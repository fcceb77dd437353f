"""Generate tests/golden/loop_toy.npz by running the REFERENCE's own lcasr/lib.py dynamic_eval
(imported read-only from /root/reference under import stubs; nothing is copied) on the toy model.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_loop_golden.py
The three behavioural stand-ins are oracle.ref_loop.OracleSpecAugment, OracleGreedy and
dae.optim.MADGRAD (SURVEY.md §8c "Whole adapt loop").
"""
import contextlib
import io
import os
import random
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/lcasr"


class AttrDict(dict):
    def __getattr__(self, k):
        return self.get(k, AttrDict())

    def __fspath__(self):
        return "/nonexistent"


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    from oracle.ref_loop import OracleGreedy, OracleSpecAugment
    from dae.optim import MADGRAD
    stub("omegaconf", OmegaConf=types.SimpleNamespace(load=lambda p: AttrDict()))
    lc = stub("lcasr")
    stub("lcasr.utils", audio_tools=types.SimpleNamespace(load_tokenizer=lambda: None))
    lc.utils = sys.modules["lcasr.utils"]
    stub("lcasr.utils.augmentation", SpecAugment=OracleSpecAugment)
    stub("lcasr.decoding")
    stub("lcasr.decoding.greedy", GreedyCTCDecoder=OracleGreedy)
    stub("lcasr.optim", madgrad=types.SimpleNamespace(MADGRAD=MADGRAD))
    stub("lming")
    stub("lming.utils", general=types.SimpleNamespace())
    stub("lming.utils.helpers", exists=lambda x: x is not None)
    stub("lming.models")
    stub("lming.models.transformer", transformer_lm=object)
    if "matplotlib" not in sys.modules:
        stub("matplotlib", pyplot=types.SimpleNamespace())
        stub("matplotlib.pyplot")
    from oracle.ref_loop import OracleEMA
    stub("torch_ema", ExponentialMovingAverage=OracleEMA)
    stub("lcasr.utils.lm_tools", add_eos=None, token_lens_to_mask=None, mark_padding=None)
    stub("lcasr.components")
    stub("lcasr.components.batchrenorm", BatchRenorm1d=object)
    stub("lcasr.eval")
    stub("lcasr.eval.wer", word_error_rate_detail=None)


def main():
    from toy import TOY, TOY_CONFIG, RecordingTokenizer, ToyModel, toy_spec
    from dae.optim import MADGRAD
    from dae.standin import SyntheticTokenizer
    from oracle.ref_loop import make_args
    install_stubs()
    sys.path.insert(0, REF)
    import lib as ref_lib                                   # the reference file itself
    assert ref_lib.__file__.startswith(REF)
    out = {}
    for online in (False, True):
        tok = RecordingTokenizer(SyntheticTokenizer(vocab_size=TOY["C"] - 1, seed=0))
        model = ToyModel(TOY["C"], seed=TOY["model_seed"])
        before = [p.detach().clone() for p in model.parameters()]
        args = make_args(TOY_CONFIG, online=online, **TOY["kwargs"])
        random.seed(TOY["seed"])
        torch.manual_seed(TOY["seed"])
        with contextlib.redirect_stdout(io.StringIO()):
            logits = ref_lib.dynamic_eval(args, model, toy_spec(TOY["spec_seed"], TOY["spec_n"]), TOY["seq_len"], TOY["overlap"], tok,
                                          use_tqdm=False, optim=MADGRAD)
        assert all(torch.equal(a, b) for a, b in zip(before, model.parameters())), "params must be restored"
        tag = "online" if online else "offline"
        out[f"logits_{tag}"] = logits.astype(np.float32)
        out[f"n_steps_{tag}"] = np.int64(len(tok.encoded))
        out[f"ids_flat_{tag}"] = np.array([i for e in tok.encoded for i in e], dtype=np.int64)
        out[f"ids_len_{tag}"] = np.array([len(e) for e in tok.encoded], dtype=np.int64)
        out[f"min_margin_{tag}"] = np.float64(model.min_margin)
        print(tag, logits.shape, len(tok.encoded), [len(e) for e in tok.encoded], 'min top-2 margin', model.min_margin)
    # AWMC baseline (lib.py:206-376) through the same stubs
    tok = RecordingTokenizer(SyntheticTokenizer(vocab_size=TOY["C"] - 1, seed=0))
    model = ToyModel(TOY["C"], seed=TOY["model_seed"])
    args = make_args(TOY_CONFIG, **dict(TOY["kwargs"], ema_decay=0.9))
    random.seed(TOY["seed"])
    torch.manual_seed(TOY["seed"])
    with contextlib.redirect_stdout(io.StringIO()):
        logits = ref_lib.AWMC(args, model, toy_spec(TOY["spec_seed"], TOY["spec_n"]), TOY["seq_len"], TOY["overlap"], tok,
                              use_tqdm=False, optim=MADGRAD)
    out["logits_awmc"] = logits.astype(np.float32)
    out["ids_flat_awmc"] = np.array([i for e in tok.encoded for i in e], dtype=np.int64)
    out["ids_len_awmc"] = np.array([len(e) for e in tok.encoded], dtype=np.int64)
    print("awmc", logits.shape, len(tok.encoded))
    # cutout (lib.py:384-417): reference function with a fixed torch seed; the rectangle table is re-drawn
    # with the same seed by dae.augment.draw_cutout_rects in the tests
    for mode in ("mean", "mean_recording", "zero"):
        spec = toy_spec(3, 700).clone()
        torch.manual_seed(77)
        res = ref_lib.cutout(spec, seq_len=512, cutout_val=mode, num_rectangles=40, max_width=60, max_height=12)
        out[f"cutout_{mode}"] = res[0].numpy().astype(np.float32)
    # frame_shuffle (lib.py:81-84) and add_random_noise (:379-382): the reference functions under a fixed torch
    # seed; the tests re-draw the permutations / the normal field with the same seed on the host
    spec = toy_spec(4, 200).clone()
    for td_, fd_ in ((True, False), (False, True), (True, True)):
        torch.manual_seed(78)
        res = ref_lib.frame_shuffle(spec.clone(), time_dimension=td_, freq_dimension=fd_)
        out[f"frame_shuffle_t{int(td_)}f{int(fd_)}"] = res[0].numpy().astype(np.float32)
    torch.manual_seed(79)
    out["add_random_noise_0p3"] = ref_lib.add_random_noise(spec.clone(), 0.3)[0].numpy().astype(np.float32)
    # chunk index vectors from the reference's prepare_chunks at the BASELINE window settings
    for spec_n in (6000, 120000, 360000, 415990):
        td, keys = ref_lib.prepare_chunks(torch.zeros(1, 1, spec_n), 16384, 14336)
        out[f"chunks_{spec_n}"] = np.array([[k, td[k].shape[-1]] for k in keys], dtype=np.int64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "loop_toy.npz"), **out)
    print("wrote tests/golden/loop_toy.npz")


if __name__ == "__main__":
    main()
