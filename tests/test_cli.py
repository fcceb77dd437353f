"""CLI surface: lib.apply_args (lcasr/lib.py:1756-1787) and the -beamsearch wiring of run_dynamic_eval_full.main
(lcasr/run_dynamic_eval_full.py:56-65)."""
import argparse

import pytest


def _parse(argv):
    from dae import lib
    from dae.run_dynamic_eval_full import build_parser
    return lib.apply_args(build_parser(), argv)


def test_apply_args_defaults_match_reference():
    a = _parse([])
    assert (a.seq_len, a.overlap, a.epochs, a.split, a.checkpoint, a.log) == (16384, 14336, 1, 'test', '', '')
    assert a.shuffle is True and a.verbose is True                    # inverses of -ds / -nv (:1783-1784)
    assert a.beamsearch is False and a.awmc is False and a.consistency is False
    assert a.freeze_subsampling is False and a.freeze_all_but_last_block_and_head is False
    assert a.train_subsampling_only is False and a.kwargs == []
    assert (a.dataset, a.repeats, a.save_path) == ('earnings22', 1, '')


def test_apply_args_flags_and_kwargs():
    a = _parse(['-dfa', '-epochs', '5', '-seq', '8192', '-o', '4096', '-d', 'tedlium', '-ds', '-nv', '-beamsearch',
                '-awmc', '-r', '3', '-s', 'out.pkl', '-log', 'l.txt', '-split', 'dev', '-c', 'ck.pt',
                '--freeze_subsampling', '-kwargs', 'optim_lr=9e-5', 'spec_augment_n_freq_masks=6',
                'spec_augment_freq_mask_param=34', 'spec_augment_n_time_masks=0', 'cutout_value=mean',
                'online=True', 'lm_alpha=0.4016', 'lm_path=/tmp/x=y.arpa', 'sizes=[1,2]'])
    assert a.disable_flash_attention and a.epochs == 5 and a.seq_len == 8192 and a.overlap == 4096
    assert a.dataset == 'tedlium' and a.shuffle is False and a.verbose is False and a.beamsearch and a.awmc
    assert a.repeats == 3 and a.save_path == 'out.pkl' and a.log == 'l.txt' and a.split == 'dev' and a.checkpoint == 'ck.pt'
    assert a.freeze_subsampling is True
    d = a.__dict__
    assert d['optim_lr'] == 9e-5 and isinstance(d['optim_lr'], float)
    assert d['spec_augment_n_freq_masks'] == 6 and isinstance(d['spec_augment_n_freq_masks'], int)
    assert d['online'] is True and d['lm_alpha'] == 0.4016 and d['sizes'] == [1, 2]
    assert d['cutout_value'] == 'mean'                                # not a literal: stays a string
    assert d['lm_path'] == '/tmp/x=y.arpa'                            # split on the first '=' only
    from dae import lib
    assert lib.get_specaugment_config_from_args(a) == {'n_time_masks': 0, 'n_freq_masks': 6, 'freq_mask_param': 34,
                                                       'time_mask_param': -1, 'min_p': 0.05, 'zero_masking': False}
    assert lib.get_lr_args_from_args(a) == {'lr': 9e-5}
    assert lib.get_cutout_params_from_args(a, 16384)['cutout_val'] == 'mean'


def test_apply_args_never_evaluates_code(tmp_path):
    marker = tmp_path / "pwned"
    a = _parse(['-kwargs', f'x=__import__("os").system("touch {marker}")'])
    assert isinstance(a.__dict__['x'], str) and not marker.exists()  # the reference eval()s this (:1778-1781)


def test_beamsearch_flag_without_lm_raises(monkeypatch, tmp_path):
    """-beamsearch must not silently decode greedy: without an LM path main() refuses."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from toy import TOY, TOY_CONFIG, ToyModel
    from dae import run_dynamic_eval_full as r
    from dae.standin import SyntheticTokenizer, synthetic_recordings
    monkeypatch.chdir(tmp_path)                                       # no paths.yaml here
    monkeypatch.delenv("DAE_PATHS", raising=False)
    args = _parse(['-beamsearch', '-epochs', '0'])
    args.config = TOY_CONFIG
    tok = SyntheticTokenizer(vocab_size=TOY["C"] - 1, seed=0)
    data = synthetic_recordings("tedlium", tokenizer=tok, scale=0.004)[:1]
    with pytest.raises(ValueError, match="lm_path"):
        r.main(args, model=ToyModel(TOY["C"]), tokenizer=tok, data=data)


def test_lm_path_from_paths_yaml(monkeypatch, tmp_path):
    from dae import lib
    p = tmp_path / "paths.yaml"
    p.write_text("checkpoints:\n  lm: /models/4gram.arpa.gz\n")
    monkeypatch.setenv("DAE_PATHS", str(p))
    assert lib.lm_path_from_paths() == "/models/4gram.arpa.gz"
    monkeypatch.setenv("DAE_PATHS", str(tmp_path / "missing.yaml"))
    assert lib.lm_path_from_paths() is None
