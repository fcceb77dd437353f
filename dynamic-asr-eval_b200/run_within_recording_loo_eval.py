"""Within-recording leave-one-out driver with the flow and result schema of lcasr/run_within_recording_loo_eval.py
(:31-237), on the dae GPU pipeline and sharded over ranks.

Two-level windowing (:103-181): the recording is cut into outer chunks of ``loo_seq_len`` frames
(``lib.prepare_chunks``); for every outer chunk i the model adapts on it (``eval_fn(..., return_params=True)``) and
then transcribes, WITHOUT adapting (``epochs=0``), every outer chunk j whose audio does not overlap chunk i; the
chunk posteriors are averaged in probability space at their downsampled positions and the covered rows are decoded.
Here every ``epochs=0`` call is `dae.lib.dynamic_eval(output='device')` (final pass, dae_stitch); the outer
accumulation stays on the device; recordings are LPT-sharded over the ranks and the two WER tuples (LOO, baseline)
come from all-reduced integer counts.
"""
import argparse
import copy
import pickle
import random
import zlib

import torch

from . import lib
from .greedy import GreedyCTCDecoder
from .lib import AWMC, dynamic_eval, prepare_chunks
from .shard import all_reduce_counts, gather_objects, init_distributed, lpt_assign
from .wer import rates_from_counts, word_error_counts


def loo_eval(args, model, audio_spec, tokenizer, eval_fn=dynamic_eval, beamsearch=None, original=None):
    """One recording (:103-181) -> (log-probs [N, C] on the device, info dict)."""
    device = model.device
    baseline_args = copy.copy(args)
    baseline_args.epochs = 0
    ds = args.config['model']['subsampling_factor']
    C = tokenizer.vocab_size() + 1
    original = original if original is not None else [p.detach().clone() for p in model.parameters()]

    def restore():
        with torch.no_grad():
            for p, u in zip(model.parameters(), original):
                p.data = u.data.clone()

    def windowed_inference(chunk):
        return eval_fn(baseline_args, model, chunk, args.seq_len, args.overlap, tokenizer, use_tqdm=False,
                       beam_search_fn=beamsearch, output='device')

    spec_n = audio_spec.shape[-1]
    chunks, keys = prepare_chunks(audio_spec, args.loo_seq_len, args.loo_overlap)
    keys = sorted(keys)
    if len(keys) <= 1:
        return windowed_inference(audio_spec), {'n_chunks': len(keys), 'mode': 'fallback_windowed_eval'}
    clen = {k: chunks[k].shape[-1] for k in keys}
    valid = {a: [e for e in keys if e >= a + clen[a] or a >= e + clen[e]] for a in keys}     # audio-disjoint (:118-121)
    if sum(len(v) for v in valid.values()) == 0:
        return windowed_inference(audio_spec), {'n_chunks': len(keys), 'mode': 'fallback_no_disjoint_pairs'}
    rows = spec_n // ds + args.loo_seq_len
    acc = torch.zeros((rows, C), dtype=torch.float32, device=device)
    cnt = torch.zeros((rows,), dtype=torch.float32, device=device)
    for a in [k for k in keys if valid[k]]:
        restore()
        _, updated = eval_fn(args, model, chunks[a], args.seq_len, args.overlap, tokenizer, use_tqdm=False,
                             beam_search_fn=beamsearch, return_params=True, output='device')
        with torch.no_grad():
            for p, u in zip(model.parameters(), updated):
                p.data = u.data.to(p.device)
        for e in valid[a]:
            lp = windowed_inference(chunks[e])
            pos, n = e // ds, lp.shape[0]
            acc[pos:pos + n] += lp.exp()
            cnt[pos:pos + n] += 1
    restore()
    covered = cnt != 0
    if not bool(covered.any()):
        raise RuntimeError('LOO stitching produced no coverage at any position.')
    out = torch.log(acc[covered] / cnt[covered][:, None])
    return out, {'n_chunks': len(keys), 'mode': 'loo'}


def main(args, model, tokenizer, data, normalize=None, beamsearch=None):
    assert args.loo_seq_len > args.loo_overlap, 'loo_seq_len must be greater than loo_overlap'
    assert args.loo_seq_len >= args.seq_len, 'loo_seq_len should be >= inner seq_len'
    rank, world, local = init_distributed()
    device = torch.device('cuda', local)
    model.device = device
    model = model.to(device).eval()
    normalize = normalize or (lambda s: s)
    blank = model.decoder.num_classes - 1
    decoder = GreedyCTCDecoder(tokenizer=tokenizer, blank_id=blank)
    beams = args.__dict__.get('lm_eval_beams', 20)
    eval_fn = dynamic_eval if not getattr(args, 'awmc', False) else AWMC
    original = [p.detach().clone() for p in model.parameters()]
    baseline_args = copy.copy(args)
    baseline_args.epochs = 0

    def transcribe(logits):
        if beamsearch is None:
            text = decoder(logits)
        else:
            bs = beamsearch(log_probs=logits, beam_width=beams)
            bs.run_search(use_tqdm=False)
            text = bs.return_text(idx=0)
        return normalize(text).lower()

    mine = lpt_assign([int(r.get('frames', 1)) for r in data], world)[rank]
    out = []
    for repeat in range(args.repeats):
        res = {}
        for idx in mine:
            rec = data[idx]
            audio_spec, gold = rec['process_fn'](rec)
            key = zlib.crc32(f"{args.__dict__.get('seed', 0)}|{repeat}|{rec['id']}".encode())
            random.seed(key)
            torch.manual_seed(key ^ 0x5bd1e995)
            base = transcribe(eval_fn(baseline_args, model, audio_spec, args.seq_len, args.overlap, tokenizer,
                                      use_tqdm=False, beam_search_fn=beamsearch, output='device'))
            stitched, info = loo_eval(args, model, audio_spec, tokenizer, eval_fn, beamsearch, original)
            res[idx] = (transcribe(stitched), base, gold, {'id': rec['id'], **info})
        ids = sorted(res)
        loo = rates_from_counts(all_reduce_counts(word_error_counts([res[i][0] for i in ids], [res[i][2] for i in ids]), device))
        basew = rates_from_counts(all_reduce_counts(word_error_counts([res[i][1] for i in ids], [res[i][2] for i in ids]), device))
        merged = {}
        for part in gather_objects(res):
            merged.update(part)
        order = sorted(merged)
        names = ('wer', 'words', 'ins_rate', 'del_rate', 'sub_rate')
        save_data = {
            'loo': dict(zip(names, loo)), 'baseline': dict(zip(names, basew)),
            'model_output': [merged[i][0] for i in order], 'baseline_model_output': [merged[i][1] for i in order],
            'gold': [merged[i][2] for i in order], 'per_recording_meta': [merged[i][3] for i in order],
            'dataset': getattr(args, 'dataset', ''), 'args_dict': {k: v for k, v in vars(args).items() if k != 'config'},
            'repeat': f'{repeat + 1}/{args.repeats}',
        }
        if rank == 0 and getattr(args, 'save_path', '') != '':
            save_path = args.save_path
            save_path = save_path.replace('.pkl', f'_{repeat + 1}.pkl') if save_path.endswith('.pkl') \
                else save_path + f'_{repeat + 1}.pkl'
            with open(save_path, 'wb') as f:
                pickle.dump(save_data, f)
        out.append(save_data)
    return out


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--dataset', '-d', type=str, default='earnings22')
    parser.add_argument('--repeats', '-r', type=int, default=1)
    parser.add_argument('--save_path', '-s', type=str, default='')
    parser.add_argument('--loo_seq_len', '-loo_s', type=int, default=65536)
    parser.add_argument('--loo_overlap', '-loo_o', type=int, default=57344)
    return parser


if __name__ == '__main__':
    from . import standin
    a = lib.apply_args(build_parser())
    tok = standin.SyntheticTokenizer()
    a.config = standin.default_config()
    main(a, standin.build_model(tok.vocab_size()), tok,
         standin.synthetic_recordings(a.dataset, tokenizer=tok, scale=a.__dict__.get('synthetic_scale', 1.0)))
