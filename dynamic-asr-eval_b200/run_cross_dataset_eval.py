"""Transfer / leave-one-out driver with the flow and result schema of lcasr/run_cross_dataset_eval.py (:31-212),
run as a batch workload over all ranks.

The reference issues O(|A|*(|A|+|B|)) `epochs=0` inference calls one after the other on one GPU (:92-94,147-195).
Here every such call is `dae.lib.dynamic_eval(epochs=0, output='greedy')` — a pure GPU pipeline (final pass, stitch,
fused argmax, collapse) — and the calls are sharded:
  * baselines on A and B (:103-137): recordings LPT-sharded over the ranks, one int64[5] all-reduce each;
  * the A-X sweep (:139-198): the adaptation recordings A[i] are dealt to the ranks; a rank adapts on its A[i]
    (`return_params=True`, :142-151), loads the updated parameters (:152-153), evaluates on every B[j] and every
    A[k != i] without adapting, and restores the original parameters (:197-198).  WER tuples are gathered on rank 0.
Integer error counts make every WER independent of the shard layout.
"""
import argparse
import pickle
import random
import zlib

import torch

from . import lib
from .lib import AWMC, dynamic_eval
from .shard import all_reduce_counts, gather_objects, init_distributed, lpt_assign
from .wer import rates_from_counts, word_error_counts


def _wer_dict(counts):
    wer, words, ins_rate, del_rate, sub_rate = rates_from_counts(counts)
    return {"wer": wer, "words": words, "ins_rate": ins_rate, "del_rate": del_rate, "sub_rate": sub_rate}


def main(args, model, tokenizer, data_a, data_b, normalize=None, beamsearch=None):
    rank, world, local = init_distributed()
    device = torch.device('cuda', local)
    model.device = device
    model = model.to(device).eval()
    normalize = normalize or (lambda s: s)
    beams = args.__dict__.get('lm_eval_beams', 20)
    eval_fn = dynamic_eval if not getattr(args, 'awmc', False) else AWMC
    adapt_overlap = args.adapt_overlap if getattr(args, 'adapt_overlap', None) is not None else args.overlap
    args_dict = vars(args).copy()
    args_dict['epochs'] = 0                                   # :92-94: no-adapt inference
    baseline_args = argparse.Namespace(**args_dict)
    original = [p.detach().clone() for p in model.parameters()]

    def transcribe(rec):
        """One `epochs=0` call + decode (:81-90,105-118)."""
        audio_spec, gold = rec['process_fn'](rec)
        if beamsearch is None:
            ids = eval_fn(baseline_args, model, audio_spec, args.seq_len, args.overlap, tokenizer, use_tqdm=False,
                          output='greedy')
            text = tokenizer.decode(ids)
        else:
            logits = eval_fn(baseline_args, model, audio_spec, args.seq_len, args.overlap, tokenizer, use_tqdm=False,
                             beam_search_fn=beamsearch, output='device')
            bs = beamsearch(log_probs=logits, beam_width=beams)
            bs.run_search(use_tqdm=False)
            text = bs.return_text(idx=0)
        return normalize(text).lower(), gold

    def sharded_baseline(data):
        mine = lpt_assign([int(r.get('frames', 1)) for r in data], world)[rank]
        pairs = [transcribe(data[i]) for i in mine]
        counts = word_error_counts([p for p, _ in pairs], [g for _, g in pairs])
        return _wer_dict(all_reduce_counts(counts, device))

    out = []
    for repeat in range(args.repeats):
        a_baseline = sharded_baseline(data_a)
        b_baseline = sharded_baseline(data_b)
        mine = list(range(rank, len(data_a), world))         # adaptation recordings of this rank
        local_res = {}
        for i in mine:
            audio_spec, _ = data_a[i]['process_fn'](data_a[i])
            key = zlib.crc32(f"{args.__dict__.get('seed', 0)}|{repeat}|{data_a[i]['id']}".encode())
            random.seed(key)
            torch.manual_seed(key ^ 0x5bd1e995)
            _, updated = eval_fn(args, model, audio_spec, args.seq_len, adapt_overlap, tokenizer, use_tqdm=False,
                                 beam_search_fn=beamsearch, return_params=True, output='device')
            with torch.no_grad():
                for p, u in zip(model.parameters(), updated):
                    p.data = u.data.to(p.device)
            pb = [transcribe(r) for r in data_b]
            pa = [transcribe(data_a[k]) for k in range(len(data_a)) if k != i]
            local_res[i] = (_wer_dict(word_error_counts([p for p, _ in pb], [g for _, g in pb])),
                            _wer_dict(word_error_counts([p for p, _ in pa], [g for _, g in pa])))
            with torch.no_grad():
                for p, u in zip(model.parameters(), original):
                    p.data = u.data.clone()
        merged = {}
        for part in gather_objects(local_res):
            merged.update(part)
        results = {
            'a_baseline': a_baseline, 'b_baseline': b_baseline,
            'a_to_b': [merged[i][0] for i in sorted(merged)], 'a_to_a_loo': [merged[i][1] for i in sorted(merged)],
            'dataset_a': getattr(args, 'dataset', ''), 'dataset_b': getattr(args, 'dataset2', ''),
            'args_dict': {k: v for k, v in vars(args).items() if k != 'config'}, 'repeat': f'{repeat + 1}/{args.repeats}',
        }
        if rank == 0 and getattr(args, 'save_path', '') != '':
            save_path = args.save_path
            save_path = save_path.replace('.pkl', f'_{repeat + 1}.pkl') if save_path.endswith('.pkl') \
                else save_path + f'_{repeat + 1}.pkl'
            with open(save_path, 'wb') as f:
                pickle.dump(results, f)
        out.append(results)
    return out


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--dataset', '-d', type=str, default='earnings22')
    parser.add_argument('--dataset2', '-d2', type=str, default='tedlium')
    parser.add_argument('--repeats', '-r', type=int, default=1)
    parser.add_argument('--save_path', '-s', type=str, default='')
    parser.add_argument('--adapt_overlap', '-ao', type=int, default=None)
    return parser


if __name__ == '__main__':
    from . import standin
    a = lib.apply_args(build_parser())
    tok = standin.SyntheticTokenizer()
    a.config = standin.default_config()
    scale = a.__dict__.get('synthetic_scale', 1.0)
    main(a, standin.build_model(tok.vocab_size()), tok, standin.synthetic_recordings(a.dataset, tokenizer=tok, scale=scale),
         standin.synthetic_recordings(a.dataset2, tokenizer=tok, scale=scale))
