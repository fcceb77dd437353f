"""Entry point with the CLI and flow of lcasr/run_dynamic_eval_full.py (:31-159), sharded over GPUs.

Differences from the reference, all outside the numerics: the checkpoint/model/dataset loaders
are pluggable (``main(args, model=..., tokenizer=..., data=...)``; without them a synthetic
stand-in is built, since neither `lcasr` nor any audio exists here), recordings are LPT-sharded
over ranks, and the per-repeat WER comes from all-reduced integer counts.
"""
import argparse
import pickle
import random
import time
import zlib

import torch

from . import lib
from .greedy import GreedyCTCDecoder
from .lib import dynamic_eval
from .shard import all_reduce_counts, gather_objects, init_distributed, lpt_assign
from .wer import rates_from_counts, word_error_counts


def _default_normalize():
    try:                                   # the reference uses whisper's EnglishTextNormalizer (:8,20)
        from transformers.models.whisper.english_normalizer import EnglishTextNormalizer
        return EnglishTextNormalizer({})
    except Exception:                      # pragma: no cover
        return lambda s: s


datasets_functions = {}                     # name -> callable(split) -> [{id,text,audio,process_fn}]


def main(args, model=None, tokenizer=None, data=None, normalize=None, beamsearch=None):
    assert args.split in ['test', 'dev'], f'Split must be either test or dev (got {args.split})'
    rank, world, local = init_distributed()
    device = torch.device('cuda', local) if torch.cuda.is_available() else torch.device('cpu')
    if model is None:
        from . import standin
        tokenizer = tokenizer or standin.SyntheticTokenizer()
        args.config = standin.default_config()
        model = standin.build_model(tokenizer.vocab_size(), device=device)
    model.device = device
    model = model.to(device)
    model.eval()
    if data is None:
        from . import standin
        data = datasets_functions[args.dataset](args.split) if args.dataset in datasets_functions else \
            standin.synthetic_recordings(args.dataset, tokenizer=tokenizer, scale=args.__dict__.get('synthetic_scale', 1.0))
    normalize = normalize or _default_normalize()
    blank = model.decoder.num_classes - 1
    decoder = GreedyCTCDecoder(tokenizer=tokenizer, blank_id=blank)
    if args.__dict__.get('beamsearch', False) and beamsearch is None:
        # run_dynamic_eval_full.py:56-64: the reference reads the LM checkpoint path from paths.yaml
        # (lib.paths.checkpoints.lm); here it is an ARPA file named by `-kwargs lm_path=...` or lib.paths
        lm_path = args.__dict__.get('lm_path', None) or lib.lm_path_from_paths()
        if not lm_path:
            raise ValueError("-beamsearch needs an ARPA language model: pass `-kwargs lm_path=/path/to/lm.arpa[.gz]` "
                             "or set checkpoints.lm in paths.yaml (DAE_PATHS)")
        beamsearch = lib.load_beamsearch(
            path=lm_path,
            alpha=args.__dict__.get('lm_alpha', 0.45),
            beta=args.__dict__.get('lm_beta', 1.53),
            prune_less_than_val=args.__dict__.get('lm_prune_less_than_val', 3.17),
            top_am_threshold=args.__dict__.get('lm_top_am_threshold', -6),
            tokenizer=tokenizer,
        )
    beams = args.__dict__.get('lm_eval_beams', 20)
    if args.awmc:
        eval_fn = lib.AWMC
    elif getattr(args, 'consistency', False):
        raise NotImplementedError("consistency variant is out of scope (SURVEY.md §2 row 7)")
    else:
        eval_fn = dynamic_eval

    costs = [int(r.get('frames', 1)) for r in data]
    mine = lpt_assign(costs, world)[rank]
    avg_wers = []
    for repeat in range(args.repeats):
        texts, golds, elapsed = {}, {}, {}
        for rec in mine:
            audio_spec, gold_text = data[rec]['process_fn'](data[rec])
            # shuffle order must not depend on the shard layout
            key = zlib.crc32(f"{args.__dict__.get('seed', 0)}|{repeat}|{data[rec]['id']}".encode())
            random.seed(key)
            torch.manual_seed(key ^ 0x5bd1e995)
            stime = time.time()
            if beamsearch is None:
                ids = eval_fn(args, model, audio_spec, args.seq_len, args.overlap, tokenizer,
                              beam_search_fn=None, use_tqdm=False, output='greedy')
                out_text = tokenizer.decode(ids)
            else:
                logits = eval_fn(args, model, audio_spec, args.seq_len, args.overlap, tokenizer,
                                 beam_search_fn=beamsearch, use_tqdm=False, output='device')
                run_beam_search = beamsearch(log_probs=logits, beam_width=beams)
                run_beam_search.run_search(use_tqdm=False)
                out_text = run_beam_search.return_text(idx=0)
            elapsed[rec] = time.time() - stime
            texts[rec] = normalize(out_text).lower()
            golds[rec] = gold_text
        counts = word_error_counts([texts[r] for r in mine], [golds[r] for r in mine])
        total = all_reduce_counts(counts, device if device.type == 'cuda' else None)
        wer, words, ins_rate, del_rate, sub_rate = rates_from_counts(total)
        if rank == 0:
            print(f'WER: {wer}')
        parts = gather_objects((texts, golds, elapsed))
        if rank == 0:
            all_t, all_g, all_e = {}, {}, {}
            for t, g, e in parts:
                all_t.update(t), all_g.update(g), all_e.update(e)
            order = sorted(all_t)
            if args.log != '':
                with open(args.log, 'a') as f:
                    f.write(f'{args.checkpoint}\t overlap: {args.overlap}\t seq_len: {args.seq_len}\t WER: {wer}\n')
            if args.save_path != '':
                save_data = {'wer': wer, 'words': words, 'ins_rate': ins_rate, 'del_rate': del_rate,
                             'sub_rate': sub_rate, 'model_output': [all_t[k] for k in order],
                             'gold': [all_g[k] for k in order], 'elapsed_times': [all_e[k] for k in order],
                             'args_dict': {k: v for k, v in vars(args).items() if k != 'config'},
                             'repeat': f'{repeat + 1}/{args.repeats}'}
                save_path = args.save_path
                save_path = save_path.replace('.pkl', f'_{repeat + 1}.pkl') if save_path.endswith('.pkl') \
                    else save_path + f'_{repeat + 1}.pkl'
                with open(save_path, 'wb') as f:
                    pickle.dump(save_data, f)
        avg_wers.append(wer)
    avg = sum(avg_wers) / len(avg_wers)
    if rank == 0:
        print(f'Average WER: {avg}')
    return avg


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--dataset', '-d', type=str, default='earnings22')
    parser.add_argument('--repeats', '-r', type=int, default=1, help='Number of times to repeat the evaluation')
    parser.add_argument('--save_path', '-s', type=str, default='', help='path to save')
    return parser


if __name__ == '__main__':
    main(lib.apply_args(build_parser()))
