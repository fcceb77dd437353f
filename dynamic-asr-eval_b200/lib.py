"""Dynamic-evaluation library with the reference's entry points (lcasr/lib.py), on the dae kernels.

``dynamic_eval`` / ``dynamic_eval_ctc_loss`` (lib.py:450-643), ``AWMC`` (:206-376),
``prepare_chunks`` (:128-145), the kwarg getters (:102-125,419-428), ``apply_args``
(:1756-1787) and ``load_beamsearch`` (:37-72) keep their signatures, so
``run_dynamic_eval_full.py`` works with ``import dae.lib as lib``.

What changed underneath (same arithmetic, different placement):
  * the spectrogram is copied to the GPU once; windows are device views (the reference builds and
    augments every window on the CPU and copies [2,80,16384] per step, lib.py:538-549);
  * the reference runs [augmented, clean] as ONE batch with autograd on (lib.py:539-550) although the loss only
    sees the augmented row (:570); here the clean (teacher) branch runs first without a graph and the augmented
    branch alone carries the gradient: same parameter gradient (the clean row's upstream gradient is exactly
    zero), one third less encoder work per step (``split_branches=False`` restores the single batched call);
  * SpecAugment masks straight from the window view (per-recording window sums + one single-pass launch);
  * pseudo-labels come from the greedy kernel on the device posteriors; only the collapsed ids
    cross PCIe for the tokenizer's decode -> re-encode round trip (lib.py:559,569 copy 33.5 MB
    twice per step to the host), and that hop runs while the GPU computes the augmented branch;
  * CTC loss/grad run in dae_ctc_lattice + dae_ctc_grad;
  * final-pass window posteriors stay on the device and are stitched by dae_stitch (the reference
    keeps ~6.6 GB of host copies and two ~2 GB host accumulators for a 69 min recording).
"""
import argparse
import ast
import random
import time
from functools import partial
from typing import Callable

import torch
import torch.nn as nn

from . import _C, prof
from .augment import SpecAugment, add_random_noise, cutout, frame_shuffle
from .ctc import CTCLoss
from .greedy import GreedyCTCDecoder, greedy_ids_device
from .optim import MADGRAD
from .stitch import stitch_flat, window_positions


# ----------------------------------------------------------------------------- kwarg getters
def get_specaugment_config_from_args(args):
    """lcasr/lib.py:102-112."""
    a = {k.replace('spec_augment_', ''): v for k, v in args.__dict__.items() if k.startswith('spec_augment')}
    return {
        'n_time_masks': a.get('n_time_masks', 0),
        'n_freq_masks': a.get('n_freq_masks', 0),
        'freq_mask_param': a.get('freq_mask_param', 42),
        'time_mask_param': a.get('time_mask_param', -1),
        'min_p': a.get('min_p', 0.05),
        'zero_masking': a.get('zero_masking', False),
    }


def get_frame_shuffle_config_from_args(args):
    """lcasr/lib.py:114-120."""
    a = {k.replace('frame_shuffle_', ''): v for k, v in args.__dict__.items() if k.startswith('frame_shuffle')}
    return {'time_dimension': a.get('time_dimension', False), 'freq_dimension': a.get('freq_dimension', False)}


def get_lr_args_from_args(args):
    """lcasr/lib.py:122-125."""
    lr_args = {k.replace('optim_', ''): v for k, v in args.__dict__.items() if k.startswith('optim_')}
    lr_args['lr'] = lr_args.get('lr', 9e-5)
    return lr_args


def get_cutout_params_from_args(args, seq_len):
    """lcasr/lib.py:419-428."""
    a = {k.replace('cutout_', ''): v for k, v in args.__dict__.items() if k.startswith('cutout')}
    return {'seq_len': seq_len, 'cutout_val': a.get('value', 'mean'), 'num_rectangles': a.get('num_rectangles', 0),
            'max_width': a.get('max_width', 100), 'max_height': a.get('max_height', 10)}


def prepare_chunks(spec, seq_len, overlap):
    """lcasr/lib.py:128-145: windows of ``seq_len`` frames every ``seq_len-overlap`` frames; the first
    window shorter than its predecessor is kept and ends the list.  Values are views of ``spec``."""
    spec_n = spec.shape[-1]
    last_ulen, kill_next = None, False
    if spec_n <= seq_len:
        return {0: spec}, [0]
    training_data = {}
    for i in range(0, spec_n, seq_len - overlap):
        audio_chunk = spec[:, :, i:i + seq_len]
        u_len = audio_chunk.shape[-1]
        if kill_next:
            break
        elif last_ulen is not None and u_len < last_ulen:
            kill_next = True
        last_ulen = u_len
        training_data[i] = audio_chunk
    return training_data, list(training_data.keys())


# frame_shuffle (lcasr/lib.py:81-84) and add_random_noise (:379-382) live in dae/augment.py: host-drawn
# randomness in the reference's order, gather / streaming kernels on the device.


def bitfit(model):
    """lcasr/lib.py:148-160: train biases only (LayerNorm, Linear and batch-renorm biases)."""
    for param in model.parameters():
        param.requires_grad = False
    for module in model.modules():
        is_norm = isinstance(module, torch.nn.LayerNorm) or type(module).__name__ in ('FusedLayerNorm', 'BatchRenorm1d')
        if (is_norm or isinstance(module, torch.nn.Linear)) and getattr(module, 'bias', None) is not None:
            module.bias.requires_grad = True
    return model


def lm_path_from_paths():
    """``lib.paths.checkpoints.lm`` of the reference (lcasr/lib.py:5, run_dynamic_eval_full.py:58): read from the
    YAML file named by $DAE_PATHS (default ./paths.yaml) if it exists; None otherwise."""
    import os
    p = os.environ.get('DAE_PATHS', 'paths.yaml')
    if not os.path.exists(p):
        return None
    import yaml
    with open(p) as f:
        cfg = yaml.safe_load(f) or {}
    return (cfg.get('checkpoints') or {}).get('lm')


def _freeze(model, args, allow_bitfit=False):
    """lcasr/lib.py:163-204 (the three CLI freeze switches); AWMC also honours ``bitfit`` (:233-234)."""
    d = args.__dict__
    if allow_bitfit and d.get('bitfit', False):
        model = bitfit(model)
    if d.get('freeze_subsampling', False):
        for p in model.subsampling.parameters():
            p.requires_grad = False
    if d.get('freeze_all_but_last_block_and_head', False):
        for p in model.parameters():
            p.requires_grad = False
        for p in list(model.layers[-1].parameters()) + list(model.decoder.parameters()):
            p.requires_grad = True
    if d.get('train_subsampling_only', False):
        for p in model.parameters():
            p.requires_grad = False
        for p in model.subsampling.parameters():
            p.requires_grad = True
    return model


class _Timer:
    """Per-phase host+device timings of one dynamic_eval call (``args.print_runtimes``)."""

    def __init__(self):
        self.t = {}

    def add(self, key, dt):
        self.t[key] = self.t.get(key, 0.0) + dt


def _pseudo_targets(lp_teacher, blank, tokenizer, beam_search_fn, beams):
    """Greedy (device kernel) or beam-search pseudo-labels -> python id list after the tokenizer's
    decode -> re-encode round trip (lcasr/lib.py:558-569)."""
    if beam_search_fn is None or beams == 0:
        _, ids, n = greedy_ids_device(lp_teacher, blank)
        k = int(n[0].item())                               # the one host sync of the step
        prof.count_d2h(n)
        prof.count_d2h(ids[0, :k])
        text = tokenizer.decode(ids[0, :k].tolist())
    else:
        bs = beam_search_fn(log_probs=lp_teacher.detach(), beam_width=beams)
        bs.run_search(use_tqdm=False)
        text = bs.return_text(idx=0)
    return text, tokenizer.encode(text)


class _PendingLabels:
    """Greedy pseudo-labels in flight: the collapse kernel and the device->host copy of (count, ids) are enqueued,
    the host picks them up later (after it has enqueued the augmented branch), so the tokenizer's decode -> re-encode
    hop (lcasr/lib.py:559,569) runs while the GPU is busy."""
    _pinned = {}

    def __init__(self, lp_teacher, blank):
        _, ids, n = greedy_ids_device(lp_teacher, blank)
        Tp = int(ids.shape[1])
        key = (lp_teacher.device.index, Tp)
        if key not in _PendingLabels._pinned:
            _PendingLabels._pinned[key] = torch.empty(Tp + 1, dtype=torch.int32).pin_memory()
        self.buf = _PendingLabels._pinned[key]
        self.buf[:1].copy_(n, non_blocking=True)
        self.buf[1:].copy_(ids[0], non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()
        prof.d2h_bytes += 4 * (Tp + 1)

    def finish(self, tokenizer):
        self.event.synchronize()                            # the one host sync of the step
        k = int(self.buf[0])
        text = tokenizer.decode(self.buf[1:1 + k].tolist())
        return text, tokenizer.encode(text)


def dynamic_eval_ctc_loss(
        args,
        model: nn.Module,
        spec: torch.Tensor,
        seq_len: int,
        overlap: int,
        tokenizer,
        use_tqdm=True,
        optim=MADGRAD,
        optimizer_state: dict = None,
        beam_search_fn: Callable = None,
        return_params: bool = False,
        output: str = "numpy",
):
    """lcasr/lib.py:450-640.  ``output``: 'numpy' (reference behaviour: [N,C] float32 log-probs on the
    host), 'device' (same tensor left on the GPU), 'greedy' (collapsed ids of the stitched posteriors as a
    python list — what run_dynamic_eval_full.py:100 computes next) or 'params' (adapt only: the updated
    parameters, no final pass / stitch)."""
    device = model.device
    if torch.device(device).type != "cuda":
        raise _C.DaeError("dae.lib.dynamic_eval needs the model on a CUDA device: there is no CPU path")
    spec_n = spec.shape[-1]
    downsampling_factor = args.config['model']['subsampling_factor']
    seq_len = seq_len if seq_len != -1 else args.config['audio_chunking']['size']
    d = args.__dict__
    verbose = d.get('verbose', False)

    spec_augment_config = get_specaugment_config_from_args(args)
    random_noise = d.get('random_noise', 0.0)
    lr_args = get_lr_args_from_args(args)
    frame_shuffle_args = get_frame_shuffle_config_from_args(args)
    cutout_args = get_cutout_params_from_args(args, seq_len)
    if d.get('entropy_augmentation_enabled', False):
        raise _C.DaeError("entropy_augmentation is not on the B200 hot path (SURVEY.md §8f-3)")
    num_negatives = 1

    # parameter snapshot: on-device clone (lib.py:482-483 clones to the CPU; restore semantics are the same)
    params = list(model.parameters())
    original = [p.detach().clone() for p in params]
    req_grad = [p.requires_grad for p in params]
    model = _freeze(model, args)

    blank = model.decoder.num_classes - 1
    ctc_loss_fn = CTCLoss(blank=blank, reduction='sum', validate=False)   # labels come from tokenizer.encode
    optimizer = optim([p for p in model.parameters() if p.requires_grad], **lr_args)
    if optimizer_state is not None:
        optimizer.load_state_dict(optimizer_state)
    augmentation = SpecAugment(**spec_augment_config)

    if seq_len > spec_n:
        seq_len, overlap = spec_n, 0
    else:
        overlap = overlap if overlap != -1 else args.config['audio_chunking']['overlap']
    assert args.config['training'].get("max_seq_len", 0) == 0, 'caching is not used anymore'
    assert overlap / downsampling_factor == overlap // downsampling_factor, \
        'Overlap must be a multiple of the downsampling factor'

    epochs = d.get('epochs', 1)
    shuffle = d.get('shuffle', False)
    online = d.get('online', False)
    beams = d.get('lm_tta_beams', 3)
    shuffle = False if online else shuffle
    print_runtimes = d.get('print_runtimes', False)
    tm = _Timer()

    # one H2D copy of the whole recording; windows are device views
    t0 = time.perf_counter()
    spec_dev = spec.to(device, non_blocking=True) if not spec.is_cuda else spec
    if not spec.is_cuda:
        prof.count_h2d(spec)
    if spec_dev.dtype != torch.float32:
        spec_dev = spec_dev.float()
    model.eval()                                            # don't update batchrenorm (lib.py:524)
    training_data, training_keys = prepare_chunks(spec_dev, seq_len, overlap)
    tm.add('h2d', time.perf_counter() - t0)

    # every window's mean (SpecAugment's fill value) in one launch per recording
    win_sums = None
    if not spec_augment_config['zero_masking'] and (spec_augment_config['n_freq_masks'] or spec_augment_config['n_time_masks']) \
            and d.get('epochs', 1) > 0 and spec_dev.dim() == 3 and spec_dev.stride(2) == 1:
        sums = SpecAugment.window_sums(spec_dev, training_keys, [int(training_data[i].shape[-1]) for i in training_keys])
        win_sums = {i: sums[k] for k, i in enumerate(training_keys)}
    # The reference runs [augmented, clean] as ONE batch with autograd on (lib.py:539-550) although the loss only sees
    # the augmented row (:570): the clean row's backward is exact zeros and costs as much as the useful half.  Here
    # the clean (teacher) branch runs first, without a graph, and the augmented branch alone carries the gradient:
    # same parameter gradient (a zero upstream gradient contributes nothing), one third less encoder work per step,
    # and the pseudo-label hop to the host overlaps the augmented forward.  `split_branches=False` restores the
    # reference's single batched call (models whose forward couples the batch items need it).
    split_branches = d.get('split_branches', True)
    C = model.decoder.num_classes
    kept = {}                                               # online: teacher posteriors of each window
    step_log = []
    for epoch in range(d.get('epochs', 1)):                 # lib.py:528 uses args.epochs even when online (:516 unused)
        keys = list(training_data.keys())
        keys = random.sample(keys, len(keys)) if shuffle else keys
        e0 = time.perf_counter()
        for i in keys:
            window = training_data[i]                       # [1,F,T] view
            u_len = window.shape[-1]
            pending = None
            if split_branches:
                with torch.no_grad():
                    teacher = model(audio_signal=window)['final_posteriors'][-1]      # [T', C], no graph
                if beam_search_fn is None or beams == 0:
                    pending = _PendingLabels(teacher, blank)                          # greedy + async D2H enqueued
            audio_chunk = augmentation(window.expand(num_negatives, -1, -1), n_clean=0 if split_branches else 1,
                                       window_sums=None if win_sums is None else win_sums[i])   # [aug..., (clean)]
            if frame_shuffle_args['time_dimension'] or frame_shuffle_args['freq_dimension']:
                audio_chunk[:num_negatives] = frame_shuffle(audio_chunk[:num_negatives], **frame_shuffle_args)
            if random_noise:
                add_random_noise(audio_chunk[:num_negatives], random_noise)   # in place on the augmented copy
            if cutout_args['num_rectangles']:
                cutout(audio_chunk[:num_negatives], **cutout_args)      # in place, lib.py:544
            out = model(audio_signal=audio_chunk)
            post = out['final_posteriors']                  # [1 or 2, T', C] log-probs
            if not split_branches:
                teacher = post[-1].detach()
            if pending is not None:
                text, ids = pending.finish(tokenizer)
            else:
                text, ids = _pseudo_targets(teacher, blank, tokenizer, beam_search_fn, beams)
            if verbose:
                noisy = GreedyCTCDecoder(tokenizer=tokenizer, blank_id=blank)(post[0].detach())
                print(f'Pseudo targets: {text}\nNoisy predictions: {noisy}\n\n--\n')
            pseudo_host = torch.tensor(ids, dtype=torch.long).unsqueeze(0)
            prof.count_h2d(pseudo_host)
            pseudo = pseudo_host.to(device, non_blocking=True).repeat(num_negatives, 1)
            augmented_outs = post[:num_negatives]
            N, B = augmented_outs.shape[1], augmented_outs.shape[0]
            total_tokens_in_loss = N * B
            loss = ctc_loss_fn.with_scale(augmented_outs.transpose(0, 1), pseudo,
                                          torch.full((B,), N, dtype=torch.long, device=device),
                                          torch.full((B,), pseudo.shape[1], dtype=torch.long, device=device),
                                          grad_scale_hint=1.0 / total_tokens_in_loss) / total_tokens_in_loss
            optimizer.zero_grad(set_to_none=True)
            loss.backward()
            optimizer.step()
            if d.get('_record_steps', False):
                step_log.append({'key': i, 'ids': ids, 'loss': loss.detach()})   # no sync here
            if online:
                # a copy when it is a view of the batched output: the view would pin the whole [2,T',C] tensor
                kept[i] = (teacher if split_branches else teacher.clone(), u_len)
        tm.add('adapt', time.perf_counter() - e0)
        if print_runtimes:
            torch.cuda.synchronize(device)
            print(f'Epoch runtime: {time.perf_counter() - e0}')

    if output == 'params':
        # adapt-only (run_half_concat_eval.py:64-160 `adapt_on_concat_only`): no final pass, no stitch
        updated_model_params = [p.clone().detach().cpu() for p in model.parameters()]
        with torch.no_grad():
            for p, p_orig, rg in zip(params, original, req_grad):
                p.data = p_orig.data
                p.requires_grad = rg
        return updated_model_params

    f0 = time.perf_counter()
    if not online:
        model.eval()
        training_data, training_keys = prepare_chunks(spec_dev, seq_len, overlap)
        starts = sorted(training_keys)
        u_lens = [int(training_data[i].shape[-1]) for i in starts]
        flat, used, offs, ds = None, 0, [], []
        with torch.no_grad():
            for i in starts:
                lp = model(audio_signal=training_data[i])['final_posteriors'][0]
                rows = int(lp.shape[0])
                if flat is None:                            # full windows come first: size from the first T'
                    flat = torch.empty((rows * len(starts), C), dtype=torch.float32, device=device)
                if used + rows > flat.shape[0]:
                    flat = torch.cat([flat[:used], torch.empty((rows * 2, C), dtype=torch.float32, device=device)], 0)
                flat[used:used + rows] = lp
                offs.append(used)
                ds.append(rows)
                used += rows
        model.train()                                       # lib.py:612
    else:
        starts = sorted(kept.keys())
        u_lens = [kept[i][1] for i in starts]
        ds = [int(kept[i][0].shape[0]) for i in starts]
        flat = torch.cat([kept[i][0] for i in starts], 0)
        offs, o = [], 0
        for n_ in ds:
            offs.append(o)
            o += n_
    tm.add('final_pass', time.perf_counter() - f0)

    s0 = time.perf_counter()
    pos = window_positions(starts, u_lens, ds, overlap)
    logits, path = stitch_flat(flat, offs, pos, ds, want_path=(output == 'greedy'))
    tm.add('stitch', time.perf_counter() - s0)

    if return_params:
        updated_model_params = [p.clone().detach().cpu() for p in model.parameters()]
    # reset model parameters (lib.py:636-637)
    with torch.no_grad():
        for p, p_orig, rg in zip(params, original, req_grad):
            p.data = p_orig.data
            p.requires_grad = rg

    if output == 'numpy':
        prof.count_d2h(logits)
        result = logits.cpu().numpy()
    elif output == 'device':
        result = logits
    elif output == 'greedy':
        from .greedy import collapse_path_device
        result = collapse_path_device(path, blank)
        prof.h2d_bytes += 0
        prof.d2h_bytes += 4 * len(result) + 4
    else:
        raise ValueError(f"unknown output mode {output!r}")
    if print_runtimes:
        torch.cuda.synchronize(device)
        print('dae runtimes:', {k: round(v, 4) for k, v in tm.t.items()})
    if d.get('_record_steps', False):
        for r in step_log:
            r['loss'] = float(r['loss'])
        args.__dict__['_step_log'] = step_log
    return result if not return_params else (result, updated_model_params)


dynamic_eval = dynamic_eval_ctc_loss


def adapt_on_concat_only(args, model, concat_spec, tokenizer, beamsearch=None, adapt_overlap=None):
    """lcasr/run_half_concat_eval.py:64-160: adapt on a (concatenated) spectrogram and return the updated
    parameters (CPU clones) without building stitched logits; the model's own parameters are restored."""
    if getattr(args, 'awmc', False):
        _, updated = AWMC(args, model, concat_spec, args.seq_len, adapt_overlap, tokenizer, use_tqdm=False,
                          beam_search_fn=beamsearch, return_params=True)
        return updated
    return dynamic_eval_ctc_loss(args, model, concat_spec, args.seq_len, adapt_overlap, tokenizer, use_tqdm=False,
                                 beam_search_fn=beamsearch, output='params')


class _EMA:
    """On-device exponential moving average of parameters with torch_ema's call surface
    (``update()``, ``average_parameters()`` context manager), as AWMC uses it (lcasr/lib.py:243-246,283-291)."""

    def __init__(self, parameters, decay):
        self.params = list(parameters)
        self.decay = float(decay)
        self.shadow = [p.detach().clone() for p in self.params]
        self.num_updates = 0

    @torch.no_grad()
    def update(self):
        self.num_updates += 1
        decay = min(self.decay, (1 + self.num_updates) / (10 + self.num_updates))   # torch_ema's warm-up
        if decay >= 1.0:
            return
        torch._foreach_lerp_(self.shadow, [p.detach() for p in self.params], 1.0 - decay)

    def average_parameters(self):
        ema = self

        class _Ctx:
            def __enter__(self_inner):
                self_inner.saved = [p.detach().clone() for p in ema.params]
                with torch.no_grad():
                    for p, s in zip(ema.params, ema.shadow):
                        p.data.copy_(s)
                return ema.shadow

            def __exit__(self_inner, *exc):
                with torch.no_grad():
                    for p, s in zip(ema.params, self_inner.saved):
                        p.data.copy_(s)
                return False
        return _Ctx()


def AWMC(
        args,
        model: nn.Module,
        spec: torch.Tensor,
        seq_len: int,
        overlap: int,
        tokenizer,
        use_tqdm: bool = True,
        optim=MADGRAD,
        optimizer_state: dict = None,
        beam_search_fn: Callable = None,
        return_params: bool = False,
        output: str = "numpy",
):
    """Anchor/leader EMA-teacher baseline, lcasr/lib.py:206-376, on the dae kernels: greedy pseudo-labels of
    the anchor (first epoch) and leader models on the device, CTC on the N=2 ragged label bank, per-window
    final posterior, overlap stitch."""
    assert beam_search_fn is None, 'Beam search function not implemented for AWMC'
    device = model.device
    if torch.device(device).type != "cuda":
        raise _C.DaeError("dae.lib.AWMC needs the model on a CUDA device: there is no CPU path")
    d = args.__dict__
    spec_augment_config = get_specaugment_config_from_args(args)
    lr_args = get_lr_args_from_args(args)
    frame_shuffle_args = get_frame_shuffle_config_from_args(args)
    spec_n = spec.shape[-1]
    downsampling_factor = args.config['model']['subsampling_factor']
    seq_len = seq_len if seq_len != -1 else args.config['audio_chunking']['size']
    params = list(model.parameters())
    original = [p.detach().clone() for p in params]
    req_grad = [p.requires_grad for p in params]
    model = _freeze(model, args, allow_bitfit=True)
    model.train()                                           # lib.py:242
    ema_leader = _EMA(model.parameters(), decay=d.get('ema_decay', 0.999))
    ema_leader.update()
    ema_anchor = _EMA(model.parameters(), decay=1.0)
    ema_anchor.update()
    blank = model.decoder.num_classes - 1
    ctc_loss_fn = CTCLoss(blank=blank, reduction='sum')
    optimizer = optim([p for p in model.parameters() if p.requires_grad], **lr_args)
    if optimizer_state is not None:
        optimizer.load_state_dict(optimizer_state)
    augmentation = SpecAugment(**spec_augment_config)
    if seq_len > spec_n:
        seq_len, overlap = spec_n, 0
    else:
        overlap = overlap if overlap != -1 else args.config['audio_chunking']['overlap']
    assert args.config['training'].get("max_seq_len", 0) == 0, 'caching is not used anymore'
    assert overlap / downsampling_factor == overlap // downsampling_factor
    epochs = d.get('epochs', 1)
    spec_dev = spec.to(device, non_blocking=True).float() if not spec.is_cuda else spec.float()
    training_data, training_keys = prepare_chunks(spec_dev, seq_len, overlap)
    C = model.decoder.num_classes

    def labels_of(lp):
        _, ids, n = greedy_ids_device(lp, blank)
        text = tokenizer.decode(ids[0, :int(n[0].item())].tolist())
        return torch.tensor(tokenizer.encode(text), dtype=torch.long, device=device)

    kept = {}
    model.eval()                                            # lib.py:276
    for i in training_keys:
        label_bank = [None, None]
        for j in range(epochs):
            audio_chunk = training_data[i]
            if j == 0:
                with ema_anchor.average_parameters(), torch.no_grad():
                    label_bank[0] = labels_of(model(audio_signal=audio_chunk)['final_posteriors'][-1])
            with ema_leader.average_parameters(), torch.no_grad():
                label_bank[1] = labels_of(model(audio_signal=audio_chunk)['final_posteriors'][-1])
            noisy = augmentation(audio_chunk)
            noisy = frame_shuffle(noisy, **frame_shuffle_args)
            out = model(audio_signal=noisy)
            post = out['final_posteriors']
            labels = [el for el in label_bank if el.shape[0] > 0]
            if len(labels) == 0:
                labels = [torch.zeros(0, dtype=torch.long, device=device)]
            lens = torch.tensor([el.shape[0] for el in labels], dtype=torch.long, device=device)
            padded = torch.nn.utils.rnn.pad_sequence(labels, batch_first=True, padding_value=0)
            N, B = post.shape[1], post.shape[0]
            total_tokens_in_loss = N * B * 2
            loss = ctc_loss_fn(post.repeat(lens.shape[0], 1, 1).transpose(0, 1), padded,
                               torch.full((lens.shape[0],), N, dtype=torch.long, device=device), lens) / total_tokens_in_loss
            optimizer.zero_grad(set_to_none=True)
            loss.backward()
            optimizer.step()
            ema_leader.update()
            if j == epochs - 1:
                with torch.no_grad():
                    kept[i] = (model(audio_signal=training_data[i])['final_posteriors'][0], int(audio_chunk.shape[-1]))
    starts = sorted(kept.keys())
    u_lens = [kept[i][1] for i in starts]
    ds = [int(kept[i][0].shape[0]) for i in starts]
    flat = torch.cat([kept[i][0] for i in starts], 0) if len(starts) > 1 else kept[starts[0]][0].contiguous()
    offs, o = [], 0
    for n_ in ds:
        offs.append(o)
        o += n_
    logits, path = stitch_flat(flat, offs, window_positions(starts, u_lens, ds, overlap), ds,
                               want_path=(output == 'greedy'))
    if return_params:
        updated_model_params = [p.clone().detach().cpu() for p in model.parameters()]
    with torch.no_grad():
        for p, p_orig, rg in zip(params, original, req_grad):
            p.data = p_orig.data
            p.requires_grad = rg
    if output == 'numpy':
        result = logits.cpu().numpy()
    elif output == 'device':
        result = logits
    else:
        from .greedy import collapse_path_device
        result = collapse_path_device(path, blank)
    return result if not return_params else (result, updated_model_params)


def load_beamsearch(path: str, alpha: float = 0.45, beta: float = 1.53, prune_less_than_val: float = 3.17,
                    top_am_threshold: float = -6, tokenizer=None, vocab_size=None, bos_id=None):
    """lcasr/lib.py:37-72 with the LM checkpoint replaced by an ARPA n-gram file: returns
    ``partial(BeamSearch, language_model=..., tokenizer=..., blank_id=vocab_size, alpha=..., ...)``."""
    from . import ctc_beam_search as beam_search
    from .ngram import NGramLM
    if tokenizer is None:
        from .standin import SyntheticTokenizer
        tokenizer = SyntheticTokenizer()
    V = vocab_size or tokenizer.vocab_size()
    if bos_id is None:
        bos_id = tokenizer.bos_id() if hasattr(tokenizer, 'bos_id') else -1   # sentencepiece without bos: -1
    language_model = NGramLM.from_arpa(path, V, bos_id=bos_id)
    return partial(beam_search.BeamSearch, language_model=language_model, tokenizer=tokenizer, blank_id=V,
                   alpha=alpha, beta=beta, debug=False, prune_less_than_val=prune_less_than_val,
                   top_am_threshold=top_am_threshold, max_cache_length=128)


# ----------------------------------------------------------------------------- CLI surface
def apply_args(parser, argv=None):
    """lcasr/lib.py:1756-1787.  ``-kwargs key=value`` values are parsed with ast.literal_eval (the
    reference eval()s them, :1778-1781); non-literals stay strings."""
    parser.add_argument('-c', '--checkpoint', type=str, default='', help='path to checkpoint')
    parser.add_argument('-split', '--split', type=str, default='test', help='test or dev split')
    parser.add_argument('-seq', '--seq_len', type=int, default=16384, help='-1 to use setting from config in checkpoint file')
    parser.add_argument('-o', '--overlap', type=int, default=14336, help='-1 to use setting from config in checkpoint file')
    parser.add_argument('-nv', '--not_verbose', action='store_true', help='verbose')
    parser.add_argument('-log', '--log', type=str, default='')
    parser.add_argument('-ds', '--dont_shuffle', action='store_true', help='dont shuffle')
    parser.add_argument('-epochs', '--epochs', type=int, default=1, help='epochs')
    parser.add_argument('-dfa', '--disable_flash_attention', action='store_true', help='disable flash attention')
    parser.add_argument('-beamsearch', '--beamsearch', action='store_true', help='use beam search')
    parser.add_argument('-kwargs', '--kwargs', nargs='+', help='kwargs')
    parser.add_argument('-awmc', '--awmc', action='store_true', help='Use AWMC instead of dynamic eval')
    parser.add_argument('--consistency', '--consistency', action='store_true', help='Use consistency training')
    parser.add_argument('--freeze_subsampling', action='store_true')
    parser.add_argument('--freeze_all_but_last_block_and_head', action='store_true')
    parser.add_argument('--train_subsampling_only', action='store_true')
    args = parser.parse_args(argv)
    if args.kwargs is None:
        args.kwargs = []
    for kwarg in args.kwargs:
        key, value = kwarg.split('=', 1)
        try:
            args.__dict__[key] = ast.literal_eval(value)
        except (ValueError, SyntaxError):
            args.__dict__[key] = value
        print(f'Overriding {key} to {value}')
    args.shuffle = not args.dont_shuffle
    args.verbose = not args.not_verbose
    return args
