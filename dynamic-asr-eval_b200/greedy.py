"""Greedy CTC decoder with the call signature the reference uses.

Drop-in for ``lcasr.decoding.greedy.GreedyCTCDecoder`` (un-vendored dependency) as called
at lcasr/lib.py:498,559,565, lcasr/run_dynamic_eval_full.py:53,100 and
earnings_finetune/train.py:248,294: argmax over classes, collapse repeats, drop blank, then
``tokenizer.decode`` (or the id list when ``decode=False`` / no tokenizer).  The reference
copies the [T,C] posteriors to the host first; here the kernel reads them where they are and
only the collapsed ids (<= T int32) cross PCIe.
"""
import torch

from . import _C, prof


_sync_words = {}


def _sync_word(dev):
    """One zero-initialised scratch word per (device, stream): the fused kernel's last-CTA ticket (self-resetting)."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    if key not in _sync_words:
        _sync_words[key] = torch.zeros(_C.lib().dae_greedy_scratch_bytes(), dtype=torch.uint8, device=dev)
    return _sync_words[key]


def greedy_ids_device(emission: torch.Tensor, blank_id: int, lengths: torch.Tensor = None):
    """emission [T,C] or [B,T,C] fp32 CUDA -> (path [B,T] i32, ids [B,T] i32, n_ids [B] i32), all on device."""
    _C.require_cuda(emission, "emission")
    x = emission.detach()
    if x.dim() == 2:
        x = x.unsqueeze(0)
    if x.dtype != torch.float32:
        x = x.float()
    if x.stride(2) != 1:
        x = x.contiguous()
    B, T, C = x.shape
    dev = x.device
    path = torch.empty((B, max(T, 1)), dtype=torch.int32, device=dev)
    ids = torch.empty((B, max(T, 1)), dtype=torch.int32, device=dev)
    n_ids = torch.empty((B,), dtype=torch.int32, device=dev)
    if lengths is not None:
        lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
    with torch.cuda.device(dev), prof.span("greedy_collapse", B * T * C * 4):
        rc = _C.lib().dae_greedy_collapse(x.data_ptr(), x.stride(0), x.stride(1), B, T, C, _C.ptr(lengths),
                                          int(blank_id), path.data_ptr(), ids.data_ptr(), n_ids.data_ptr(),
                                          _sync_word(dev).data_ptr(), _C.stream_ptr(dev))
    _C.check(rc, "dae_greedy_collapse")
    return path, ids, n_ids


def collapse_path_device(path: torch.Tensor, blank_id: int):
    """Collapse an already computed per-frame argmax path [T] int32 (dae_stitch's fused output)."""
    _C.require_cuda(path, "path")
    p = path.detach().to(torch.int32).contiguous().reshape(1, -1)
    T = int(p.shape[1])
    ids = torch.empty((1, max(T, 1)), dtype=torch.int32, device=p.device)
    n = torch.empty((1,), dtype=torch.int32, device=p.device)
    with torch.cuda.device(p.device):
        rc = _C.lib().dae_collapse_path(p.data_ptr(), 1, T, None, int(blank_id), ids.data_ptr(), n.data_ptr(),
                                        _C.stream_ptr(p.device))
    _C.check(rc, "dae_collapse_path")
    return ids[0, :int(n[0].item())].tolist()


def greedy_ids(emission: torch.Tensor, blank_id: int):
    """[T,C] -> python list of collapsed ids (one small D2H copy)."""
    _, ids, n = greedy_ids_device(emission, blank_id)
    k = int(n[0].item())
    return ids[0, :k].tolist()


class GreedyCTCDecoder(torch.nn.Module):
    def __init__(self, tokenizer=None, blank_id: int = 0):
        super().__init__()
        self.tokenizer = tokenizer
        self.blank = blank_id

    def forward(self, emission: torch.Tensor, decode: bool = True):
        if not emission.is_cuda:
            # the reference hands over a host copy (lib.py:559); send it back rather than decode on the CPU
            if not torch.cuda.is_available():
                raise _C.DaeError("GreedyCTCDecoder needs a CUDA device: there is no CPU path")
            emission = emission.cuda(non_blocking=True)
        ids = greedy_ids(emission, self.blank)
        if decode and self.tokenizer is not None:
            return self.tokenizer.decode(ids)
        return ids
