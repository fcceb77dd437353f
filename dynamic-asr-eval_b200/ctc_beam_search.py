"""Prefix beam search with the reference class API (lcasr/ctc_beam_search.py), computed on the GPU.

``BeamSearch(tokenizer, beam_width, log_probs, language_model, alpha, beta, blank_id, ...)`` with
``run_search(use_tqdm)``, ``step()``, ``return_text(idx)`` and ``.beams[i].{score, am_sequence,
lm_sequence, stimes}`` (:90-145,193-200) keeps its signature; ``language_model`` is a
``dae.ngram.NGramLM`` (flat back-off trie in HBM) instead of the Transformer-LM wrapper (:45-87).
``beam_search_batch`` decodes many independent segments in one launch (one CTA each).
"""
import numpy as np
import torch

from . import _C, prof
from .ngram import NGramLM

MAX_BEAMS = 128


class Beam:
    """Read-only view of one hypothesis, field names as lcasr/ctc_beam_search.py:15-42."""

    def __init__(self, score, tokens, times, blank_end, bos_id, blank_id):
        self.score = np.float32(score)
        self.lm_sequence = [bos_id] + list(tokens)
        self.stimes = [0] + list(times)
        self.am_sequence = [None] + list(tokens) + ([blank_id] if blank_end else [])
        self.state = None
        self.next_lm_token_lps = None

    def __str__(self):
        return f"{self.am_sequence}"

    __repr__ = __str__


LanguageModel = NGramLM           # implements get_initial_state() / __call__ of ctc_beam_search.py:45-87


class _Search:
    """Device state of a (possibly multi-segment) search; advanced by dae_beam_search launches."""

    def __init__(self, log_probs, seg_offsets, lm, beam_width, alpha, beta, blank_id, blank_penalty,
                 repetition_penalty, top_am_threshold, prune_less_than_val, n_best=None, arena_per_frame=16,
                 device=None, dense_lm=True):
        if not torch.cuda.is_available():
            raise _C.DaeError("beam search needs a CUDA device: there is no CPU path")
        if not isinstance(lm, NGramLM):
            raise _C.DaeError("language_model must be a dae.ngram.NGramLM")
        lp = torch.as_tensor(log_probs)
        dev = lp.device if lp.is_cuda else torch.device(device or "cuda")
        self.lp = lp.to(device=dev, dtype=torch.float32).contiguous()
        self.dev = dev
        T, C = self.lp.shape
        if blank_id != C - 1:
            raise _C.DaeError(f"blank_id must be the last class (got {blank_id}, C={C}); for blank-first posteriors "
                              "(wav2vec2, blank 0) use dae.ctc_beam_search.blank_first_to_last(log_probs)")
        if not 1 <= beam_width <= MAX_BEAMS:
            raise _C.DaeError(f"beam_width must be in 1..{MAX_BEAMS}")
        self.T, self.C = T, C
        so = [0, T] if seg_offsets is None else [int(x) for x in seg_offsets]
        self.seg_host = so
        self.n_seg = len(so) - 1
        self.max_T = max((b - a for a, b in zip(so[:-1], so[1:])), default=0)
        self.seg = torch.tensor(so, dtype=torch.int32, device=dev)
        if lm.vocab_size != blank_id:
            raise _C.DaeError(f"language model vocabulary ({lm.vocab_size}) must equal blank_id ({blank_id})")
        self.lm, self.arr = lm, lm.device_arrays(dev)
        self.row, self.nxt = lm.dense_tables(dev) if dense_lm else (None, None)
        self.W, self.alpha, self.beta = int(beam_width), float(alpha), float(beta)
        self.blank, self.bpen, self.rpen = int(blank_id), float(blank_penalty), float(repetition_penalty)
        self.thr = float(top_am_threshold)
        self.prune = prune_less_than_val
        self.n_best = int(n_best or beam_width)
        self.arena_cap = int(min(self.max_T * self.W, self.max_T * arena_per_frame + 1024)) + 1
        self._alloc()
        self.started = False

    def _alloc(self):
        lib = _C.lib()
        self.nbytes = lib.dae_beam_scratch_bytes(self.n_seg, self.arena_cap)
        self.scratch = torch.empty(self.nbytes, dtype=torch.uint8, device=self.dev)
        nb, cap = self.n_best, max(self.max_T, 1)
        self.out_cap = cap
        self.o_score = torch.empty((self.n_seg, nb), dtype=torch.float32, device=self.dev)
        self.o_len = torch.empty((self.n_seg, nb), dtype=torch.int32, device=self.dev)
        self.o_flag = torch.empty((self.n_seg, nb), dtype=torch.int32, device=self.dev)
        self.o_tok = torch.zeros((self.n_seg, nb, cap), dtype=torch.int32, device=self.dev)
        self.o_time = torch.zeros((self.n_seg, nb, cap), dtype=torch.int32, device=self.dev)
        self.o_n = torch.zeros((self.n_seg, 4), dtype=torch.int32, device=self.dev)

    def advance(self, n_frames, finalize=True):
        a = self.arr
        with torch.cuda.device(self.dev), prof.span("beam_search", self.T * self.C * 4):
            rc = _C.lib().dae_beam_search(
                self.lp.data_ptr(), self.seg.data_ptr(), self.n_seg, self.C, self.blank, self.W, self.alpha, self.beta,
                self.thr, float(self.prune) if self.prune is not None else 0.0, int(self.prune is not None),
                self.bpen, self.rpen, a["tok"].data_ptr(), a["logp"].data_ptr(), a["bo"].data_ptr(),
                a["fail"].data_ptr(), a["cb"].data_ptr(), a["depth"].data_ptr(), self.lm.n_nodes, self.lm.order,
                self.lm.state_of([self.lm.bos_id]), float(self.lm.unk_lp), _C.ptr(self.row), _C.ptr(self.nxt),
                self.scratch.data_ptr(), self.nbytes,
                self.arena_cap, 1 if self.started else 0, int(n_frames), int(finalize), self.n_best, self.out_cap,
                self.o_score.data_ptr(), self.o_len.data_ptr(), self.o_flag.data_ptr(), self.o_tok.data_ptr(),
                self.o_time.data_ptr(), self.o_n.data_ptr(), _C.stream_ptr(self.dev))
        _C.check(rc, "dae_beam_search")
        self.started = True

    def run_all(self):
        """Whole search in one launch; grows the backpointer arena and reruns if it overflowed."""
        while True:
            self.started = False
            self.advance(self.max_T, finalize=True)
            err = self.o_n[:, 1].cpu()
            if int((err == -3).any()) and self.arena_cap < self.max_T * self.W + 1:
                self.arena_cap = self.max_T * self.W + 1
                self._alloc()
                continue
            if int((err != 0).any()):
                code = int(err[err != 0][0])
                raise _C.DaeError(f"beam search failed on a segment with code {code}: "
                                  f"{_C.lib().dae_error_string(code).decode()} (more than 4096 candidates in one frame: "
                                  "lower beam_width or raise top_am_threshold)")
            return

    def stats(self):
        """(candidates scored, loads against the LM arrays) of the last launch, summed over segments."""
        s = self.o_n[:, 2:4].to(torch.int64).sum(0).cpu().tolist()
        return {"candidates": int(s[0]), "lm_loads": int(s[1])}

    def results(self):
        """-> per segment: list of (score, tokens, times, blank_end), best first."""
        n = self.o_n[:, 0].cpu().tolist()
        sc, ln, fl = self.o_score.cpu().numpy(), self.o_len.cpu().numpy(), self.o_flag.cpu().numpy()
        tok, tim = self.o_tok.cpu().numpy(), self.o_time.cpu().numpy()
        out = []
        for g in range(self.n_seg):
            beams = []
            for r in range(min(n[g], self.n_best)):
                L = int(ln[g, r])
                beams.append((sc[g, r], tok[g, r, :L].tolist(), tim[g, r, :L].tolist(), bool(fl[g, r])))
            out.append(beams)
        return out


def blank_first_to_last(log_probs):
    """Re-lay blank-first posteriors (HF wav2vec2: blank/pad id 0, wav2vec2/tedlium/run.py:123) for this class,
    which assumes ``blank_id == vocab_size`` (lcasr/lib.py:64) and never expands class 0 (ctc_beam_search.py:242):
    [T,C] -> [T,C+1] = [-inf, lp[:,1:], lp[:,0]].  Token ids keep their meaning (class i stays class i for i >= 1),
    the new blank id is C, so ``tokenizer.vocab_size()`` must report C."""
    lp = torch.as_tensor(log_probs)
    pad = torch.full_like(lp[..., :1], float("-inf"))
    return torch.cat([pad, lp[..., 1:], lp[..., :1]], dim=-1).contiguous()


def beam_search_batch(log_probs, seg_offsets, language_model, beam_width, alpha=0.4, beta=0.4, blank_id=None,
                      blank_penalty=0.0, repitition_penalty=0.0, top_am_threshold=-6, prune_less_than_val=None,
                      n_best=1, dense_lm=True):
    """Decode independent segments of ``log_probs`` [total_T, C] in ONE launch (one CTA per segment).
    Returns, per segment, the n_best hypotheses as (score, token ids, start frames, ends_in_blank)."""
    lp = torch.as_tensor(log_probs)
    s = _Search(lp, seg_offsets, language_model, beam_width, alpha, beta,
                lp.shape[-1] - 1 if blank_id is None else blank_id, blank_penalty, repitition_penalty,
                top_am_threshold, prune_less_than_val, n_best=n_best, dense_lm=dense_lm)
    s.run_all()
    return s.results()


class BeamSearch:
    def __init__(self, tokenizer, beam_width, log_probs, language_model, alpha=0.4, beta=0.4, blank_id=128,
                 blank_penalty=0.0, repitition_penalty=0.0, top_am_threshold=-6, max_cache_length=-1, debug=False,
                 prune_less_than_val=None, cache_init=None, dense_lm=True):
        """Signature of lcasr/ctc_beam_search.py:90-125 plus ``dense_lm``: True = serve LM queries from the dense
        row/next expansion when it fits in 2 GB of HBM (else walk the trie), False = always walk the trie."""
        self.tokenizer = tokenizer
        self.beam_width = beam_width
        self.vocab_size = tokenizer.vocab_size()
        self.log_probs = log_probs
        self.language_model = language_model
        self.blank_id = blank_id
        self.alpha, self.beta = alpha, beta
        self._beams, self._stale = [], False
        self.position = 0
        self.blank_penalty, self.repitition_penalty = blank_penalty, repitition_penalty
        self.top_am_threshold, self.prune_less_than_val = top_am_threshold, prune_less_than_val
        self.max_cache_length, self.debug, self.cache_init = max_cache_length, debug, cache_init
        self.dense_lm = dense_lm
        if blank_id != self.vocab_size:
            raise _C.DaeError("BeamSearch assumes blank_id == tokenizer.vocab_size() (lcasr/lib.py:64)")
        self._s = None

    def _search(self):
        if self._s is None:
            self._s = _Search(self.log_probs, None, self.language_model, self.beam_width, self.alpha, self.beta,
                              self.blank_id, self.blank_penalty, self.repitition_penalty, self.top_am_threshold,
                              self.prune_less_than_val, dense_lm=self.dense_lm)
        return self._s

    def _refresh(self):
        res = self._s.results()[0]
        lm = self.language_model
        self._beams = [Beam(sc, tok, tim, fl, lm.bos_id, self.blank_id) for sc, tok, tim, fl in res]
        self._stale = False

    @property
    def beams(self):
        """Current hypotheses, best first (ctc_beam_search.py:118).  After step() they are only materialised (one
        zero-frame launch that writes the n-best lists + the device-to-host copies) when somebody looks."""
        if self._stale:
            self._s.advance(0, finalize=True)
            self._refresh()
        return self._beams

    @beams.setter
    def beams(self, value):
        self._beams, self._stale = value, False

    def run_search(self, use_tqdm=True):
        s = self._search()
        if self.position == 0:
            s.run_all()
        else:
            s.advance(s.max_T - self.position, finalize=True)
        self.position = s.max_T
        self._refresh()

    def step(self):
        s = self._search()
        if self.position == s.max_T:
            return False
        s.advance(1, finalize=False)                        # the search state stays on the device
        self._stale = True
        if self.position == s.max_T - 1:                   # ctc_beam_search.py:280-282: last frame, stay there
            self.position = s.max_T
            return False
        self.position += 1
        return True

    def return_text(self, idx):
        beams = self.beams
        if idx >= len(beams):
            print('Beam index out of range')
            return
        return self.tokenizer.decode(beams[idx].lm_sequence[1:])

    def print_beams(self):
        for i, beam in enumerate(self.beams):
            print(f'{i}: {self.return_text(i)} | {beam.score}')
