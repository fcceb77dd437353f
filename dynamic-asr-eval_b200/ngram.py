"""Token-level back-off n-gram language model held as a flat trie in HBM.

This replaces the Transformer LM behind ``ctc_beam_search.LanguageModel`` (lcasr/ctc_beam_search.py:45-87,
built by lcasr/lib.py:37-72) with the ARPA n-gram the north star asks for.  The reference has no n-gram
code of its own (pyctcdecode/kenlm are absent and unpinned), so the semantics are defined
here and restated independently in oracle/beam_oracle.py:

  score(h, w), h = last (order-1) tokens of the LM sequence (bos included):
      acc = 0
      for ctx in (h, h[1:], ..., ()):            # longest context first
          if ctx+(w,) is an n-gram: return acc + logp(ctx+(w,))
          acc = acc + backoff(ctx)               # 0 when ctx has no entry
      return acc + unk_logp
  all values are natural-log fp32 (ARPA log10 * ln 10, rounded once), all adds fp32 in that order.

Device layout (one node per n-gram, root = node 0, nodes ordered by (depth, parent, token) so the
children of a node are one sorted run): tok[n], logp[n], bo[n], fail[n] (node of the longest proper
suffix), cb[n]..cb[n+1] (child run), depth[n].  A beam's LM state is the node of its longest context
suffix present in the trie; scoring walks fail links with one binary search per level.
"""
import gzip
import math
import random

import numpy as np
import torch

LN10 = math.log(10.0)


def write_synthetic_arpa(path, vocab_size, order=4, counts=(None, 2000, 4000, 4000), seed=4, bos_id=0, fast=False):
    """Write a random but well-formed ARPA file (prefix- and suffix-closed) over token ids 1..vocab_size-1
    (words are decimal ids, ``<s>`` is the bos token).  counts[k-1] = number of k-grams (None = all unigrams)."""
    rng = random.Random(seed)
    toks = list(range(1, vocab_size))
    levels = []
    uni = {(bos_id,): (-99.0, rng.uniform(-1.0, -0.05))}
    for t in toks:
        uni[(t,)] = (rng.uniform(-4.0, -0.5), rng.uniform(-1.0, -0.05))
    levels.append(uni)
    for k in range(2, order + 1):
        prev = levels[-1]
        prev_keys = list(prev.keys())
        by_prefix = {}
        for g in prev_keys:
            by_prefix.setdefault(g[:-1], []).append(g)
        want = counts[k - 1] if counts[k - 1] is not None else len(prev_keys)
        cur, tries = {}, 0
        if fast:
            # enumerate every admissible extension once and sample without replacement (benchmark-sized LMs;
            # the rejection loop below is kept as the default because the golden fixtures were drawn with it)
            space = [g + (h[-1],) for g in prev_keys for h in by_prefix.get(g[1:], ()) if h[-1] != bos_id]
            for ng in (space if len(space) <= want else rng.sample(space, want)):
                cur[ng] = (rng.uniform(-3.0, -0.05), rng.uniform(-1.0, -0.02) if k < order else None)
            tries = want * 50
        while len(cur) < want and tries < want * 50:
            tries += 1
            g = rng.choice(prev_keys)
            cands = by_prefix.get(g[1:], None)          # (k-1)-grams starting with g's suffix
            if not cands:
                continue
            h = rng.choice(cands)
            ng = g + (h[-1],)
            if ng[-1] == bos_id or ng in cur:
                continue
            cur[ng] = (rng.uniform(-3.0, -0.05), rng.uniform(-1.0, -0.02) if k < order else None)
        levels.append(cur)
    name = lambda t: "<s>" if t == bos_id else str(t)
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wt") as f:
        f.write("\\data\\\n")
        for k, lv in enumerate(levels, 1):
            f.write(f"ngram {k}={len(lv)}\n")
        for k, lv in enumerate(levels, 1):
            f.write(f"\n\\{k}-grams:\n")
            for ng in sorted(lv):
                lp, bo = lv[ng]
                words = " ".join(name(t) for t in ng)
                f.write(f"{lp:.6f}\t{words}" + (f"\t{bo:.6f}\n" if bo is not None else "\n"))
        f.write("\n\\end\\\n")
    return sum(len(lv) for lv in levels)


def read_arpa(path, word_to_id=None, bos_id=0):
    """-> (order, {tuple_of_ids: (logp10, backoff10 or None)}).  Unknown words are skipped."""
    def default_map(w):
        if w == "<s>":
            return bos_id
        if w in ("</s>", "<unk>"):
            return None
        try:
            return int(w)
        except ValueError:
            return None
    wmap = word_to_id or default_map
    grams, order, cur = {}, 0, 0
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rt") as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("ngram ") or line in ("\\data\\", "\\end\\"):
                continue
            if line.startswith("\\") and line.endswith("-grams:"):
                cur = int(line[1:line.index("-")])
                order = max(order, cur)
                continue
            parts = line.split("\t") if "\t" in line else line.split()
            if "\t" in line:
                lp, words = float(parts[0]), parts[1].split()
                bo = float(parts[2]) if len(parts) > 2 else None
            else:
                lp, words = float(parts[0]), parts[1:1 + cur]
                bo = float(parts[1 + cur]) if len(parts) > 1 + cur else None
            ids = tuple(wmap(w) for w in words)
            if any(i is None for i in ids):
                continue
            grams[ids] = (lp, bo)
    return order, grams


def bos_token(vocab_size, bos_id):
    """Token id that stands for ``<s>`` inside the trie (see NGramLM.__init__)."""
    return int(bos_id) if bos_id is not None and 0 <= int(bos_id) < int(vocab_size) else int(vocab_size)


class NGramLM:
    """Flat-trie n-gram LM.  ``language_model=`` argument of dae.ctc_beam_search.BeamSearch; also implements the
    reference's ``LanguageModel`` duck type (lcasr/ctc_beam_search.py:45-87): ``bos_id``, ``get_initial_state()``
    and ``__call__(input_ids, input_lengths, states)``, so code written against that interface (including the
    reference's own BeamSearch class) can be driven by the HBM trie.  Neural LMs are not supported."""

    def __init__(self, grams, order, vocab_size, bos_id=0, unk_logprob10=-10.0, device=None):
        """``grams``: {token-id tuple: (log10 p, log10 backoff or None)} with the sentence-start token written as
        ``bos_token(vocab_size, bos_id)``.  ``bos_id`` is what the caller's tokenizer reports (``Beam.lm_sequence[0]``,
        lcasr/ctc_beam_search.py:133); a tokenizer without bos (sentencepiece returns -1) gets the reserved id
        ``vocab_size`` inside the trie, which no search candidate can ever produce (it is the blank class)."""
        self.order, self.vocab_size, self.bos_id = int(order), int(vocab_size), int(bos_id)
        self.bos_tok = bos_token(vocab_size, bos_id)
        self.device = torch.device(device) if device is not None else None
        self.unk_lp = np.float32(unk_logprob10 * LN10)
        by_depth = [[] for _ in range(order + 1)]
        for ng in grams:
            if 1 <= len(ng) <= order:
                by_depth[len(ng)].append(ng)
        ids = {(): 0}
        tok, logp, bo, depth, parent = [0], [0.0], [0.0], [0], [0]
        for k in range(1, order + 1):
            # parents already numbered; order this level by (parent id, token): children of a node are one run
            level = [ng for ng in by_depth[k] if ng[:-1] in ids]
            level.sort(key=lambda ng: (ids[ng[:-1]], ng[-1]))
            for ng in level:
                lp10, bo10 = grams[ng]
                ids[ng] = len(tok)
                tok.append(ng[-1])
                logp.append(lp10 * LN10)
                bo.append((bo10 or 0.0) * LN10)
                depth.append(k)
                parent.append(ids[ng[:-1]])
        n = len(tok)
        par = np.asarray(parent, dtype=np.int64)
        cnt = np.bincount(par[1:], minlength=n)
        cb = (1 + np.concatenate(([0], np.cumsum(cnt)))).astype(np.int32)   # root's children start at node 1
        fail = np.zeros(n, dtype=np.int32)
        for ng, i in ids.items():
            for s in range(1, len(ng)):
                j = ids.get(ng[s:])
                if j is not None:
                    fail[i] = j
                    break
        self.tok = np.asarray(tok, dtype=np.int32)
        self.logp = np.asarray(logp, dtype=np.float64).astype(np.float32)
        self.bo = np.asarray(bo, dtype=np.float64).astype(np.float32)
        self.depth = np.asarray(depth, dtype=np.int32)
        self.fail, self.cb, self.n_nodes = fail, cb, n
        self._ids = ids
        self._dev = {}

    @classmethod
    def from_arpa(cls, path, vocab_size, bos_id=0, word_to_id=None, unk_logprob10=-10.0, device=None):
        order, grams = read_arpa(path, word_to_id, bos_token(vocab_size, bos_id))
        return cls(grams, order, vocab_size, bos_id, unk_logprob10, device=device)

    @classmethod
    def synthetic(cls, vocab_size, order, last_level_nodes, seed=4, bos_id=0):
        """A large random but well-formed LM built directly as trie arrays (no ARPA text, no python dict): every
        context of fewer than ``order-1`` tokens over ids 0..V-1 exists (a full V-ary tree, so every fail link
        does too) and ``last_level_nodes`` random n-grams of the highest order hang below it.  For benchmarks that
        need a trie larger than L2 (bench.py); scoring semantics are those of the class docstring."""
        V, rng = int(vocab_size), np.random.default_rng(seed)
        self = cls.__new__(cls)
        self.order, self.vocab_size, self.bos_id, self.bos_tok = int(order), V, int(bos_id), bos_token(V, bos_id)
        self.device, self.unk_lp, self._dev, self._ids = None, np.float32(-10.0 * LN10), {}, None
        full_depth = order - 1
        level_start = [0]
        for k in range(full_depth + 1):
            level_start.append(level_start[-1] + V ** k)            # nodes of depth k occupy [start[k], start[k+1])
        n_full = level_start[-1]
        # last level: a sorted random sample of (parent at full_depth, token) pairs
        n_par = V ** full_depth
        pairs = np.unique(rng.integers(0, n_par * V, size=int(last_level_nodes * 1.05), dtype=np.int64))[:last_level_nodes]
        n = n_full + len(pairs)
        tok = np.zeros(n, np.int32); depth = np.zeros(n, np.int32); fail = np.zeros(n, np.int32)
        cb = np.zeros(n + 1, np.int32)
        for k in range(1, full_depth + 1):
            idx = np.arange(V ** k, dtype=np.int64)                  # position inside level k = base-V digits t1..tk
            tok[level_start[k]:level_start[k + 1]] = (idx % V).astype(np.int32)
            depth[level_start[k]:level_start[k + 1]] = k
            fail[level_start[k]:level_start[k + 1]] = (level_start[k - 1] + idx % (V ** (k - 1))).astype(np.int32)
        tok[n_full:] = (pairs % V).astype(np.int32)
        depth[n_full:] = order
        par_pos = pairs // V
        fail[n_full:] = (level_start[full_depth] + (par_pos % (V ** (full_depth - 1))) * V + pairs % V).astype(np.int32) \
            if full_depth >= 1 else 0
        # child runs: full levels have exactly V children each; the deepest full level gets the sampled runs
        for k in range(full_depth):
            cnt_idx = np.arange(V ** k, dtype=np.int64)
            cb[level_start[k]:level_start[k + 1]] = (level_start[k + 1] + cnt_idx * V).astype(np.int32)
        counts = np.bincount(par_pos, minlength=n_par)
        cb[level_start[full_depth]:level_start[full_depth + 1]] = (n_full + np.concatenate(([0], np.cumsum(counts)[:-1]))).astype(np.int32)
        cb[n_full:] = n
        cb[n] = n
        self.tok, self.depth, self.fail, self.cb, self.n_nodes = tok, depth, fail, cb, n
        self.logp = (rng.uniform(-4.0, -0.05, size=n) * LN10).astype(np.float32)
        self.bo = (rng.uniform(-1.0, -0.02, size=n) * LN10).astype(np.float32)
        self.logp[0] = 0.0
        self.bo[n_full:] = 0.0
        self._level_start, self._full_depth = level_start, full_depth
        return self

    # ---- LanguageModel duck type (lcasr/ctc_beam_search.py:45-87) -------------------------------------------
    # The "KV cache" of the reference is the token history here: state = {'cache': float32 [1,1,B,1,N,1] holding
    # the LM sequence of every beam (what BeamSearch.step pads, rearranges and slices, :284-312,172-191),
    # 'cache_lengths': int64 [B]}.  Rows are computed on the device by dae_ngram_rows and returned on the CPU,
    # as the reference's wrapper does (:74,87).
    def to(self, device):
        self.device = torch.device(device)
        return self

    def _rows(self, histories):
        from . import _C
        dev = self.device if self.device is not None else torch.device("cuda")
        if dev.type != "cuda" or not torch.cuda.is_available():
            raise _C.DaeError("NGramLM scoring needs a CUDA device: there is no CPU path")
        a = self.device_arrays(dev)
        st = torch.tensor([self.state_of(h) for h in histories], dtype=torch.int32).to(dev)
        row = torch.empty((len(histories), self.vocab_size), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _C.lib().dae_ngram_rows(a["tok"].data_ptr(), a["logp"].data_ptr(), a["bo"].data_ptr(),
                                         a["fail"].data_ptr(), a["cb"].data_ptr(), a["depth"].data_ptr(),
                                         self.n_nodes, self.order, float(self.unk_lp), self.vocab_size,
                                         st.data_ptr(), len(histories), row.data_ptr(), None, _C.stream_ptr(dev))
        _C.check(rc, "dae_ngram_rows")
        return row.cpu()

    @staticmethod
    def _pack_state(histories):
        n = max(len(h) for h in histories)
        cache = torch.zeros(1, 1, len(histories), 1, n, 1)
        for b, h in enumerate(histories):
            cache[0, 0, b, 0, :len(h), 0] = torch.tensor(h, dtype=torch.float32)
        return {'cache': cache, 'cache_lengths': torch.LongTensor([len(h) for h in histories])}

    def get_initial_state(self):
        """-> (log p(. | <s>) [V] on the CPU, state)   (ctc_beam_search.py:70-74)."""
        h = [self.bos_tok]
        return self._rows([h])[0], self._pack_state([h])

    def __call__(self, input_ids, input_lengths=None, states=None):
        """input_ids [nb,1] = the token each beam just took; states = padded history batch ->
        (log-probs [nb,1,V] on the CPU, new states)   (ctc_beam_search.py:83-87)."""
        nb = int(input_ids.shape[0])
        hists = []
        for b in range(nb):
            past = []
            if states is not None:
                n = int(states['cache_lengths'][b])
                past = [int(x) for x in states['cache'][0, 0, b, 0, :n, 0].tolist()]
            hists.append(past + [int(input_ids[b, -1])])
        return self._rows(hists)[:, None, :], self._pack_state(hists)

    def nbytes(self):
        return sum(a.nbytes for a in (self.tok, self.logp, self.bo, self.depth, self.fail, self.cb))

    def device_arrays(self, device):
        """Upload once per device; returns dict of CUDA tensors (the trie stays resident in HBM)."""
        key = str(device)
        if key not in self._dev:
            self._dev[key] = {k: torch.from_numpy(getattr(self, k)).to(device)
                              for k in ("tok", "logp", "bo", "depth", "fail", "cb")}
        return self._dev[key]

    def dense_tables(self, device, max_bytes=2 << 30):
        """Dense (row, next) expansion of the context nodes on ``device`` (dae_ngram_expand), or (None, None) when
        n_ctx * vocab * 8 bytes would exceed ``max_bytes`` (large vocabularies keep walking the trie)."""
        from . import _C
        key = "dense:" + str(device)
        if key not in self._dev:
            n_ctx = int((self.depth < self.order).sum())           # nodes are ordered by depth: a prefix
            V = self.vocab_size
            if n_ctx * V * 8 > max_bytes:
                self._dev[key] = (None, None)
            else:
                a = self.device_arrays(device)
                row = torch.empty((n_ctx, V), dtype=torch.float32, device=device)
                nxt = torch.empty((n_ctx, V), dtype=torch.int32, device=device)
                with torch.cuda.device(device):
                    rc = _C.lib().dae_ngram_expand(a["tok"].data_ptr(), a["logp"].data_ptr(), a["bo"].data_ptr(),
                                                   a["fail"].data_ptr(), a["cb"].data_ptr(), a["depth"].data_ptr(),
                                                   self.n_nodes, self.order, float(self.unk_lp), V, n_ctx,
                                                   row.data_ptr(), nxt.data_ptr(), _C.stream_ptr(device))
                _C.check(rc, "dae_ngram_expand")
                self._dev[key] = (row, nxt)
        return self._dev[key]

    # host-side state helpers (index arithmetic only; scoring happens on the device)
    def state_of(self, history):
        """Trie node of the longest suffix of ``history`` (at most order-1 tokens) that is a node."""
        h = tuple(self.bos_tok if (t == self.bos_id and t != self.bos_tok) else t for t in history)
        h = h[-(self.order - 1):] if self.order > 1 else ()
        if self._ids is None:                                       # synthetic(): full tree, positional index
            h = tuple(t for t in h if 0 <= t < self.vocab_size)
            pos = 0
            for t in h:
                pos = pos * self.vocab_size + t
            return self._level_start[len(h)] + pos
        for s in range(len(h) + 1):
            j = self._ids.get(h[s:])
            if j is not None:
                return j
        return 0
