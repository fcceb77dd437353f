"""Prefix beam search: oracle pinned to the reference class (CPU), CUDA kernel vs golden/oracle (GPU).
Scores are compared as raw fp32 bits, hypotheses and start times exactly."""
import os

import numpy as np
import pytest
import torch

from oracle.beam_oracle import BeamSearchOracle, NGramOracle, NGramOracleLM, peaky_log_probs

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "beam_ref.npz"))
ALL = ["v31_b8", "v31_b100", "v31_b20_noprune", "v128_b3", "v31_b5_flat", "v12_b10_pen", "cfg3_v31_b100_t2400"]
FAST = ["v31_b8", "v128_b3", "v31_b5_flat", "v12_b10_pen"]


def _case(name, tmp_path):
    from dae.ngram import read_arpa, write_synthetic_arpa
    T, V, W, alpha, beta, thr, prune, seed, sharp, bpen, rpen = GOLD[f"{name}_meta"]
    T, V, W, seed = int(T), int(V), int(W), int(seed)
    arpa = str(tmp_path / f"{name}.arpa")
    write_synthetic_arpa(arpa, V, order=4, counts=(None, 40 * V, 60 * V, 60 * V), seed=seed)
    order, grams = read_arpa(arpa)
    lp = peaky_log_probs(T, V + 1, V, seed, sharp=float(sharp))
    kw = dict(alpha=float(alpha), beta=float(beta), blank_id=V, top_am_threshold=float(thr),
              prune_less_than_val=None if prune < 0 else float(prune), blank_penalty=float(bpen),
              repitition_penalty=float(rpen))
    return lp, V, W, order, grams, kw


def _gold(name):
    lens = GOLD[f"{name}_lens"].tolist()
    seqs, st = GOLD[f"{name}_seqs"].tolist(), GOLD[f"{name}_stimes"].tolist()
    out, o = [], 0
    for r, L in enumerate(lens):
        out.append((GOLD[f"{name}_scores"][r], seqs[o:o + L], st[o:o + L], bool(GOLD[f"{name}_blankend"][r])))
        o += L
    return out


def _same(got, gold):
    assert len(got) == len(gold)
    for (s, seq, st, fl), (gs, gseq, gst, gfl) in zip(got, gold):
        assert np.float32(s).tobytes() == np.float32(gs).tobytes(), (float(s), float(gs))
        assert list(seq) == list(gseq) and list(st) == list(gst) and bool(fl) == bool(gfl)


@pytest.mark.parametrize("name", FAST)
def test_oracle_matches_reference_class(name, tmp_path):
    lp, V, W, order, grams, kw = _case(name, tmp_path)
    res = BeamSearchOracle(V, W, lp, NGramOracle(grams, order, V), **kw).run_search().result()
    _same(res, _gold(name))


@pytest.mark.parametrize("name", ["v31_b8", "v12_b10_pen"])
def test_oracle_lm_protocol_matches_reference_class(name, tmp_path):
    """The oracle driven through the LanguageModel duck type (get_initial_state / batched __call__ over padded
    caches, ctc_beam_search.py:70-87,284-312) gives the reference class's beams; the golden script asserts the
    same against the class itself for every case."""
    lp, V, W, order, grams, kw = _case(name, tmp_path)
    res = BeamSearchOracle(V, W, lp, NGramOracleLM(NGramOracle(grams, order, V)), lm_protocol=True,
                           max_cache_length=128, **kw).run_search().result()
    _same(res, _gold(name))


def test_bos_less_tokenizer_gets_reserved_id(tmp_path):
    """sentencepiece without bos reports -1 (lcasr/lib.py:56): `<s>` then lives at the reserved id V in the trie,
    the Beam view still starts with the tokenizer's own bos id."""
    from dae.ngram import NGramLM, bos_token, read_arpa, write_synthetic_arpa
    V = 20
    assert bos_token(V, -1) == V and bos_token(V, 0) == 0 and bos_token(V, None) == V
    arpa = str(tmp_path / "nb.arpa")
    write_synthetic_arpa(arpa, V, order=3, counts=(None, 200, 300), seed=2)
    lm = NGramLM.from_arpa(arpa, V, bos_id=-1)
    assert lm.bos_id == -1 and lm.bos_tok == V
    assert int(lm.tok[lm.state_of([-1])]) == V and lm.state_of([-1]) == lm.state_of([V]) != 0
    ref = NGramLM.from_arpa(arpa, V, bos_id=0)
    assert (lm.logp == ref.logp).all() or sorted(lm.logp.tolist()) == sorted(ref.logp.tolist())


def test_blank_first_to_last_layout():
    from dae.ctc_beam_search import blank_first_to_last
    lp = torch.randn(7, 5).log_softmax(-1)
    out = blank_first_to_last(lp)
    assert out.shape == (7, 6) and torch.isinf(out[:, 0]).all()
    assert torch.equal(out[:, 1:5], lp[:, 1:]) and torch.equal(out[:, 5], lp[:, 0])


def test_ngram_trie_matches_oracle_scoring(tmp_path):
    """Host walk of the flat trie (same fail-link algorithm as the kernel) == dict-based oracle, bitwise."""
    from dae.ngram import NGramLM, read_arpa, write_synthetic_arpa
    V = 40
    arpa = str(tmp_path / "lm.arpa.gz")
    write_synthetic_arpa(arpa, V, order=4, counts=(None, 900, 1500, 1500), seed=11)
    order, grams = read_arpa(arpa)
    lm, orc = NGramLM(grams, order, V), NGramOracle(grams, order, V)

    def find(node, w):
        lo, hi = lm.cb[node], lm.cb[node + 1]
        j = lo + np.searchsorted(lm.tok[lo:hi], w)
        return int(j) if j < hi and lm.tok[j] == w else -1

    def score(state, w):
        acc, cur = np.float32(0), state
        while True:
            c = find(cur, w)
            if c >= 0:
                return np.float32(acc + lm.logp[c])
            acc = np.float32(acc + lm.bo[cur])
            if cur == 0:
                return np.float32(acc + lm.unk_lp)
            cur = int(lm.fail[cur])
    rng = np.random.default_rng(0)
    keys = [k for k in grams if len(k) == 3]
    for _ in range(300):
        h = [0] + [int(x) for x in rng.integers(1, V, size=rng.integers(0, 5))]
        if rng.random() < 0.5:
            h = h + list(keys[rng.integers(len(keys))])        # make deep matches likely
        st = lm.state_of(h)
        for w in rng.integers(0, V, size=6):
            assert score(st, int(w)).tobytes() == orc.score(h, int(w)).tobytes()


# ------------------------------------------------------------------------------------ GPU
class _Tok:
    def __init__(self, V):
        self.V = V

    def vocab_size(self):
        return self.V

    def decode(self, ids):
        return " ".join(map(str, ids))


@pytest.mark.gpu
@pytest.mark.parametrize("dense_lm", [True, False], ids=["dense", "triewalk"])
@pytest.mark.parametrize("name", ALL)
def test_cuda_matches_reference_golden(cuda, name, dense_lm, tmp_path):
    """Both LM paths of beam_search_kernel (dense row/next tables; fail-link walk of the trie in HBM) against
    the beams the reference class itself produced, incl. the cfg3-shaped 2400-frame beam-100 case."""
    from dae.ctc_beam_search import BeamSearch
    from dae.ngram import NGramLM
    lp, V, W, order, grams, kw = _case(name, tmp_path)
    bs = BeamSearch(_Tok(V), W, torch.from_numpy(lp).to(cuda), NGramLM(grams, order, V), dense_lm=dense_lm, **kw)
    bs.run_search(use_tqdm=False)
    assert (bs._s.row is not None) == dense_lm
    got = [(b.score, b.lm_sequence, b.stimes, b.am_sequence[-1] == V) for b in bs.beams]
    _same(got, _gold(name))
    assert bs.return_text(0) == " ".join(map(str, _gold(name)[0][1][1:]))


@pytest.mark.gpu
def test_cuda_step_api_and_numpy_input(cuda, tmp_path):
    from dae.ctc_beam_search import BeamSearch
    from dae.ngram import NGramLM
    lp, V, W, order, grams, kw = _case("v12_b10_pen", tmp_path)
    lm = NGramLM(grams, order, V)
    a = BeamSearch(_Tok(V), W, lp, lm, **kw)                   # numpy log-probs, as run_dynamic_eval_full.py:102
    a.run_search(use_tqdm=False)
    b = BeamSearch(_Tok(V), W, lp, lm, **kw)
    n = 0
    while b.step():
        n += 1
    assert n == len(lp) - 1
    assert [(x.score.tobytes(), x.lm_sequence, x.stimes) for x in a.beams] == \
           [(x.score.tobytes(), x.lm_sequence, x.stimes) for x in b.beams]
    _same([(x.score, x.lm_sequence, x.stimes, x.am_sequence[-1] == V) for x in a.beams], _gold("v12_b10_pen"))


@pytest.mark.gpu
@pytest.mark.parametrize("dense_lm", [True, False], ids=["dense", "triewalk"])
def test_cuda_batch_segments_match_oracle(cuda, dense_lm, tmp_path):
    """Ragged independent segments in one launch == the oracle run per segment (incl. a 1-frame segment)."""
    from dae.ctc_beam_search import beam_search_batch
    from dae.ngram import NGramLM, read_arpa, write_synthetic_arpa
    V, W = 31, 6
    arpa = str(tmp_path / "b.arpa")
    write_synthetic_arpa(arpa, V, order=3, counts=(None, 500, 800), seed=21)
    order, grams = read_arpa(arpa)
    lens = [57, 1, 120, 33, 80]
    lps = [peaky_log_probs(n, V + 1, V, 100 + k, sharp=4.0) for k, n in enumerate(lens)]
    offs = np.concatenate(([0], np.cumsum(lens)))
    kw = dict(alpha=0.45, beta=1.53, top_am_threshold=-6, prune_less_than_val=3.17)
    res = beam_search_batch(torch.from_numpy(np.concatenate(lps)).to(cuda), offs, NGramLM(grams, order, V), W,
                            blank_id=V, n_best=W, dense_lm=dense_lm, **kw)
    ng = NGramOracle(grams, order, V)
    for g, lp in enumerate(lps):
        ref = BeamSearchOracle(V, W, lp, ng, blank_id=V, **kw).run_search().result()
        _same([(s, [0] + t, [0] + tm, fl) for s, t, tm, fl in res[g]], ref)


@pytest.mark.gpu
def test_cuda_hot_path_vocab_walks_the_trie(cuda, tmp_path):
    """V=4095 (+blank = 4096 classes, the adapt loop's vocabulary) with an LM whose dense expansion would not fit
    (n_ctx*V*8 > 2 GB): the search must take the trie walk on its own, and match the oracle bit for bit at the
    in-loop beam width (3, lib.py:515) and at the eval width (20, run_dynamic_eval_full.py:65)."""
    from dae.ctc_beam_search import BeamSearch
    from dae.ngram import NGramLM, read_arpa, write_synthetic_arpa
    V = 4095
    arpa = str(tmp_path / "big.arpa")
    write_synthetic_arpa(arpa, V, order=3, counts=(None, 70000, 90000), seed=31)
    order, grams = read_arpa(arpa)
    lm = NGramLM(grams, order, V)
    assert int((lm.depth < lm.order).sum()) * V * 8 > (2 << 30)
    assert lm.dense_tables(cuda) == (None, None)
    ng = NGramOracle(grams, order, V)
    for W, T, seed in ((3, 160, 41), (20, 60, 42)):
        lp = peaky_log_probs(T, V + 1, V, seed, sharp=12.0)
        kw = dict(alpha=0.45, beta=1.53, blank_id=V, top_am_threshold=-6, prune_less_than_val=3.17)
        bs = BeamSearch(_Tok(V), W, torch.from_numpy(lp).to(cuda), lm, **kw)
        bs.run_search(use_tqdm=False)
        assert bs._s.row is None
        ref = BeamSearchOracle(V, W, lp, ng, **kw).run_search().result()
        _same([(b.score, b.lm_sequence, b.stimes, b.am_sequence[-1] == V) for b in bs.beams], ref)


@pytest.mark.gpu
def test_cuda_language_model_duck_type(cuda, tmp_path):
    """NGramLM as the reference's LanguageModel (ctc_beam_search.py:45-87): rows from dae_ngram_rows equal the
    dict oracle bitwise, and the oracle search driven through get_initial_state()/__call__ (the way the reference
    class drives its LM, pinned on the CPU above) reproduces the reference class's golden beams."""
    from dae.ngram import NGramLM
    name = "v12_b10_pen"
    lp, V, W, order, grams, kw = _case(name, tmp_path)
    lm, ng = NGramLM(grams, order, V).to(cuda), NGramOracle(grams, order, V)
    lps0, st0 = lm.get_initial_state()
    assert lps0.shape == (V,) and not lps0.is_cuda and st0['cache'].shape == (1, 1, 1, 1, 1, 1)
    assert lps0.numpy().tobytes() == ng.row([0]).tobytes()
    rng = np.random.default_rng(5)
    hists = [[0] + [int(x) for x in rng.integers(1, V, size=n)] for n in (0, 1, 2, 3, 5, 9)]
    batch = lm._pack_state([h[:-1] if len(h) > 1 else h for h in hists])
    ids = torch.tensor([[h[-1]] for h in hists])
    rows, new = lm(ids, torch.ones(len(hists), dtype=torch.long), batch)
    assert rows.shape == (len(hists), 1, V) and new['cache_lengths'].tolist() == [max(len(h), 2) for h in hists]
    for b, h in enumerate(hists):
        full = (h[:-1] if len(h) > 1 else h) + [h[-1]]
        assert rows[b, 0].numpy().tobytes() == ng.row(full).tobytes()
    res = BeamSearchOracle(V, W, lp, lm, lm_protocol=True, max_cache_length=128, **kw).run_search().result()
    _same(res, _gold(name))


@pytest.mark.gpu
def test_cuda_blank_first_posteriors(cuda, tmp_path):
    """wav2vec2 layout (blank 0): blank_first_to_last + a tokenizer reporting C gives the same hypothesis as the
    blank-last layout with every token id shifted by one."""
    from dae.ctc_beam_search import BeamSearch, blank_first_to_last
    from dae.ngram import NGramLM
    lp, V, W, order, grams, kw = _case("v31_b8", tmp_path)           # blank-last [T, V+1]
    C = V + 1
    w2v = np.concatenate([lp[:, V:], lp[:, 1:V]], axis=1)            # blank first, tokens 1..V-1 kept
    relaid = blank_first_to_last(torch.from_numpy(w2v).to(cuda))     # [T, C+1]... class 0 = -inf
    assert relaid.shape[1] == C and torch.isinf(relaid[:, 0]).all()
    kw2 = dict(kw, blank_id=V)
    a = BeamSearch(_Tok(V), W, relaid, NGramLM(grams, order, V), **kw2)
    a.run_search(use_tqdm=False)
    lp0 = lp.copy()
    lp0[:, 0] = -np.inf                                              # class 0 is never expanded anyway
    b = BeamSearch(_Tok(V), W, torch.from_numpy(lp0).to(cuda), NGramLM(grams, order, V), **kw2)
    b.run_search(use_tqdm=False)
    assert [(x.score.tobytes(), x.lm_sequence) for x in a.beams] == [(x.score.tobytes(), x.lm_sequence) for x in b.beams]


@pytest.mark.gpu
def test_cuda_errors(cuda, tmp_path):
    import dae._C as C
    from dae.ctc_beam_search import BeamSearch
    from dae.ngram import NGramLM, read_arpa, write_synthetic_arpa
    arpa = str(tmp_path / "e.arpa")
    write_synthetic_arpa(arpa, 8, order=2, counts=(None, 20), seed=1)
    order, grams = read_arpa(arpa)
    lm = NGramLM(grams, order, 8)
    with pytest.raises(C.DaeError):
        BeamSearch(_Tok(8), 4, np.zeros((5, 9), np.float32), lm, blank_id=3)            # blank must be V
    with pytest.raises(C.DaeError):
        BeamSearch(_Tok(8), 1000, np.zeros((5, 9), np.float32), lm, blank_id=8).run_search()
