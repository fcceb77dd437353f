"""Generate tests/golden/beam_ref.npz by running the REFERENCE's own BeamSearch class
(/root/reference/lcasr/ctc_beam_search.py, imported read-only under a 2-symbol `lming` stub) with an
n-gram LM adapter behind its duck-typed LanguageModel interface (token history carried in the
'cache' tensor).  Also asserts that oracle/beam_oracle.py reproduces every case
bit-for-bit in this container.  Build container only.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def stub(name, **a):
    m = types.ModuleType(name)
    m.__dict__.update(a)
    sys.modules[name] = m


stub("lming"); stub("lming.utils"); stub("lming.utils.helpers", exists=lambda x: x is not None)
stub("lming.models"); stub("lming.models.transformer", transformer_lm=object)
sys.path.insert(0, "/root/reference/lcasr")
import ctc_beam_search as ref  # noqa: E402

from dae.ngram import read_arpa, write_synthetic_arpa  # noqa: E402
from oracle.beam_oracle import BeamSearchOracle, NGramOracle, NGramOracleLM, peaky_log_probs  # noqa: E402


RefLMAdapter = NGramOracleLM   # LanguageModel duck type (ctc_beam_search.py:45-87): history rides in the cache tensor


class Tok:
    def __init__(self, V):
        self.V = V

    def vocab_size(self):
        return self.V

    def decode(self, ids):
        return " ".join(map(str, ids))


CASES = [  # name, T, V, beam, alpha, beta, thr, prune, seed, sharp
    ("v31_b8", 300, 31, 8, 0.45, 1.53, -6, 3.17, 3, 5.0),
    ("v31_b100", 200, 31, 100, 0.45, 1.53, -6, 3.17, 4, 4.0),
    ("v31_b20_noprune", 150, 31, 20, 0.3, 0.8, -6, None, 5, 3.0),
    ("v128_b3", 400, 128, 3, 0.4016, 1.625, -6, 3.221, 6, 6.0),
    ("v31_b5_flat", 80, 31, 5, 0.45, 1.53, -3, 3.17, 7, 1.5),
    ("v12_b10_pen", 120, 12, 10, 0.5, 0.2, -8, 5.0, 8, 2.5),
    # BASELINE.json configs[2] shape (wav2vec2-like V=31+blank, beam 100, class-default LM weights) on a
    # 2400-frame (48 s at 50 fps) stretch: the longest the pure-Python reference class finishes in minutes
    ("cfg3_v31_b100_t2400", 2400, 31, 100, 0.45, 1.53, -6, 3.17, 9, 5.0),
]


def main():
    out = {}
    check_only = "--check-only" in sys.argv          # small cases only, nothing written
    for name, T, V, W, alpha, beta, thr, prune, seed, sharp in CASES:
        if check_only and T > 400:
            continue
        arpa = f"/tmp/beam_{name}.arpa"
        write_synthetic_arpa(arpa, V, order=4, counts=(None, 40 * V, 60 * V, 60 * V), seed=seed)
        order, grams = read_arpa(arpa)
        ng = NGramOracle(grams, order, V)
        lp = peaky_log_probs(T, V + 1, V, seed, sharp=sharp)
        pen = dict(blank_penalty=-0.1, repitition_penalty=-0.05) if name.endswith("pen") else {}
        bs = ref.BeamSearch(Tok(V), W, lp, RefLMAdapter(ng), alpha=alpha, beta=beta, blank_id=V, top_am_threshold=thr,
                            prune_less_than_val=prune, **pen)
        bs.run_search(use_tqdm=False)
        got = [(float(b.score), list(b.lm_sequence), list(b.stimes), b.am_sequence[-1] == V) for b in bs.beams]
        orc = BeamSearchOracle(V, W, lp, ng, alpha=alpha, beta=beta, blank_id=V, top_am_threshold=thr,
                               prune_less_than_val=prune, **pen).run_search().result()
        assert len(got) == len(orc), (name, len(got), len(orc))
        for g, o in zip(got, orc):
            assert np.float32(g[0]).tobytes() == np.float32(o[0]).tobytes() and g[1:] == o[1:], (name, g[:1], o[:1])
        if T <= 400:                                  # the oracle's LanguageModel-protocol mode == the class, too
            orc2 = BeamSearchOracle(V, W, lp, NGramOracleLM(ng), alpha=alpha, beta=beta, blank_id=V,
                                    top_am_threshold=thr, prune_less_than_val=prune, lm_protocol=True, **pen
                                    ).run_search().result()
            assert [(np.float32(o[0]).tobytes(),) + tuple(o[1:]) for o in orc2] == \
                   [(np.float32(g[0]).tobytes(),) + tuple(g[1:]) for g in got], name
        out[f"{name}_meta"] = np.array([T, V, W, alpha, beta, thr, -1.0 if prune is None else prune, seed, sharp,
                                        pen.get("blank_penalty", 0.0), pen.get("repitition_penalty", 0.0)])
        out[f"{name}_scores"] = np.array([g[0] for g in got], dtype=np.float32)
        out[f"{name}_lens"] = np.array([len(g[1]) for g in got], dtype=np.int64)
        out[f"{name}_seqs"] = np.array([t for g in got for t in g[1]], dtype=np.int64)
        out[f"{name}_stimes"] = np.array([t for g in got for t in g[2]], dtype=np.int64)
        out[f"{name}_blankend"] = np.array([g[3] for g in got], dtype=np.bool_)
        print(name, "beams", len(got), "best", got[0][0], "len", len(got[0][1]))
    if check_only:
        print("oracle (both LM modes) == reference on the small cases; nothing written")
        return
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "beam_ref.npz"), **out)
    print("oracle == reference on all cases; wrote tests/golden/beam_ref.npz")


if __name__ == "__main__":
    main()
