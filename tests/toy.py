"""Tiny deterministic model/tokenizer shared by the loop-parity tests and the golden generator."""
import torch
import torch.nn as nn


class _Dec(nn.Module):
    def __init__(self, d, C):
        super().__init__()
        self.num_classes = C
        self.ff = nn.Linear(d, C)

    def forward(self, x):
        return self.ff(x)


class ToyModel(nn.Module):
    """x8 strided conv -> tanh MLP -> peaky log-softmax; satisfies the model contract of lcasr/lib.py."""

    def __init__(self, C, F=80, d=32, seed=0, sharp=12.0, blank_bias=0.25):
        super().__init__()
        torch.manual_seed(seed)
        self.subsampling = nn.Conv1d(F, d, kernel_size=8, stride=8)
        self.layers = nn.ModuleList([nn.Linear(d, d)])
        self.decoder = _Dec(d, C)
        self.sharp = sharp
        with torch.no_grad():
            self.decoder.ff.bias.zero_()
            self.decoder.ff.bias[C - 1] = blank_bias
        self.device = torch.device("cpu")

    def forward(self, audio_signal, length=None):
        x = self.subsampling(audio_signal).transpose(1, 2)
        x = torch.tanh(self.layers[0](x))
        z = self.decoder(x) * self.sharp
        lp = z.log_softmax(-1)
        with torch.no_grad():                               # smallest top-2 margin seen (tie-risk indicator)
            top2 = lp.detach().topk(2, -1).values
            self.min_margin = min(getattr(self, "min_margin", 1e9), float((top2[..., 0] - top2[..., 1]).min()))
        return {"final_posteriors": lp, "length": None}


class RecordingTokenizer:
    """Wraps a tokenizer and records every encode() result (the per-step pseudo-label ids)."""

    def __init__(self, tok):
        self.tok, self.encoded = tok, []

    def vocab_size(self):
        return self.tok.vocab_size()

    def decode(self, ids):
        return self.tok.decode(ids)

    def encode(self, text):
        out = self.tok.encode(text)
        self.encoded.append(list(out))
        return out


TOY = dict(C=65, spec_n=3000, seq_len=1024, overlap=512, seed=1234, model_seed=5, spec_seed=0,
           kwargs=dict(epochs=2, shuffle=True, spec_augment_n_freq_masks=2, spec_augment_freq_mask_param=10,
                       optim_lr=1e-3))
TOY_CONFIG = {"model": {"subsampling_factor": 8}, "audio_chunking": {"size": 1024, "overlap": 0},
              "training": {"max_seq_len": 0}}


def toy_spec(seed=0, spec_n=3000, F=80):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, F, spec_n, generator=g)
