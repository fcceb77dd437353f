"""Stand-ins for what the reference loads from disk or from the absent `lcasr` package:
an LCASR-shaped CTC encoder (PyTorch; NOT part of the product), a synthetic
word-piece tokenizer, and Earnings22/TED-LIUM/Rev16-shaped synthetic recordings.

Model contract consumed by lib.py: ``model(audio_signal=[B,80,T]) ->
{'final_posteriors': [B,T',C] log-probs, 'length': ...}``, ``model.device``,
``model.decoder.num_classes`` (= C, blank = C-1), ``model.layers``, ``model.subsampling``.
Shapes follow earnings_finetune/lcasr160rb1.yaml:1-29.
"""
import math
import random

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

LCASR160RB1 = dict(feat_in=80, n_layers=6, d_model=768, n_heads=6, head_dim=128, subsampling_factor=8,
                   subsampling_conv_channels=256, conv_kernel_size=9, self_conditioning=True,
                   rotary_base_freq=1500000, ff_mult=4)
LCASR_SMALL = dict(feat_in=80, n_layers=6, d_model=256, n_heads=4, head_dim=64, subsampling_factor=8,
                   subsampling_conv_channels=64, conv_kernel_size=9, self_conditioning=True,
                   rotary_base_freq=1500000, ff_mult=4)


def default_config(model_cfg=None):
    """The ``args.config`` dict lib.py reads (lcasr/lib.py:464-465,501-507)."""
    cfg = dict(LCASR160RB1 if model_cfg is None else model_cfg)
    return {"model": cfg, "audio_chunking": {"size": 16384, "overlap": 0}, "training": {"max_seq_len": 0}}


class DwStridingSubsampling(nn.Module):
    """x8 'dw_striding' subsampling: three k3/s2/p1 conv stages, T' = floor((T-1)/2)+1 thrice."""

    def __init__(self, feat_in, channels, d_model):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(1, channels, 3, 2, 1), nn.SiLU(),
            nn.Conv2d(channels, channels, 3, 2, 1, groups=channels), nn.Conv2d(channels, channels, 1), nn.SiLU(),
            nn.Conv2d(channels, channels, 3, 2, 1, groups=channels), nn.Conv2d(channels, channels, 1), nn.SiLU())
        f = feat_in
        for _ in range(3):
            f = (f - 1) // 2 + 1
        self.out = nn.Linear(channels * f, d_model)

    def forward(self, x):                       # [B, F, T]
        x = self.conv(x.transpose(1, 2).unsqueeze(1))   # [B, ch, T', F']
        b, c, t, f = x.shape
        return self.out(x.permute(0, 2, 1, 3).reshape(b, t, c * f))

    @staticmethod
    def out_len(n):
        for _ in range(3):
            n = (n - 1) // 2 + 1
        return n


def _rotary(x, base):                            # x [B,H,T,D]
    t, d = x.shape[-2], x.shape[-1]
    inv = 1.0 / (base ** (torch.arange(0, d, 2, device=x.device, dtype=torch.float32) / d))
    ang = torch.arange(t, device=x.device, dtype=torch.float32)[:, None] * inv[None]
    cos, sin = ang.cos()[None, None], ang.sin()[None, None]
    x1, x2 = x[..., 0::2], x[..., 1::2]
    return torch.stack((x1 * cos - x2 * sin, x1 * sin + x2 * cos), -1).flatten(-2)


class ConformerBlock(nn.Module):
    def __init__(self, d, heads, head_dim, ff_mult, kernel, rotary_base):
        super().__init__()
        self.heads, self.head_dim, self.base = heads, head_dim, rotary_base
        self.ff1 = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, d * ff_mult, bias=False), nn.SiLU(),
                                 nn.Linear(d * ff_mult, d, bias=False))
        self.ff2 = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, d * ff_mult, bias=False), nn.SiLU(),
                                 nn.Linear(d * ff_mult, d, bias=False))
        self.attn_norm = nn.LayerNorm(d)
        self.qkv = nn.Linear(d, 3 * heads * head_dim, bias=False)
        self.attn_out = nn.Linear(heads * head_dim, d, bias=False)
        self.conv_norm = nn.LayerNorm(d)
        self.pw1 = nn.Conv1d(d, 2 * d, 1)
        self.dw = nn.Conv1d(d, d, kernel, padding=kernel // 2, groups=d)
        self.dw_norm = nn.LayerNorm(d)
        self.pw2 = nn.Conv1d(d, d, 1)
        self.out_norm = nn.LayerNorm(d)

    def forward(self, x):                        # [B,T,D]
        x = x + 0.5 * self.ff1(x)
        b, t, _ = x.shape
        q, k, v = self.qkv(self.attn_norm(x)).view(b, t, 3, self.heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k = _rotary(q, self.base), _rotary(k, self.base)
        a = F.scaled_dot_product_attention(q, k, v)
        x = x + self.attn_out(a.transpose(1, 2).reshape(b, t, -1))
        c = self.pw1(self.conv_norm(x).transpose(1, 2))
        c = F.glu(c, dim=1)
        c = self.dw(c)
        c = F.silu(self.dw_norm(c.transpose(1, 2))).transpose(1, 2)
        x = x + self.pw2(c).transpose(1, 2)
        x = x + 0.5 * self.ff2(x)
        return self.out_norm(x)


class CTCDecoder(nn.Module):
    def __init__(self, d, num_classes):
        super().__init__()
        self.num_classes = num_classes
        self.norm = nn.LayerNorm(d)
        self.ff = nn.Linear(d, num_classes)

    def forward(self, x, logits=False):
        z = self.ff(self.norm(x))
        return z if logits else F.log_softmax(z, dim=-1)


class StandInSCConformer(nn.Module):
    """Self-conditioned conformer CTC encoder with the I/O contract of lcasr's SCConformerXL."""

    def __init__(self, vocab_size, feat_in=80, n_layers=6, d_model=768, n_heads=6, head_dim=128,
                 subsampling_factor=8, subsampling_conv_channels=256, conv_kernel_size=9, self_conditioning=True,
                 rotary_base_freq=1500000, ff_mult=4, **_):
        super().__init__()
        assert subsampling_factor == 8
        self.subsampling = DwStridingSubsampling(feat_in, subsampling_conv_channels, d_model)
        self.layers = nn.ModuleList([ConformerBlock(d_model, n_heads, head_dim, ff_mult, conv_kernel_size,
                                                    rotary_base_freq) for _ in range(n_layers)])
        self.decoder = CTCDecoder(d_model, vocab_size + 1)
        self.self_conditioning = self_conditioning
        self.sc_proj = nn.Linear(vocab_size + 1, d_model) if self_conditioning else None
        self.device = torch.device("cpu")
        self.register_buffer("spike_prior", None, persistent=False)

    def set_spike_prior(self, n_frames=2048, nonblank_frac=0.3, strength=12.0, seed=0):
        """Synthetic-data shaping (not part of the architecture): a fixed, non-trainable logit pattern that puts a
        spike on one class per output frame — blank on ~70 % of the frames, a pseudo-random label elsewhere — so
        that random-init weights yield speech-like, stable pseudo-labels (L ~ 600 per 2048-frame window) instead
        of collapsing to all-blank after the first adaptation steps."""
        g = torch.Generator().manual_seed(seed)
        C = self.decoder.num_classes
        cls = torch.randint(0, C - 1, (n_frames,), generator=g)
        cls[torch.rand(n_frames, generator=g) >= nonblank_frac] = C - 1
        prior = torch.zeros(n_frames, C)
        prior[torch.arange(n_frames), cls] = strength
        self.spike_prior = prior.to(next(self.parameters()).device)
        return int((cls != C - 1).sum())

    def print_total_params(self):
        print(f"Total params: {sum(p.numel() for p in self.parameters()) / 1e6:.1f}M")

    def forward(self, audio_signal, length=None):
        x = self.subsampling(audio_signal)
        n = len(self.layers)
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if self.self_conditioning and i != n - 1:
                x = x + self.sc_proj(self.decoder(x).exp())
        if self.spike_prior is not None:
            z = self.decoder(x, logits=True)
            lp = F.log_softmax(z + self.spike_prior[:z.shape[1]], dim=-1)
        else:
            lp = self.decoder(x)
        out_len = None
        if length is not None:
            out_len = torch.as_tensor([DwStridingSubsampling.out_len(int(l)) for l in length], device=lp.device)
        return {"final_posteriors": lp, "length": out_len}


@torch.no_grad()
def calibrate_blank_prior(model, spec, blank_frac=0.7, spike=8.0):
    """Shape random-init posteriors like speech: raise the blank bias so ~blank_frac of the frames
    decode to blank (random weights would otherwise emit a label on nearly every frame, L ~ T')."""
    model.decoder.ff.bias.zero_()
    z = model.decoder(_encode(model, spec), logits=True)
    blank = model.decoder.num_classes - 1
    nb = z[..., :blank].max(-1).values - z[..., blank]
    model.decoder.ff.bias[blank] = torch.quantile(nb.flatten().float(), blank_frac)
    return float(model.decoder.ff.bias[blank])


def _encode(model, spec):
    x = model.subsampling(spec)
    n = len(model.layers)
    for i, layer in enumerate(model.layers):
        x = layer(x)
        if model.self_conditioning and i != n - 1:
            x = x + model.sc_proj(model.decoder(x).exp())
    return x


class SyntheticTokenizer:
    """Word-piece tokenizer with the SentencePiece call surface lib.py uses
    (``vocab_size() / encode(str) -> ids / decode(ids) -> str / bos_id()``).

    Pieces are generated deterministically from a seed: `▁`-prefixed word-initial pieces and
    word-internal pieces over a small alphabet.  encode() is greedy longest-match, so
    ``encode(decode(ids))`` is NOT the identity in general — the same property that forces the
    reference's decode -> re-encode round trip (lcasr/lib.py:559,569).
    """

    def __init__(self, vocab_size=4095, seed=0, max_piece=6):
        rng = random.Random(seed)
        alphabet = "abcdefghijklmnopqrstuvwxyz'"
        pieces = ["<unk>"] + ["▁" + ch for ch in alphabet] + list(alphabet)
        seen = set(pieces)
        while len(pieces) < vocab_size:
            n = rng.randint(2, max_piece)
            body = "".join(rng.choice(alphabet) for _ in range(n))
            piece = ("▁" + body) if rng.random() < 0.5 else body
            if piece not in seen:
                seen.add(piece)
                pieces.append(piece)
        self.pieces = pieces[:vocab_size]
        self.index = {p: i for i, p in enumerate(self.pieces)}
        self.max_len = max(len(p) for p in self.pieces)

    def vocab_size(self):
        return len(self.pieces)

    def bos_id(self):
        return 0

    def decode(self, ids):
        text = "".join(self.pieces[int(i)] for i in ids if 0 < int(i) < len(self.pieces))
        return text.replace("▁", " ").strip()

    def encode(self, text):
        text = text.strip()
        if not text:
            return []
        s = "▁" + "▁".join(text.split())
        out, i, n = [], 0, len(s)
        while i < n:
            for ln in range(min(self.max_len, n - i), 0, -1):
                j = self.index.get(s[i:i + ln])
                if j is not None:
                    out.append(j)
                    i += ln
                    break
            else:
                out.append(0)
                i += 1
        return out


def synthetic_recordings(kind="earnings22", seed=5, frames_per_second=100, tokenizer=None, scale=1.0):
    """List of {id, text, audio, process_fn} like the dataset adapters return
    (lcasr/earnings22/run.py:63-75), with ``process_fn(rec) -> (spec [1,80,T] fp32 CPU, gold_text)``.

    Durations follow SURVEY.md §8d config 5: earnings22 6 x 20-70 min, tedlium 11 x 8-25 min,
    rev16 16 x 30-70 min.  ``scale`` shrinks the durations (tests, smoke).
    """
    shapes = {"earnings22": (6, 20, 70), "tedlium": (11, 8, 25), "rev16": (16, 30, 70)}
    n, lo, hi = shapes[kind]
    rng = random.Random(seed)
    recs = []
    for i in range(n):
        minutes = rng.uniform(lo, hi) * scale
        frames = max(64, int(minutes * 60 * frames_per_second))
        recs.append(_make_rec(f"{kind}_{i:02d}", frames, seed * 1000 + i, tokenizer))
    return recs


def _make_rec(rid, frames, seed, tokenizer):
    def process_fn(rec):
        g = torch.Generator().manual_seed(rec["seed"])
        spec = torch.randn(1, 80, rec["frames"], generator=g)     # normalised log-mel stand-in, N(0,1)
        return spec, rec["text"]
    text = ""
    if tokenizer is not None:
        rng = random.Random(seed)
        n_words = max(1, int(frames / 100 * 2.5))                  # ~150 words per minute
        ids = [rng.randrange(1, tokenizer.vocab_size()) for _ in range(n_words)]
        text = " ".join(w for w in (tokenizer.decode([i]) for i in ids) if w)
    return {"id": rid, "text": text, "audio": None, "frames": frames, "seed": seed, "process_fn": process_fn}


def build_model(vocab_size, model_cfg=None, device="cuda", seed=0):
    torch.manual_seed(seed)
    cfg = dict(LCASR160RB1 if model_cfg is None else model_cfg)
    model = StandInSCConformer(vocab_size, **cfg)
    model.device = torch.device(device)
    return model.to(model.device).eval()


def peaky_log_probs(T, C, blank, seed, p_blank=0.6, sharp=5.0, rng=None):
    """Synthetic CTC posteriors: one dominant class per frame with runs, fp32 log-softmax."""
    rng = np.random.default_rng(seed) if rng is None else rng
    z = rng.standard_normal((T, C)).astype(np.float32)
    cls = rng.integers(1, C - 1, size=T)
    cls[rng.random(T) < p_blank] = blank
    for t in range(1, T):
        if rng.random() < 0.4:
            cls[t] = cls[t - 1]
    z[np.arange(T), cls] += sharp
    second = rng.integers(1, C - 1, size=T)                # a competitor so the beam actually branches
    z[np.arange(T), second] += sharp * rng.random(T).astype(np.float32)
    z = z - z.max(-1, keepdims=True)
    lse = np.log(np.exp(z.astype(np.float64)).sum(-1, keepdims=True))
    return (z - lse).astype(np.float32)
