"""Seeded synthetic inputs (there is no network: no checkpoints, no audio).

Shapes follow the checkpoints norma names in `monolingual::ModelType`
(/root/reference/src/models/whisper/monolingual.rs:32-46) and the HF tensor names candle's `Whisper::load`
reads (/root/reference/src/models/whisper/monolingual.rs:371-373; SURVEY.md §8 b).  Generators are the ones
SURVEY.md §8 d fixes: matrices N(0, 0.02²), biases N(0, 0.02²), LN γ = 1 + N(0, 0.1²), β = N(0, 0.1²),
`embed_tokens` scaled so greedy decoding has usable logit margins.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

# name -> (n_mel, d_model, heads, enc_layers, dec_layers, vocab)
MODEL_SHAPES = {
    "test-micro": (80, 128, 2, 2, 2, 51864),  # not a checkpoint: smallest shape the kernels accept
    "tiny.en": (80, 384, 6, 4, 4, 51864),
    "base.en": (80, 512, 8, 6, 6, 51864),
    "small.en": (80, 768, 12, 12, 12, 51864),
    "medium.en": (80, 1024, 16, 24, 24, 51864),
    "distil-medium.en": (80, 1024, 16, 24, 2, 51865),
    "distil-large-v2": (80, 1280, 20, 32, 2, 51865),
    "distil-large-v3": (128, 1280, 20, 32, 2, 51866),
    "large-v3": (128, 1280, 20, 32, 32, 51866),
}


def model_config(name: str) -> Dict[str, int]:
    n_mel, d, h, le, ld, v = MODEL_SHAPES[name]
    return dict(
        num_mel_bins=n_mel,
        max_source_positions=1500,
        d_model=d,
        encoder_attention_heads=h,
        encoder_layers=le,
        vocab_size=v,
        max_target_positions=448,
        decoder_attention_heads=h,
        decoder_layers=ld,
    )


def special_tokens(vocab_size: int) -> Dict[str, int]:
    """Public Whisper special-token layouts by vocab size (norma reads them from tokenizer.json,
    /root/reference/src/models/whisper/monolingual.rs:376-384,419-420; no tokenizer file exists offline)."""
    table = {
        51864: dict(sot=50257, eot=50256, task=50358, lang=50258, no_speech=50361, no_timestamps=50362, ts_zero=50363, ts_one=50413),
        51865: dict(sot=50258, eot=50257, task=50359, lang=50259, no_speech=50362, no_timestamps=50363, ts_zero=50364, ts_one=50414),
        51866: dict(sot=50258, eot=50257, task=50360, lang=50259, no_speech=50363, no_timestamps=50364, ts_zero=50365, ts_one=50415),
    }
    return dict(table[vocab_size])


def synth_weights(cfg: Dict[str, int], seed: int = 1, embed_scale: float = 8.0, sigma: float = 0.02,
                  decoder: bool = True) -> Dict[str, torch.Tensor]:
    """Random-init weights of the named architecture, keyed by HF tensor name, fp32 on CPU."""
    g = torch.Generator().manual_seed(seed)
    d, n_mel = cfg["d_model"], cfg["num_mel_bins"]
    w: Dict[str, torch.Tensor] = {}

    def mat(*shape, s=sigma):
        return torch.randn(*shape, generator=g) * s

    def ln(prefix):
        w[prefix + ".weight"] = 1.0 + torch.randn(d, generator=g) * 0.1
        w[prefix + ".bias"] = torch.randn(d, generator=g) * 0.1

    def lin(prefix, out_f, in_f, bias=True):
        w[prefix + ".weight"] = mat(out_f, in_f)
        if bias:
            w[prefix + ".bias"] = mat(out_f)

    def attn(prefix):
        lin(prefix + ".q_proj", d, d)
        lin(prefix + ".k_proj", d, d, bias=False)
        lin(prefix + ".v_proj", d, d)
        lin(prefix + ".out_proj", d, d)

    e = "model.encoder."
    w[e + "conv1.weight"] = mat(d, n_mel, 3)
    w[e + "conv1.bias"] = mat(d)
    w[e + "conv2.weight"] = mat(d, d, 3)
    w[e + "conv2.bias"] = mat(d)
    for i in range(cfg["encoder_layers"]):
        p = f"{e}layers.{i}"
        attn(p + ".self_attn")
        ln(p + ".self_attn_layer_norm")
        lin(p + ".fc1", 4 * d, d)
        lin(p + ".fc2", d, 4 * d)
        ln(p + ".final_layer_norm")
    ln(e + "layer_norm")
    if decoder:
        dd = "model.decoder."
        w[dd + "embed_tokens.weight"] = mat(cfg["vocab_size"], d, s=sigma * embed_scale)
        w[dd + "embed_positions.weight"] = mat(cfg["max_target_positions"], d)
        for i in range(cfg["decoder_layers"]):
            p = f"{dd}layers.{i}"
            attn(p + ".self_attn")
            ln(p + ".self_attn_layer_norm")
            attn(p + ".encoder_attn")
            ln(p + ".encoder_attn_layer_norm")
            lin(p + ".fc1", 4 * d, d)
            lin(p + ".fc2", d, 4 * d)
            ln(p + ".final_layer_norm")
        ln(dd + "layer_norm")
    return w


def synth_pcm(kind: str, seed: int, n: int = 480_000) -> np.ndarray:
    """Synthetic 16 kHz PCM (f32).  `gauss`: N(0, 0.1²) (the broadband gate signal); `uniform`:
    U[-1, 1]; `bursts`: 1 s-gated noise (exercises the 1e-10 clamp and the max-8 dB floor); `chirp`."""
    rng = np.random.default_rng(seed)
    if kind == "gauss":
        return (0.1 * rng.standard_normal(n)).astype(np.float32)
    if kind == "uniform":
        return rng.uniform(-1.0, 1.0, n).astype(np.float32)
    if kind == "bursts":
        x = (0.1 * rng.standard_normal(n)).astype(np.float32)
        t = np.arange(n) // 16000
        x[(t % 2) == 1] = 0.0
        return x
    if kind == "chirp":
        t = np.arange(n) / 16000.0
        f = 100.0 + (3000.0 / 30.0) * t
        return (0.5 * np.sin(2 * np.pi * np.cumsum(f) / 16000.0) + 0.01 * rng.standard_normal(n)).astype(np.float32)
    if kind == "zeros":
        return np.zeros(n, np.float32)
    raise ValueError(kind)


def synth_pcm_window(seed: int, n: int = 480_000) -> np.ndarray:
    """Config-3 mix (SURVEY §8 d): Gaussian / uniform / gated bursts by window index."""
    return synth_pcm(("gauss", "uniform", "bursts")[seed % 3], seed, n)


def plant_decoder_plan(w: Dict[str, torch.Tensor], cfg: Dict[str, int], plan: Dict[int, int], seed: int = 7, pos_gain: float = 6.0,
                       tok_gain: float = 2.0) -> Dict[str, torch.Tensor]:
    """Make a random-init decoder CONFIDENT: after position p it predicts token plan[p] with probability ~1.

    Random weights give near-uniform logits, so norma's `decode_with_fallback` always falls through to entropy-seeded
    sampling.  Here the learned positional embedding of position p is a large vector along a random unit
    direction u_p and the (tied) embedding of plan[p] is aligned with u_p, so the logit of plan[p] at position p
    dominates; everything else stays random-init.  Tokens in `plan` must be distinct."""
    assert len(set(plan.values())) == len(plan)
    g = torch.Generator().manual_seed(seed)
    d = cfg["d_model"]
    w = dict(w)
    pos = w["model.decoder.embed_positions.weight"].clone()
    emb = w["model.decoder.embed_tokens.weight"].clone()
    q, _ = torch.linalg.qr(torch.randn(d, len(plan) + 4, generator=g))
    for k, (p, tok) in enumerate(sorted(plan.items())):
        u = q[:, k]
        pos[p] = pos_gain * u
        emb[tok] = tok_gain * u
    w["model.decoder.embed_positions.weight"] = pos
    w["model.decoder.embed_tokens.weight"] = emb
    return w
