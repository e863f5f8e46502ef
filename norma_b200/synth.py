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
    # multilingual checkpoints (multilingual::ModelType, /root/reference/src/models/whisper/multilingual.rs:47-57)
    "tiny": (80, 384, 6, 4, 4, 51865),
    "base": (80, 512, 8, 6, 6, 51865),
    "small": (80, 768, 12, 12, 12, 51865),
    "medium": (80, 1024, 16, 24, 24, 51865),
    "large": (80, 1280, 20, 32, 32, 51865),
    "large-v2": (80, 1280, 20, 32, 32, 51865),
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


def special_tokens(vocab_size: int, task: str = "transcribe", lang: "int | None" = 0) -> Dict[str, int]:
    """Public Whisper special-token layouts by vocab size (norma reads them from tokenizer.json,
    /root/reference/src/models/whisper/monolingual.rs:376-384,419-420; no tokenizer file exists offline).
    `lang` = index into the 99 `Language` variants (0 = English) or None (multilingual `LanguageState::Detect`:
    no language token until one is detected); `task` = "transcribe" | "translate" (multilingual.rs:239-242)."""
    table = {
        51864: dict(sot=50257, eot=50256, translate=50357, task=50358, lang=50258, no_speech=50361, no_timestamps=50362, ts_zero=50363, ts_one=50413),
        51865: dict(sot=50258, eot=50257, translate=50358, task=50359, lang=50259, no_speech=50362, no_timestamps=50363, ts_zero=50364, ts_one=50414),
        51866: dict(sot=50258, eot=50257, translate=50359, task=50360, lang=50259, no_speech=50363, no_timestamps=50364, ts_zero=50365, ts_one=50415),
    }
    t = dict(table[vocab_size])
    translate = t.pop("translate")
    if task == "translate":
        t["task"] = translate
    elif task != "transcribe":
        raise ValueError(task)
    t["lang"] = None if lang is None else t["lang"] + int(lang)
    return t


def language_tokens(vocab_size: int):
    """ids of the 99 `Language` tokens in declaration order (languages.rs:7-107): consecutive from <|en|>."""
    en = {51864: 50258, 51865: 50259, 51866: 50259}[vocab_size]
    return [en + i for i in range(99)]


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


# ---- synthetic checkpoint FILES (there is no network: no hub download) -------------------------------------------------
def bytes_to_unicode() -> Dict[int, str]:
    """GPT-2 byte alphabet used by byte-level BPE vocabularies (tokenizers `bytes_char`)."""
    bs = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return {b: chr(c) for b, c in zip(bs, cs)}


LANGUAGE_CODES = (
    "en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur hr bg lt la mi ml cy sk te fa lv bn "
    "sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si km sn yo so af oc ka be tg sd gu am yi lo uz fo ht ps tk nn mt sa lb my "
    "bo tl mg as tt haw ln ha ba jw su").split()


def whisper_added_tokens(vocab_size: int):
    """[(id, content, special)] in the public Whisper layout for this vocab size."""
    eot = {51864: 50256, 51865: 50257, 51866: 50257}[vocab_size]
    names = ["<|endoftext|>", "<|startoftranscript|>"] + [f"<|{c}|>" for c in LANGUAGE_CODES]
    if vocab_size == 51866:
        names.append("<|yue|>")
    names += ["<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>",
              "<|nospeech|>" if vocab_size == 51866 else "<|nocaptions|>", "<|notimestamps|>"]
    out = [(eot + i, n, True) for i, n in enumerate(names)]
    t0 = eot + len(names)
    out += [(t0 + i, f"<|{i * 0.02:.2f}|>", False) for i in range(1501)]
    assert out[-1][0] == vocab_size - 1, (out[-1], vocab_size)
    return out


def synth_tokenizer_json(vocab_size: int, seed: int = 3, added=None) -> str:
    """A byte-level-BPE tokenizer.json with Whisper's special-token layout and a synthetic text vocabulary: the 256 byte
    tokens, then seeded random byte strings (1-6 bytes, so some are partial UTF-8 sequences) up to <|endoftext|>.
    `added` overrides the added-token list [(id, content, special)]; text ids fill [0, added[0].id)."""
    import json

    rng = np.random.default_rng(seed)
    b2u = bytes_to_unicode()
    added = whisper_added_tokens(vocab_size) if added is None else added
    n_text = added[0][0]
    vocab: Dict[str, int] = {}
    for b in range(256):
        vocab[b2u[b]] = len(vocab)
    words = [w.encode() for w in (" the", " a", " hello", " world", " café", " 你好", " \U0001f642", "ing", "ed", ".", ",")]
    for w in words:
        vocab.setdefault("".join(b2u[x] for x in w), len(vocab))
    while len(vocab) < n_text:
        k = int(rng.integers(2, 7))
        bs = rng.integers(0, 256, k) if rng.random() < 0.3 else rng.integers(97, 123, k)
        vocab.setdefault("".join(b2u[int(x)] for x in bs), len(vocab))
    j = {
        "version": "1.0", "truncation": None, "padding": None,
        "added_tokens": [dict(id=i, content=c, single_word=False, lstrip=False, rstrip=False, normalized=False, special=s) for i, c, s in added],
        "normalizer": None,
        "pre_tokenizer": {"type": "ByteLevel", "add_prefix_space": False, "trim_offsets": True, "use_regex": True},
        "post_processor": None,
        "decoder": {"type": "ByteLevel", "add_prefix_space": True, "trim_offsets": True, "use_regex": True},
        "model": {"type": "BPE", "dropout": None, "unk_token": None, "continuing_subword_prefix": "", "end_of_word_suffix": "",
                  "fuse_unk": False, "byte_fallback": False, "ignore_merges": False, "vocab": vocab, "merges": []},
    }
    return json.dumps(j, ensure_ascii=False)


def write_safetensors(path: str, tensors: Dict[str, "np.ndarray | torch.Tensor"], metadata: "Dict[str, str] | None" = None):
    """safetensors 0.4 layout: u64 LE header length, JSON header (name -> dtype / shape / data_offsets), raw little-endian data."""
    import json

    names = {np.dtype(np.float32): "F32", np.dtype(np.float16): "F16", np.dtype(np.float64): "F64"}
    header, blobs, off = {}, [], 0
    if metadata:
        header["__metadata__"] = metadata
    for k, v in tensors.items():
        if isinstance(v, torch.Tensor) and v.dtype == torch.bfloat16:
            raw, dt, shape = v.contiguous().view(torch.int16).numpy().tobytes(), "BF16", list(v.shape)
        else:
            a = np.ascontiguousarray(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v)
            raw, dt, shape = a.tobytes(), names[a.dtype], list(a.shape)
        header[k] = {"dtype": dt, "shape": shape, "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    h = json.dumps(header, separators=(",", ":")).encode()
    h += b" " * (-len(h) % 8)
    with open(path, "wb") as f:
        f.write(len(h).to_bytes(8, "little"))
        f.write(h)
        for b in blobs:
            f.write(b)


def write_checkpoint(directory: str, cfg: Dict[str, int], weights: Dict[str, torch.Tensor], suppress_tokens=(), seed: int = 3):
    """config.json + tokenizer.json + model.safetensors as the hub would serve them -> their three paths."""
    import json
    import os

    os.makedirs(directory, exist_ok=True)
    hf = dict(cfg)
    hf.update(architectures=["WhisperForConditionalGeneration"], model_type="whisper", suppress_tokens=list(suppress_tokens),
              activation_function="gelu", dropout=0.0, torch_dtype="float32")  # extra keys serde ignores
    paths = [os.path.join(directory, n) for n in ("config.json", "tokenizer.json", "model.safetensors")]
    with open(paths[0], "w") as f:
        json.dump(hf, f)
    with open(paths[1], "w", encoding="utf-8") as f:
        f.write(synth_tokenizer_json(cfg["vocab_size"], seed))
    w = dict(weights)
    w["model.encoder.embed_positions.weight"] = torch.zeros(cfg["max_source_positions"], cfg["d_model"])  # present in real files, ignored
    if "model.decoder.embed_tokens.weight" in w:
        w["proj_out.weight"] = w["model.decoder.embed_tokens.weight"]  # tied copy some checkpoints carry
    write_safetensors(paths[2], w, {"format": "pt"})
    return tuple(paths)


def quantize_q8_0(a: np.ndarray):
    """ggml `quantize_row_q8_0`: blocks of 32 along the last axis, d = max|x| / 127 stored as f16, q = round(x / d).
    -> (raw block bytes, the f32 values those blocks dequantise to)."""
    x = np.ascontiguousarray(a, np.float32).reshape(-1, 32)
    amax = np.abs(x).max(axis=1)
    d = (amax / 127.0).astype(np.float32)
    inv = np.where(d > 0, 1.0 / np.where(d > 0, d, 1.0), 0.0).astype(np.float32)
    q = np.rint(x * inv[:, None]).astype(np.int8)
    d16 = d.astype(np.float16)
    blocks = np.zeros((x.shape[0], 34), np.uint8)
    blocks[:, :2] = d16.view(np.uint8).reshape(-1, 2)
    blocks[:, 2:] = q.view(np.uint8)
    deq = (d16.astype(np.float32)[:, None] * q.astype(np.float32)).reshape(a.shape)
    return blocks.tobytes(), deq


def write_gguf(path: str, tensors: Dict[str, "np.ndarray | torch.Tensor"], version: int = 3, alignment: int = 32):
    """GGUF like candle's `model-{ext}-q80.gguf`: matrices (rank >= 2, last dimension a multiple of 32) as Q8_0, everything else F32.
    Returns {name: the f32 values a reader dequantises to}."""
    import struct

    def gstr(s: str) -> bytes:
        b = s.encode()
        return struct.pack("<Q", len(b)) + b

    infos, blobs, deq, off = [], [], {}, 0
    for name, v in tensors.items():
        a = np.ascontiguousarray(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v, np.float32)
        if a.ndim >= 2 and a.shape[-1] % 32 == 0:
            raw, d = quantize_q8_0(a)
            typ = 8
        else:
            raw, d, typ = a.tobytes(), a, 0
        deq[name] = d
        pad = (-len(raw)) % alignment
        infos.append((name, a.shape, typ, off))
        blobs.append(raw + b"\0" * pad)
        off += len(raw) + pad
    kv = gstr("general.architecture") + struct.pack("<I", 8) + gstr("whisper")
    kv += gstr("general.alignment") + struct.pack("<II", 4, alignment)
    kv += gstr("test.array") + struct.pack("<IIQ", 9, 5, 3) + struct.pack("<iii", 1, 2, 3)
    kv += gstr("test.strings") + struct.pack("<IIQ", 9, 8, 2) + gstr("a") + gstr("bc")
    head = struct.pack("<IIQQ", 0x46554747, version, len(infos), 4) + kv
    for name, shape, typ, o in infos:
        head += gstr(name) + struct.pack("<I", len(shape)) + b"".join(struct.pack("<Q", int(x)) for x in reversed(shape)) + struct.pack("<IQ", typ, o)
    head += b"\0" * ((-len(head)) % alignment)
    with open(path, "wb") as f:
        f.write(head)
        for b in blobs:
            f.write(b)
    return deq
