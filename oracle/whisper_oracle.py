"""oracle/whisper_oracle.py — TEST INFRASTRUCTURE ONLY.

CPU (torch, fp32 or fp64) restatement of the Whisper graph and of norma's greedy decode loop.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import this module; the
product path (`norma_b200/`) never does and fails loudly without its CUDA library.

What is restated, and from where:
  * encoder / decoder / final_linear — the graph norma reaches through its 5-method seam
    `/root/reference/src/models/whisper/model.rs:442-491` (`Type::{encoder_forward, decoder_forward,
    decoder_final_linear, reset_kv_cache, config}`).  The arithmetic lives in the un-vendored crate
    candle-transformers 0.7.2 (`/root/reference/Cargo.lock:293-294`, `models::whisper::model`), restated from
    its published algorithm: tanh-GELU, LayerNorm eps 1e-5, k_proj without bias, q and k both
    scaled by head_dim^-0.25, computed (not loaded) sinusoids as [sin | cos] halves, cross-attention-only KV
    cache, tied-embedding logits.
  * decode(t = 0) and the suppression rules — `/root/reference/src/models/whisper/model.rs:212-390`
    (masks are added to *probabilities*, greedy arg-max takes the LAST index among ties, stop rule at
    `max_target_positions - 1`, trailing-timestamp strip), masks built as in
    `/root/reference/src/models/whisper/monolingual.rs:386-430`.
  * log-mel — a vectorised numpy-f32 restatement of the same recursion tree as `oracle/mel_oracle.c`
    (used to cross-check the C restatement) and an independent fp64 STFT evaluation.

PARITY UNPINNED: the reference's tests never run a forward pass, so no golden vector exists at
this boundary.  The restatement is pinned structurally against HF `transformers` WhisperModel
(tests/test_oracle_model.py) and the pinned filterbank fixtures.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = 480_000
N_FRAMES = 3000
NO_SPEECH_THRESHOLD = 0.6
LOGPROB_THRESHOLD = -1.0
TEMPERATURES = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0)


# ----------------------------------------------------------------------------------------------------------
# log-mel (numpy restatements; the C file is the primary oracle, these cross-check it)
# ----------------------------------------------------------------------------------------------------------
def mel_n_len(n_samples: int) -> int:
    """candle `log_mel_spectrogram_` frame-count rule (SURVEY §8 c-1 rule 3)."""
    n_len = n_samples // HOP_LENGTH
    pad = 100 * CHUNK_LENGTH // 2
    if n_len % pad != 0:
        n_len = (n_len // pad + 1) * pad
    return n_len + pad


def _fft_rec_f32(x: np.ndarray) -> np.ndarray:
    """Batched restatement of candle `fft`/`dft` on rows of x (f32 [F, n]) -> complex parts (re, im) f32."""
    n = x.shape[1]
    f32 = np.float32
    two_pi = f32(np.pi) + f32(np.pi)
    if n == 1:
        return x.copy(), np.zeros_like(x)
    if n % 2 == 1:
        k = np.arange(n, dtype=f32)[:, None]
        j = np.arange(n, dtype=f32)[None, :]
        angle = (two_pi * k * j / f32(n)).astype(f32)
        c, s = np.cos(angle).astype(f32), np.sin(angle).astype(f32)
        re = np.zeros((x.shape[0], n), f32)
        im = np.zeros((x.shape[0], n), f32)
        for jj in range(n):  # sequential accumulation order as in the scalar loop
            re += x[:, jj : jj + 1] * c[None, :, jj]
            im -= x[:, jj : jj + 1] * s[None, :, jj]
        return re, im
    er, ei = _fft_rec_f32(np.ascontiguousarray(x[:, 0::2]))
    orr, oi = _fft_rec_f32(np.ascontiguousarray(x[:, 1::2]))
    k = np.arange(n // 2, dtype=f32)
    theta = (two_pi * k / f32(n)).astype(f32)
    re_t, im_t = np.cos(theta).astype(f32), (-np.sin(theta)).astype(f32)
    out_re = np.concatenate([er + re_t * orr - im_t * oi, er - re_t * orr + im_t * oi], axis=1)
    out_im = np.concatenate([ei + re_t * oi + im_t * orr, ei - re_t * oi - im_t * orr], axis=1)
    return out_re.astype(f32), out_im.astype(f32)


def pcm_to_mel_numpy(pcm: np.ndarray, filters: np.ndarray) -> np.ndarray:
    """numpy-f32 restatement of candle pcm_to_mel -> [n_mel, n_len] (mel-major, all n_len frames)."""
    f32 = np.float32
    pcm = np.asarray(pcm, f32)
    n_mel = filters.shape[0]
    n_len = mel_n_len(len(pcm))
    samples = np.zeros(n_len * HOP_LENGTH + N_FFT, f32)
    samples[: len(pcm)] = pcm
    two_pi = f32(np.pi) + f32(np.pi)
    hann = (f32(0.5) * (f32(1.0) - np.cos((two_pi * np.arange(N_FFT, dtype=f32)) / f32(N_FFT)).astype(f32))).astype(f32)
    idx = np.arange(n_len)[:, None] * HOP_LENGTH + np.arange(N_FFT)[None, :]
    frames = (samples[idx] * hann[None, :]).astype(f32)
    # candle zero-fills past n_len*160 (samples buffer is exactly n_len*160 long)
    frames[idx >= n_len * HOP_LENGTH] = 0
    re, im = _fft_rec_f32(frames)
    p = (re * re + im * im).astype(f32)
    p[:, 1 : N_FFT // 2] += p[:, N_FFT - 1 : N_FFT // 2 : -1]
    p = p[:, : N_FFT // 2 + 1]
    s = (p @ filters.T.astype(f32)).astype(f32)  # summation order differs from the scalar loop (cross-check only)
    mel = np.log10(np.maximum(s, f32(1e-10))).astype(f32).T
    mmax = mel.max() - f32(8.0)
    return (np.maximum(mel, mmax) / f32(4.0) + f32(1.0)).astype(f32)


def pcm_to_mel_fp64(pcm: np.ndarray, filters: np.ndarray) -> np.ndarray:
    """Independent fp64 evaluation (exact DFT via torch.stft, center=False) of the same spec."""
    pcm64 = torch.as_tensor(np.asarray(pcm, np.float64))
    n_len = mel_n_len(len(pcm))
    x = torch.zeros(n_len * HOP_LENGTH + N_FFT, dtype=torch.float64)
    x[: len(pcm64)] = pcm64
    x[n_len * HOP_LENGTH :] = 0
    win = torch.hann_window(N_FFT, periodic=True, dtype=torch.float64)
    st = torch.stft(x, N_FFT, HOP_LENGTH, window=win, center=False, return_complex=True)[:, :n_len]
    p = st.real**2 + st.imag**2  # [201, n_len]
    p[1:200] *= 2.0
    s = torch.as_tensor(np.asarray(filters, np.float64)) @ p
    mel = torch.log10(torch.clamp(s, min=float(np.float32(1e-10))))
    mmax = mel.max() - 8.0
    return (torch.maximum(mel, mmax) / 4.0 + 1.0).numpy()


def slaney_mel_filters(n_mel: int) -> np.ndarray:
    """Regenerate the Whisper filterbank ([n_mel, 201] f32): Slaney scale, Slaney norm, 0..8 kHz at 16 kHz."""
    sr, n_fft = 16000, N_FFT
    fftfreqs = np.linspace(0, sr / 2, n_fft // 2 + 1)

    def hz_to_mel(f):
        f = np.asarray(f, np.float64)
        mels = f / (200.0 / 3)
        min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
        min_log_mel = min_log_hz / (200.0 / 3)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)

    def mel_to_hz(m):
        m = np.asarray(m, np.float64)
        freqs = m * (200.0 / 3)
        min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
        min_log_mel = min_log_hz / (200.0 / 3)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)

    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(8000.0), n_mel + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mel, n_fft // 2 + 1))
    for i in range(n_mel):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mel + 2] - mel_f[:n_mel])
    return (w * enorm[:, None]).astype(np.float32)


# ----------------------------------------------------------------------------------------------------------
# model
# ----------------------------------------------------------------------------------------------------------
@dataclass
class Config:
    num_mel_bins: int
    max_source_positions: int
    d_model: int
    encoder_attention_heads: int
    encoder_layers: int
    vocab_size: int
    max_target_positions: int
    decoder_attention_heads: int
    decoder_layers: int
    suppress_tokens: tuple = ()


def sinusoids(length: int, channels: int) -> torch.Tensor:
    """candle `sinusoids` (f32): [sin | cos] halves, inv_t[i] = exp(-i * ln(10000)/(channels/2-1)).
    Evaluated with libm (oracle/mel_oracle.c) like Rust's f32 intrinsics; at t*inv ~ 1500 rad one ulp of `inv`
    is 1e-4 in the angle, so the numpy variant below (kept for cross-checks) can differ by that much."""
    from . import mel_c
    return torch.from_numpy(mel_c.sinusoids(length, channels))


def sinusoids_numpy(length: int, channels: int) -> torch.Tensor:
    inc = np.float32(np.log(np.float32(10000.0))) / np.float32(channels // 2 - 1)
    inv = np.exp((np.arange(channels // 2, dtype=np.float32) * (-inc)).astype(np.float32)).astype(np.float32)
    t = np.arange(length, dtype=np.float32)[:, None] * inv[None, :]
    return torch.from_numpy(np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32))


def _gelu(x):
    return F.gelu(x, approximate="tanh")


def _ln(x, w, p):
    return F.layer_norm(x, (x.shape[-1],), w[p + ".weight"], w[p + ".bias"], 1e-5)


def _lin(x, w, p, bias=True):
    return F.linear(x, w[p + ".weight"], w[p + ".bias"] if bias else None)


def _qkv_attention(q, k, v, n_head, mask=None):
    B, n_ctx, n_state = q.shape
    hd = n_state // n_head
    scale = float(hd) ** -0.25
    q = q.view(B, n_ctx, n_head, hd).transpose(1, 2) * scale
    k = k.view(B, k.shape[1], n_head, hd).transpose(1, 2).transpose(2, 3) * scale
    v = v.view(B, v.shape[1], n_head, hd).transpose(1, 2)
    qk = q @ k
    if mask is not None:
        qk = qk + mask[:n_ctx, :n_ctx]
    w_ = torch.softmax(qk, dim=-1)
    return (w_ @ v).transpose(1, 2).flatten(2)


class WhisperOracle:
    """Functional restatement over a dict of HF-named tensors (`model.encoder...`, `model.decoder...`)."""

    def __init__(self, cfg: Config, weights: Dict[str, torch.Tensor], dtype=torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        self.w = {k: torch.as_tensor(v).to(dtype) for k, v in weights.items()}
        self.pos = sinusoids(cfg.max_source_positions, cfg.d_model).to(dtype)
        n = cfg.max_target_positions
        self.mask = torch.triu(torch.full((n, n), float("-inf"), dtype=dtype), diagonal=1)
        self.cross_kv: Optional[List] = None  # per decoder layer (k, v); cross-attention-only cache

    # -- seam (5) ------------------------------------------------------------------------------------------
    def reset_kv_cache(self):
        self.cross_kv = None

    # -- seam (2) ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def encoder_forward(self, mel: torch.Tensor, return_stages: bool = False):
        """mel [B, n_mel, T] -> [B, T/2, d]."""
        w, c = self.w, self.cfg
        x = mel.to(self.dtype)
        p = "model.encoder."
        x = _gelu(F.conv1d(x, w[p + "conv1.weight"], w[p + "conv1.bias"], stride=1, padding=1))
        x = _gelu(F.conv1d(x, w[p + "conv2.weight"], w[p + "conv2.bias"], stride=2, padding=1))
        x = x.transpose(1, 2)
        x = x + self.pos[: x.shape[1]]
        stages = {"stem": x.clone()} if return_stages else None
        for i in range(c.encoder_layers):
            lp = f"{p}layers.{i}."
            h = _ln(x, w, lp + "self_attn_layer_norm")
            q = _lin(h, w, lp + "self_attn.q_proj")
            k = _lin(h, w, lp + "self_attn.k_proj", bias=False)
            v = _lin(h, w, lp + "self_attn.v_proj")
            x = x + _lin(_qkv_attention(q, k, v, c.encoder_attention_heads), w, lp + "self_attn.out_proj")
            h = _ln(x, w, lp + "final_layer_norm")
            x = x + _lin(_gelu(_lin(h, w, lp + "fc1")), w, lp + "fc2")
            if return_stages and i == 0:
                stages["layer0"] = x.clone()
        x = _ln(x, w, p + "layer_norm")
        if return_stages:
            return x, stages
        return x

    # -- seam (3) ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def decoder_forward(self, tokens: torch.Tensor, xa: torch.Tensor, flush: bool) -> torch.Tensor:
        """tokens [B, n] (all tokens so far: no self-attention cache), xa [B, 1500, d] -> [B, n, d]."""
        w, c = self.w, self.cfg
        p = "model.decoder."
        n = tokens.shape[-1]
        x = w[p + "embed_tokens.weight"][tokens.long()] + w[p + "embed_positions.weight"][:n]
        if flush:
            self.cross_kv = None
        if self.cross_kv is None:
            kv = []
            for i in range(c.decoder_layers):
                lp = f"{p}layers.{i}.encoder_attn."
                kv.append((_lin(xa.to(self.dtype), w, lp + "k_proj", bias=False), _lin(xa.to(self.dtype), w, lp + "v_proj")))
            self.cross_kv = kv
        for i in range(c.decoder_layers):
            lp = f"{p}layers.{i}."
            h = _ln(x, w, lp + "self_attn_layer_norm")
            q = _lin(h, w, lp + "self_attn.q_proj")
            k = _lin(h, w, lp + "self_attn.k_proj", bias=False)
            v = _lin(h, w, lp + "self_attn.v_proj")
            x = x + _lin(_qkv_attention(q, k, v, c.decoder_attention_heads, self.mask), w, lp + "self_attn.out_proj")
            h = _ln(x, w, lp + "encoder_attn_layer_norm")
            q = _lin(h, w, lp + "encoder_attn.q_proj")
            ck, cv = self.cross_kv[i]
            x = x + _lin(_qkv_attention(q, ck, cv, c.decoder_attention_heads), w, lp + "encoder_attn.out_proj")
            h = _ln(x, w, lp + "final_layer_norm")
            x = x + _lin(_gelu(_lin(h, w, lp + "fc1")), w, lp + "fc2")
        return _ln(x, w, p + "layer_norm")

    # -- seam (4) ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def final_linear(self, x: torch.Tensor) -> torch.Tensor:
        return x @ self.w["model.decoder.embed_tokens.weight"].t()


@dataclass
class SpecialTokens:
    sot: int
    eot: int
    task: int  # <|transcribe|>
    lang: Optional[int]  # None => prompt is [sot, task]
    no_speech: int
    no_timestamps: int
    ts_zero: int  # <|0.00|>
    ts_one: int  # <|1.00|>


def special_tokens_for_vocab(vocab_size: int) -> SpecialTokens:
    """Public Whisper token layouts; norma looks these up from tokenizer.json
    (/root/reference/src/models/whisper/monolingual.rs:376-384,419-420)."""
    if vocab_size == 51864:  # EnV1: *.en checkpoints
        return SpecialTokens(50257, 50256, 50358, 50258, 50361, 50362, 50363, 50413)
    if vocab_size == 51865:  # V1
        return SpecialTokens(50258, 50257, 50359, 50259, 50362, 50363, 50364, 50414)
    if vocab_size == 51866:  # V2: large-v3 / distil-large-v3
        return SpecialTokens(50258, 50257, 50360, 50259, 50363, 50364, 50365, 50415)
    raise ValueError(f"unknown vocab size {vocab_size}")


@dataclass
class DecodingResult:
    tokens: List[int]
    avg_logprob: float
    no_speech_prob: float
    compression_ratio: float = float("nan")
    # oracle-only diagnostics: top-2 probability margin at every step (to gate token parity, H5)
    margins: Optional[List[float]] = None


class GreedyDecoder:
    """norma `Model::decode` at t = 0 (/root/reference/src/models/whisper/model.rs:279-390)."""

    def __init__(self, model: WhisperOracle, st: SpecialTokens):
        self.m, self.st = model, st
        V = model.cfg.vocab_size
        ninf = float("-inf")
        i = torch.arange(V)
        sup = torch.zeros(V)
        for t in model.cfg.suppress_tokens:
            sup[t] = ninf
        sup[st.no_timestamps] = ninf
        self.suppress_tokens = sup  # monolingual.rs:386-395
        self.supress_non_timestamps = torch.where(i > st.no_timestamps, 0.0, ninf)  # :397-406
        self.supress_timestamps = torch.where(i > st.no_timestamps, ninf, 0.0)  # :408-417
        self.first_token_supress = torch.where((i < st.ts_zero) | (i > st.ts_one), ninf, 0.0)  # :421-430

    def _supress_past_timestamps(self, p, last_ts):  # model.rs:225-243
        i = torch.arange(p.shape[-1])
        return p + torch.where((i > self.st.no_timestamps) & (i <= last_ts), float("-inf"), 0.0)

    def _supress_non_timestamps(self, p, last_ts):  # model.rs:216-223
        return self._supress_past_timestamps(p, last_ts) + self.supress_non_timestamps

    def _supress_tokens(self, p, tokens, last_ts):  # model.rs:245-277
        st = self.st
        p = p + self.suppress_tokens
        l_token = tokens[-1]
        sl_token = tokens[-2] if len(tokens) >= 2 else None
        if l_token > st.no_timestamps:
            if sl_token is not None and sl_token >= st.eot:
                return p + self.supress_timestamps
            return self._supress_non_timestamps(p, last_ts)
        sum_prob_timestamp = float(p[st.no_timestamps + 1 :].sum())
        prob_non_timestamp = float(p[: st.no_timestamps].max())
        if sum_prob_timestamp >= prob_non_timestamp:
            return self._supress_non_timestamps(p, last_ts)
        return self._supress_past_timestamps(p, last_ts)

    @torch.no_grad()
    def decode(self, audio_features: torch.Tensor, max_steps: Optional[int] = None) -> DecodingResult:
        m, st = self.m, self.st
        sum_logprob = 0.0
        tokens = [st.sot]
        if st.lang is not None:
            tokens.append(st.lang)
        tokens.append(st.task)
        last_timestamp = None
        ys = m.decoder_forward(torch.tensor([tokens]), audio_features, True)
        logits = m.final_linear(ys[:1])[0, 0]
        no_speech_prob = float(torch.softmax(logits.float(), 0)[st.no_speech])
        if no_speech_prob > NO_SPEECH_THRESHOLD:
            return DecodingResult(tokens, 0.0, no_speech_prob, margins=[])
        margins = []
        steps = 0
        while tokens[-1] != st.eot:
            ys = m.decoder_forward(torch.tensor([tokens]), audio_features, False)
            logits = m.final_linear(ys[:1, -1:])[0, 0]
            p = torch.softmax(logits.float(), -1)
            if last_timestamp is not None:
                p = self._supress_tokens(p, tokens, last_timestamp)
            else:
                p = p + self.first_token_supress
            # Rust `max_by` keeps the LAST maximal element
            mx = p.max()
            next_token = int(torch.nonzero(p == mx)[-1])
            top2 = torch.topk(p, 2).values
            margins.append(float(top2[0] - top2[1]))
            if next_token > st.no_timestamps:
                last_timestamp = next_token
            tokens.append(next_token)
            sum_logprob += math.log(float(p[next_token])) if float(p[next_token]) > 0 else float("-inf")
            steps += 1
            if len(tokens) >= m.cfg.max_target_positions - 1 or (max_steps is not None and steps >= max_steps):
                tokens.append(st.eot)
                break
        avg_logprob = sum_logprob / len(tokens)
        while len(tokens) >= 2 and tokens[-2] > st.no_timestamps:
            del tokens[-2]
        return DecodingResult(tokens, avg_logprob, no_speech_prob, margins=margins)


    def select(self, p: torch.Tensor, prefix: List[int]):
        """One pass of the loop body of model.rs:333-356 on a given probability vector: suppression rules for `prefix` (prompt + tokens
        sampled so far), then the LAST maximal index.  -> (token, masked p, top-2 margin of the masked p)"""
        st = self.st
        plen = 3 if st.lang is not None else 2
        last_timestamp = next((t for t in reversed(prefix[plen:]) if t > st.no_timestamps), None)
        if last_timestamp is not None:
            p = self._supress_tokens(p, prefix, last_timestamp)
        else:
            p = p + self.first_token_supress
        mx = p.max()
        tok = int(torch.nonzero(p == mx)[-1])
        top2 = torch.topk(p, 2).values
        return tok, p, float(top2[0] - top2[1])

    @torch.no_grad()
    def teacher_forced(self, audio_features: torch.Tensor, tokens: List[int]):
        """The probability vectors the loop of model.rs:317-371 sees when it produces `tokens` (prompt + sampled tokens, no final eot
        needed), from ONE causal pass: position i of `decoder_forward(tokens[:-1])` equals the last position of
        `decoder_forward(tokens[:i + 1])` because decoder self-attention is masked, so this is the reference's per-step recomputation
        without its O(n^2) cost.  -> (raw softmax probabilities [n_sampled, V], no_speech_prob)"""
        m, st = self.m, self.st
        plen = 3 if st.lang is not None else 2
        ys = m.decoder_forward(torch.tensor([tokens[:-1]]), audio_features, True)
        no_speech_prob = float(torch.softmax(m.final_linear(ys[:1, :1])[0, 0].float(), 0)[st.no_speech])
        logits = m.final_linear(ys[:1, plen - 1:])[0]
        return torch.softmax(logits.float(), -1), no_speech_prob


@torch.no_grad()
def detect_language(model: WhisperOracle, sot: int, language_tokens: List[int], audio_features: torch.Tensor):
    """norma `Model::detect_language` (/root/reference/src/models/whisper/model.rs:194-210): one decoder pass over [[sot]] with
    flush = true, the logits of the language tokens only, softmax over those, then a STABLE descending sort by total_cmp — so the
    first of equal probabilities wins.  Returns (token id, probabilities in `language_tokens` order)."""
    ys = model.decoder_forward(torch.tensor([[sot]]), audio_features, True)
    logits = model.final_linear(ys[:1, :1])[0, 0]
    probs = torch.softmax(logits.float()[torch.tensor(language_tokens)], -1)
    order = sorted(range(len(language_tokens)), key=lambda i: -float(probs[i]))  # Python's sort is stable, like slice::sort_by
    return language_tokens[order[0]], probs.numpy()
