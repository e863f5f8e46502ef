"""Generates the committed golden fixtures from the CPU oracle (run once, here; the script is the provenance).

The reference is Rust and cannot run in this image, and its tests pin no numbers, so these are
ORACLE outputs ("parity unpinned"): they freeze the restatement so that (a) the oracle cannot drift silently and
(b) the GPU tests at BASELINE.json's full size compare against a committed vector instead of a 30 s CPU run.

    python tests/golden/make_golden.py [--full]      (--full also regenerates the distil-large-v3 vector)
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from norma_b200 import filters, synth  # noqa: E402
from oracle import mel_c  # noqa: E402
from oracle.whisper_oracle import Config, GreedyDecoder, WhisperOracle, special_tokens_for_vocab  # noqa: E402

FRAME_STRIDE = 25  # mel fixtures keep frames 0, 25, 50, ... (and all of the first 8)
ROWS = [0, 1, 2, 100, 749, 1498, 1499]  # encoder fixture rows


def mel_fixture(n_mel: int, kind: str, seed: int, n: int):
    f = filters.mel_filters(n_mel)
    pcm = synth.synth_pcm(kind, seed, n)
    mel = mel_c.pcm_to_mel(pcm, f)
    return dict(n_len=np.int64(mel.shape[1]), first8=mel[:, :8].copy(), strided=mel[:, :3000:FRAME_STRIDE].copy(),
                tail=mel[:, -1].copy(), checksum=np.float64(mel.astype(np.float64).sum()))


def encoder_fixture(name: str, seed_w: int, seed_pcm: int):
    c = synth.model_config(name)
    w = synth.synth_weights(c, seed=seed_w, decoder=False)
    orc = WhisperOracle(Config(**c), w)
    f = filters.mel_filters(c["num_mel_bins"])
    pcm = synth.synth_pcm("gauss", seed_pcm)
    mel = mel_c.pcm_to_mel(pcm, f)[:, :3000]
    t = time.time()
    y = orc.encoder_forward(torch.from_numpy(mel[None]))[0].numpy()
    print(f"{name}: oracle encoder {time.time() - t:.1f} s")
    return dict(rows=np.asarray(ROWS), values=y[ROWS].copy(), fro=np.float64(np.linalg.norm(y.astype(np.float64))),
                absmax=np.float32(np.abs(y).max()), col_mean=y.mean(0).astype(np.float32))


def decode_fixture(name: str, max_steps: int):
    c = synth.model_config(name)
    w = synth.synth_weights(c, seed=1)
    orc = WhisperOracle(Config(**c), w)
    f = filters.mel_filters(c["num_mel_bins"])
    out = {}
    for i, (kind, seed) in enumerate((("gauss", 0), ("uniform", 1))):
        mel = mel_c.pcm_to_mel(synth.synth_pcm(kind, seed), f)[:, :3000]
        xa = orc.encoder_forward(torch.from_numpy(mel[None]))
        st = special_tokens_for_vocab(c["vocab_size"])
        dr = GreedyDecoder(orc, st).decode(xa, max_steps=max_steps)
        out[f"tokens{i}"] = np.asarray(dr.tokens, np.int64)
        out[f"avg_logprob{i}"] = np.float64(dr.avg_logprob)
        out[f"no_speech_prob{i}"] = np.float64(dr.no_speech_prob)
        out[f"margins{i}"] = np.asarray(dr.margins, np.float64)
    out["max_steps"] = np.int64(max_steps)
    return out


def main():
    full = "--full" in sys.argv
    np.savez(os.path.join(HERE, "mel_gauss0_80.npz"), **mel_fixture(80, "gauss", 0, 480000))
    np.savez(os.path.join(HERE, "mel_gauss0_128.npz"), **mel_fixture(128, "gauss", 0, 480000))
    np.savez(os.path.join(HERE, "mel_short_80.npz"), **mel_fixture(80, "gauss", 3, 100000))
    np.savez(os.path.join(HERE, "enc_tiny_en.npz"), **encoder_fixture("tiny.en", 1, 0))
    np.savez(os.path.join(HERE, "decode_tiny_en.npz"), **decode_fixture("tiny.en", 16))
    if full:
        np.savez(os.path.join(HERE, "enc_distil_large_v3.npz"), **encoder_fixture("distil-large-v3", 1, 0))


if __name__ == "__main__":
    main()
