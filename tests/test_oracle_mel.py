"""The mel oracle against everything the reference pins for this path (SURVEY.md §8 c): the two filterbank byte
files, candle's two shape known-answer tests, the frame-count table; plus an independent fp64 evaluation and the
committed golden vectors."""
import hashlib
import os

import numpy as np
import pytest

from conftest import REFERENCE, golden
from norma_b200 import filters, synth
from oracle import mel_c
from oracle.whisper_oracle import mel_n_len, pcm_to_mel_fp64, pcm_to_mel_numpy, slaney_mel_filters

SHA = {80: "85818f156f7e1894", 128: "2a5f9822897750e0"}  # SURVEY §8 c: pinned artefacts of the reference


@pytest.mark.parametrize("n_mel", [80, 128])
def test_filterbank_matches_reference_bytes(n_mel):
    path = f"{REFERENCE}/src/models/whisper/whisper_mel_bytes/{n_mel}.bytes"
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    raw = open(path, "rb").read()
    assert hashlib.sha256(raw).hexdigest().startswith(SHA[n_mel])
    ref = np.frombuffer(raw, "<f4").reshape(n_mel, 201)
    for gen in (filters.mel_filters(n_mel), slaney_mel_filters(n_mel)):
        assert np.abs(gen - ref).max() < 5e-9
        assert ((gen != 0) == (ref != 0)).all()  # same sparsity pattern (the banded kernel relies on it)
    nz = ref != 0
    assert nz.sum() == {80: 391, 128: 394}[n_mel]
    assert not nz[:, 0].any() and not nz[:, 200].any()
    span = [np.flatnonzero(r).max() - np.flatnonzero(r).min() + 1 for r in nz]
    assert max(span) <= 16  # MEL_BAND in csrc/mel.cu


def test_candle_shape_known_answers():
    # candle-transformers whisper::audio tests: log_mel_spectrogram_(zeros(1000), zeros(1000), 100, 10, 10, false)
    assert mel_c.log_mel_spectrogram(np.zeros(1000, np.float32), np.zeros(1000, np.float32), 100, 10, 10).size == 30_000
    assert mel_c.log_mel_spectrogram(np.zeros(100, np.float32), np.zeros(100, np.float32), 20, 2, 2).size == 6_000


@pytest.mark.parametrize("n,expect", [(0, 1500), (159, 1500), (160, 3000), (16000, 3000), (240_159, 3000), (240_160, 4500), (480_000, 4500)])
def test_frame_count_table(n, expect):
    assert mel_n_len(n) == expect
    f = filters.mel_filters(80)
    assert mel_c.pcm_to_mel(np.zeros(n, np.float32), f).shape == (80, expect)


def test_silence_is_the_clamp_floor():
    f = filters.mel_filters(80)
    mel = mel_c.pcm_to_mel(np.zeros(4000, np.float32), f)
    assert np.all(mel == np.float32(-10.0 / 4 + 1))  # log10(1e-10) = -10 everywhere, max-8 floor inactive


def test_fft_restatement_matches_numpy_fft():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(400).astype(np.float32)
    got = mel_c.fft(x)
    ref = np.fft.fft(x.astype(np.float64))
    assert np.abs(got - ref).max() < 2e-4 * np.abs(ref).max()


@pytest.mark.parametrize("n_mel", [80, 128])
def test_c_vs_numpy_vs_fp64(n_mel):
    f = filters.mel_filters(n_mel)
    pcm = synth.synth_pcm("gauss", 0, 160_000)
    a = mel_c.pcm_to_mel(pcm, f)
    b = pcm_to_mel_numpy(pcm, f)
    c = pcm_to_mel_fp64(pcm, f)
    assert a.shape == b.shape == c.shape
    assert np.abs(a - b).max() < 3e-5  # same recursion tree, different summation order in the projection
    assert np.abs(a - c).max() < 1e-4  # the f32 recursive FFT itself sits 2e-5..5e-5 from exact


def test_thread_count_does_not_change_result():
    f = filters.mel_filters(80)
    pcm = synth.synth_pcm("uniform", 5, 50_000)
    assert np.array_equal(mel_c.pcm_to_mel(pcm, f, n_threads=2), mel_c.pcm_to_mel(pcm, f, n_threads=12))


@pytest.mark.parametrize("fix,n_mel,kind,seed,n", [("mel_gauss0_80.npz", 80, "gauss", 0, 480_000), ("mel_gauss0_128.npz", 128, "gauss", 0, 480_000),
                                                   ("mel_short_80.npz", 80, "gauss", 3, 100_000)])
def test_golden_mel(fix, n_mel, kind, seed, n):
    g = golden(fix)
    mel = mel_c.pcm_to_mel(synth.synth_pcm(kind, seed, n), filters.mel_filters(n_mel))
    assert mel.shape[1] == int(g["n_len"])
    assert np.abs(mel[:, :8] - g["first8"]).max() < 1e-6
    assert np.abs(mel[:, :3000:25] - g["strided"]).max() < 1e-6
    assert np.abs(mel[:, -1] - g["tail"]).max() < 1e-6
