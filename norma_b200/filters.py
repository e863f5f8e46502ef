"""Whisper mel filterbank ([n_mel][201] f32): Slaney mel scale, Slaney area normalisation, 0-8 kHz at 16 kHz,
n_fft = 400.  norma embeds the same tables as raw bytes
(/root/reference/src/models/whisper/whisper_mel_bytes/{80,128}.bytes, read at monolingual.rs:351-362); they are
regenerated here rather than copied (max abs difference 3.7e-9, tests/test_oracle_mel.py)."""
from __future__ import annotations

import numpy as np


def mel_filters(n_mel: int) -> np.ndarray:
    if n_mel not in (80, 128):
        raise ValueError(f"Unexpected number of mel bins (num_mel_bins), got: {n_mel}")  # whisper::Error::MelBins
    sr, n_fft = 16000, 400
    fftfreqs = np.linspace(0, sr / 2, n_fft // 2 + 1)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asarray(f, np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), m * f_sp)

    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(8000.0), n_mel + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mel + 2] - mel_f[:n_mel])
    return (w * enorm[:, None]).astype(np.float32)
