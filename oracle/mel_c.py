"""ctypes loader for oracle/mel_oracle.c — TEST INFRASTRUCTURE ONLY (see that file's header)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmel_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mel_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        _lib.oracle_log_mel_spectrogram.argtypes = [fp, ctypes.c_size_t, fp, ctypes.c_size_t, ctypes.c_size_t,
                                                    ctypes.c_size_t, ctypes.c_int, fp, ctypes.POINTER(ctypes.c_size_t)]
        _lib.oracle_log_mel_spectrogram.restype = ctypes.c_int
        _lib.oracle_fft.argtypes = [fp, ctypes.c_size_t, fp]
        _lib.oracle_fft.restype = None
        _lib.oracle_sinusoids.argtypes = [ctypes.c_size_t, ctypes.c_size_t, fp]
        _lib.oracle_sinusoids.restype = None
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def log_mel_spectrogram(samples, filters, fft_size, fft_step, n_mel, n_threads=None) -> np.ndarray:
    """candle `log_mel_spectrogram_` -> flat f32 array of n_mel * n_len (mel-major)."""
    samples = np.ascontiguousarray(samples, np.float32)
    filters = np.ascontiguousarray(filters, np.float32)
    n_threads = n_threads or (os.cpu_count() or 2)
    n_len = ctypes.c_size_t(0)
    dummy = np.zeros(1, np.float32)
    sp = _fp(samples) if samples.size else _fp(dummy)
    lib().oracle_log_mel_spectrogram(sp, samples.size, _fp(filters), fft_size, fft_step, n_mel, n_threads, None, ctypes.byref(n_len))
    out = np.empty(n_mel * n_len.value, np.float32)
    lib().oracle_log_mel_spectrogram(sp, samples.size, _fp(filters), fft_size, fft_step, n_mel, n_threads, _fp(out), ctypes.byref(n_len))
    return out


def pcm_to_mel(pcm, filters, n_threads=None) -> np.ndarray:
    """candle `pcm_to_mel` -> [n_mel, n_len] f32 (all n_len frames; norma narrows to min(3000, n_len))."""
    filters = np.ascontiguousarray(filters, np.float32)
    n_mel = filters.shape[0]
    return log_mel_spectrogram(pcm, filters, 400, 160, n_mel, n_threads).reshape(n_mel, -1)


def fft(x) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(2 * x.size, np.float32)
    lib().oracle_fft(_fp(x), x.size, _fp(out))
    return out[0::2] + 1j * out[1::2]


def sinusoids(length: int, channels: int) -> np.ndarray:
    out = np.empty((length, channels), np.float32)
    lib().oracle_sinusoids(length, channels, _fp(out))
    return out
