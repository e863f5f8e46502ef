"""ctypes binding of include/norma_b200.h — the same declarations the Rust `norma-b200-sys` crate would make
(INTEGRATION.md).  There is no fallback: if libnorma_b200.so is missing or fails to load, importing callers get a
loud error, and every non-zero status raises `Nb200Error` carrying `nb200_last_error`."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NB200_LIB_PATH", os.path.join(_HERE, "libnorma_b200.so"))  # override: instrumented builds (scripts/probes)

NB200_OK = 0
STATUS_NAMES = {0: "OK", 1: "INVALID_ARG", 2: "CUDA_ERROR", 3: "OOM", 4: "NOT_LOADED", 5: "UNSUPPORTED_SHAPE", 6: "ARCH_MISMATCH", 7: "IO_ERROR", 8: "PARSE_ERROR", 9: "NOT_FOUND", 10: "BUFFER_TOO_SMALL"}
TASKS = {"transcribe": 0, "translate": 1}
DTYPES = {"f32": 0, "bf16": 1, "f16": 2, "f64": 3, "u8": 4, "u32": 5}
KERNEL_CLASSES = ["mel", "mel_norm", "gemm", "attn", "layernorm", "decode_gemv", "decode_attn", "decode_select", "misc"]
Q = dict(n_frames=0, enc_len=1, d_model=2, vocab=3, max_batch=4, kernel_launches=5, device_bytes=6, compute_dtype=7, max_target_positions=8)

# every symbol include/norma_b200.h declares (tests/test_abi.py checks the header against this and the .so)
SYMBOLS = [
    "nb200_device_count", "nb200_create", "nb200_destroy", "nb200_last_error", "nb200_query", "nb200_sync",
    "nb200_load_tensor", "nb200_finalize_weights", "nb200_set_mel_filters", "nb200_set_tokens", "nb200_set_suppress",
    "nb200_pcm_to_mel", "nb200_pcm_to_mel_batch", "nb200_encoder_forward", "nb200_transcode_batch", "nb200_stage_pcm",
    "nb200_run_resident", "nb200_fetch_features", "nb200_fetch_mel", "nb200_decoder_forward", "nb200_final_linear",
    "nb200_reset_kv_cache", "nb200_decode_greedy", "nb200_timer_start", "nb200_timer_stop", "nb200_profile_enable",
    "nb200_profile_read", "nb200_profile_reset", "nb200_flush_l2", "nb200_test_gemm", "nb200_test_gemm_perf", "nb200_test_attention", "nb200_test_attention_perf",
    "nb200_decode", "nb200_model_create", "nb200_model_destroy", "nb200_model_last_error", "nb200_model_set_vocab", "nb200_model_transcribe",
    "nb200_model_state", "nb200_model_script_push", "nb200_model_script_log",
    "nb200_stream_reset", "nb200_stream_push", "nb200_stream_drain", "nb200_stream_features",
    "nb200_transcode_submit", "nb200_transcode_collect",
    "nb200_detect_language", "nb200_model_set_language_detection", "nb200_model_language", "nb200_model_script_push_language",
    "nb200_config_from_file", "nb200_mel_filters", "nb200_tokenizer_from_file", "nb200_tokenizer_destroy", "nb200_tokenizer_token_to_id",
    "nb200_tokenizer_decode", "nb200_tokenizer_special_tokens", "nb200_tokenizer_language_tokens", "nb200_load_safetensors", "nb200_safetensors_read", "nb200_load_gguf", "nb200_gguf_read", "nb200_set_decode_mode",
    "nb200_model_set_tokenizer", "nb200_model_from_files",
    "nb200_set_audio_features", "nb200_decode_begin", "nb200_decode_advance", "nb200_decode_peek_logits", "nb200_decode_end",
    "nb200_model_last_result", "nb200_model_no_progress_windows",
]


class Nb200Error(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"nb200 status {status} ({STATUS_NAMES.get(status, '?')}): {msg}")
        self.status = status


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "num_mel_bins", "max_source_positions", "d_model", "encoder_attention_heads", "encoder_layers", "vocab_size",
        "max_target_positions", "decoder_attention_heads", "decoder_layers", "max_batch")]


class SpecialTokens(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("sot", "eot", "task", "lang", "no_speech", "no_timestamps", "ts_zero", "ts_one")]


_lib = None


def load_library() -> C.CDLL:
    """Load libnorma_b200.so (built in-tree by norma_b200/build.py).  Raises if it is missing: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m norma_b200.build` (there is no CPU/PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    p, sz, i, f32p, u32p = C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    sigs = {
        "nb200_device_count": ([C.POINTER(C.c_int)], i),
        "nb200_create": ([i, C.POINTER(Config), i, C.POINTER(p)], i),
        "nb200_destroy": ([p], None),
        "nb200_last_error": ([p], C.c_char_p),
        "nb200_query": ([p, i, C.POINTER(C.c_int64)], i),
        "nb200_sync": ([p], i),
        "nb200_load_tensor": ([p, C.c_char_p, p, i, C.POINTER(C.c_int64), i], i),
        "nb200_finalize_weights": ([p], i),
        "nb200_set_mel_filters": ([p, f32p, i], i),
        "nb200_set_tokens": ([p, C.POINTER(SpecialTokens)], i),
        "nb200_set_suppress": ([p, u32p, sz], i),
        "nb200_pcm_to_mel": ([p, f32p, sz, f32p, C.POINTER(sz)], i),
        "nb200_pcm_to_mel_batch": ([p, f32p, sz, sz, C.POINTER(sz), f32p], i),
        "nb200_encoder_forward": ([p, f32p, sz, f32p], i),
        "nb200_transcode_batch": ([p, f32p, sz, sz, C.POINTER(sz), f32p], i),
        "nb200_stage_pcm": ([p, f32p, sz, sz, C.POINTER(sz)], i),
        "nb200_run_resident": ([p, sz, i, i], i),
        "nb200_fetch_features": ([p, sz, f32p, sz], i),
        "nb200_fetch_mel": ([p, sz, f32p, sz], i),
        "nb200_decoder_forward": ([p, sz, u32p, sz, i, f32p], i),
        "nb200_final_linear": ([p, f32p, f32p], i),
        "nb200_reset_kv_cache": ([p], i),
        "nb200_decode_greedy": ([p, sz, sz, u32p, C.POINTER(sz), C.POINTER(C.c_double), C.POINTER(C.c_double)], i),
        "nb200_timer_start": ([p], i),
        "nb200_timer_stop": ([p, f32p], i),
        "nb200_profile_enable": ([p, i], i),
        "nb200_profile_read": ([p, f32p, C.POINTER(C.c_int64), C.POINTER(C.c_double)], i),
        "nb200_profile_reset": ([p], i),
        "nb200_flush_l2": ([p], i),
        "nb200_test_gemm": ([p, p, p, f32p, i, i, i, i, f32p], i),
        "nb200_test_gemm_perf": ([p, i, i, i, i, i, f32p], i),
        "nb200_test_attention": ([p, f32p, i, i, i, f32p], i),
        "nb200_test_attention_perf": ([p, f32p, i, i, i, i, f32p], i),
        "nb200_decode": ([p, sz, C.c_float, C.c_uint64, sz, u32p, C.POINTER(sz), C.POINTER(C.c_double), C.POINTER(C.c_double)], i),
        "nb200_model_create": ([p, C.POINTER(SpecialTokens), sz, C.c_uint64, C.POINTER(p)], i),
        "nb200_model_destroy": ([p], None),
        "nb200_model_last_error": ([p], C.c_char_p),
        "nb200_model_set_vocab": ([p, C.c_uint32, C.c_char_p, sz], i),
        "nb200_model_transcribe": ([p, f32p, sz, i, C.c_char_p, sz, C.POINTER(sz), u32p, sz, C.POINTER(sz)], i),
        "nb200_model_state": ([p, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)], i),
        "nb200_model_script_push": ([p, C.c_double, C.c_double, u32p, sz], i),
        "nb200_model_script_log": ([p, sz, C.POINTER(sz), C.POINTER(C.c_double)], i),
        "nb200_transcode_submit": ([p, f32p, sz, sz, C.POINTER(sz), f32p], i),
        "nb200_transcode_collect": ([p], i),
        "nb200_stream_reset": ([p], i),
        "nb200_stream_push": ([p, f32p, sz], i),
        "nb200_stream_drain": ([p, sz], i),
        "nb200_stream_features": ([p, i, f32p, f32p], i),
        "nb200_detect_language": ([p, sz, u32p, sz, u32p, f32p], i),
        "nb200_model_set_language_detection": ([p, u32p, sz], i),
        "nb200_model_language": ([p, u32p, C.POINTER(sz)], i),
        "nb200_model_script_push_language": ([p, C.c_uint32], i),
        "nb200_config_from_file": ([C.c_char_p, C.POINTER(Config), u32p, sz, C.POINTER(sz)], i),
        "nb200_mel_filters": ([i, f32p], i),
        "nb200_tokenizer_from_file": ([C.c_char_p, C.POINTER(p)], i),
        "nb200_tokenizer_destroy": ([p], None),
        "nb200_tokenizer_token_to_id": ([p, C.c_char_p, u32p], i),
        "nb200_tokenizer_decode": ([p, u32p, sz, i, C.c_char_p, sz, C.POINTER(sz)], i),
        "nb200_tokenizer_special_tokens": ([p, C.c_char_p, i, C.POINTER(SpecialTokens)], i),
        "nb200_tokenizer_language_tokens": ([p, u32p], i),
        "nb200_load_safetensors": ([p, C.c_char_p, C.POINTER(sz)], i),
        "nb200_safetensors_read": ([C.c_char_p, C.c_char_p, f32p, sz, C.POINTER(C.c_int64), C.POINTER(i)], i),
        "nb200_set_decode_mode": ([p, i], i),
        "nb200_load_gguf": ([p, C.c_char_p, C.POINTER(sz)], i),
        "nb200_gguf_read": ([C.c_char_p, C.c_char_p, f32p, sz, C.POINTER(C.c_int64), C.POINTER(i), C.POINTER(i)], i),
        "nb200_model_set_tokenizer": ([p, p], i),
        "nb200_set_audio_features": ([p, f32p, sz], i),
        "nb200_decode_begin": ([p, sz, C.c_float, C.c_uint64, sz], i),
        "nb200_decode_advance": ([p, sz, C.POINTER(i)], i),
        "nb200_decode_peek_logits": ([p, sz, f32p], i),
        "nb200_decode_end": ([p, u32p, C.POINTER(sz), C.POINTER(C.c_double), C.POINTER(C.c_double)], i),
        "nb200_model_last_result": ([p, C.c_char_p, sz, C.POINTER(sz), u32p, sz, C.POINTER(sz)], i),
        "nb200_model_no_progress_windows": ([p, C.POINTER(sz)], i),
        "nb200_model_from_files": ([i, C.c_char_p, C.c_char_p, C.c_char_p, i, C.c_char_p, i, sz, C.c_uint64, C.POINTER(p), C.POINTER(p)], i),
    }
    for name, (args, res) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    return lib


def _f32p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def f32_to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """round-to-nearest-even f32 -> bf16 bit pattern (uint16)"""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    return r


def bf16_round(a: np.ndarray) -> np.ndarray:
    return (f32_to_bf16_bits(a).astype(np.uint32) << 16).view(np.float32)


def _ck_global(lib, st: int):
    if st != NB200_OK:
        raise Nb200Error(st, (lib.nb200_last_error(None) or b"").decode())


def config_from_file(path: str):
    """`serde_json::from_str::<Config>` (monolingual.rs:347): returns (config dict, suppress_tokens list)."""
    lib = load_library()
    c = Config()
    n = C.c_size_t()
    sup = np.zeros(1 << 16, np.uint32)
    _ck_global(lib, lib.nb200_config_from_file(os.fsencode(path), C.byref(c), sup.ctypes.data_as(C.POINTER(C.c_uint32)), sup.size, C.byref(n)))
    cfg = {k: int(getattr(c, k)) for k, _ in Config._fields_ if k != "max_batch"}
    return cfg, sup[: n.value].tolist()


def mel_filters(n_mel: int) -> np.ndarray:
    lib = load_library()
    out = np.empty((max(n_mel, 1), 201), np.float32)
    _ck_global(lib, lib.nb200_mel_filters(n_mel, _f32p(out)))
    return out


def safetensors_read(path: str, name: str) -> np.ndarray:
    """One tensor of a .safetensors file converted to f32 (what `VarBuilder::from_mmaped_safetensors(.., F32, ..)` yields)."""
    lib = load_library()
    shape = (C.c_int64 * 8)()
    rank = C.c_int()
    _ck_global(lib, lib.nb200_safetensors_read(os.fsencode(path), name.encode(), None, 0, shape, C.byref(rank)))
    shp = tuple(int(shape[k]) for k in range(rank.value))
    out = np.empty(shp, np.float32)
    _ck_global(lib, lib.nb200_safetensors_read(os.fsencode(path), name.encode(), _f32p(out), out.size, None, None))
    return out


def gguf_read(path: str, name: str):
    """One tensor of a GGUF file dequantised to f32 -> (array, ggml type id)."""
    lib = load_library()
    shape = (C.c_int64 * 8)()
    rank, typ = C.c_int(), C.c_int()
    _ck_global(lib, lib.nb200_gguf_read(os.fsencode(path), name.encode(), None, 0, shape, C.byref(rank), C.byref(typ)))
    out = np.empty(tuple(int(shape[k]) for k in range(rank.value)), np.float32)
    _ck_global(lib, lib.nb200_gguf_read(os.fsencode(path), name.encode(), _f32p(out), out.size, None, None, None))
    return out, typ.value


class Tokenizer:
    """nb200_tokenizer: `tokenizers::Tokenizer::from_file` + token_to_id + decode (monolingual.rs:348; model.rs:147)."""

    def __init__(self, path: str):
        self.lib = load_library()
        h = C.c_void_p()
        _ck_global(self.lib, self.lib.nb200_tokenizer_from_file(os.fsencode(path), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.nb200_tokenizer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def token_to_id(self, token: str) -> int:
        v = C.c_uint32()
        _ck_global(self.lib, self.lib.nb200_tokenizer_token_to_id(self.h, token.encode(), C.byref(v)))
        return v.value

    def decode(self, ids: Sequence[int], skip_special_tokens: bool = True) -> str:
        t = np.ascontiguousarray(np.asarray(list(ids), np.uint32))
        n = C.c_size_t()
        ptr = t.ctypes.data_as(C.POINTER(C.c_uint32))
        _ck_global(self.lib, self.lib.nb200_tokenizer_decode(self.h, ptr, t.size, int(skip_special_tokens), None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value + 1)
        _ck_global(self.lib, self.lib.nb200_tokenizer_decode(self.h, ptr, t.size, int(skip_special_tokens), buf, n.value + 1, None))
        return buf.raw[: n.value].decode("utf-8")

    def special_tokens(self, language_token: Optional[str], task: str = "transcribe") -> Dict[str, int]:
        st = SpecialTokens()
        lt = None if language_token is None else language_token.encode()
        _ck_global(self.lib, self.lib.nb200_tokenizer_special_tokens(self.h, lt, TASKS[task], C.byref(st)))
        return {k: int(getattr(st, k)) for k, _ in SpecialTokens._fields_}

    def language_tokens(self):
        out = np.zeros(99, np.uint32)
        _ck_global(self.lib, self.lib.nb200_tokenizer_language_tokens(self.h, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out.tolist()


class Context:
    """One nb200_ctx: one GPU ordinal, one stream.  Mirrors what norma's whisper `Model` owns
    (/root/reference/src/models/whisper/model.rs:16-42) on the device side."""

    def __init__(self, cfg: Dict[str, int], ordinal: int = 0, compute: str = "bf16", max_batch: int = 1, handle=None):
        self.lib = load_library()
        self.cfg = dict(cfg)
        self.max_batch = max_batch
        self.compute = compute
        if handle is None:
            c = Config(**{k: int(cfg[k]) for k, _ in Config._fields_ if k != "max_batch"}, max_batch=max_batch)
            handle = C.c_void_p()
            st = self.lib.nb200_create(ordinal, C.byref(c), DTYPES[compute], C.byref(handle))
            if st != NB200_OK:
                raise Nb200Error(st, (self.lib.nb200_last_error(None) or b"").decode())
        self.h = handle  # handle != None: adopt a ctx made by nb200_model_from_files
        self.d = cfg["d_model"]
        self.n_mel = cfg["num_mel_bins"]
        self.V = cfg["vocab_size"]
        self.T = cfg["max_source_positions"]
        self.P = cfg["max_target_positions"]

    # -- plumbing ------------------------------------------------------------------------------------------
    def _ck(self, st: int):
        if st != NB200_OK:
            raise Nb200Error(st, (self.lib.nb200_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.nb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, key: str) -> int:
        v = C.c_int64()
        self._ck(self.lib.nb200_query(self.h, Q[key], C.byref(v)))
        return v.value

    def sync(self):
        self._ck(self.lib.nb200_sync(self.h))

    # -- load ----------------------------------------------------------------------------------------------
    def load_tensor(self, name: str, arr):
        a = np.ascontiguousarray(arr.detach().cpu().numpy() if hasattr(arr, "detach") else arr)
        if a.dtype == np.float32:
            dt = DTYPES["f32"]
        elif a.dtype == np.float16:
            dt = DTYPES["f16"]
        elif a.dtype == np.float64:
            dt = DTYPES["f64"]
        else:
            raise TypeError(f"unsupported dtype {a.dtype} for {name}")
        shape = (C.c_int64 * a.ndim)(*a.shape)
        self._ck(self.lib.nb200_load_tensor(self.h, name.encode(), a.ctypes.data_as(C.c_void_p), dt, shape, a.ndim))

    def load_safetensors(self, path: str, finalize: bool = True) -> int:
        """`VarBuilder::from_mmaped_safetensors` + `Whisper::load` (monolingual.rs:371-373)."""
        n = C.c_size_t()
        self._ck(self.lib.nb200_load_safetensors(self.h, os.fsencode(path), C.byref(n)))
        if finalize:
            self._ck(self.lib.nb200_finalize_weights(self.h))
        return n.value

    def load_gguf(self, path: str, finalize: bool = True) -> int:
        """`VarBuilder::from_gguf` + `quantized_model::Whisper::load` (monolingual.rs:364-369), dequantised at load."""
        n = C.c_size_t()
        self._ck(self.lib.nb200_load_gguf(self.h, os.fsencode(path), C.byref(n)))
        if finalize:
            self._ck(self.lib.nb200_finalize_weights(self.h))
        return n.value

    def load_weights(self, weights: Dict[str, object]):
        for k, v in weights.items():
            self.load_tensor(k, v)
        self._ck(self.lib.nb200_finalize_weights(self.h))

    def set_mel_filters(self, filters: np.ndarray):
        f = np.ascontiguousarray(filters, np.float32)
        self._ck(self.lib.nb200_set_mel_filters(self.h, _f32p(f), f.shape[0]))

    def set_tokens(self, sot, eot, task, lang, no_speech, no_timestamps, ts_zero, ts_one):
        t = SpecialTokens(sot, eot, task, 0xFFFFFFFF if lang is None else lang, no_speech, no_timestamps, ts_zero, ts_one)
        self._ck(self.lib.nb200_set_tokens(self.h, C.byref(t)))

    def set_suppress(self, ids: Sequence[int]):
        a = np.ascontiguousarray(np.asarray(list(ids), np.uint32))
        self._ck(self.lib.nb200_set_suppress(self.h, a.ctypes.data_as(C.POINTER(C.c_uint32)), a.size))

    # -- seam (1) ------------------------------------------------------------------------------------------
    def pcm_to_mel(self, pcm: np.ndarray) -> np.ndarray:
        """-> [n_mel, n_len] f32 exactly as candle's pcm_to_mel returns it."""
        pcm = np.ascontiguousarray(pcm, np.float32)
        n_len = C.c_size_t()
        self._ck(self.lib.nb200_pcm_to_mel(self.h, None, pcm.size, None, C.byref(n_len)))
        out = np.empty((self.n_mel, n_len.value), np.float32)
        src = pcm if pcm.size else np.zeros(1, np.float32)
        self._ck(self.lib.nb200_pcm_to_mel(self.h, _f32p(src), pcm.size, _f32p(out), C.byref(n_len)))
        return out

    def pcm_to_mel_batch(self, pcm: np.ndarray, lens=None, want_output=True) -> Optional[np.ndarray]:
        pcm = np.ascontiguousarray(pcm, np.float32)
        nw, stride = pcm.shape
        out = np.empty((nw, self.n_mel, 3000), np.float32) if want_output else None
        lp = None
        if lens is not None:
            lp = (C.c_size_t * nw)(*[int(x) for x in lens])
        self._ck(self.lib.nb200_pcm_to_mel_batch(self.h, _f32p(pcm), nw, stride, lp, _f32p(out)))
        return out

    # -- seam (2) ------------------------------------------------------------------------------------------
    def encoder_forward(self, mel: Optional[np.ndarray], n_windows: Optional[int] = None, want_output=True):
        if mel is not None:
            mel = np.ascontiguousarray(mel, np.float32)
            assert mel.shape[1:] == (self.n_mel, 3000), mel.shape
            n_windows = mel.shape[0]
        out = np.empty((n_windows, self.T, self.d), np.float32) if want_output else None
        self._ck(self.lib.nb200_encoder_forward(self.h, _f32p(mel), n_windows, _f32p(out)))
        return out

    def transcode_batch(self, pcm: np.ndarray, lens=None, out: Optional[np.ndarray] = None, want_output=True):
        pcm = np.ascontiguousarray(pcm, np.float32) if not (isinstance(pcm, np.ndarray) and pcm.flags.c_contiguous and pcm.dtype == np.float32) else pcm
        nw, stride = pcm.shape
        if out is None and want_output:
            out = np.empty((nw, self.T, self.d), np.float32)
        lp = None
        if lens is not None:
            lp = (C.c_size_t * nw)(*[int(x) for x in lens])
        self._ck(self.lib.nb200_transcode_batch(self.h, _f32p(pcm), nw, stride, lp, _f32p(out)))
        return out

    def transcode_submit(self, pcm: np.ndarray, out: np.ndarray, lens=None):
        """Pipelined form: `pcm` / `out` must be C-contiguous f32 (ideally pinned) and stay alive until collected."""
        assert pcm.dtype == np.float32 and pcm.flags.c_contiguous and out.dtype == np.float32 and out.flags.c_contiguous
        nw, stride = pcm.shape
        lp = None
        if lens is not None:
            lp = (C.c_size_t * nw)(*[int(x) for x in lens])
        self._ck(self.lib.nb200_transcode_submit(self.h, _f32p(pcm), nw, stride, lp, _f32p(out)))

    def transcode_collect(self):
        self._ck(self.lib.nb200_transcode_collect(self.h))

    def stage_pcm(self, pcm: np.ndarray, lens=None):
        pcm = np.ascontiguousarray(pcm, np.float32)
        nw, stride = pcm.shape
        lp = None
        if lens is not None:
            lp = (C.c_size_t * nw)(*[int(x) for x in lens])
        self._ck(self.lib.nb200_stage_pcm(self.h, _f32p(pcm), nw, stride, lp))

    def run_resident(self, n_windows: int, do_mel=True, do_encoder=True):
        self._ck(self.lib.nb200_run_resident(self.h, n_windows, int(do_mel), int(do_encoder)))

    def fetch_features(self, window: int, n: Optional[int] = None) -> np.ndarray:
        n = self.T * self.d if n is None else n
        out = np.empty(n, np.float32)
        self._ck(self.lib.nb200_fetch_features(self.h, window, _f32p(out), n))
        return out.reshape(self.T, self.d) if n == self.T * self.d else out

    def fetch_mel(self, window: int) -> np.ndarray:
        out = np.empty((self.n_mel, 3000), np.float32)
        self._ck(self.lib.nb200_fetch_mel(self.h, window, _f32p(out), out.size))
        return out

    # -- streaming (window 0) ----------------------------------------------------------------------------------
    def stream_reset(self):
        self._ck(self.lib.nb200_stream_reset(self.h))

    def stream_push(self, chunk: np.ndarray):
        chunk = np.ascontiguousarray(chunk, np.float32)
        if chunk.size:
            self._ck(self.lib.nb200_stream_push(self.h, _f32p(chunk), chunk.size))

    def stream_drain(self, n: int):
        self._ck(self.lib.nb200_stream_drain(self.h, n))

    def stream_features(self, run_encoder=True, want_mel=False, want_features=False):
        mel = np.empty((self.n_mel, 3000), np.float32) if want_mel else None
        feat = np.empty((self.T, self.d), np.float32) if want_features else None
        self._ck(self.lib.nb200_stream_features(self.h, int(run_encoder), _f32p(mel), _f32p(feat)))
        return mel, feat

    # -- seams (3)-(5) ---------------------------------------------------------------------------------------
    def decoder_forward(self, tokens: Sequence[int], flush: bool, window: int = 0) -> np.ndarray:
        t = np.ascontiguousarray(np.asarray(list(tokens), np.uint32))
        out = np.empty((t.size, self.d), np.float32)
        self._ck(self.lib.nb200_decoder_forward(self.h, window, t.ctypes.data_as(C.POINTER(C.c_uint32)), t.size, int(flush), _f32p(out)))
        return out

    def final_linear(self, hidden: np.ndarray) -> np.ndarray:
        hdn = np.ascontiguousarray(hidden, np.float32).reshape(-1)
        assert hdn.size == self.d
        out = np.empty(self.V, np.float32)
        self._ck(self.lib.nb200_final_linear(self.h, _f32p(hdn), _f32p(out)))
        return out

    def reset_kv_cache(self):
        self._ck(self.lib.nb200_reset_kv_cache(self.h))

    def detect_language(self, lang_tokens: Sequence[int], window: int = 0):
        """`Model::detect_language` (model.rs:194-210): returns (token id, probabilities over lang_tokens)."""
        t = np.ascontiguousarray(np.asarray(list(lang_tokens), np.uint32))
        probs = np.empty(t.size, np.float32)
        tok = C.c_uint32()
        self._ck(self.lib.nb200_detect_language(self.h, window, t.ctypes.data_as(C.POINTER(C.c_uint32)), t.size, C.byref(tok), _f32p(probs)))
        return tok.value, probs

    def set_decode_mode(self, separate: bool):
        """False: fused cooperative step kernel when supported (default); True: per-operation kernels replayed as a CUDA graph."""
        self._ck(self.lib.nb200_set_decode_mode(self.h, int(bool(separate))))

    def decode(self, n_windows: int = 1, temperature: float = 0.0, seed: int = 0, max_new_tokens: int = 0):
        toks = np.zeros((n_windows, self.P), np.uint32)
        n = (C.c_size_t * n_windows)()
        alp = (C.c_double * n_windows)()
        nsp = (C.c_double * n_windows)()
        self._ck(self.lib.nb200_decode(self.h, n_windows, temperature, seed, max_new_tokens, toks.ctypes.data_as(C.POINTER(C.c_uint32)), n, alp, nsp))
        return [dict(tokens=toks[b, : n[b]].tolist(), avg_logprob=alp[b], no_speech_prob=nsp[b]) for b in range(n_windows)]

    # the loop of model.rs:317-371 opened up: begin / advance / peek / end
    def set_audio_features(self, xa: np.ndarray):
        """`audio_features` handed in from the host ([n_windows, 1500, d] f32) instead of produced by the encoder."""
        xa = np.ascontiguousarray(xa, np.float32)
        assert xa.ndim == 3 and xa.shape[1:] == (self.T, self.d), xa.shape
        self._ck(self.lib.nb200_set_audio_features(self.h, _f32p(xa), xa.shape[0]))

    def decode_begin(self, n_windows: int = 1, temperature: float = 0.0, seed: int = 0, max_new_tokens: int = 0):
        self._run_windows = n_windows
        self._ck(self.lib.nb200_decode_begin(self.h, n_windows, temperature, seed, max_new_tokens))

    def decode_advance(self, n_steps: int = 1) -> bool:
        """-> True when every window had already finished (nothing was launched)."""
        done = C.c_int()
        self._ck(self.lib.nb200_decode_advance(self.h, n_steps, C.byref(done)))
        return bool(done.value)

    def decode_peek_logits(self, window: int = 0) -> np.ndarray:
        out = np.empty(self.V, np.float32)
        self._ck(self.lib.nb200_decode_peek_logits(self.h, window, _f32p(out)))
        return out

    def decode_end(self):
        nw = self._run_windows
        toks = np.zeros((nw, self.P), np.uint32)
        n = (C.c_size_t * nw)()
        alp = (C.c_double * nw)()
        nsp = (C.c_double * nw)()
        self._ck(self.lib.nb200_decode_end(self.h, toks.ctypes.data_as(C.POINTER(C.c_uint32)), n, alp, nsp))
        return [dict(tokens=toks[b, : n[b]].tolist(), avg_logprob=alp[b], no_speech_prob=nsp[b]) for b in range(nw)]

    def decode_greedy(self, n_windows: int = 1, max_new_tokens: int = 0):
        toks = np.zeros((n_windows, self.P), np.uint32)
        n = (C.c_size_t * n_windows)()
        alp = (C.c_double * n_windows)()
        nsp = (C.c_double * n_windows)()
        self._ck(self.lib.nb200_decode_greedy(self.h, n_windows, max_new_tokens, toks.ctypes.data_as(C.POINTER(C.c_uint32)), n, alp, nsp))
        return [dict(tokens=toks[b, : n[b]].tolist(), avg_logprob=alp[b], no_speech_prob=nsp[b]) for b in range(n_windows)]

    # -- measurement -----------------------------------------------------------------------------------------
    def timer_start(self):
        self._ck(self.lib.nb200_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.nb200_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile_enable(self, on: bool):
        self._ck(self.lib.nb200_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self._ck(self.lib.nb200_profile_reset(self.h))

    def profile_read(self):
        ms = (C.c_float * len(KERNEL_CLASSES))()
        ln = (C.c_int64 * len(KERNEL_CLASSES))()
        fl = C.c_double()
        self._ck(self.lib.nb200_profile_read(self.h, ms, ln, C.byref(fl)))
        return {k: dict(ms=ms[i], launches=ln[i]) for i, k in enumerate(KERNEL_CLASSES)}, fl.value

    def flush_l2(self):
        self._ck(self.lib.nb200_flush_l2(self.h))

    # -- kernel self-tests -----------------------------------------------------------------------------------
    def test_gemm(self, a: np.ndarray, w: np.ndarray, bias: Optional[np.ndarray] = None, gelu=False, out_bf16=False) -> np.ndarray:
        M, K = a.shape
        N = w.shape[0]
        if self.compute == "bf16":
            a_, w_ = f32_to_bf16_bits(a), f32_to_bf16_bits(w)
        else:
            a_, w_ = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(w, np.float32)
        b_ = None if bias is None else np.ascontiguousarray(bias, np.float32)
        out = np.empty((M, N), np.float32)
        self._ck(self.lib.nb200_test_gemm(self.h, a_.ctypes.data_as(C.c_void_p), w_.ctypes.data_as(C.c_void_p), _f32p(b_), M, N, K, int(gelu) | (2 if out_bf16 else 0), _f32p(out)))
        return out

    def test_gemm_perf(self, M: int, N: int, K: int, epi_kind: int, iters: int = 20) -> float:
        ms = C.c_float()
        self._ck(self.lib.nb200_test_gemm_perf(self.h, M, N, K, epi_kind, iters, C.byref(ms)))
        return ms.value

    def test_attention_perf(self, qkv: np.ndarray, B: int, T: int, n_heads: int, iters: int = 20) -> float:
        qkv = np.ascontiguousarray(qkv, np.float32)
        ms = C.c_float()
        self._ck(self.lib.nb200_test_attention_perf(self.h, _f32p(qkv), B, T, n_heads, iters, C.byref(ms)))
        return ms.value

    def test_attention(self, qkv: np.ndarray, B: int, T: int, n_heads: int) -> np.ndarray:
        qkv = np.ascontiguousarray(qkv, np.float32)
        out = np.empty((B * T, n_heads * 64), np.float32)
        self._ck(self.lib.nb200_test_attention(self.h, _f32p(qkv), B, T, n_heads, _f32p(out)))
        return out
