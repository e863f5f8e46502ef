"""Host-side mirror of norma's public API for this path, over the C ABI (names as in the reference):

  SelectedDevice                      /root/reference/src/models/mod.rs:38-55
  CommonModelParams                   /root/reference/src/models/mod.rs:57-117
  monolingual.ModelType / Definition  /root/reference/src/models/whisper/monolingual.rs:32-174
  multilingual.ModelType / Task / Definition
                                      /root/reference/src/models/whisper/multilingual.rs:16-175
  Language                            /root/reference/src/models/whisper/languages.rs:7-222
  Model.transcribe(data, final_chunk) /root/reference/src/models/whisper/model.rs:55-160 (implemented in C++,
                                      norma_b200/csrc/host/whisper_host.cc, reached through nb200_model_*)

What is NOT mirrored (out of scope): the hf-hub download, the Transcriber thread and cpal capture.
`Definition.blocking_try_to_model_from_files` is `blocking_try_to_model` from the point where the three checkpoint
files are on disk (config.json, tokenizer.json, model.safetensors; parsed by norma_b200/csrc/host/loader.cc);
`Definition.blocking_try_to_model` takes weights / vocabulary from the caller instead (synthetic checkpoints).
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import ffi, filters, synth

SAMPLE_RATE = 16_000
MIN_CHUNK_LEN = 100       # models/mod.rs:59
MIN_STRING_BUF_SIZE = 1   # models/mod.rs:63


class WhisperError(Exception):
    """whisper::Error (/root/reference/src/models/whisper/mod.rs:65-84)"""


class Respnsivness(WhisperError):  # sic: the reference's spelling
    def __init__(self):
        super().__init__("The respnsivness must be over 1 second and under 30")


@dataclass(frozen=True)
class SelectedDevice:
    kind: str = "cpu"
    ordinal: int = 0

    @staticmethod
    def Cpu():
        return SelectedDevice("cpu")

    @staticmethod
    def Cuda(n: int):
        return SelectedDevice("cuda", n)

    @staticmethod
    def Metal():
        return SelectedDevice("metal")


class CommonModelParams:
    def __init__(self, max_chunk_len: int, data_buffer_size: int, string_buffer_size: int):
        self._max_chunk_len = max(max_chunk_len, MIN_CHUNK_LEN)
        self._data_buffer_size = data_buffer_size + 2          # thingbuf's usable capacity is n - 2 (models/mod.rs:81)
        self._string_buffer_size = max(string_buffer_size, MIN_STRING_BUF_SIZE)

    def max_chunk_len(self) -> int:
        return max(self._max_chunk_len, MIN_CHUNK_LEN)

    def data_buffer_size(self) -> int:
        return self._data_buffer_size

    def string_buffer_size(self) -> int:
        return self._string_buffer_size

    def set_max_chunk_len(self, n: int):
        self._max_chunk_len = max(n, MIN_CHUNK_LEN)

    def set_data_buffer_size(self, n: int):
        self._data_buffer_size = n + 2

    def set_string_buffer_size(self, n: int):
        self._string_buffer_size = max(n, MIN_STRING_BUF_SIZE)


LANGUAGE_CODES = (
    "en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur hr bg lt la mi ml cy sk te fa lv bn "
    "sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si km sn yo so af oc ka be tg sd gu am yi lo uz fo ht ps tk nn mt sa lb my "
    "bo tl mg as tt haw ln ha ba jw su").split()


@dataclass(frozen=True)
class Language:
    """whisper::Language (languages.rs:7-107): the 99 variants in declaration order; `index` is the position."""
    index: int

    @staticmethod
    def from_code(code: str) -> "Language":
        return Language(LANGUAGE_CODES.index(code))

    @staticmethod
    def iter():
        return [Language(i) for i in range(len(LANGUAGE_CODES))]

    def code(self) -> str:
        return LANGUAGE_CODES[self.index]

    def token(self) -> str:  # languages.rs:120-222
        return f"<|{self.code()}|>"


Language.English = Language(0)


class Task(enum.Enum):
    """multilingual::Task (multilingual.rs:16-25)"""
    Transcribe = "transcribe"
    Translate = "translate"


class MultilingualModelType(enum.Enum):
    """multilingual::ModelType (multilingual.rs:47-116): (hub id, revision, architecture shape, vocab version)"""
    QuantizedTiny = ("lmz/candle-whisper", "main", "tiny", "V1")
    Tiny = ("openai/whisper-tiny", "main", "tiny", "V1")
    Base = ("openai/whisper-base", "refs/pr/22", "base", "V1")
    Small = ("openai/whisper-small", "main", "small", "V1")
    Medium = ("openai/whisper-medium", "main", "medium", "V1")
    Large = ("openai/whisper-large", "refs/pr/36", "large", "V1")
    LargeV2 = ("openai/whisper-large-v2", "refs/pr/57", "large-v2", "V1")
    LargeV3 = ("openai/whisper-large-v3", "main", "large-v3", "V2")

    @classmethod
    def default(cls):
        return cls.Medium  # #[default], multilingual.rs:53-54

    def id(self) -> str:
        return self.value[0]

    def rev(self) -> str:
        return self.value[1]

    def shape(self) -> str:
        return self.value[2]

    def vocab_version(self) -> str:
        return self.value[3]

    def quantized_ext(self) -> Optional[str]:
        return "tiny" if self is MultilingualModelType.QuantizedTiny else None


@dataclass(frozen=True)
class MultiAsMono:
    """monolingual::ModelType::MultiAsMono { model, lang } (monolingual.rs:42-45): a multilingual checkpoint pinned to one language."""
    model: MultilingualModelType
    lang: Language

    def id(self) -> str:
        return self.model.id()

    def rev(self) -> str:
        return self.model.rev()

    def shape(self) -> str:
        return self.model.shape()

    def vocab_version(self) -> str:
        return self.model.vocab_version()

    def quantized_ext(self) -> Optional[str]:
        return self.model.quantized_ext()

    def language(self) -> Language:
        return self.lang


class ModelType(enum.Enum):
    """monolingual::ModelType (monolingual.rs:32-111): (hub id, revision, architecture shape, vocab version)"""
    QuantizedTinyEn = ("lmz/candle-whisper", "main", "tiny.en", "EnV1")
    TinyEn = ("openai/whisper-tiny.en", "refs/pr/15", "tiny.en", "EnV1")
    BaseEn = ("openai/whisper-base.en", "refs/pr/13", "base.en", "EnV1")
    SmallEn = ("openai/whisper-small.en", "refs/pr/10", "small.en", "EnV1")
    MediumEn = ("openai/whisper-medium.en", "main", "medium.en", "EnV1")
    DistilMediumEn = ("distil-whisper/distil-medium.en", "main", "distil-medium.en", "V1")
    DistilLargeEnV2 = ("distil-whisper/distil-large-v2", "main", "distil-large-v2", "V1")
    DistilLargeEnV3 = ("distil-whisper/distil-large-v3", "main", "distil-large-v3", "V2")

    @classmethod
    def default(cls):
        return cls.DistilLargeEnV3  # #[default], monolingual.rs:40-41

    def id(self) -> str:
        return self.value[0]

    def rev(self) -> str:
        return self.value[1]

    def shape(self) -> str:
        return self.value[2]

    def vocab_version(self) -> str:
        return self.value[3]

    def quantized_ext(self) -> Optional[str]:
        return "tiny-en" if self is ModelType.QuantizedTinyEn else None

    def language(self) -> Language:
        return Language.English  # monolingual.rs:85-96


ModelType.MultiAsMono = MultiAsMono


class Definition:
    """monolingual::Definition (monolingual.rs:116-174).  `task` / `detect` carry multilingual::Definition
    (multilingual.rs:108-175) on the same class: see `MultilingualDefinition`."""

    def __init__(self, model, device: SelectedDevice, task: Task = Task.Transcribe, detect: bool = False):
        self.model, self.device, self.task, self.detect = model, device, task, detect
        self.common_params = CommonModelParams(SAMPLE_RATE * 25, 3, 3)  # monolingual.rs:128, multilingual.rs:126

    @staticmethod
    def new(model, device: SelectedDevice) -> "Definition":
        return Definition(model, device)

    def set_responsiveness(self, period_ms: int):
        if 1_000 <= period_ms <= 30_000:  # monolingual.rs:147-156
            self.common_params.set_max_chunk_len((SAMPLE_RATE * period_ms) // 1000)
        else:
            raise Respnsivness()

    def set_data_buffer_size(self, size: int):
        self.common_params.set_data_buffer_size(size)

    def set_string_buffer_size(self, size: int):
        self.common_params.set_string_buffer_size(size)

    def blocking_try_to_model(self, weights: Dict[str, object], compute: str = "bf16", vocab: Optional[Dict[int, bytes]] = None,
                              suppress_tokens: Sequence[int] = (), seed: int = 0) -> "Model":
        """monolingual.rs:320-451 minus the hub download: the caller supplies what the reference reads from the hub."""
        if self.device.kind != "cuda":
            raise WhisperError("this build only implements SelectedDevice::Cuda(ord) (sm_100a, no CPU fallback)")
        if self.model.quantized_ext() is not None:
            raise WhisperError("quantized (q8_0) checkpoints are loaded from their GGUF file: use blocking_try_to_model_from_files")
        cfg = synth.model_config(self.model.shape())
        ctx = ffi.Context(cfg, ordinal=self.device.ordinal, compute=compute, max_batch=1)
        ctx.set_mel_filters(filters.mel_filters(cfg["num_mel_bins"]))  # Error::MelBins for anything but 80 | 128
        ctx.load_weights(weights)
        tok = synth.special_tokens(cfg["vocab_size"], task=self.task.value, lang=None if self.detect else self.model.language().index)
        ctx.set_tokens(**tok)
        ctx.set_suppress(suppress_tokens)
        m = Model(ctx, tok, self.common_params.max_chunk_len(), vocab, seed)
        if self.detect:  # LanguageState::Detect (multilingual.rs:319-322)
            m.set_language_detection(synth.language_tokens(cfg["vocab_size"]))
        return m

    def blocking_try_to_model_from_files(self, config_json: str, tokenizer_json: str, safetensors: str, compute: str = "bf16",
                                         seed: int = 0) -> "Model":
        """monolingual.rs:347-451 / multilingual.rs:225-323: everything `blocking_try_to_model` does once hf-hub has put the
        three files on disk (parsing, upload, token-id lookups, masks), done natively by nb200_model_from_files."""
        if self.device.kind != "cuda":
            raise WhisperError("this build only implements SelectedDevice::Cuda(ord) (sm_100a, no CPU fallback)")
        # `Quantized*` model types pass config-{ext}.json, tokenizer-{ext}.json and model-{ext}-q80.gguf (monolingual.rs:198-203): the GGUF
        # file is recognised by its magic and dequantised at load (csrc/host/loader.h)
        import os

        lib = ffi.load_library()
        ch, mh = C.c_void_p(), C.c_void_p()
        lang = None if self.detect else self.model.language().token().encode()
        st = lib.nb200_model_from_files(self.device.ordinal, os.fsencode(config_json), os.fsencode(tokenizer_json), os.fsencode(safetensors),
                                        ffi.DTYPES[compute], lang, ffi.TASKS[self.task.value], self.common_params.max_chunk_len(), seed,
                                        C.byref(ch), C.byref(mh))
        if st != 0:
            raise ffi.Nb200Error(st, (lib.nb200_last_error(None) or b"").decode())
        cfg, _ = ffi.config_from_file(config_json)
        ctx = ffi.Context(cfg, ordinal=self.device.ordinal, compute=compute, max_batch=1, handle=ch)
        return Model(ctx, None, 0, handle=mh)


class MultilingualDefinition(Definition):
    """multilingual::Definition::new(model, device, task) (multilingual.rs:108-127): language detected per transcription."""

    def __init__(self, model: MultilingualModelType, device: SelectedDevice, task: Task = Task.Transcribe):
        super().__init__(model, device, task, detect=True)

    @staticmethod
    def new(model: MultilingualModelType, device: SelectedDevice, task: Task = Task.Transcribe) -> "MultilingualDefinition":
        return MultilingualDefinition(model, device, task)


class Model:
    """whisper::Model (model.rs:16-160): `Data = f32`, `SAMPLE_RATE = 16_000`."""
    SAMPLE_RATE = SAMPLE_RATE

    def __init__(self, ctx: Optional[ffi.Context], tok: Optional[Dict[str, int]], max_chunk_len: int, vocab: Optional[Dict[int, bytes]] = None,
                 seed: int = 0, handle=None):
        self.lib = ffi.load_library()
        self.ctx = ctx
        self._owns_ctx = handle is not None
        if handle is not None:  # adopt an nb200_model made by nb200_model_from_files
            self.h = handle
            return
        t = ffi.SpecialTokens(tok["sot"], tok["eot"], tok["task"], 0xFFFFFFFF if tok.get("lang") is None else tok["lang"], tok["no_speech"],
                              tok["no_timestamps"], tok["ts_zero"], tok["ts_one"])
        h = C.c_void_p()
        st = self.lib.nb200_model_create(ctx.h if ctx is not None else None, C.byref(t), max_chunk_len, seed, C.byref(h))
        if st != 0:
            raise ffi.Nb200Error(st, "nb200_model_create failed")
        self.h = h
        for i, b in (vocab or {}).items():
            self.lib.nb200_model_set_vocab(self.h, i, b, len(b))

    def close(self):
        if getattr(self, "h", None):
            self.lib.nb200_model_destroy(self.h)
            self.h = None
        if getattr(self, "_owns_ctx", False) and self.ctx is not None:
            self.ctx.close()

    def set_tokenizer(self, tokenizer: "ffi.Tokenizer"):
        self.lib.nb200_model_set_tokenizer(self.h, tokenizer.h)

    def set_language_detection(self, lang_tokens: Sequence[int]):
        t = np.ascontiguousarray(np.asarray(list(lang_tokens), np.uint32))
        st = self.lib.nb200_model_set_language_detection(self.h, t.ctypes.data_as(C.POINTER(C.c_uint32)), t.size)
        if st != 0:
            raise ffi.Nb200Error(st, (self.lib.nb200_model_last_error(self.h) or b"").decode())

    def language(self) -> Tuple[Optional[int], int]:
        """-> (`LanguageState::language_token()`, number of detect_language calls so far)"""
        t, n = C.c_uint32(), C.c_size_t()
        self.lib.nb200_model_language(self.h, C.byref(t), C.byref(n))
        return (None if t.value == 0xFFFFFFFF else t.value), n.value

    def script_push_language(self, token: int):
        st = self.lib.nb200_model_script_push_language(self.h, token)
        if st != 0:
            raise ffi.Nb200Error(st, "script_push_language on a non-scripted model")

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transcribe(self, data: np.ndarray, final_chunk: bool) -> Tuple[str, List[List[int]]]:
        """-> (text, emitted token segments).  `data` is consumed, as the reference's `&mut Vec<f32>` is."""
        d = np.ascontiguousarray(data, np.float32)
        text = C.create_string_buffer(1 << 16)
        tl, sl = C.c_size_t(), C.c_size_t()
        seg = np.zeros(1 << 14, np.uint32)
        ptr = d.ctypes.data_as(C.POINTER(C.c_float)) if d.size else None
        st = self.lib.nb200_model_transcribe(self.h, ptr, d.size, int(final_chunk), text, len(text), C.byref(tl),
                                             seg.ctypes.data_as(C.POINTER(C.c_uint32)), seg.size, C.byref(sl))
        if st == 10:  # NB200_BUFFER_TOO_SMALL: the audio is consumed and the result kept; fetch it into buffers of the reported sizes
            text = C.create_string_buffer(tl.value + 1)
            seg = np.zeros(max(sl.value, 1), np.uint32)
            st = self.lib.nb200_model_last_result(self.h, text, len(text), C.byref(tl), seg.ctypes.data_as(C.POINTER(C.c_uint32)), seg.size, C.byref(sl))
        if st != 0:
            raise ffi.Nb200Error(st, (self.lib.nb200_model_last_error(self.h) or b"").decode())
        segs, i = [], 1
        for _ in range(int(seg[0])):
            n = int(seg[i])
            segs.append(seg[i + 1 : i + 1 + n].tolist())
            i += 1 + n
        return text.value.decode("utf-8", "replace"), segs

    def state(self) -> Dict[str, int]:
        a, b, c, d = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        self.lib.nb200_model_state(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        e = C.c_size_t()
        self.lib.nb200_model_no_progress_windows(self.h, C.byref(e))
        return dict(buffered=a.value, n_encodes=b.value, n_decodes=c.value, n_resets=d.value, n_no_progress=e.value)

    # scripted backend (ctx is None): canned results for the host-logic tests
    def script_push(self, tokens: Sequence[int], avg_logprob: float, no_speech_prob: float):
        t = np.ascontiguousarray(np.asarray(list(tokens), np.uint32))
        st = self.lib.nb200_model_script_push(self.h, avg_logprob, no_speech_prob, t.ctypes.data_as(C.POINTER(C.c_uint32)), t.size)
        if st != 0:
            raise ffi.Nb200Error(st, "script_push on a non-scripted model")

    def script_log(self, i: int) -> Tuple[int, float]:
        n, t = C.c_size_t(), C.c_double()
        self.lib.nb200_model_script_log(self.h, i, C.byref(n), C.byref(t))
        return (-1 if n.value == 2**64 - 1 else n.value), t.value
