"""oracle/norma_host_oracle.py — TEST INFRASTRUCTURE ONLY.

Python restatement of the host logic of norma's whisper `Model` (buffering, 30 s slicing, temperature fallback,
timestamp segmentation, seek), statement by statement from
  /root/reference/src/models/whisper/model.rs:55-160   Model::transcribe
  /root/reference/src/models/whisper/model.rs:164-191  Model::decode_with_fallback
  /root/reference/src/models/whisper/model.rs:392-440  LanguageState (Detect / ConstLang)
  /root/reference/src/utils.rs:1-76                     SliceExt::inclusive_boxed_by
`encode` / `decode` are callbacks so the same restatement runs over scripted results (CPU tests of the C++ mirror) or
over the model oracle.  PARITY UNPINNED: the reference's only tests of this code are two `#[ignore]`d microphone tests
against a mock model (tests/transcriber.rs), which pin no token sequence."""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Tuple

N_SAMPLES = 480_000
NO_SPEECH_THRESHOLD = 0.6
LOGPROB_THRESHOLD = -1.0
COMPRESSION_RATIO_THRESHOLD = 2.4
TEMPERATURES = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0)


def inclusive_boxed_by(v: List[int], pred: Callable[[int], bool]):
    """utils.rs:30-54"""
    out = []
    while True:
        s_idx = next((i for i, x in enumerate(v) if pred(x)), None)
        if s_idx is None:
            return out
        e_rel = next((i for i, x in enumerate(v[s_idx + 1:]) if pred(x)), None)
        if e_rel is None:
            return out
        e_idx = s_idx + e_rel + 2
        out.append(v[s_idx:e_idx])
        v = v[e_idx:]


class ReferenceWouldHang(RuntimeError):
    """model.rs:68-150 on a decoding result without a drainable segment: nothing is drained, the same slice is decoded forever."""


class HostModelOracle:
    """`skip_no_progress=False` is the reference as written (the endless loop is raised as ReferenceWouldHang).  True is the ONE documented
    divergence of the product's host mirror (include/norma_b200.h, nb200_model_transcribe): such a window is dropped like the no-speech
    skip of model.rs:95-98 and counted in `n_no_progress`."""

    def __init__(self, encode, decode, reset_kv_cache, no_timestamps: int, eot: int, detok=None, detect_language=None, const_lang=None,
                 skip_no_progress: bool = False):
        self.skip_no_progress, self.n_no_progress = skip_no_progress, 0
        self.encode, self.decode, self.reset_kv_cache = encode, decode, reset_kv_cache
        self.nts, self.eot = no_timestamps, eot
        self.detok = detok or (lambda toks: "")
        self.buf: List[float] = []
        # LanguageState: Detect { language_token: None, .. } when a detector is given (multilingual.rs:319-322), else ConstLang
        self.detect_language = detect_language
        self.language_token: Optional[int] = None if detect_language else const_lang

    def decode_with_fallback(self):
        if self.detect_language is not None and self.language_token is None:     # model.rs:170 `self.lang.is_none()`
            self.language_token = self.detect_language()                         # model.rs:171-172
        for t in TEMPERATURES:                                                   # model.rs:175
            dr = self.decode(t)                                                  # (tokens, avg_logprob, no_speech_prob)
            compression_ratio = float("nan")
            needs_fallback = compression_ratio > COMPRESSION_RATIO_THRESHOLD or dr[1] < LOGPROB_THRESHOLD
            if not needs_fallback or dr[2] > NO_SPEECH_THRESHOLD:                # model.rs:179
                return dr
        return None

    def transcribe(self, data: List[float], final_chunk: bool) -> Tuple[str, List[List[int]]]:
        self.buf.extend(data)                                                    # model.rs:60-64
        res, segs = "", []
        new_chunk = False
        while self.buf and not new_chunk:                                        # model.rs:68
            len_before = len(self.buf)
            slice_len = min(len(self.buf), N_SAMPLES)
            self.encode(self.buf[:slice_len])                                    # model.rs:74-88, 168
            dr = self.decode_with_fallback()
            if dr is None:                                                       # model.rs:90-93
                del self.buf[:slice_len]
                continue
            tokens, avg_logprob, no_speech_prob = dr
            if no_speech_prob > NO_SPEECH_THRESHOLD and avg_logprob < LOGPROB_THRESHOLD:  # model.rs:95-98
                del self.buf[:slice_len]
                continue
            for seg in inclusive_boxed_by(list(tokens), lambda t: t > self.nts or t == self.eot):
                s_timestamp = (seg[0] - self.nts - 1) & 0xFFFFFFFF               # u32 arithmetic, model.rs:103
                if seg[-1] == self.eot:                                          # model.rs:107
                    if s_timestamp == 0 or final_chunk:
                        if slice_len == N_SAMPLES or final_chunk:
                            del self.buf[:slice_len]                             # model.rs:110
                        else:
                            new_chunk = True                                     # model.rs:122
                            break
                    else:
                        pre = len(self.buf)
                        del self.buf[: min(s_timestamp * 320, slice_len)]        # model.rs:126-127
                        if pre > slice_len:
                            break                                                # model.rs:135
                        new_chunk = True                                         # model.rs:143
                        break
                segs.append(seg)
                res += self.detok(seg[1:-1])                                     # model.rs:147-149
            if not new_chunk and len(self.buf) == len_before:                    # nothing drained and no break: the reference loops forever
                if not self.skip_no_progress:
                    raise ReferenceWouldHang("decoding result made no progress (no timestamp-delimited segment)")
                self.n_no_progress += 1
                del self.buf[:slice_len]
        if final_chunk:
            if self.detect_language is not None:
                self.language_token = None                                       # model.rs:154 `self.lang.clear()`
            self.reset_kv_cache()                                                # model.rs:153-156
        return res, segs
