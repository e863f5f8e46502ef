"""The C++ host mirror of norma's `Model::transcribe` / `decode_with_fallback` / `inclusive_boxed_by`
(norma_b200/csrc/host/whisper_host.cc, reached through the C ABI) against the Python restatement of the same
reference lines (oracle/norma_host_oracle.py), over a scripted backend: no GPU involved."""
import random

import numpy as np
import pytest

from norma_b200 import synth, whisper
from oracle.norma_host_oracle import HostModelOracle, ReferenceWouldHang, inclusive_boxed_by

TOK = synth.special_tokens(51864)
NTS, EOT, SOT, LANG, TASK = TOK["no_timestamps"], TOK["eot"], TOK["sot"], TOK["lang"], TOK["task"]
TS = lambda sec: NTS + 1 + int(round(sec / 0.02))  # <|t|> token
PROMPT = [SOT, LANG, TASK]


def run_both(script, chunks):
    """script: list of (tokens, avg_logprob, no_speech_prob); chunks: list of (n_samples, final_chunk)."""
    m = whisper.Model(None, TOK, 400_000, vocab={i: bytes([97 + i % 26]) for i in range(1000)})
    for s in script:
        m.script_push(*s)
    it = iter(script)
    enc, temps, resets = [], [], []
    orc = HostModelOracle(lambda sl: enc.append(len(sl)), lambda t: (temps.append(t), next(it))[1], lambda: resets.append(1), NTS, EOT,
                          detok=lambda toks: "".join(chr(97 + t % 26) for t in toks if t < EOT), skip_no_progress=True)
    outs = []
    for n, final in chunks:
        data = np.zeros(n, np.float32)
        got = m.transcribe(data, final)
        ref = orc.transcribe([0.0] * n, final)
        assert got[1] == ref[1]
        assert got[0] == ref[0]
        assert m.state()["buffered"] == len(orc.buf)
        assert m.state()["n_no_progress"] == orc.n_no_progress
        outs.append(got)
    st = m.state()
    assert st["n_encodes"] == len(enc) and st["n_decodes"] == len(temps) and st["n_resets"] == len(resets)
    for i, (e, t) in enumerate(zip(enc, temps[: len(enc)])):
        assert m.script_log(i)[0] == e
    for i, t in enumerate(temps):
        assert abs(m.script_log(i)[1] - t) < 1e-12
    m.close()
    return outs


def test_inclusive_boxed_by_matches_reference_semantics():
    pred = lambda t: t > NTS or t == EOT
    v = PROMPT + [TS(0), 5, 6, TS(2), TS(2), 7, TS(4), EOT]
    assert inclusive_boxed_by(v, pred) == [[TS(0), 5, 6, TS(2)], [TS(2), 7, TS(4)]]  # the trailing eot alone is not yielded
    assert inclusive_boxed_by(PROMPT + [TS(0), 5, EOT], pred) == [[TS(0), 5, EOT]]
    assert inclusive_boxed_by(PROMPT + [5, 6, EOT], pred) == []
    assert inclusive_boxed_by([], pred) == []


def test_full_window_emits_closed_segments_and_keeps_the_open_tail():
    # the tail segment [<|3.00|>, 3, eot] is unfinished: the reference drains 3.00 s (150 * 320 samples) and waits
    script = [(PROMPT + [TS(0), 1, 2, TS(3), TS(3), 3, EOT], -0.3, 0.01)]
    (text, segs), = run_both(script, [(480_000, False)])
    assert segs == [[TS(0), 1, 2, TS(3)]]
    assert text == "bc"


def test_single_segment_window_is_consumed_whole():
    script = [(PROMPT + [TS(0), 1, 2, 3, EOT], -0.3, 0.01)]
    (text, segs), = run_both(script, [(480_000, False)])
    assert segs == [[TS(0), 1, 2, 3, EOT]] and text == "bcd"


def test_partial_chunk_waits_for_more_data_then_seeks():
    # 10 s chunk: the unfinished tail segment starts at 4.00 s -> the reference keeps the audio from 4.00 s on
    script = [(PROMPT + [TS(0), 1, TS(4), TS(4), 2, EOT], -0.2, 0.0),
              (PROMPT + [TS(0), 9, EOT], -0.2, 0.0)]
    outs = run_both(script, [(160_000, False), (0, True)])
    assert outs[0][1] == [[TS(0), 1, TS(4)]]
    assert outs[1][1] == [[TS(0), 9, EOT]]


def test_temperature_fallback_and_none():
    bad = (PROMPT + [TS(0), 1, EOT], -1.5, 0.1)  # avg_logprob < -1 -> needs fallback
    good = (PROMPT + [TS(0), 2, EOT], -0.5, 0.1)
    outs = run_both([bad, bad, good], [(480_000, False)])
    assert outs[0][1] == [[TS(0), 2, EOT]]
    # all six temperatures fail -> None -> the slice is dropped without output
    outs = run_both([bad] * 6, [(480_000, False)])
    assert outs[0] == ("", [])


def test_no_speech_gates():
    # no_speech > 0.6 stops the fallback loop (model.rs:179) and, with avg_logprob < -1, drops the window (model.rs:95)
    silent = (PROMPT, 0.0, 0.9)  # decode()'s early return: prompt only, avg_logprob 0
    quiet = (PROMPT + [TS(0), 1, EOT], -2.0, 0.7)
    assert run_both([quiet], [(480_000, False)])[0] == ("", [])
    # the silent window passes the gate of model.rs:95 (0 is not < -1) and has no segment: the reference never drains it and loops forever
    it = iter([silent])
    ref = HostModelOracle(lambda sl: None, lambda t: next(it), lambda: None, NTS, EOT)
    with pytest.raises(ReferenceWouldHang):
        ref.transcribe([0.0] * 480_000, False)
    # the documented divergence: the window is dropped (like the no-speech skip) and the stream keeps going
    outs = run_both([silent, (PROMPT + [TS(0), 4, EOT], -0.2, 0.0)], [(480_000, False), (480_000, False)])
    assert outs == [("", []), ("e", [[TS(0), 4, EOT]])]


def test_result_larger_than_the_callers_buffers_is_kept_not_truncated():
    import ctypes as C

    m = whisper.Model(None, TOK, 400_000, vocab={i: b"\xc3\xa9" for i in range(1000)})  # every token is a 2-byte UTF-8 character
    toks = PROMPT + [TS(0)] + [5] * 40 + [EOT]
    m.script_push(toks, -0.2, 0.0)
    d = np.zeros(480_000, np.float32)
    text = C.create_string_buffer(16)
    tl, sl = C.c_size_t(), C.c_size_t()
    seg = np.zeros(8, np.uint32)
    st = m.lib.nb200_model_transcribe(m.h, d.ctypes.data_as(C.POINTER(C.c_float)), d.size, 0, text, len(text), C.byref(tl),
                                      seg.ctypes.data_as(C.POINTER(C.c_uint32)), seg.size, C.byref(sl))
    assert st == 10 and tl.value == 80 and sl.value == 2 + 42 and text.value == b"" and not seg.any()  # nothing partial
    assert m.state()["buffered"] == 0                                                                   # the audio WAS consumed
    text = C.create_string_buffer(tl.value + 1)
    seg = np.zeros(sl.value, np.uint32)
    st = m.lib.nb200_model_last_result(m.h, text, len(text), C.byref(tl), seg.ctypes.data_as(C.POINTER(C.c_uint32)), seg.size, C.byref(sl))
    assert st == 0 and text.value.decode() == "\u00e9" * 40 and seg[0] == 1 and seg[1] == 42
    m.close()
    # the Python mirror grows its buffers by itself
    m = whisper.Model(None, TOK, 400_000, vocab={i: b"x" * 1000 for i in range(1000)})
    m.script_push(PROMPT + [TS(0)] + [5] * 100 + [EOT], -0.2, 0.0)
    text2, segs = m.transcribe(d, False)
    assert len(text2) == 100_000 and len(segs[0]) == 102
    m.close()


def test_long_buffer_is_processed_in_30s_slices():
    a = (PROMPT + [TS(0), 1, TS(29.98), TS(29.98), 2, EOT], -0.1, 0.0)  # tail starts at 29.98 s: seek there
    b = (PROMPT + [TS(0), 3, EOT], -0.1, 0.0)
    outs = run_both([a, b, b], [(480_000 + 200_000, True)])
    assert [s[:1] for s in outs[0][1]] == [[TS(0)], [TS(0)]] or len(outs[0][1]) >= 2


@pytest.mark.parametrize("seed", range(40))
def test_random_scripts_match_reference_restatement(seed):
    rng = random.Random(seed)

    def result():
        toks = list(PROMPT)
        t = 0.0
        toks.append(TS(t))
        for _ in range(rng.randint(0, 4)):
            for _ in range(rng.randint(0, 3)):
                toks.append(rng.randint(0, 999))
            t += rng.choice([0.0, 0.5, 2.0, 7.5])
            toks.append(TS(min(t, 29.98)))
            if rng.random() < 0.7:
                toks.append(TS(min(t, 29.98)))
        for _ in range(rng.randint(0, 2)):
            toks.append(rng.randint(0, 999))
        toks.append(EOT)
        while len(toks) >= 2 and toks[-2] > NTS:  # decode() strips trailing timestamps (model.rs:375-381)
            del toks[-2]
        return toks, rng.choice([-0.2, -0.9, -1.2, -3.0]), rng.choice([0.0, 0.3, 0.7])

    # every pass over a slice either drains samples or ends the call, so the script below cannot run out: 4 chunks of at most 700 000
    # samples, at least 0.5 s (8 000 samples) drained per decoding result that seeks, and no exception is tolerated
    script = [result() for _ in range(2000)]
    chunks = [(rng.choice([16_000, 160_000, 400_000, 480_000, 700_000]), rng.random() < 0.3) for _ in range(4)]
    outs = run_both(script, chunks)
    assert len(outs) == 4


def test_public_api_mirror():
    d = whisper.Definition.new(whisper.ModelType.default(), whisper.SelectedDevice.Cuda(0))
    assert whisper.ModelType.default() is whisper.ModelType.DistilLargeEnV3
    assert d.common_params.max_chunk_len() == 400_000 and d.common_params.data_buffer_size() == 5  # 16 kHz * 25 s; +2 quirk
    d.set_responsiveness(10_000)
    assert d.common_params.max_chunk_len() == 160_000
    for bad in (999, 30_001):
        with pytest.raises(whisper.Respnsivness):
            d.set_responsiveness(bad)
    d.common_params.set_max_chunk_len(3)
    assert d.common_params.max_chunk_len() == 100  # MIN_CHUNK_LEN
    assert whisper.ModelType.TinyEn.rev() == "refs/pr/15" and whisper.ModelType.DistilLargeEnV3.vocab_version() == "V2"
    with pytest.raises(whisper.WhisperError):
        whisper.Definition.new(whisper.ModelType.TinyEn, whisper.SelectedDevice.Cpu()).blocking_try_to_model({})


# ---- LanguageState (model.rs:392-440; multilingual.rs:319-322) --------------------------------------------------------
MTOK = synth.special_tokens(51865, lang=None)
LANGS = synth.language_tokens(51865)


def _multilingual_pair(script, languages):
    m = whisper.Model(None, MTOK, 400_000)
    m.set_language_detection(LANGS)
    for s in script:
        m.script_push(*s)
    for lang in languages:
        m.script_push_language(lang)
    it, lit = iter(script), iter(languages)
    orc = HostModelOracle(lambda sl: None, lambda t: next(it), lambda: None, MTOK["no_timestamps"], MTOK["eot"], detect_language=lambda: next(lit))
    return m, orc


def test_language_is_detected_once_per_transcription_and_cleared_on_final_chunk():
    nts, eot = MTOK["no_timestamps"], MTOK["eot"]
    ts = lambda sec: nts + 1 + int(round(sec / 0.02))
    whole = ([MTOK["sot"], LANGS[2], MTOK["task"], ts(0), 1, 2, eot], -0.2, 0.01)
    m, orc = _multilingual_pair([whole] * 4, [LANGS[2], LANGS[6]])
    assert m.language() == (None, 0)
    for k, (n, final, want_lang, want_detects) in enumerate([(480_000, False, LANGS[2], 1), (480_000, False, LANGS[2], 1),
                                                              (480_000, True, None, 1), (480_000, False, LANGS[6], 2)]):
        got = m.transcribe(np.zeros(n, np.float32), final)
        ref = orc.transcribe([0.0] * n, final)
        assert got[1] == ref[1]
        assert m.language() == (want_lang, want_detects), k
        assert orc.language_token == want_lang
    m.close()


def test_const_lang_models_never_detect():
    m = whisper.Model(None, TOK, 400_000)
    m.script_push(PROMPT + [TS(0), 1, EOT], -0.2, 0.01)
    m.transcribe(np.zeros(480_000, np.float32), True)
    assert m.language() == (LANG, 0)  # LanguageState::ConstLang: clear() and set_language_token() are no-ops
    m.close()


def test_detection_failure_is_reported_not_swallowed():
    m = whisper.Model(None, MTOK, 400_000)
    m.set_language_detection(LANGS)
    m.script_push([MTOK["sot"], LANGS[0], MTOK["task"], MTOK["no_timestamps"] + 1, 1, MTOK["eot"]], -0.2, 0.01)
    with pytest.raises(Exception) as e:  # the scripted backend has no language queued
        m.transcribe(np.zeros(16_000, np.float32), False)
    assert "ran out of detected languages" in str(e.value)
    m.script_push_language(12345)        # not one of the language tokens
    with pytest.raises(Exception) as e:
        m.transcribe(np.zeros(0, np.float32), False)
    assert "not one of the language tokens" in str(e.value)
    m.close()


def test_public_api_mirror_of_multilingual_definitions():
    d = whisper.MultilingualDefinition.new(whisper.MultilingualModelType.LargeV3, whisper.SelectedDevice.Cuda(0), whisper.Task.Translate)
    assert d.detect and d.task is whisper.Task.Translate and d.model.id() == "openai/whisper-large-v3" and d.model.vocab_version() == "V2"
    assert whisper.MultilingualModelType.default() is whisper.MultilingualModelType.Medium
    assert whisper.MultilingualModelType.Base.rev() == "refs/pr/22" and whisper.MultilingualModelType.QuantizedTiny.quantized_ext() == "tiny"
    mono = whisper.ModelType.MultiAsMono(whisper.MultilingualModelType.Small, whisper.Language.from_code("de"))
    dd = whisper.Definition.new(mono, whisper.SelectedDevice.Cuda(1))
    assert not dd.detect and dd.model.language().token() == "<|de|>" and dd.model.id() == "openai/whisper-small"
    assert [l.token() for l in whisper.Language.iter()][:3] == ["<|en|>", "<|zh|>", "<|de|>"] and len(whisper.Language.iter()) == 99
    assert whisper.ModelType.TinyEn.language() == whisper.Language.English
    with pytest.raises(whisper.Respnsivness):
        d.set_responsiveness(500)
