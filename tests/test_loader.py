"""Checkpoint side files: config.json, tokenizer.json, model.safetensors and the mel filterbank, parsed natively
(norma_b200/csrc/host/loader.cc) the way `blocking_try_to_model` does (/root/reference/src/models/whisper/monolingual.rs:347-430).
CPU only: no compute calls."""
from __future__ import annotations

import json
import os
import struct

import numpy as np
import pytest
import torch

from norma_b200 import ffi, filters, synth

from conftest import GOLDEN, REFERENCE


@pytest.fixture(scope="module")
def small_tok(lib):
    return ffi.Tokenizer(os.path.join(GOLDEN, "tokenizer_small.json"))


@pytest.fixture(scope="module")
def decodes():
    with open(os.path.join(GOLDEN, "tokenizer_small_decodes.json"), encoding="utf-8") as f:
        return json.load(f)


# ---- config.json ---------------------------------------------------------------------------------------------------------
def _write(tmp_path, name, obj):
    p = tmp_path / name
    p.write_text(obj if isinstance(obj, str) else json.dumps(obj))
    return str(p)


def test_config_roundtrip_with_hf_extras(lib, tmp_path):
    c = synth.model_config("distil-large-v3")
    hf = dict(c, architectures=["WhisperForConditionalGeneration"], suppress_tokens=[1, 2, 7, 50359], begin_suppress_tokens=[220, 50257],
              dropout=0.0, scale_embedding=False, forced_decoder_ids=None, nested={"a": [1, {"b": "x\\u00e9\\ud83d\\ude42"}]})
    cfg, sup = ffi.config_from_file(_write(tmp_path, "config.json", hf))
    assert cfg == c
    assert sup == [1, 2, 7, 50359]


def test_config_suppress_defaults_to_empty(lib, tmp_path):
    cfg, sup = ffi.config_from_file(_write(tmp_path, "config.json", synth.model_config("tiny.en")))
    assert sup == [] and cfg["d_model"] == 384


@pytest.mark.parametrize("missing", ["num_mel_bins", "d_model", "decoder_layers", "vocab_size"])
def test_config_missing_field_is_a_parse_error(lib, tmp_path, missing):
    c = synth.model_config("tiny.en")
    del c[missing]
    with pytest.raises(ffi.Nb200Error) as e:
        ffi.config_from_file(_write(tmp_path, "config.json", c))
    assert e.value.status == 8 and f"missing field `{missing}`" in str(e.value)


@pytest.mark.parametrize("text", ["{", "[1,2", '{"d_model": 1.5}', '{"a": tru}', '{"a": "\\ud800"}', "", '{"a":1} x'])
def test_config_malformed_json(lib, tmp_path, text):
    with pytest.raises(ffi.Nb200Error) as e:
        ffi.config_from_file(_write(tmp_path, "config.json", text))
    assert e.value.status == 8


def test_config_wrong_type(lib, tmp_path):
    c = dict(synth.model_config("tiny.en"), d_model="384")
    with pytest.raises(ffi.Nb200Error) as e:
        ffi.config_from_file(_write(tmp_path, "config.json", c))
    assert e.value.status == 8 and "d_model" in str(e.value)


def test_config_missing_file_is_an_io_error(lib, tmp_path):
    with pytest.raises(ffi.Nb200Error) as e:
        ffi.config_from_file(str(tmp_path / "nope.json"))
    assert e.value.status == 7


# ---- mel filterbank ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_mel", [80, 128])
def test_mel_filters_match_python_and_reference_bytes(lib, n_mel):
    a = ffi.mel_filters(n_mel)
    assert np.array_equal(a, filters.mel_filters(n_mel))
    ref = os.path.join(REFERENCE, "src/models/whisper/whisper_mel_bytes", f"{n_mel}.bytes")
    if os.path.exists(ref):  # the table norma embeds (monolingual.rs:351-362); absent on the GPU box
        r = np.fromfile(ref, "<f4").reshape(n_mel, 201)
        assert np.abs(a - r).max() < 1e-8
        assert np.array_equal(a != 0, r != 0)


def test_mel_filters_other_sizes_are_melbins_errors(lib):
    for n in (0, 64, 81, 256):
        with pytest.raises(ffi.Nb200Error) as e:
            ffi.mel_filters(n)
        assert e.value.status == 5 and "Unexpected number of mel bins" in str(e.value)


# ---- tokenizer.json ------------------------------------------------------------------------------------------------------
def test_tokenizer_decode_matches_tokenizers_crate_golden(small_tok, decodes):
    """813 decodes produced by the `tokenizers` library itself (tests/golden/make_tokenizer_golden.py)."""
    for c in decodes["cases"]:
        assert small_tok.decode(c["ids"], c["skip"]) == c["text"], c


def test_tokenizer_token_to_id_matches_golden(small_tok, decodes):
    for tok, i in decodes["token_to_id"].items():
        if i is None:
            with pytest.raises(ffi.Nb200Error) as e:
                small_tok.token_to_id(tok)
            assert e.value.status == 9 and "Failed to get the id for token" in str(e.value)
        else:
            assert small_tok.token_to_id(tok) == i
    for tok, i in decodes["text_token_to_id"].items():
        assert small_tok.token_to_id(tok) == i


def test_tokenizer_live_against_tokenizers_package(lib, tmp_path):
    """Full-size Whisper layout, compared live with the Python binding of the reference's tokenizers crate."""
    tokenizers = pytest.importorskip("tokenizers")
    p = tmp_path / "tokenizer.json"
    p.write_text(synth.synth_tokenizer_json(51866), encoding="utf-8")
    mine, ref = ffi.Tokenizer(str(p)), tokenizers.Tokenizer.from_file(str(p))
    rng = np.random.default_rng(2)
    for it in range(600):
        n = int(rng.integers(0, 16))
        ids = (rng.integers(0, 51866 + 40, n) if it % 2 else rng.integers(0, 600, n)).tolist()
        for skip in (True, False):
            assert mine.decode(ids, skip) == ref.decode(ids, skip_special_tokens=skip)
    for name in ("<|startoftranscript|>", "<|endoftext|>", "<|nospeech|>", "<|notimestamps|>", "<|0.00|>", "<|1.00|>", "<|30.00|>", "<|su|>", "<|yue|>"):
        assert mine.token_to_id(name) == ref.token_to_id(name)


@pytest.mark.parametrize("vocab", [51864, 51865, 51866])
def test_special_token_lookup_matches_public_layouts(lib, tmp_path, vocab):
    """monolingual.rs:376-384,419-420: ids come from tokenizer.json; they must equal the public Whisper layouts."""
    p = tmp_path / "tokenizer.json"
    p.write_text(synth.synth_tokenizer_json(vocab), encoding="utf-8")
    t = ffi.Tokenizer(str(p))
    assert t.special_tokens("<|en|>") == synth.special_tokens(vocab)
    assert t.special_tokens("<|de|>", "translate") == synth.special_tokens(vocab, task="translate", lang=2)
    none = t.special_tokens(None)
    assert none["lang"] == 0xFFFFFFFF and none["sot"] == synth.special_tokens(vocab)["sot"]
    assert t.language_tokens() == synth.language_tokens(vocab)  # multilingual.rs:251-254
    with pytest.raises(ffi.Nb200Error) as e:
        t.special_tokens("<|xx|>")
    assert e.value.status == 9


def test_tokenizer_without_nospeech_token_is_reported(lib, tmp_path):
    j = json.loads(synth.synth_tokenizer_json(51864))
    j["added_tokens"] = [a for a in j["added_tokens"] if a["content"] != "<|nocaptions|>"]
    p = tmp_path / "tokenizer.json"
    p.write_text(json.dumps(j), encoding="utf-8")
    with pytest.raises(ffi.Nb200Error) as e:
        ffi.Tokenizer(str(p)).special_tokens("<|en|>")
    assert e.value.status == 9 and "<|nocaptions|> nor <|nospeech|>" in str(e.value)  # Error::TokenId(NO_SPEECH_TOKENS.join(" nor "))


def test_tokenizer_rejects_other_decoders_and_garbage(lib, tmp_path):
    j = json.loads(synth.synth_tokenizer_json(0, added=[(300, "<|endoftext|>", True)]))
    j["decoder"] = {"type": "WordPiece", "prefix": "##", "cleanup": True}
    p = tmp_path / "t.json"
    p.write_text(json.dumps(j), encoding="utf-8")
    with pytest.raises(ffi.Nb200Error) as e:
        ffi.Tokenizer(str(p))
    assert e.value.status == 8 and "Failed to load the tokenizer" in str(e.value)
    p.write_text('{"model": 3}')
    with pytest.raises(ffi.Nb200Error):
        ffi.Tokenizer(str(p))
    with pytest.raises(ffi.Nb200Error) as e:
        ffi.Tokenizer(str(tmp_path / "absent.json"))
    assert e.value.status == 7


def test_tokenizer_without_decoder_joins_with_spaces(lib, tmp_path):
    j = json.loads(synth.synth_tokenizer_json(0, added=[(300, "<|endoftext|>", True)]))
    j["decoder"] = None
    p = tmp_path / "t.json"
    p.write_text(json.dumps(j), encoding="utf-8")
    t = ffi.Tokenizer(str(p))
    tokenizers = pytest.importorskip("tokenizers")
    ref = tokenizers.Tokenizer.from_file(str(p))
    for ids in ([65, 66, 300, 67], [], [299], [300], [70, 9999, 71]):
        for skip in (True, False):
            assert t.decode(ids, skip) == ref.decode(ids, skip_special_tokens=skip)


# ---- safetensors ---------------------------------------------------------------------------------------------------------
def test_safetensors_dtypes_convert_to_f32(lib, tmp_path):
    g = torch.Generator().manual_seed(0)
    a = torch.randn(3, 5, generator=g)
    tensors = {"a.f32": a, "a.f16": a.half(), "a.bf16": a.bfloat16(), "a.f64": a.double(), "vec": torch.arange(7, dtype=torch.float32),
               "sub.f16": torch.tensor([6e-8, -6e-5, 65504.0, float("inf"), 0.0], dtype=torch.float16)}
    p = str(tmp_path / "m.safetensors")
    synth.write_safetensors(p, tensors, {"format": "pt"})
    for k, v in tensors.items():
        got = ffi.safetensors_read(p, k)
        assert got.shape == tuple(v.shape)
        assert np.array_equal(got, v.float().numpy()), k
    with pytest.raises(ffi.Nb200Error) as e:
        ffi.safetensors_read(p, "absent")
    assert e.value.status == 9


def test_safetensors_file_is_readable_by_the_safetensors_package(lib, tmp_path):
    """The writer used for the synthetic checkpoints produces what the real library reads (so the reader is tested on real layout)."""
    st = pytest.importorskip("safetensors.torch")
    w = synth.synth_weights(synth.model_config("test-micro"), seed=4, decoder=False)
    p = str(tmp_path / "m.safetensors")
    synth.write_safetensors(p, w)
    back = st.load_file(p)
    assert set(back) == set(w)
    for k in list(w)[:10]:
        assert torch.equal(back[k], w[k])
        assert np.array_equal(ffi.safetensors_read(p, k), w[k].numpy())
    # and the other direction: a file written by the library, read natively
    q = str(tmp_path / "lib.safetensors")
    st.save_file({k: v for k, v in list(w.items())[:6]}, q, metadata={"format": "pt"})
    for k in list(w)[:6]:
        assert np.array_equal(ffi.safetensors_read(q, k), w[k].numpy())


def _raw_safetensors(path, header: dict, data: bytes, hlen=None):
    h = json.dumps(header).encode()
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h) if hlen is None else hlen))
        f.write(h)
        f.write(data)


@pytest.mark.parametrize("case", ["short", "hlen", "json", "range", "size", "dtype", "offsets", "overflow"])
def test_safetensors_malformed_files_are_parse_errors(lib, tmp_path, case):
    p = str(tmp_path / "bad.safetensors")
    ok = {"t": {"dtype": "F32", "shape": [2, 2], "data_offsets": [0, 16]}}
    if case == "short":
        open(p, "wb").write(b"\x01\x02")
    elif case == "hlen":
        _raw_safetensors(p, ok, b"\0" * 16, hlen=1 << 40)
    elif case == "json":
        open(p, "wb").write(struct.pack("<Q", 5) + b"{oops" + b"\0" * 16)
    elif case == "range":
        _raw_safetensors(p, ok, b"\0" * 8)
    elif case == "size":
        _raw_safetensors(p, {"t": {"dtype": "F32", "shape": [3, 2], "data_offsets": [0, 16]}}, b"\0" * 16)
    elif case == "dtype":
        _raw_safetensors(p, {"t": {"dtype": "I64", "shape": [2], "data_offsets": [0, 16]}}, b"\0" * 16)
    elif case == "offsets":
        _raw_safetensors(p, {"t": {"dtype": "F32", "shape": [2, 2], "data_offsets": [16, 0]}}, b"\0" * 16)
    elif case == "overflow":  # 2^62 x 4 elements x 4 bytes wraps round to 0 bytes: must not pass as an empty tensor
        _raw_safetensors(p, {"t": {"dtype": "F32", "shape": [1 << 62, 4], "data_offsets": [0, 0]}}, b"\0" * 16)
    with pytest.raises(ffi.Nb200Error) as e:
        ffi.safetensors_read(p, "t")
    assert e.value.status == 8


def test_write_checkpoint_produces_the_three_hub_files(lib, tmp_path):
    c = synth.model_config("test-micro")
    w = synth.synth_weights(c, seed=1)
    cj, tj, sf = synth.write_checkpoint(str(tmp_path / "ckpt"), c, w, suppress_tokens=[11, 13])
    cfg, sup = ffi.config_from_file(cj)
    assert cfg == c and sup == [11, 13]
    t = ffi.Tokenizer(tj)
    assert t.special_tokens("<|en|>") == synth.special_tokens(c["vocab_size"])
    assert np.array_equal(ffi.safetensors_read(sf, "model.decoder.embed_tokens.weight"), w["model.decoder.embed_tokens.weight"].numpy())
    assert ffi.safetensors_read(sf, "model.encoder.embed_positions.weight").shape == (1500, c["d_model"])


# ---- GGUF (q8_0 `Quantized*` checkpoints) ------------------------------------------------------------------------------------
def test_gguf_q8_0_dequantises_exactly(lib, tmp_path):
    c = synth.model_config("test-micro")
    w = synth.synth_weights(c, seed=2, decoder=False)
    w["extra.f16"] = torch.randn(4, 64).half()
    p = str(tmp_path / "m.gguf")
    deq = synth.write_gguf(p, {k: v for k, v in w.items() if k != "extra.f16"})
    for k in list(deq)[:12]:
        got, typ = ffi.gguf_read(p, k)
        assert got.shape == tuple(w[k].shape)
        assert typ == (8 if (w[k].ndim >= 2 and w[k].shape[-1] % 32 == 0) else 0)
        assert np.array_equal(got, deq[k]), k                          # d (f16) * q (int8), exactly
        if typ == 8:
            amax = np.abs(w[k].numpy()).reshape(-1, 32).max(1)
            assert np.all(np.abs(got - w[k].numpy()).reshape(-1, 32).max(1) <= amax / 127 * 0.51 + amax * 1e-3)  # half a step + f16 scale rounding
    with pytest.raises(ffi.Nb200Error) as e:
        ffi.gguf_read(p, "absent")
    assert e.value.status == 9


def test_gguf_v2_and_alignment(lib, tmp_path):
    t = {"model.a.weight": np.arange(64, dtype=np.float32).reshape(2, 32) / 7, "model.b": np.arange(5, dtype=np.float32)}
    for version, align in ((2, 32), (3, 64)):
        p = str(tmp_path / f"v{version}.gguf")
        deq = synth.write_gguf(p, t, version=version, alignment=align)
        for k in t:
            assert np.array_equal(ffi.gguf_read(p, k)[0], deq[k])


@pytest.mark.parametrize("case", ["magic", "version", "truncated", "type", "row", "short", "offset_wrap", "numel_wrap"])
def test_gguf_malformed_files_are_parse_errors(lib, tmp_path, case):
    p = str(tmp_path / "bad.gguf")
    synth.write_gguf(p, {"t": np.ones((2, 32), np.float32)})
    raw = bytearray(open(p, "rb").read())
    if case == "magic":
        raw[0:4] = b"GGML"
    elif case == "version":
        raw[4:8] = struct.pack("<I", 1)
    elif case == "truncated":
        raw = raw[:-40]
    elif case == "type":
        i = raw.index(b"\x01\x00\x00\x00\x00\x00\x00\x00t")  # the tensor-info record of "t"
        j = i + 9 + 4 + 16                                      # name, rank, two dims
        raw[j:j + 4] = struct.pack("<I", 2)                     # Q4_0: not read
    elif case == "row":
        i = raw.index(b"\x01\x00\x00\x00\x00\x00\x00\x00t")
        j = i + 9 + 4
        raw[j:j + 8] = struct.pack("<Q", 31)                    # innermost dimension no longer a multiple of 32
    elif case == "short":
        raw = raw[:10]
    elif case == "offset_wrap":  # data offset close to 2^64: data_off + offset + bytes wraps round
        i = raw.index(b"\x01\x00\x00\x00\x00\x00\x00\x00t")
        j = i + 9 + 4 + 16 + 4                                  # name, rank, two dims, type
        raw[j:j + 8] = struct.pack("<Q", (1 << 64) - 64)
    elif case == "numel_wrap":  # 2^40 x 2^40 elements
        i = raw.index(b"\x01\x00\x00\x00\x00\x00\x00\x00t")
        j = i + 9 + 4
        raw[j:j + 16] = struct.pack("<QQ", 1 << 40, 1 << 40)
    open(p, "wb").write(bytes(raw))
    with pytest.raises(ffi.Nb200Error) as e:
        ffi.gguf_read(p, "t")
    assert e.value.status == 8
