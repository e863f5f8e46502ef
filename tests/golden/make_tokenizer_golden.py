"""Generates tests/golden/tokenizer_small.json and tokenizer_small_decodes.json (run once, here).

The decodes are produced by the Python `tokenizers` package (0.22.2 in this image) — the binding of the same Rust
`tokenizers` crate norma links (0.20.0, /root/reference/Cargo.lock:2586-2587; `Tokenizer::from_file`, `token_to_id`,
`decode(ids, skip_special_tokens)` at monolingual.rs:348, mod.rs:86-90, model.rs:147).  So unlike the numeric fixtures these
ARE reference-library outputs: they pin the C++ restatement in norma_b200/csrc/host/loader.cc.

    python tests/golden/make_tokenizer_golden.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from norma_b200 import synth  # noqa: E402

N_TEXT = 700


def small_added():
    names = ["<|endoftext|>", "<|startoftranscript|>"] + [f"<|{c}|>" for c in synth.LANGUAGE_CODES]
    names += ["<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>", "<|nospeech|>", "<|notimestamps|>"]
    out = [(N_TEXT + i, n, True) for i, n in enumerate(names)]
    t0 = N_TEXT + len(names)
    out += [(t0 + i, f"<|{i * 0.02:.2f}|>", False) for i in range(51)]  # <|0.00|> .. <|1.00|>
    # added tokens whose characters leave the byte alphabet (ByteLevel keeps such a token verbatim)
    out += [(t0 + 51, "hello world", False), (t0 + 52, "café \U0001f642", False), (t0 + 53, "<|weird special|>", True)]
    return out


def main():
    import tokenizers

    added = small_added()
    text = synth.synth_tokenizer_json(0, seed=5, added=added)
    path = os.path.join(HERE, "tokenizer_small.json")
    with open(path, "w", encoding="utf-8") as f:
        f.write(text)
    ref = tokenizers.Tokenizer.from_file(path)
    rng = np.random.default_rng(11)
    top = added[-1][0] + 1
    cases = []
    for it in range(400):
        n = int(rng.integers(0, 14))
        if it % 4 == 0:
            ids = rng.integers(0, 256, n)           # raw bytes: mostly invalid UTF-8 -> U+FFFD handling
        elif it % 4 == 1:
            ids = rng.integers(0, N_TEXT, n)
        elif it % 4 == 2:
            ids = rng.integers(N_TEXT - 20, top + 30, n)  # specials, timestamps, odd added tokens, ids nobody owns
        else:
            ids = rng.integers(0, top + 5, n)
        ids = [int(x) for x in ids]
        for skip in (True, False):
            cases.append(dict(ids=ids, skip=skip, text=ref.decode(ids, skip_special_tokens=skip)))
    # hand-written UTF-8 edge cases (byte tokens have id == position in the GPT-2 alphabet order, so look them up)
    vocab = json.loads(text)["model"]["vocab"]
    b2u = synth.bytes_to_unicode()
    for bs in (b"\xe4\xbd\xa0\xe5\xa5\xbd", b"\xe4\xbd", b"\xf0\x9f\x99", b"\xf0\x9f\x99\x82", b"\xed\xa0\x80", b"\xc0\xaf", b"\xf4\x90\x80\x80",
               b"\xe0\x80\x80", b"a\x80b", b"\xf0\x9fab", b"\xe4\xbd\xe4\xbd\xa0", b"\xff\xfe", b"\xc3"):
        ids = [vocab[b2u[b]] for b in bs]
        cases.append(dict(ids=ids, skip=True, text=ref.decode(ids, skip_special_tokens=True)))
    ids_of = {c: ref.token_to_id(c) for _, c, _ in added}
    ids_of["<|nocaptions|>"] = ref.token_to_id("<|nocaptions|>")  # None: absent
    some_text = {k: v for k, v in list(vocab.items())[250:270]}
    with open(os.path.join(HERE, "tokenizer_small_decodes.json"), "w", encoding="utf-8") as f:
        json.dump(dict(tokenizers_version=tokenizers.__version__, cases=cases, token_to_id=ids_of, text_token_to_id=some_text), f, ensure_ascii=True)
    print(len(cases), "cases;", os.path.getsize(path), "B tokenizer")


if __name__ == "__main__":
    main()
