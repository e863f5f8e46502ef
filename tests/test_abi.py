"""The C-ABI library: loads without a GPU, exports exactly what include/norma_b200.h declares, and refuses to
compute (loudly, with an error string) when no sm_100 device is present.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT
from norma_b200 import ffi


def header_symbols():
    src = open(os.path.join(ROOT, "include", "norma_b200.h")).read()
    return re.findall(r"^NB200_API\s+[\w\s\*]+?\b(nb200_\w+)\s*\(", src, flags=re.M)


def test_header_and_binding_agree():
    hs = header_symbols()
    assert len(hs) == len(set(hs)) and len(hs) >= 30
    assert sorted(hs) == sorted(ffi.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", ffi.LIB_PATH], text=True)
    exported = set(re.findall(r" T (nb200_\w+)", out))
    assert exported == set(header_symbols())  # nothing missing, nothing extra leaks (hidden visibility)
    for s in header_symbols():
        assert hasattr(lib, s)


def test_rust_binding_block_is_generated_from_the_header():
    """INTEGRATION.md's `extern "C"` block (the norma-b200-sys crate) is scripts/gen_rust_bindings.py's output for the CURRENT header:
    nothing hand-written, nothing omitted."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("gen_rust_bindings", os.path.join(ROOT, "scripts", "gen_rust_bindings.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    block = gen.generate(open(os.path.join(ROOT, "include", "norma_b200.h")).read())
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    a, b = doc.index(gen.BEGIN) + len(gen.BEGIN), doc.index(gen.END)
    assert doc[a:b].strip() == block.strip(), "run `python scripts/gen_rust_bindings.py`"
    fns = re.findall(r"pub fn (nb200_\w+)\(", block)
    assert sorted(fns) == sorted(header_symbols())
    for const in ("NB200_OK", "NB200_BUFFER_TOO_SMALL", "NB200_BF16", "NB200_DECODE_SEPARATE", "NB200_TASK_TRANSLATE"):
        assert f"pub const {const}: c_int" in block
    assert "pub struct nb200_config {" in block and "pub max_batch: i32," in block and "pub ts_one: u32," in block


def test_header_compiles_as_c():
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "norma_b200.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_library_is_sm100a_only_and_native():
    sass = subprocess.run(["cuobjdump", "-lelf", ffi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert not re.search(r"sm_(5|6|7|8|9)\d", sass)  # no multi-arch dispatch


def test_no_gpu_fails_loudly(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    n = C.c_int(-1)
    st = lib.nb200_device_count(C.byref(n))
    assert st != 0 and n.value == 0
    assert lib.nb200_last_error(None)
    from norma_b200 import synth

    with pytest.raises(ffi.Nb200Error):
        ffi.Context(synth.model_config("test-micro"))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "norma_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src or f.endswith(".md"), f


def test_bf16_round_helper():
    import numpy as np

    x = np.array([1.0, 1.00390625, 3.14159265, -2.71828, 1e-30, 65504.0], np.float32)
    import torch

    ref = torch.from_numpy(x).to(torch.bfloat16).float().numpy()
    assert np.array_equal(ffi.bf16_round(x), ref)
