"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/seqkit_b200.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from seqkit_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "seqkit_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sk_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    declared = _declared_symbols()
    assert len(declared) >= 25
    assert sorted(L.SIGNATURES) == declared


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.sk_abi_version() == 2


def test_struct_sizes_match_header():
    # sizes implied by the field lists in include/seqkit_b200.h
    assert C.sizeof(L.Limits) == 32
    assert C.sizeof(L.Event) == 20
    assert C.sizeof(L.DemuxOpts) == 24
    assert C.sizeof(L.SynthSpec) == 56
    assert C.sizeof(L.Result) == 4 + 4 + 8 + 8 + 32 + 32 + 16 + 16 + 8 + 8 + 8 + 4 + 4 + 4 + 4 + 16


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    lim = L.Limits(1 << 20, 1 << 12, 1, 8, 1, 0)
    ctx = C.c_void_p()
    rc = L.lib().sk_ctx_create(0, C.byref(lim), C.byref(ctx))
    assert rc == -2 and not ctx.value
    assert b"no CPU fallback" in L.lib().sk_last_error(None)
    from seqkit_b200 import Engine
    with pytest.raises(L.SkError):
        Engine()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under seqkit_b200/ may import, include, dlopen or
    execute anything under oracle/."""
    pat = re.compile(r"^\s*(from|import)\s+\S*oracle|#\s*include\s*[<\"][^>\"]*oracle|liboracle|fasta_oracle|pyoracle|restatement",
                     re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "seqkit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not pat.search(text), f


def test_sheet_parser_matches_reference_rules():
    from seqkit_b200.engine import Engine
    names, bcs, err = Engine.parse_sheet(b"# c\nA\tACGT\textra\n\nB \tTTTT \n\tGGGG\nonlyname\n")
    assert names == [b"A", b"B "] and bcs == [b"ACGT", b"TTTT"] and err is None
    assert Engine.parse_sheet(b"A\tAC\nB\tACG\n")[2] == b"Barcodes in sample sheet must all be of same length."
    assert Engine.parse_sheet(b"A\t\tx\n")[2] == b"Sample A has no barcode."
    assert Engine.parse_sheet(b"A\tAC\nA\tGG\n")[2] == b"Sample A is listed multiple times in sample sheet."
