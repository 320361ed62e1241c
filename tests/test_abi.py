"""CPU tier: the C-ABI library loads, exports every symbol include/kaptive_b200.h declares, fails loudly
without a device (no CPU fallback), and its host-only entry points (params, FASTA) behave."""

import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as ol
from kaptive_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def header_functions():
    txt = (ROOT / "include" / "kaptive_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/kaptive_b200.h but not exported"
    assert set(_lib.declared_symbols()) == set(names)


def test_params_default_equals_oracle_and_minimap2_defaults():
    p, o = _lib.default_params(), ol.default_params()
    for f, _ in _lib.KbParams._fields_:
        assert getattr(p, f) == getattr(o, f), f
    assert (p.k, p.w, p.a, p.b, p.q, p.e, p.q2, p.e2) == (15, 10, 2, 4, 4, 2, 24, 1)
    assert (p.min_cnt, p.min_chain_score, p.bw, p.max_gap, p.zdrop, p.min_dp_max) == (3, 40, 500, 5000, 400, 80)


def test_no_cpu_fallback_without_device():
    L = _lib.load()
    if L.kb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from kaptive_b200 import mapper

    with pytest.raises(_lib.KbError, match="no CUDA device"):
        mapper.GeneIndex([b"ACGT" * 100])
    with pytest.raises(_lib.KbError, match="no CUDA device"):
        mapper.AssemblyBatch.from_contigs([[b"ACGT" * 100]])


def test_argument_errors_have_messages():
    L = _lib.load()
    p = _lib.default_params()
    h = C.c_void_p(0)
    assert L.kb_index_create(None, None, None, 1, C.byref(p), 0, C.byref(h)) == -1
    assert b"null" in L.kb_last_error()
    p.max_gap = 9000
    g = np.frombuffer(b"ACGT" * 10, dtype=np.uint8)
    off, ln = np.zeros(1, np.int64), np.array([40], np.int32)
    assert L.kb_index_create(_lib.ptr(g), _lib.ptr(off), _lib.ptr(ln), 1, C.byref(p), 0, C.byref(h)) == -3
    assert b"max_gap" in L.kb_last_error()


def test_documented_limits_are_reported_not_crossed():
    """DESIGN.md section 3 limits: reported through kb_last_error before any device work (so checkable without a GPU)."""
    L = _lib.load()
    p = _lib.default_params()
    h = C.c_void_p(0)
    # > 32768 genes
    n = 32769
    g = np.frombuffer(b"ACGTACGTACGTACGTACGT" * n, dtype=np.uint8)
    ln = np.full(n, 20, np.int32)
    off = (np.arange(n, dtype=np.int64) * 20)
    assert L.kb_index_create(_lib.ptr(g), _lib.ptr(off), _lib.ptr(ln), n, C.byref(p), 0, C.byref(h)) == -3
    assert b"too many genes" in L.kb_last_error()
    # other k / w than the compiled-in minimap2 defaults
    p2 = _lib.default_params()
    p2.k = 19
    assert L.kb_index_create(_lib.ptr(g), _lib.ptr(off), _lib.ptr(ln), 4, C.byref(p2), 0, C.byref(h)) == -3
    assert b"k=15" in L.kb_last_error()
    # a gene minimizer that occurs in more than 2047 gene positions (one k-mer repeated over many genes)
    unit = np.frombuffer(b"ACGGTCATTGCAAGCTTGACCATGCAAGTC", dtype=np.uint8)  # 30 bases, no internal repeat
    n = 2100
    g = np.tile(unit, n)
    ln = np.full(n, len(unit), np.int32)
    off = (np.arange(n, dtype=np.int64) * len(unit))
    rc = L.kb_index_create(_lib.ptr(g), _lib.ptr(off), _lib.ptr(ln), n, C.byref(p), 0, C.byref(h))
    assert rc == -3 and b"2047" in L.kb_last_error()


def parse(data: bytes):
    import sys

    sys.path.insert(0, str(ROOT / "kaptive_b200" / "shim"))
    import importlib

    fasta = importlib.import_module("rammappy.fasta")
    return fasta.parse_fasta_bytes(data)


def test_fasta_parse():
    data = b">c1 desc here\nACGT\nacgtNN\n\n>c2\r\nTT\r\nGG\r\n>empty\n>c4\tx\nA"
    assert parse(data) == [("c1", b"ACGTacgtNN"), ("c2", b"TTGG"), ("empty", b""), ("c4", b"A")]
    assert parse(b"") == []
    assert parse(b"no header line\nACGT\n") == []
    big = b">x\n" + b"ACGT" * 25000 + b"\n"
    assert parse(big) == [("x", b"ACGT" * 25000)]


def test_shim_surface_matches_reference_call_sites():
    """Everything the reference touches on the module exists with the right shape."""
    import sys

    sys.path.insert(0, str(ROOT / "kaptive_b200" / "shim"))
    import rammappy
    from rammappy.align import Aligner

    assert hasattr(rammappy, "Preset") and hasattr(rammappy.Index, "build") and hasattr(rammappy.fasta, "parse_fasta_bytes")
    a = Aligner.__new__(Aligner)
    from rammappy._objects import Hit, Options, Strand, check_supported

    o = Options()
    assert (o.filtering.best_n, o.filtering.pri_ratio) == (5, 0.8)
    with pytest.raises(NotImplementedError):
        check_supported(o, True, False, False, None)
    o.filtering.best_n, o.filtering.pri_ratio = 50000, 0.0
    check_supported(o, True, False, False, None)
    assert "Forward" in repr(Strand.Forward) and "Forward" not in repr(Strand.Reverse)
    h = Hit(b"c", 0, 10, 100, 5, 15, Strand.Forward, 10, 9, 1, 14, 60, True, b"10M")
    for f in ("target_name", "query_start", "query_end", "target_len", "target_start", "target_end", "strand", "block_len", "matches",
              "edit_distance", "score", "mapq", "is_primary", "is_supplementary", "is_spliced", "divergence", "cs", "md", "cigar"):
        assert hasattr(h, f)
    assert a is not None
