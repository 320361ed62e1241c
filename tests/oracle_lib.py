"""ctypes binding of the CPU oracle (``oracle/kb_oracle.c``) -- test infrastructure only.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import this
module; the product package ``kaptive_b200`` never does.
"""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
_SO = ROOT / "oracle" / "_build" / "libkb_oracle.so"


class Params(C.Structure):
    _fields_ = [
        ("k", C.c_int32), ("w", C.c_int32),
        ("min_cnt", C.c_int32), ("min_chain_score", C.c_int32), ("bw", C.c_int32), ("max_gap", C.c_int32),
        ("max_chain_skip", C.c_int32), ("max_chain_iter", C.c_int32),
        ("chain_gap_scale", C.c_float),
        ("a", C.c_int32), ("b", C.c_int32), ("q", C.c_int32), ("e", C.c_int32), ("q2", C.c_int32), ("e2", C.c_int32),
        ("sc_ambi", C.c_int32),
        ("zdrop", C.c_int32), ("min_dp_max", C.c_int32), ("min_ksw_len", C.c_int32),
        ("mid_occ", C.c_int32), ("min_mid_occ", C.c_int32), ("max_mid_occ", C.c_int32),
        ("mid_occ_frac", C.c_float), ("q_occ_frac", C.c_float), ("mask_level", C.c_float),
        ("mask_len", C.c_int32), ("seed", C.c_int32), ("ext_bw", C.c_int32), ("max_sw_cells", C.c_int32),
    ]  # fmt: skip


HIT_DTYPE = np.dtype(
    [(n, np.int32) for n in (
        "gene", "q_start", "q_end", "t_ctg", "t_len", "t_start", "t_end", "strand", "score", "matches",
        "block_len", "edit_distance", "mapq", "is_primary", "dp_max", "chain_score", "chain_cnt",
        "cigar_off", "n_cigar")]
)  # fmt: skip
ANCHOR_DTYPE = np.dtype([(n, np.int32) for n in ("gene", "rev", "rid", "tpos", "qpos", "flags")])
CHAIN_DTYPE = np.dtype([(n, np.int32) for n in ("gene", "score", "cnt", "rev", "rid", "rs", "re", "qs", "qe")])


class Result(C.Structure):
    _fields_ = [
        ("n_hits", C.c_int32), ("hits", C.c_void_p),
        ("n_cigar", C.c_int32), ("cigar", C.c_void_p),
        ("mid_occ", C.c_int32), ("n_minimizers", C.c_int64),
        ("n_anchors", C.c_int64), ("anchors", C.c_void_p),
        ("n_chains", C.c_int32), ("chains", C.c_void_p),
    ]  # fmt: skip


def build() -> Path:
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle")], check=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        src_mtime = max(p.stat().st_mtime for p in (ROOT / "oracle").glob("*.[ch]"))
        if not _SO.exists() or _SO.stat().st_mtime < src_mtime:
            build()
        L = C.CDLL(str(_SO))
        L.kbo_params_default.argtypes = [C.POINTER(Params)]
        L.kbo_db_create.restype = C.c_void_p
        L.kbo_db_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(Params)]
        L.kbo_db_destroy.argtypes = [C.c_void_p]
        L.kbo_map_assembly.restype = C.POINTER(Result)
        L.kbo_map_assembly.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        L.kbo_result_free.argtypes = [C.POINTER(Result)]
        L.kbo_sketch.restype = C.c_int64
        L.kbo_sketch.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
        L.kbo_logf.restype = C.c_float
        L.kbo_logf.argtypes = [C.c_float]
        L.kbo_log2_fast.restype = C.c_float
        L.kbo_log2_fast.argtypes = [C.c_float]
        L.kbo_hash32.restype = C.c_uint32
        L.kbo_hash32.argtypes = [C.c_uint32, C.c_uint32]
        _lib = L
    return _lib


def default_params(**over) -> Params:
    p = Params()
    lib().kbo_params_default(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def _ptr(a: np.ndarray) -> C.c_void_p:
    return a.ctypes.data_as(C.c_void_p)


def _copy(ptr, n, dtype) -> np.ndarray:
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


class OracleDB:
    """Gene queries (``Serotyper._gene_seqs``, reference serotyping/core.py:111-121)."""

    def __init__(self, seqs: np.ndarray, offsets: np.ndarray, lengths: np.ndarray, params: Params | None = None):
        self.params = params or default_params()
        self._seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        self._off = np.ascontiguousarray(offsets, dtype=np.int64)
        self._len = np.ascontiguousarray(lengths, dtype=np.int32)
        self.n_genes = len(self._len)
        self._h = lib().kbo_db_create(_ptr(self._seqs), _ptr(self._off), _ptr(self._len), self.n_genes, C.byref(self.params))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().kbo_db_destroy(self._h)
            self._h = None

    def map(self, seqs: np.ndarray, offsets: np.ndarray, lengths: np.ndarray, keep_stages: bool = False) -> dict:
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        r = lib().kbo_map_assembly(self._h, _ptr(seqs), _ptr(offsets), _ptr(lengths), len(lengths), int(keep_stages))
        rc = r.contents
        out = {
            "hits": _copy(rc.hits, rc.n_hits, HIT_DTYPE),
            "cigar": _copy(rc.cigar, rc.n_cigar, np.dtype(np.uint32)),
            "mid_occ": int(rc.mid_occ),
            "n_minimizers": int(rc.n_minimizers),
            "anchors": _copy(rc.anchors, rc.n_anchors, ANCHOR_DTYPE),
            "chains": _copy(rc.chains, rc.n_chains, CHAIN_DTYPE),
        }
        lib().kbo_result_free(r)
        return out


def sketch(seq: bytes | np.ndarray, w: int = 10, k: int = 15) -> tuple[np.ndarray, np.ndarray]:
    s = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, np.uint8)
    cap = max(16, len(s))
    x = np.zeros(cap, dtype=np.uint64)
    y = np.zeros(cap, dtype=np.uint32)
    n = lib().kbo_sketch(_ptr(s), len(s), w, k, _ptr(x), _ptr(y), cap)
    return x[:n].copy(), y[:n].copy()


def cigar_string(cig: np.ndarray) -> str:
    return "".join(f"{int(c) >> 4}{'MIDNSHP=X'[int(c) & 0xF]}" for c in cig)
