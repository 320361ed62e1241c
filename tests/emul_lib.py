"""ctypes binding of tests/host_emul (the product's KB_HD device logic run sequentially on the host).

Test infrastructure only: lets the CPU-only tier compare the code the CUDA kernels are built from
with the oracle.  The product never loads it.
"""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from oracle_lib import ANCHOR_DTYPE, CHAIN_DTYPE, HIT_DTYPE, Params, _ptr

ROOT = Path(__file__).resolve().parent.parent
_DIR = ROOT / "tests" / "host_emul"
_SO = _DIR / "libkb_host_emul.so"
_SRCS = [_DIR / "kb_host_emul.cpp", ROOT / "kaptive_b200/csrc/kb_index.cpp", ROOT / "kaptive_b200/csrc/kb_params.cpp"]


def build() -> Path:
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", str(_SO)]
    subprocess.run(cmd + [str(s) for s in _SRCS], check=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        deps = list(_SRCS) + list((ROOT / "kaptive_b200/csrc").glob("*.cuh")) + list((ROOT / "kaptive_b200/csrc").glob("*.h"))
        if not _SO.exists() or _SO.stat().st_mtime < max(p.stat().st_mtime for p in deps):
            build()
        L = C.CDLL(str(_SO))
        L.kbe_params_default.argtypes = [C.POINTER(Params)]
        L.kbe_index_create.restype = C.c_void_p
        L.kbe_index_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(Params)]
        L.kbe_index_destroy.argtypes = [C.c_void_p]
        L.kbe_map_assembly.restype = C.c_void_p
        L.kbe_map_assembly.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.kbe_result_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.kbe_result_fetch.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.kbe_result_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def default_params(**over) -> Params:
    p = Params()
    lib().kbe_params_default(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


class EmulIndex:
    def __init__(self, seqs, offsets, lengths, params: Params | None = None):
        self.params = params or default_params()
        self._seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        self._off = np.ascontiguousarray(offsets, dtype=np.int64)
        self._len = np.ascontiguousarray(lengths, dtype=np.int32)
        self._h = lib().kbe_index_create(_ptr(self._seqs), _ptr(self._off), _ptr(self._len), len(self._len), C.byref(self.params))
        if not self._h:
            raise RuntimeError("kbe_index_create failed")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().kbe_index_destroy(self._h)
            self._h = None

    def map(self, seqs, offsets, lengths, lane_bases: int = 256, keep_stages: bool = False) -> dict:
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        r = lib().kbe_map_assembly(self._h, _ptr(seqs), _ptr(offsets), _ptr(lengths), len(lengths), lane_bases, int(keep_stages))
        cnt = np.zeros(8, dtype=np.int64)
        lib().kbe_result_counts(r, _ptr(cnt))
        hits = np.zeros(cnt[0], dtype=HIT_DTYPE)
        cigar = np.zeros(cnt[1], dtype=np.uint32)
        anchors = np.zeros(cnt[2], dtype=ANCHOR_DTYPE)
        chains = np.zeros(cnt[3], dtype=CHAIN_DTYPE)
        mzh = np.zeros(cnt[6], dtype=np.uint32)
        mzc = np.zeros(cnt[6], dtype=np.int32)
        mzp = np.zeros(cnt[6], dtype=np.uint32)
        lib().kbe_result_fetch(r, _ptr(hits), _ptr(cigar), _ptr(anchors), _ptr(chains), _ptr(mzh), _ptr(mzc), _ptr(mzp))
        lib().kbe_result_free(r)
        return {"hits": hits, "cigar": cigar, "anchors": anchors, "chains": chains, "mid_occ": int(cnt[4]),
                "n_minimizers": int(cnt[5]), "mz_hash": mzh, "mz_ctg": mzc, "mz_pos": mzp,
                "stage": {"ok": int(cnt[7]) & 0x1FFFFF, "fallback": (int(cnt[7]) >> 21) & 0x1FFFFF, "mismatch": int(cnt[7]) >> 42}}
