"""``rammappy.fasta.parse_fasta_bytes`` (reference call site core/genome.py:45-46) on the C-ABI's FASTA reader."""

from __future__ import annotations

import ctypes as C

import numpy as np

from kaptive_b200 import _lib


def parse_fasta_bytes(data: bytes) -> list[tuple[str, bytes]]:
    L = _lib.load()
    buf = np.frombuffer(data, dtype=np.uint8)
    n_rec, n_bytes = C.c_int64(0), C.c_int64(0)
    _lib.check(L.kb_fasta_count(_lib.ptr(buf), len(buf), C.byref(n_rec), C.byref(n_bytes)))
    n = n_rec.value
    name_off = np.zeros(max(n, 1), dtype=np.int64)
    name_len = np.zeros(max(n, 1), dtype=np.int32)
    seq_off = np.zeros(max(n, 1), dtype=np.int64)
    seq_len = np.zeros(max(n, 1), dtype=np.int32)
    seq = np.zeros(max(n_bytes.value, 1), dtype=np.uint8)
    _lib.check(L.kb_fasta_parse(_lib.ptr(buf), len(buf), n, _lib.ptr(name_off), _lib.ptr(name_len), _lib.ptr(seq), len(seq),
                                _lib.ptr(seq_off), _lib.ptr(seq_len)))
    out = []
    sb = seq.tobytes()
    for i in range(n):
        name = data[name_off[i] : name_off[i] + name_len[i]].decode("ascii", "replace")
        out.append((name, sb[seq_off[i] : seq_off[i] + seq_len[i]]))
    return out
