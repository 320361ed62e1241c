"""ctypes binding of oracle/kb_post_oracle.c (post-mapping stages) -- test infrastructure only."""

from __future__ import annotations

import ctypes as C

import numpy as np

import oracle_lib as ol
from oracle_lib import _ptr


def _lib():
    L = ol.lib()
    if not getattr(L, "_post_bound", False):
        L.kbo_extract.argtypes = [C.c_void_p] * 6 + [C.c_int32] + [C.c_void_p] * 3
        L.kbo_translate.restype = C.c_int64
        L.kbo_translate.argtypes = [C.c_void_p] * 4 + [C.c_int32, C.c_int32] + [C.c_void_p] * 3
        L.kbo_protein_align.argtypes = [C.c_void_p] * 6 + [C.c_int32] * 4 + [C.c_void_p]
        L.kbo_cull_overlaps.argtypes = [C.c_void_p] * 5 + [C.c_double, C.c_void_p, C.c_int32, C.c_void_p]
        L.kbo_cluster.argtypes = [C.c_void_p] * 3 + [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L._post_bound = True
    return L


def offsets_of(lengths) -> np.ndarray:
    off = np.zeros(len(lengths), dtype=np.int64)
    if len(lengths) > 1:
        np.cumsum(np.asarray(lengths[:-1], dtype=np.int64), out=off[1:])
    return off


def extract(seqs, parent_off, indices, starts, ends, strands):
    n = len(indices)
    out = np.zeros(max(int((ends - starts).sum()), 1), dtype=np.uint8)
    oo, ol_ = np.zeros(n, np.int64), np.zeros(n, np.int32)
    a = [np.ascontiguousarray(x) for x in (seqs.astype(np.uint8), parent_off.astype(np.int64), indices.astype(np.int32),
                                           starts.astype(np.int32), ends.astype(np.int32), strands.astype(np.int8))]
    _lib().kbo_extract(*[_ptr(x) for x in a], n, _ptr(out), _ptr(oo), _ptr(ol_))
    return out[: int(ol_.sum())], oo, ol_


def translate(seqs, off, lengths, frames, to_stop: bool):
    n = len(lengths)
    out = np.zeros(int(lengths.sum()) // 3 + n + 1, dtype=np.uint8)
    oo, ol_ = np.zeros(n, np.int64), np.zeros(n, np.int32)
    a = [np.ascontiguousarray(x) for x in (seqs.astype(np.uint8), off.astype(np.int64), lengths.astype(np.int32), frames.astype(np.int8))]
    tot = _lib().kbo_translate(*[_ptr(x) for x in a], n, int(to_stop), _ptr(out), _ptr(oo), _ptr(ol_))
    return out[:tot], oo, ol_


def protein_align(q, q_len, t, t_len, k=20, gap_open=11, gap_extend=1):
    n = len(q_len)
    res = np.zeros((n, 8), dtype=np.int32)
    a = [np.ascontiguousarray(q, np.uint8), offsets_of(q_len), np.ascontiguousarray(q_len, np.int32),
         np.ascontiguousarray(t, np.uint8), offsets_of(t_len), np.ascontiguousarray(t_len, np.int32)]
    _lib().kbo_protein_align(*[_ptr(x) for x in a], n, k, gap_open, gap_extend, _ptr(res))
    return res


def cull_overlaps(order, g1, g2, starts, ends, frac, seg_off):
    a = [np.ascontiguousarray(x, np.int32) for x in (order, g1, g2, starts, ends)]
    seg_off = np.ascontiguousarray(seg_off, np.int64)
    kept = np.zeros(max(int(seg_off[-1]), 1), np.uint8)
    _lib().kbo_cull_overlaps(*[_ptr(x) for x in a], float(frac), _ptr(seg_off), len(seg_off) - 1, _ptr(kept))
    return kept[: int(seg_off[-1])]


def cluster(starts, ends, groups, tol, order, seg_off):
    a = [np.ascontiguousarray(x, np.int32) for x in (starts, ends, groups)]
    order = np.ascontiguousarray(order, np.int32)
    seg_off = np.ascontiguousarray(seg_off, np.int64)
    ids = np.zeros(max(int(seg_off[-1]), 1), np.int32)
    _lib().kbo_cluster(_ptr(a[0]), _ptr(a[1]), _ptr(a[2]), int(tol), _ptr(order), _ptr(seg_off), len(seg_off) - 1, _ptr(ids))
    return ids[: int(seg_off[-1])]
