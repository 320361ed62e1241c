"""Post-mapping numerics on the GPU: batched replacements of the reference's numba kernels
(``core/seq.py:612-668`` extract, ``core/seq.py:671-741`` translate, ``core/pairwise.py:395-584`` protein Gotoh,
``core/interval.py:698-751`` overlap cull, ``core/interval.py:595-639`` spatial clustering).
Arrays in, one C-ABI call, arrays out; no CPU fallback."""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ptr


def _check(rc: int) -> None:
    if rc != 0:
        L = _lib.load()
        msg = L.kb_post_last_error() or L.kb_last_error()
        raise _lib.KbError(f"libkaptive_b200 status {rc}: {msg.decode() if msg else ''}")


def _offsets(lengths: np.ndarray) -> np.ndarray:
    off = np.zeros(len(lengths), dtype=np.int64)
    if len(lengths) > 1:
        np.cumsum(lengths[:-1].astype(np.int64), out=off[1:])
    return off


def extract(seqs, parent_off, indices, starts, ends, strands):
    L = _lib.load()
    seqs = np.ascontiguousarray(seqs, np.uint8)
    parent_off = np.ascontiguousarray(parent_off, np.int64)
    indices, starts, ends = (np.ascontiguousarray(a, np.int32) for a in (indices, starts, ends))
    strands = np.ascontiguousarray(strands, np.int8)
    n = len(indices)
    cap = int(np.maximum(ends - starts, 0).sum())
    out = np.zeros(max(cap, 1), np.uint8)
    oo, ol = np.zeros(max(n, 1), np.int64), np.zeros(max(n, 1), np.int32)
    _check(L.kb_post_extract(ptr(seqs), len(seqs), ptr(parent_off), len(parent_off), ptr(indices), ptr(starts), ptr(ends), ptr(strands), n,
                             ptr(out), cap, ptr(oo), ptr(ol)))
    return out[:cap], oo[:n], ol[:n]


def translate(seqs, offsets, lengths, frames, to_stop: bool = True):
    L = _lib.load()
    seqs = np.ascontiguousarray(seqs, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int64)
    lengths = np.ascontiguousarray(lengths, np.int32)
    frames = np.ascontiguousarray(frames, np.int8)
    n = len(lengths)
    cap = int(lengths.sum()) // 3 + n + 1
    out = np.zeros(cap, np.uint8)
    oo, ol = np.zeros(max(n, 1), np.int64), np.zeros(max(n, 1), np.int32)
    tot = C.c_int64(0)
    _check(L.kb_post_translate(ptr(seqs), len(seqs), ptr(offsets), ptr(lengths), ptr(frames), n, int(to_stop), ptr(out), cap, ptr(oo), ptr(ol),
                               C.byref(tot)))
    return out[: tot.value], oo[:n], ol[:n]


def protein_align(q, q_len, t, t_len, k: int = 20, gap_open: int = 11, gap_extend: int = 1) -> np.ndarray:
    """n x 8 int32: score, matches, mismatches, gaps, q_start, q_end, t_start, t_end
    (``PairwiseAlignments`` fields, reference core/pairwise.py:35-60); percent identity = matches*100/(m+mm+gaps)."""
    L = _lib.load()
    q, t = np.ascontiguousarray(q, np.uint8), np.ascontiguousarray(t, np.uint8)
    q_len, t_len = np.ascontiguousarray(q_len, np.int32), np.ascontiguousarray(t_len, np.int32)
    n = len(q_len)
    res = np.zeros((max(n, 1), 8), np.int32)
    _check(L.kb_post_protein_align(ptr(q), ptr(_offsets(q_len)), ptr(q_len), ptr(t), ptr(_offsets(t_len)), ptr(t_len), n, k, gap_open, gap_extend,
                                   ptr(res)))
    return res[:n]


def _local_order(keys: tuple, seg_off: np.ndarray) -> np.ndarray:
    """np.lexsort within every segment (the reference sorts one assembly at a time); indices local to the segment."""
    order = np.zeros(int(seg_off[-1]), np.int32)
    for s in range(len(seg_off) - 1):
        a, b = int(seg_off[s]), int(seg_off[s + 1])
        if b > a:
            order[a:b] = np.lexsort(tuple(k[a:b] for k in keys)).astype(np.int32)
    return order


def cull_order(scores, matches, mapq, seg_off, priority_mask=None) -> np.ndarray:
    """Evaluation order of ``Alignments.cull_overlaps`` per segment: ``np.lexsort((-qualities, -matches, -scores))`` (reference
    core/alignment.py:669-675).  ``qualities`` is uint8 there (core/alignment.py:466), so its negation WRAPS: among hits that tie on
    score and matches, mapq 0 sorts first, then 255, 254, ... 1 -- reproduced with the same uint8 arithmetic.  Host-only."""
    sc = np.asarray(scores, np.float64).copy()
    if priority_mask is not None:
        sc[np.asarray(priority_mask, bool)] += 1e9
    return _local_order((-np.asarray(mapq).astype(np.uint8), -np.asarray(matches).astype(np.int32), -sc), np.asarray(seg_off, np.int64))


def cull_overlaps(starts, ends, group1, group2, scores, matches, mapq, seg_off, max_overlap_fraction: float = 0.1,
                  priority_mask=None) -> np.ndarray:
    """``Alignments.cull_overlaps`` (reference core/alignment.py:643-686) for the hits of many assemblies at once: evaluation order
    = :func:`cull_order` (score + 1e9 where ``priority_mask``, then matches, then the wrapped uint8 mapq).  Returns the boolean
    kept mask."""
    L = _lib.load()
    seg_off = np.ascontiguousarray(seg_off, np.int64)
    order = cull_order(scores, matches, mapq, seg_off, priority_mask)
    a = [np.ascontiguousarray(x, np.int32) for x in (group1, group2, starts, ends)]
    kept = np.zeros(max(int(seg_off[-1]), 1), np.uint8)
    _check(L.kb_post_cull_overlaps(ptr(order), ptr(a[0]), ptr(a[1]), ptr(a[2]), ptr(a[3]), float(max_overlap_fraction), ptr(seg_off),
                                   len(seg_off) - 1, ptr(kept)))
    return kept[: int(seg_off[-1])].astype(bool)


def cluster_spatial(starts, ends, groups, seg_off, tolerance: int = 0) -> np.ndarray:
    """``Intervals.cluster_spatial`` (reference core/interval.py:471-493) per segment: order = lexsort((ends, starts, groups))."""
    L = _lib.load()
    seg_off = np.ascontiguousarray(seg_off, np.int64)
    st, en, g = (np.ascontiguousarray(x, np.int32) for x in (starts, ends, groups))
    order = _local_order((en, st, g), seg_off)
    ids = np.zeros(max(int(seg_off[-1]), 1), np.int32)
    _check(L.kb_post_cluster(ptr(st), ptr(en), ptr(g), int(tolerance), ptr(order), ptr(seg_off), len(seg_off) - 1, ptr(ids)))
    return ids[: int(seg_off[-1])]
