"""Batch FASTA ingest: many assemblies' FASTA bytes -> the host arrays ``kb_map_assemblies`` /
``mapper.AssemblyBatch`` take, parsed by a pool of host threads inside the C library (replaces, for batches, the reference's
per-genome ``parse_fasta_bytes`` + array copy, ``core/genome.py:45`` / ``core/seq.py:307-325``)."""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import ptr


@dataclass
class IngestedBatch:
    seqs: np.ndarray              # uint8, all contigs of all assemblies, concatenated
    contig_off: np.ndarray        # int64, offset of every contig in seqs
    contig_len: np.ndarray        # int32
    asm_contig_start: np.ndarray  # int32, n_asm + 1
    names: list[list[str]]        # contig names per assembly (header up to the first whitespace)


def ingest_fasta(files: list[bytes], threads: int | None = None, out: np.ndarray | None = None) -> IngestedBatch:
    """``files[i]`` = the FASTA bytes of assembly i.  ``out`` may be a preallocated (e.g. pinned) uint8 buffer for the sequences."""
    L = _lib.load()
    n = len(files)
    threads = threads or min(os.cpu_count() or 1, 32)
    bufs = [np.frombuffer(f, dtype=np.uint8) for f in files]
    ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data if len(b) else None for b in bufs])
    lens = np.array([len(b) for b in bufs], dtype=np.int64)
    n_rec, n_seq = np.zeros(max(n, 1), np.int64), np.zeros(max(n, 1), np.int64)
    _lib.check(L.kb_fasta_ingest_count(ptrs, ptr(lens), n, threads, ptr(n_rec), ptr(n_seq)))
    rec_base = np.concatenate([[0], np.cumsum(n_rec[:n])]).astype(np.int64)
    seq_base = np.concatenate([[0], np.cumsum(n_seq[:n])]).astype(np.int64)
    tot_r, tot_s = int(rec_base[-1]), int(seq_base[-1])
    seqs = out if out is not None else np.empty(max(tot_s, 1), np.uint8)
    if len(seqs) < tot_s:
        raise ValueError("output buffer smaller than the total sequence length")
    off, ln = np.zeros(max(tot_r, 1), np.int64), np.zeros(max(tot_r, 1), np.int32)
    acs = np.zeros(n + 1, np.int32)
    name_off, name_len = np.zeros(max(tot_r, 1), np.int64), np.zeros(max(tot_r, 1), np.int32)
    _lib.check(L.kb_fasta_ingest_parse(ptrs, ptr(lens), n, threads, ptr(rec_base), ptr(seq_base), ptr(seqs), ptr(off), ptr(ln), ptr(acs),
                                       ptr(name_off), ptr(name_len)))
    names = []
    for i in range(n):
        r0, r1 = int(rec_base[i]), int(rec_base[i + 1])
        names.append([files[i][name_off[k] : name_off[k] + name_len[k]].decode("ascii", "replace") for k in range(r0, r1)])
    return IngestedBatch(seqs[:tot_s], off[:tot_r], ln[:tot_r], acs, names)
