"""Batch FASTA ingest: many assemblies' FASTA bytes -> the host arrays ``kb_map_assemblies`` /
``mapper.AssemblyBatch`` take, parsed by a pool of host threads inside the C library (replaces, for batches, the reference's
per-genome ``parse_fasta_bytes`` + array copy, ``core/genome.py:45`` / ``core/seq.py:307-325``)."""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from .parallel import host_threads

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
    threads = threads or host_threads()
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


# ---------------------------------------------------------------------------------------------- packed ingest (2 bit + N mask)
_SEQUENCE_FILE_REGEX = __import__("re").compile(r"\.(?P<ext>f(asta|a|na|fn|as))(\.(?P<compression>gz|bz2|xz))?$")


def read_fasta_files(paths, threads: int | None = None) -> tuple[list[bytes], list[str]]:
    """Whole-file reads with the reference's own rules (``GenomeAssembly.from_file``, core/genome.py:105-106,194-214): the name must
    end in .fasta/.fa/.fna/.ffn/.fas, optionally .gz/.bz2/.xz (opened with gzip / bz2 / lzma); anything else raises
    ``NotImplementedError``.  Returns (file bytes, assembly ids = file name without the matched suffix).  Decompression runs on a
    thread pool (zlib, bz2 and lzma release the GIL)."""
    import bz2
    import gzip
    import lzma
    from concurrent.futures import ThreadPoolExecutor
    from pathlib import Path

    openers = {"gz": gzip.open, "bz2": bz2.open, "xz": lzma.open}
    jobs = []
    for p in paths:
        p = Path(p)
        m = _SEQUENCE_FILE_REGEX.search(p.name)
        if not m:
            raise NotImplementedError(f"Unsupported format: {p}")
        jobs.append((p, openers.get(m.group("compression"), open), p.name.removesuffix(m.group())))

    def read(job):
        p, opener, _ = job
        with opener(p, mode="rb") as fh:
            return fh.read()

    threads = threads or host_threads()
    with ThreadPoolExecutor(max(1, min(threads, len(jobs) or 1))) as pool:
        data = list(pool.map(read, jobs))
    return data, [j[2] for j in jobs]


@dataclass
class PackedBatch:
    """Host-side packed assemblies in the device layout (``kb_packed_layout``): what ``kb_map_assemblies_packed`` takes."""

    seq2: np.ndarray              # uint32, 16 bases per word
    nmask: np.ndarray             # uint32, 32 bases per word, bit set = ambiguous / padding
    contig_len: np.ndarray        # int32
    contig_soff: np.ndarray       # int64, storage offset (bases) of every contig
    asm_contig_start: np.ndarray  # int32, n_asm + 1
    storage_bases: int
    names: list[list[str]]

    @property
    def packed_bytes(self) -> int:
        return int(self.storage_bases // 4 + self.storage_bases // 8)


def ingest_fasta_packed(files: list[bytes], threads: int | None = None, out: tuple[np.ndarray, np.ndarray] | None = None,
                        use_simd: bool = True, want_names: bool = True) -> PackedBatch:
    """``files[i]`` = the FASTA bytes of assembly i -> 2-bit packed contigs + ambiguity mask, packed by the library's host threads
    straight into ``out = (seq2, nmask)`` (e.g. pinned uint32 arrays, reused from call to call) or fresh arrays."""
    L = _lib.load()
    n = len(files)
    threads = threads or host_threads()
    bufs = [np.frombuffer(f, dtype=np.uint8) for f in files]
    ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data if len(b) else None for b in bufs])
    lens = np.array([len(b) for b in bufs], dtype=np.int64)
    n_rec = np.zeros(max(n, 1), np.int64)
    _lib.check(L.kb_fasta_ingest_count_records(ptrs, ptr(lens), n, threads, ptr(n_rec)))
    rec_base = np.concatenate([[0], np.cumsum(n_rec[:n])]).astype(np.int64)
    tot_r = int(rec_base[-1])
    ln = np.zeros(max(tot_r, 1), np.int32)
    acs = np.zeros(n + 1, np.int32)
    name_off, name_len = np.zeros(max(tot_r, 1), np.int64), np.zeros(max(tot_r, 1), np.int32)
    _lib.check(L.kb_fasta_ingest_lengths(ptrs, ptr(lens), n, threads, ptr(rec_base), ptr(ln), ptr(acs), ptr(name_off), ptr(name_len)))
    soff = np.zeros(max(tot_r, 1), np.int64)
    storage = C.c_int64(0)
    _lib.check(L.kb_packed_layout(ptr(ln), tot_r, ptr(soff), C.byref(storage)))
    ws, wm = storage.value // 16, storage.value // 32
    if out is not None:
        seq2, nmask = out
        if len(seq2) < ws or len(nmask) < wm or seq2.dtype != np.uint32 or nmask.dtype != np.uint32:
            raise ValueError("packed output buffers too small (or not uint32)")
    else:
        seq2, nmask = np.empty(ws, np.uint32), np.empty(wm, np.uint32)
    _lib.check(L.kb_fasta_ingest_pack(ptrs, ptr(lens), n, threads, ptr(rec_base), ptr(ln), ptr(soff), storage.value, ptr(seq2), ptr(nmask),
                                      int(use_simd)))
    names: list[list[str]] = []
    if want_names:
        for i in range(n):
            r0, r1 = int(rec_base[i]), int(rec_base[i + 1])
            names.append([files[i][name_off[k] : name_off[k] + name_len[k]].decode("ascii", "replace") for k in range(r0, r1)])
    return PackedBatch(seq2[:ws], nmask[:wm], ln[:tot_r], soff[:tot_r], acs, int(storage.value), names)
