"""Host-side mirror of the reference's mapper interface for the gene -> contig mapping path.

Reference call sites this replaces (``/root/reference/src/kaptive``):

* ``core/genome.py:188-189``     ``rammappy.Index.build(contigs)``          -> :class:`AssemblyBatch`
* ``serotyping/core.py:111-121`` the ``(str(i).encode(), gene_bytes)`` query list -> :class:`GeneIndex`
* ``serotyping/core.py:148-154`` ``Aligner(...).map_batch(gene_seqs)``      -> :meth:`GeneIndex.map`
* ``core/alignment.py:392-474``  per-hit drain into the ``Alignments`` SoA  -> :class:`MapResult` (already SoA)

Everything here is thin: arrays in, one C-ABI call, arrays out.  No compute happens in Python and
there is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import HIT_FIELDS, KB_N_STAGES, STAGE_NAMES, KbHits, KbParams, check, ptr


import os

_ALLOW_LIMIT_DROPS = os.environ.get("KAPTIVE_B200_ALLOW_LIMIT_DROPS", "0") == "1"


def _flat(seqs: list[bytes]) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    lengths = np.fromiter((len(s) for s in seqs), dtype=np.int32, count=len(seqs))
    offsets = np.zeros(len(seqs), dtype=np.int64)
    if len(seqs) > 1:
        np.cumsum(lengths[:-1].astype(np.int64), out=offsets[1:])
    data = np.frombuffer(b"".join(seqs), dtype=np.uint8) if seqs else np.zeros(0, dtype=np.uint8)
    return data, offsets, lengths


@dataclass
class MapResult:
    """Alignment records of one mapping call, one numpy array per field, ordered by (assembly, gene, rank)."""

    hits: dict[str, np.ndarray]
    cigar: np.ndarray  # uint32 BAM-encoded (len << 4 | op), indexed by hits['cigar_off'] / hits['n_cigar']
    stage_ms: dict[str, float]
    counters: dict[str, int]
    mid_occ: np.ndarray
    anchors: np.ndarray | None = None  # stage dumps (KAPTIVE_B200_KEEP_STAGES=1): n x 7 int32
    chains: np.ndarray | None = None  # n x 10 int32

    def __len__(self) -> int:
        return len(self.hits["gene"])

    def cigar_of(self, i: int) -> np.ndarray:
        o, n = int(self.hits["cigar_off"][i]), int(self.hits["n_cigar"][i])
        return self.cigar[o : o + n]

    def cigar_bytes(self, i: int) -> bytes:
        return b"".join(b"%d%c" % (int(c) >> 4, b"MIDNSHP=X"[int(c) & 0xF]) for c in self.cigar_of(i))


class AssemblyBatch:
    """Device-resident 2-bit packed contigs of one or more assemblies (replaces ``rammappy.Index.build``)."""

    def __init__(self, seqs, offsets, lengths, asm_contig_start, device: int = 0):
        L = _lib.load()
        self.device = device
        self.n_asm = len(asm_contig_start) - 1
        self._h = C.c_void_p(0)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        acs = np.ascontiguousarray(asm_contig_start, dtype=np.int32)
        if isinstance(seqs, np.ndarray):
            seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
            sp = ptr(seqs)
        else:  # raw device / host pointer (e.g. torch tensor .data_ptr())
            sp = C.c_void_p(int(seqs))
        check(L.kb_batch_create(sp, ptr(offsets), ptr(lengths), ptr(acs), self.n_asm, device, C.byref(self._h)))
        self.contig_lengths = lengths
        self.asm_contig_start = acs

    @classmethod
    def from_contigs(cls, assemblies: list[list[bytes]], device: int = 0) -> "AssemblyBatch":
        flat = [c for a in assemblies for c in a]
        data, off, ln = _flat(flat)
        acs = np.zeros(len(assemblies) + 1, dtype=np.int32)
        np.cumsum([len(a) for a in assemblies], out=acs[1:])
        return cls(data, off, ln, acs, device)

    @classmethod
    def from_fasta(cls, files: list[bytes], device: int = 0, threads: int | None = None) -> "AssemblyBatch":
        """One assembly per FASTA buffer, parsed by the library's thread pool (``kaptive_b200.ingest``)."""
        from . import ingest

        b = ingest.ingest_fasta(files, threads=threads)
        batch = cls(b.seqs if len(b.seqs) else np.zeros(1, np.uint8), b.contig_off, b.contig_len, b.asm_contig_start, device)
        batch.contig_names = b.names
        return batch

    @classmethod
    def from_packed(cls, pb, device: int = 0, first_soff: int = 128) -> "AssemblyBatch":
        """Host-packed contigs (``ingest.ingest_fasta_packed``) -> device batch: the 2-bit words go over PCIe as they are.
        ``first_soff``: storage offset (bases) of the first contig of ``pb.contig_len`` inside ``pb.seq2`` / ``pb.nmask`` when ``pb`` is a
        slice of a larger ingest call."""
        L = _lib.load()
        self = cls.__new__(cls)
        self.device, self.n_asm, self._h = device, len(pb.asm_contig_start) - 1, C.c_void_p(0)
        check(L.kb_batch_create_packed(ptr(pb.seq2), ptr(pb.nmask), int(first_soff), ptr(pb.contig_len), ptr(pb.asm_contig_start), self.n_asm,
                                       device, C.byref(self._h)))
        self.contig_lengths, self.asm_contig_start, self.contig_names = pb.contig_len, pb.asm_contig_start, pb.names
        return self

    def download_packed(self, out: tuple[np.ndarray, np.ndarray] | None = None) -> tuple[np.ndarray, np.ndarray]:
        """The batch's 2-bit words and mask words as host arrays (layout of ``kb_packed_layout`` over the batch's contigs)."""
        L = _lib.load()
        st = C.c_int64(0)
        check(L.kb_batch_download_packed(self._h, None, None, C.byref(st)))
        seq2, nmask = out if out is not None else (np.empty(st.value // 16, np.uint32), np.empty(st.value // 32, np.uint32))
        if len(seq2) < st.value // 16 or len(nmask) < st.value // 32:
            raise ValueError("buffers smaller than the batch's storage")
        check(L.kb_batch_download_packed(self._h, ptr(seq2), ptr(nmask), None))
        return seq2[: st.value // 16], nmask[: st.value // 32]

    @property
    def total_bases(self) -> int:
        return int(_lib.load().kb_batch_total_bases(self._h))

    @property
    def packed_bytes(self) -> int:
        return int(_lib.load().kb_batch_packed_bytes(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            _lib.load().kb_batch_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GeneIndex:
    """Device-resident minimizer hash of the DB genes (the query side of ``Aligner.map_batch``)."""

    def __init__(self, genes: list[bytes] | None = None, params: KbParams | None = None, device: int = 0, _handle=None):
        L = _lib.load()
        self.device = device
        self._h = C.c_void_p(0)
        if _handle is not None:
            self._h = _handle
        else:
            self.params = params or _lib.default_params()
            data, off, ln = _flat(list(genes or []))
            check(L.kb_index_create(ptr(data), ptr(off), ptr(ln), len(ln), C.byref(self.params), device, C.byref(self._h)))
        self.n_genes = int(L.kb_index_n_genes(self._h))
        self.n_minimizers = int(L.kb_index_n_minimizers(self._h))

    # ---- one-off broadcast support (multi-GPU: rank 0 builds, NCCL-broadcasts the image)
    def serialize(self) -> np.ndarray:
        L = _lib.load()
        n = int(L.kb_index_serialized_size(self._h))
        buf = np.zeros(n, dtype=np.uint8)
        check(L.kb_index_serialize(self._h, ptr(buf), n))
        return buf

    @classmethod
    def deserialize(cls, buf: np.ndarray, device: int = 0) -> "GeneIndex":
        L = _lib.load()
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        h = C.c_void_p(0)
        check(L.kb_index_deserialize(ptr(buf), len(buf), device, C.byref(h)))
        return cls(device=device, _handle=h)

    def map(self, batch: AssemblyBatch, fetch: bool = True, out=None) -> MapResult:
        """Map every gene onto every assembly of `batch`.  `out` = (KbHits, arrays, cigar) from `alloc_hits` plus a uint32
        cigar pool lets a caller that maps batch after batch reuse its host arrays (a fresh 30 MB of numpy arrays per call
        costs more in page faults than the copy itself); the result's arrays are then views of them."""
        L = _lib.load()
        r = C.c_void_p(0)
        check(L.kb_map_batch(self._h, batch._h, C.byref(r)))
        try:
            return _drain(L, r, batch.n_asm, fetch, out)
        finally:
            L.kb_result_destroy(r)

    def map_contigs(self, assemblies: list[list[bytes]]) -> MapResult:
        b = AssemblyBatch.from_contigs(assemblies, self.device)
        try:
            return self.map(b)
        finally:
            b.close()

    def map_packed(self, pb, out=None) -> MapResult:
        """Host-packed assemblies in, hits out, through the one-call C-ABI entry ``kb_map_assemblies_packed`` (slabs, copies inside)."""
        L = _lib.load()
        n_asm = len(pb.asm_contig_start) - 1
        if out is None:
            h, arrays = alloc_hits(1024 * max(n_asm, 1))
            cig = np.zeros(1024 * max(n_asm, 1) * 16, dtype=np.uint32)
        else:
            h, arrays, cig = out
        nh, nc = C.c_int64(0), C.c_int64(0)
        check(L.kb_map_assemblies_packed(self._h, ptr(pb.seq2), ptr(pb.nmask), ptr(pb.contig_len), ptr(pb.asm_contig_start), n_asm, C.byref(h),
                                         C.byref(nh), ptr(cig), len(cig), C.byref(nc)))
        return MapResult(hits={k: v[: nh.value] for k, v in arrays.items()}, cigar=cig[: nc.value], stage_ms={}, counters={},
                         mid_occ=np.zeros(0, np.int32))

    def bench_scan(self, batch: AssemblyBatch, iters: int = 5) -> tuple[float, int]:
        ms = C.c_float(0)
        na = C.c_int64(0)
        check(_lib.load().kb_bench_scan(self._h, batch._h, iters, C.byref(ms), C.byref(na)))
        return float(ms.value), int(na.value)

    def scan_minimizers(self, batch: AssemblyBatch, asm_id: int, cap: int) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        h = np.zeros(cap, dtype=np.uint32)
        c = np.zeros(cap, dtype=np.int32)
        p = np.zeros(cap, dtype=np.uint32)
        n = C.c_int64(0)
        check(_lib.load().kb_scan_minimizers(self._h, batch._h, asm_id, ptr(h), ptr(c), ptr(p), cap, C.byref(n)))
        m = min(int(n.value), cap)
        return h[:m], c[:m], p[:m]

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            _lib.load().kb_index_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def alloc_hits(n: int) -> tuple[KbHits, dict[str, np.ndarray]]:
    arrays = {name: np.zeros(max(n, 1), dtype=dt) for name, dt in HIT_FIELDS}
    h = KbHits()
    h.capacity = max(n, 1)
    for name, _ in HIT_FIELDS:
        setattr(h, name, arrays[name].ctypes.data)
    return h, arrays


def _drain(L, r, n_asm: int, fetch: bool, out=None) -> MapResult:
    nh, nc = C.c_int64(0), C.c_int64(0)
    check(L.kb_result_size(r, C.byref(nh), C.byref(nc)))
    ms = np.zeros(KB_N_STAGES, dtype=np.float32)
    cnt = np.zeros(16, dtype=np.int64)
    check(L.kb_result_stage_ms(r, ptr(ms)))
    check(L.kb_result_counters(r, ptr(cnt)))
    mid = np.zeros(max(n_asm, 1), dtype=np.int32)
    check(L.kb_result_mid_occ(r, ptr(mid)))
    hits: dict[str, np.ndarray] = {name: np.zeros(0, dtype=dt) for name, dt in HIT_FIELDS}
    cigar = np.zeros(0, dtype=np.uint32)
    if fetch:
        if out is not None:
            h, arrays, cigar = out  # too small: kb_result_fetch reports KB_ERR_CAPACITY
        else:
            h, arrays = alloc_hits(nh.value)
            cigar = np.zeros(max(nc.value, 1), dtype=np.uint32)
        check(L.kb_result_fetch(r, C.byref(h), ptr(cigar), len(cigar)))
        hits = {k: v[: nh.value] for k, v in arrays.items()}
        cigar = cigar[: nc.value]
    anchors = chains = None
    n = C.c_int64(0)
    check(L.kb_result_fetch_anchors(r, None, 0, C.byref(n)))
    if n.value:
        anchors = np.zeros((n.value, 7), dtype=np.int32)
        check(L.kb_result_fetch_anchors(r, ptr(anchors), n.value, C.byref(n)))
    check(L.kb_result_fetch_chains(r, None, 0, C.byref(n)))
    if n.value:
        chains = np.zeros((n.value, 10), dtype=np.int32)
        check(L.kb_result_fetch_chains(r, ptr(chains), n.value, C.byref(n)))
    names = ("minimizers", "anchors", "groups", "chains", "raw_hits", "launches", "dp_cells", "slow_chains", "limit_drops")
    if int(cnt[8]) and not _ALLOW_LIMIT_DROPS:
        raise _lib.KbError(f"{int(cnt[8])} alignment(s) ran into an internal limit (chain window > 65536 bases or CIGAR > 8192 operations) "
                           "and were dropped; set KAPTIVE_B200_ALLOW_LIMIT_DROPS=1 to accept incomplete results")
    return MapResult(
        hits=hits,
        cigar=cigar,
        stage_ms={k: float(v) for k, v in zip(STAGE_NAMES, ms)},
        counters={k: int(v) for k, v in zip(names, cnt)},
        mid_occ=mid[:n_asm],
        anchors=anchors,
        chains=chains,
    )
