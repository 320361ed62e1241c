"""Batched typing of mapped assemblies: ``type_many``.

The reference types one genome per call (``Serotyper.__call__``, ``/root/reference/src/kaptive/serotyping/core.py:124-486``) and
spends ~40 ms per assembly after the mapper in numba wake-ups and Python.  ``type_many`` takes the hits of a whole batch
(:class:`kaptive_b200.mapper.MapResult`) and the device-resident batch they were mapped from, and returns every assembly's call:

* scoring (core.py:157-207): per-locus sums in the C library (``kb_type_score``), ``completeness ** 3`` and the first-maximum
  ``argmax`` in numpy -- the very float32 power the reference evaluates;
* reconstruction, gene states, confidence (core.py:209-459): ``kb_type_call`` -- array logic on host threads with the reference's
  tie rules, and ONE device pass that extracts, translates and protein-aligns every retained hit of every assembly from the
  2-bit batch (replaces the per-genome extract / translate / Gotoh calls at core.py:333,352,360,378);
* report rows (``KaptiveRow.from_result`` + ``__bytes__``, serotyping/io.py:192-296) for the whole batch: :meth:`TypedBatch.rows`.

There is no CPU fallback for the numerics: without the CUDA library the calls raise.
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from .parallel import host_threads

from . import _lib
from ._lib import ptr

PROBLEM_SYMBOLS = tuple(b"".join(s for bit, s in ((1, b"?"), (2, b"+"), (4, b"-"), (8, b"*"), (16, b"!")) if i & bit) for i in range(32))
STATE_SUFFIX = (None, b"partial", b"truncated", b"below_id_threshold")


def _check(rc: int) -> None:
    if rc != 0:
        L = _lib.load()
        msg = L.kb_type_last_error() or L.kb_last_error()
        raise _lib.KbError(f"libkaptive_b200 status {rc}: {msg.decode() if msg else ''}")


class TypingDB:
    """The database tables the post-mapping logic reads (``kaptive.db.Database`` fields, db/core.py:82-98), device-resident where
    the numerics need them (translations)."""

    def __init__(self, gene_len, gene_locus, extra, gene_pos, gene_strand, locus_len, translations: list[bytes], locus_names, gene_names,
                 serotypes=None, id_threshold: float = 82.5, device: int = 0, name: str = "Synthetic K", version: str = "1.0.0",
                 kaptive_version: str = "3.3.2.dev12"):
        L = _lib.load()
        self.gene_len = np.ascontiguousarray(gene_len, np.int32)
        self.gene_locus = np.ascontiguousarray(gene_locus, np.int32)
        self.extra = np.ascontiguousarray(extra, np.uint8)
        self.gene_pos = np.ascontiguousarray(gene_pos, np.int32)
        self.gene_strand = np.ascontiguousarray(gene_strand, np.int8)
        self.locus_len = np.ascontiguousarray(locus_len, np.int32)
        self.n_genes, self.n_loci = len(self.gene_len), len(self.locus_len)
        self.locus_names, self.gene_names = list(locus_names), list(gene_names)
        self.serotypes = list(serotypes) if serotypes is not None else list(locus_names)
        self.id_threshold, self.device = float(id_threshold), device
        self.name, self.version, self.kaptive_version = name, version, kaptive_version
        # expected genes per locus, float32, at least 1 (serotyping/core.py:101-107)
        exp = np.zeros(self.n_loci, np.float32)
        np.add.at(exp, self.gene_locus[self.extra == 0], 1.0)
        self.expected_per_locus = np.maximum(exp, np.float32(1.0))
        tl = np.array([len(t) for t in translations], np.int32)
        flat = np.frombuffer(b"".join(translations), np.uint8) if len(translations) else np.zeros(0, np.uint8)
        self._h = C.c_void_p(0)
        _check(L.kb_typedb_create(self.n_genes, ptr(self.gene_len), ptr(self.gene_locus), ptr(self.extra), ptr(self.gene_pos), ptr(self.gene_strand),
                                  self.n_loci, ptr(self.locus_len), int(self.locus_len.max()) if self.n_loci else 0, ptr(np.ascontiguousarray(flat)),
                                  ptr(tl), self.id_threshold, device, C.byref(self._h)))

    @classmethod
    def from_synth(cls, db, id_threshold: float = 82.5, device: int = 0) -> "TypingDB":
        """From a :class:`kaptive_b200.synth.SynthDB`; translations by the library's own table-11 kernel (no stop truncation,
        like ``genes.translate()`` at db/core.py:455)."""
        from . import post

        lens = np.array([len(g) for g in db.genes], np.int32)
        flat = np.frombuffer(b"".join(db.genes), np.uint8)
        off = np.zeros(len(lens), np.int64)
        if len(lens) > 1:
            np.cumsum(lens[:-1].astype(np.int64), out=off[1:])
        aa, ao, al = post.translate(flat, off, lens, np.zeros(len(lens), np.int8), to_stop=False)
        tr = [aa[int(o) : int(o) + int(n)].tobytes() for o, n in zip(ao, al)]
        return cls(lens, db.gene_locus, db.extra, db.gene_pos, db.gene_strand, [len(s) for s in db.loci], tr, db.locus_names, db.gene_names,
                   id_threshold=id_threshold, device=device)

    @classmethod
    def from_kaptive(cls, database, device: int = 0) -> "TypingDB":
        """From a compiled ``kaptive.db.Database`` (e.g. the pickle under ``~/.kaptive``)."""
        tr = database.translations
        trs = [bytes(tr.seqs[int(o) : int(o) + int(n)]) for o, n in zip(tr.offsets, tr.lengths)]
        return cls(database.genes.lengths, database.gene_locus_indices, database.extra_genes, database.gene_positions,
                   database.gene_intervals.strands, database.loci.lengths, trs, list(database.loci.ids), list(database.genes.ids),
                   serotypes=list(database.serotypes), id_threshold=database.metadata.id_threshold, device=device,
                   name=database.metadata.name, version=database.metadata.version)

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            _lib.load().kb_typedb_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class TypedBatch:
    """Calls of a batch of assemblies, one array entry per assembly, plus the gene hits / locus pieces / missing genes of all of
    them back to back (offset arrays).  Field meanings are those of ``SerotypingResult`` (serotyping/models.py:513-536)."""

    db: TypingDB
    best_locus: np.ndarray          # int32
    best_locus_score: np.ndarray    # float64, un-penalised sum of q_cov (core.py:471)
    completeness: np.ndarray        # float64, after reconstruction
    percent_coverage: np.ndarray
    length_discrepancy: np.ndarray  # NaN unless the locus is in one piece
    typeable: np.ndarray            # bool
    problems: np.ndarray            # uint8 bit set (SerotypingProblem)
    n_pieces: np.ndarray
    gene_hit_off: np.ndarray
    piece_off: np.ndarray
    missing_off: np.ndarray
    gene_hits: dict[str, np.ndarray]
    pieces: dict[str, np.ndarray]
    missing: np.ndarray

    def __len__(self) -> int:
        return len(self.best_locus)

    def percent_identity(self, a: int) -> float:
        """Mean protein identity of the NORMAL genes of assembly `a` (core.py:395-396: ``np.mean`` of a float32 array)."""
        lo, hi = int(self.gene_hit_off[a]), int(self.gene_hit_off[a + 1])
        v = self.gene_hits["prot_ident"][lo:hi][self.gene_hits["state"][lo:hi] == 0]
        return float(np.mean(v)) if v.size > 0 else 0.0

    def best_locus_name(self, a: int) -> str:
        return self.db.locus_names[int(self.best_locus[a])]

    def row(self, a: int, assembly: str) -> bytes:
        """The ``KaptiveRow`` line of assembly `a` (serotyping/io.py:192-296 + ReportRow.__bytes__): tab separated, newline ended."""
        db, gh = self.db, self.gene_hits
        lo, hi = int(self.gene_hit_off[a]), int(self.gene_hit_off[a + 1])
        gene, state = gh["gene"][lo:hi], gh["state"][lo:hi]
        ins, exp, ext = gh["is_inside"][lo:hi].astype(bool), gh["is_expected"][lo:hi].astype(bool), gh["is_extra"][lo:hi].astype(bool)
        ident, cov = gh["prot_ident"][lo:hi], gh["coverage"][lo:hi]
        unexp = ~exp & ~ext

        def fmt(mask) -> bytes:
            out = []
            for i in np.nonzero(mask)[0]:
                parts = [db.gene_names[int(gene[i])].encode("utf-8"), b"%.2f%%" % ident[i], b"%.2f%%" % cov[i]]
                if STATE_SUFFIX[int(state[i])]:
                    parts.append(STATE_SUFFIX[int(state[i])])
                out.append(b",".join(parts))
            return b";".join(out)

        miss = self.missing[int(self.missing_off[a]) : int(self.missing_off[a + 1])]
        n_exp_in, n_exp_out = len(np.unique(gene[ins & exp])), len(np.unique(gene[~ins & exp]))
        total = n_exp_in + n_exp_out + len(miss)
        e_in = b"%d / %d (%.2f%%)" % (n_exp_in, total, n_exp_in / total * 100.0) if total else b"0 / 0 (0.00%)"
        e_out = b"%d / %d (%.2f%%)" % (n_exp_out, total, n_exp_out / total * 100.0) if total else b"0 / 0 (0.00%)"
        ld = self.length_discrepancy[a]
        cols = [
            db.kaptive_version.encode(), db.name.encode(), db.version.encode(), assembly.encode(), self.best_locus_name(a).encode(),
            db.serotypes[int(self.best_locus[a])].encode(), b"Typeable" if self.typeable[a] else b"Untypeable",
            PROBLEM_SYMBOLS[int(self.problems[a])], b"%.2f%%" % self.percent_identity(a), b"%.2f%%" % self.percent_coverage[a],
            b"n/a" if np.isnan(ld) else b"%d" % int(ld), e_in, fmt(ins & exp), b";".join(db.gene_names[int(g)].encode("utf-8") for g in miss),
            b"%d" % len(np.unique(gene[ins & unexp])), fmt(ins & unexp), e_out, fmt(~ins & exp), b"%d" % len(np.unique(gene[~ins & unexp])),
            fmt(~ins & unexp), fmt((state == 2) | (state == 1)), fmt(ext),
        ]  # fmt: skip
        return b"\t".join(cols) + b"\n"

    def rows(self, assemblies: list[str]) -> bytes:
        return b"".join(self.row(a, n) for a, n in enumerate(assemblies))


def score_loci(db: TypingDB, hits: dict[str, np.ndarray], n_asm: int, min_gene_coverage: float = 0.20, threads: int | None = None):
    """core.py:157-207 for a batch: (best locus per assembly, its un-penalised score, final scores, completeness)."""
    L = _lib.load()
    threads = threads or host_threads()
    n = len(hits["gene"])
    scores = np.zeros((n_asm, db.n_loci), np.float64)
    counts = np.zeros((n_asm, db.n_loci), np.float32)
    a = {k: np.ascontiguousarray(hits[k], np.int32) for k in ("asm_id", "gene", "q_start", "q_end", "score")}
    _check(L.kb_type_score(db._h, ptr(a["asm_id"]), ptr(a["gene"]), ptr(a["q_start"]), ptr(a["q_end"]), ptr(a["score"]), n, n_asm,
                           float(min_gene_coverage), threads, ptr(scores), ptr(counts)))
    completeness = counts / db.expected_per_locus            # float32 / float32
    final = scores * (completeness ** 3)                     # float64 * float32 ** 3: the reference's own expression (core.py:200-201)
    best = np.argmax(final, axis=1).astype(np.int32) if n_asm else np.zeros(0, np.int32)
    best_score = scores[np.arange(n_asm), best] if n_asm else np.zeros(0, np.float64)
    return best, np.ascontiguousarray(best_score), final, completeness


def type_many(db: TypingDB, batch, res, max_other_genes: int = 1, min_completeness: float = 0.5, allow_below_threshold: bool = False,
              min_gene_coverage: float = 0.20, partial_edge_tolerance: int = 5, threads: int | None = None) -> TypedBatch:
    """Type every assembly of ``batch`` (a :class:`kaptive_b200.mapper.AssemblyBatch`) from the hits ``res`` of mapping it."""
    L = _lib.load()
    threads = threads or host_threads()
    n_asm = batch.n_asm
    h = res.hits
    best, best_score, _, _ = score_loci(db, h, n_asm, min_gene_coverage, threads)
    i32 = {k: np.ascontiguousarray(h[k], np.int32) for k in ("asm_id", "gene", "q_start", "q_end", "t_ctg", "t_len", "t_start", "t_end", "score", "matches")}
    strand, mapq = np.ascontiguousarray(h["strand"], np.int8), np.ascontiguousarray(h["mapq"], np.uint8)
    r = C.c_void_p(0)
    _check(L.kb_type_call(db._h, batch._h, ptr(i32["asm_id"]), ptr(i32["gene"]), ptr(i32["q_start"]), ptr(i32["q_end"]), ptr(i32["t_ctg"]), ptr(i32["t_len"]),
                          ptr(i32["t_start"]), ptr(i32["t_end"]), ptr(strand), ptr(i32["score"]), ptr(i32["matches"]), ptr(mapq), len(strand), n_asm,
                          ptr(best), ptr(best_score), int(max_other_genes), float(min_completeness), int(allow_below_threshold),
                          int(partial_edge_tolerance), threads, C.byref(r)))
    try:
        ng, npc, nm = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        _check(L.kb_typed_sizes(r, C.byref(ng), C.byref(npc), C.byref(nm)))
        f64 = lambda: np.zeros(max(n_asm, 1), np.float64)  # noqa: E731
        score, comp, pcov, ld = f64(), f64(), f64(), f64()
        typeable, problems = np.zeros(max(n_asm, 1), np.uint8), np.zeros(max(n_asm, 1), np.uint8)
        n_pieces = np.zeros(max(n_asm, 1), np.int32)
        gho, pco, mso = (np.zeros(n_asm + 1, np.int64) for _ in range(3))
        _check(L.kb_typed_fetch_assemblies(r, ptr(score), ptr(comp), ptr(pcov), ptr(ld), ptr(typeable), ptr(problems), ptr(n_pieces), ptr(gho), ptr(pco), ptr(mso)))
        g = {k: np.zeros(max(ng.value, 1), dt) for k, dt in (("gene", np.int32), ("q_start", np.int32), ("q_end", np.int32), ("t_ctg", np.int32),
                                                              ("t_start", np.int32), ("t_end", np.int32), ("strand", np.int8), ("state", np.int8),
                                                              ("is_expected", np.uint8), ("is_inside", np.uint8), ("is_extra", np.uint8),
                                                              ("prot_ident", np.float32), ("coverage", np.float32))}
        _check(L.kb_typed_fetch_gene_hits(r, *(ptr(g[k]) for k in ("gene", "q_start", "q_end", "t_ctg", "t_start", "t_end", "strand", "state", "is_expected",
                                                                    "is_inside", "is_extra", "prot_ident", "coverage"))))
        pc = {k: np.zeros(max(npc.value, 1), dt) for k, dt in (("ctg", np.int32), ("start", np.int32), ("end", np.int32), ("strand", np.int8))}
        missing = np.zeros(max(nm.value, 1), np.int32)
        _check(L.kb_typed_fetch_pieces(r, ptr(pc["ctg"]), ptr(pc["start"]), ptr(pc["end"]), ptr(pc["strand"]), ptr(missing)))
    finally:
        L.kb_typed_destroy(r)
    return TypedBatch(db=db, best_locus=best, best_locus_score=score[:n_asm], completeness=comp[:n_asm], percent_coverage=pcov[:n_asm],
                      length_discrepancy=ld[:n_asm], typeable=typeable[:n_asm].astype(bool), problems=problems[:n_asm], n_pieces=n_pieces[:n_asm],
                      gene_hit_off=gho, piece_off=pco, missing_off=mso, gene_hits={k: v[: ng.value] for k, v in g.items()},
                      pieces={k: v[: npc.value] for k, v in pc.items()}, missing=missing[: nm.value])


def slab_bounds(n_asm: int, first: int = 512, second: int = 1536, big: int = 2560) -> list[int]:
    """The slab plan of the packed host-buffer path (kb_api.cu:map_slabs): a short first slab so that little copy time is exposed,
    then slabs as large as possible."""
    if n_asm < 1024:
        return [0, n_asm]
    b = [0, min(first, n_asm // 4)]
    if n_asm - b[-1] > 2 * second:
        b.append(b[-1] + second)
    base, rest = b[-1], n_asm - b[-1]
    n = (rest + big - 1) // big
    each = (rest + n - 1) // n
    b += [min(n_asm, base + k * each) for k in range(1, n + 1)]
    return b


def type_packed(gi, db: TypingDB, pb, device: int = 0, threads: int | None = None, bounds: list[int] | None = None, **params) -> list[TypedBatch]:
    """End-to-end typed path from host-packed assemblies (``ingest.ingest_fasta_packed``): slabs are copied to the device by a
    producer thread (at most two ahead), mapped and typed in order.  Returns one :class:`TypedBatch` per slab (assembly order)."""
    import queue
    import threading

    from . import ingest, mapper

    bnd = bounds or slab_bounds(len(pb.asm_contig_start) - 1)
    q: queue.Queue = queue.Queue(maxsize=2)

    def produce():
        try:
            for k in range(len(bnd) - 1):
                a0, a1 = bnd[k], bnd[k + 1]
                c0, c1 = int(pb.asm_contig_start[a0]), int(pb.asm_contig_start[a1])
                sub = ingest.PackedBatch(pb.seq2, pb.nmask, pb.contig_len[c0:c1], pb.contig_soff[c0:c1],
                                         np.ascontiguousarray(pb.asm_contig_start[a0 : a1 + 1] - c0), pb.storage_bases, [])
                first = int(pb.contig_soff[c0]) if c1 > c0 else 128
                q.put(mapper.AssemblyBatch.from_packed(sub, device=device, first_soff=first))
        except BaseException as e:  # noqa: BLE001  (handed to the consumer)
            q.put(e)

    from concurrent.futures import ThreadPoolExecutor

    th = threading.Thread(target=produce, daemon=True)
    th.start()
    futs = []
    # three stages in flight: the producer copies slab k + 2, this thread maps slab k + 1, the typing worker types slab k (host
    # threads + one short device pass that shares the GPU with the mapping kernels)
    with ThreadPoolExecutor(1) as typer:
        def type_and_close(b, res):
            try:
                return type_many(db, b, res, threads=threads, **params)
            finally:
                b.close()

        for _ in range(len(bnd) - 1):
            b = q.get()
            if isinstance(b, BaseException):
                raise b
            res = gi.map(b)
            futs.append(typer.submit(type_and_close, b, res))
        out = [f.result() for f in futs]
    th.join()
    return out
