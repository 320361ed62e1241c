"""type_many (kaptive_b200.serotype) against the UNMODIFIED reference pipeline: tests/golden/typing_golden.json holds what
kaptive.serotyping.Serotyper.__call__ + KaptiveRow.from_result produced for every assembly of tests/typing_cases.py (script:
tests/golden/make_golden_typing.py).  CPU tier: locus scoring on the oracle's hits.  GPU tier: the whole call -- mapping on the
device, reconstruction, translated hits and protein alignments from the packed batch, gene states, confidence, the TSV row."""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as ol
import typing_cases as tc
from kaptive_b200 import synth

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "typing_golden.json").read_text())


def _translations_py(db):
    code = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
    idx = {c: i for i, c in enumerate("TCAG")}
    out = []
    for g in db.genes:
        s = g.decode().upper()
        out.append("".join(code[idx[s[i]] * 16 + idx[s[i + 1]] * 4 + idx[s[i + 2]]] for i in range(0, len(s) - 2, 3)).encode())
    return out


def test_locus_scoring_equals_reference_on_oracle_hits():
    """kb_type_score + the numpy finish (serotyping/core.py:157-207): best locus and its un-penalised score for every assembly."""
    from kaptive_b200 import serotype

    db, n_k, n_o = tc.make_db()
    asms = tc.make_assemblies(db, n_k, n_o)
    tdb = serotype.TypingDB(np.array([len(g) for g in db.genes]), db.gene_locus, db.extra, db.gene_pos, db.gene_strand,
                            [len(s) for s in db.loci], _translations_py(db), db.locus_names, db.gene_names, device=-1)
    odb = ol.OracleDB(*db.flat())
    parts = []
    for ai, a in enumerate(asms):
        h = odb.map(*a.flat())["hits"]
        parts.append({"asm_id": np.full(len(h), ai, np.int32), **{k: h[k] for k in ("gene", "q_start", "q_end", "score")}})
    hits = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    best, best_score, _, _ = serotype.score_loci(tdb, hits, len(asms), threads=3)
    for ai, a in enumerate(asms):
        g = GOLD[a.name]
        assert int(best[ai]) == g["best_locus_idx"], a.name
        assert best_score[ai] == g["best_locus_score"], a.name  # bit-identical float64 (same sums in the same order)


@pytest.mark.gpu
def test_type_many_equals_the_unmodified_reference_pipeline():
    from kaptive_b200 import mapper, serotype

    db, n_k, n_o = tc.make_db()
    asms = tc.make_assemblies(db, n_k, n_o)
    gi = mapper.GeneIndex(db.genes)
    batch = mapper.AssemblyBatch.from_contigs([[s for _, s in a.contigs] for a in asms])
    res = gi.map(batch)
    tdb = serotype.TypingDB.from_synth(db)
    tdb.kaptive_version = "unknown"  # what kaptive.__version__ says when the reference runs from its source tree (the golden run)
    assert [t for t in _translations_py(db)] == [bytes(x) for x in _trans_of(tdb, db)]
    typed = serotype.type_many(tdb, batch, res, threads=4)
    assert len(typed) == len(asms)
    n_rows = 0
    for ai, a in enumerate(asms):
        g = GOLD[a.name]
        assert int(typed.best_locus[ai]) == g["best_locus_idx"], a.name
        assert typed.best_locus_score[ai] == g["best_locus_score"] and typed.completeness[ai] == g["completeness"], a.name
        assert bool(typed.typeable[ai]) == g["typeable"] and int(typed.problems[ai]) == g["problems"], a.name
        assert typed.percent_coverage[ai] == g["percent_coverage"], a.name
        ld = typed.length_discrepancy[ai]
        assert (g["length_discrepancy"] is None and np.isnan(ld)) or ld == g["length_discrepancy"], a.name
        lo, hi = int(typed.gene_hit_off[ai]), int(typed.gene_hit_off[ai + 1])
        for k in ("gene", "t_ctg", "t_start", "t_end", "strand", "is_expected", "is_inside", "is_extra", "state"):
            assert typed.gene_hits[k][lo:hi].astype(np.int64).tolist() == g[k], (a.name, k)
        assert typed.gene_hits["prot_ident"][lo:hi].tolist() == [float(np.float32(x)) for x in g["prot_ident"]], a.name
        assert typed.gene_hits["coverage"][lo:hi].tolist() == [float(np.float32(x)) for x in g["coverage"]], a.name
        plo, phi = int(typed.piece_off[ai]), int(typed.piece_off[ai + 1])
        for k, gk in (("ctg", "piece_ctg"), ("start", "piece_start"), ("end", "piece_end"), ("strand", "piece_strand")):
            assert typed.pieces[k][plo:phi].astype(np.int64).tolist() == g[gk], (a.name, gk)
        miss = typed.missing[int(typed.missing_off[ai]) : int(typed.missing_off[ai + 1])]
        assert [db.gene_names[int(x)] for x in miss] == g["missing"], a.name
        assert typed.percent_identity(ai) == g["percent_identity"], a.name
        assert typed.row(ai, a.name).decode() == g["row"], a.name
        n_rows += 1
    assert n_rows == 48 and sum(1 for a in asms if GOLD[a.name]["typeable"]) > 30


@pytest.mark.gpu
def test_type_packed_from_fasta_bytes_in_slabs_equals_type_many():
    """FASTA bytes -> host-packed buffers -> type_packed (three slabs, producer thread) gives the calls of one type_many over the batch."""
    import cases
    from kaptive_b200 import ingest, mapper, serotype

    db, n_k, n_o = tc.make_db()
    asms = tc.make_assemblies(db, n_k, n_o)[:20]
    gi = mapper.GeneIndex(db.genes)
    tdb = serotype.TypingDB.from_synth(db)
    tdb.kaptive_version = "unknown"
    pb = ingest.ingest_fasta_packed([cases.fasta_bytes(a.contigs) for a in asms], threads=3)
    parts = serotype.type_packed(gi, tdb, pb, bounds=[0, 3, 11, 20], threads=2)
    assert [len(p) for p in parts] == [3, 8, 9]
    k = 0
    for p in parts:
        for a in range(len(p)):
            g = GOLD[asms[k].name]
            assert p.row(a, asms[k].name).decode() == g["row"], asms[k].name
            k += 1


def _trans_of(tdb, db):
    from kaptive_b200 import post

    lens = np.array([len(g) for g in db.genes], np.int32)
    off = np.concatenate([[0], np.cumsum(lens[:-1])]).astype(np.int64)
    aa, ao, al = post.translate(np.frombuffer(b"".join(db.genes), np.uint8), off, lens, np.zeros(len(lens), np.int8), to_stop=False)
    return [aa[int(o) : int(o) + int(n)].tobytes() for o, n in zip(ao, al)]


@pytest.mark.gpu
def test_warp_gotoh_kernels_equal_the_thread_kernel(monkeypatch):
    """The protein DP of type_many runs one warp per pair (row state in registers for narrow bands, in shared memory for the wide
    bands of frameshifted hits); KAPTIVE_B200_GOTOH_WARP=0 sends every pair through the thread-per-pair kernel, which is the one the
    reference-pinned tests above were first written against.  Same scores and counts, so the same identities and gene states."""
    from kaptive_b200 import mapper, serotype, workload

    db, ranges = synth.make_ko_db(k_loci=12, k_genes=14, k_core=2, seed=3)
    wl = workload.make_device_workload(db, 40, 400_000, mean_contigs=5, seed=21, device="cuda:0", locus_ranges=ranges, indel=(0.0, 0.01))
    gi = mapper.GeneIndex(db.genes)
    batch = mapper.AssemblyBatch(wl.ascii.data_ptr(), wl.contig_off, wl.contig_len, wl.asm_contig_start, device=0)
    res = gi.map(batch)
    tdb = serotype.TypingDB.from_synth(db)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("KAPTIVE_B200_GOTOH_WARP", mode)
        t = serotype.type_many(tdb, batch, res, threads=2)
        out[mode] = ({k: np.array(v) for k, v in t.gene_hits.items()}, np.array(t.typeable), np.array(t.problems), np.array(t.best_locus))
    a, b = out["0"], out["1"]
    assert len(a[0]["gene"]) > 500
    wide = np.abs(a[0]["t_end"] - a[0]["t_start"]).min() < 300  # truncated hits: wide bands
    assert wide
    for k in a[0]:
        assert np.array_equal(a[0][k], b[0][k], equal_nan=True) if a[0][k].dtype.kind == "f" else np.array_equal(a[0][k], b[0][k]), k
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
