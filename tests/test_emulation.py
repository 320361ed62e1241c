"""CPU tier: the product's device logic (kaptive_b200/csrc/*.cuh, compiled for the host by tests/host_emul) must
agree with the oracle and with the golden vectors on every parity case, for several lane-slice sizes."""

import numpy as np
import pytest

import cases
import emul_lib as el
import oracle_lib as ol

GOLD = np.load(cases.__file__.replace("cases.py", "golden/mapping_golden.npz"))
_idx = {}


def emul_index(db):
    k = id(db)
    if k not in _idx:
        _idx[k] = el.EmulIndex(*db.flat())
    return _idx[k]


def assert_same(r, hits, cigar, chains=None):
    assert len(r["hits"]) == len(hits)
    for f in hits.dtype.names:
        if f != "cigar_off":
            assert np.array_equal(r["hits"][f], hits[f]), f
    for a, b in zip(r["hits"], hits):
        assert np.array_equal(r["cigar"][a["cigar_off"] : a["cigar_off"] + a["n_cigar"]], cigar[b["cigar_off"] : b["cigar_off"] + b["n_cigar"]])
    if chains is not None:
        assert np.array_equal(r["chains"], chains)


@pytest.mark.parametrize("name", list(cases.CASES))
def test_emulation_matches_golden(name):
    db, contigs = cases.CASES[name]()
    r = emul_index(db).map(*cases.flat_contigs(contigs), lane_bases=256, keep_stages=True)
    assert_same(r, GOLD[f"{name}/hits"], GOLD[f"{name}/cigar"], GOLD[f"{name}/chains"])
    assert [r["mid_occ"], r["n_minimizers"], len(r["anchors"])] == list(GOLD[f"{name}/meta"])
    assert r["stage"]["mismatch"] == 0  # staged plan -> jobs -> assemble path == kb_align1 wherever it claims the chain


@pytest.mark.parametrize("lane_bases", [1, 7, 24, 25, 64, 1000, -1, -9, -10, -11, -256, -100000])
def test_scan_slicing_is_exact(lane_bases):
    """Sketching independent slices with a w+k-1 look-back gives exactly the sequential minimizer list
    (negative sizes: the branch-free sliding-window sketch the scan kernel runs)."""
    for name in ("n_rich", "tiny_contigs", "boundaries", "repeat_gene"):
        db, contigs = cases.CASES[name]()
        flat = cases.flat_contigs(contigs)
        r = emul_index(db).map(*flat, lane_bases=lane_bases, keep_stages=True)
        want = []
        for ci, (_, s) in enumerate(contigs):
            x, y = ol.sketch(s)
            want += [(int(a >> 8), ci, int(b)) for a, b in zip(x, y)]
        got = sorted(zip(r["mz_hash"].tolist(), r["mz_ctg"].tolist(), r["mz_pos"].tolist()))
        assert got == sorted(want)


def test_anchors_match_oracle():
    for name in ("mutated0", "repeat_gene", "two_loci"):
        db, contigs = cases.CASES[name]()
        flat = cases.flat_contigs(contigs)
        ro = ol.OracleDB(*db.flat()).map(*flat, keep_stages=True)
        re = emul_index(db).map(*flat, keep_stages=True)
        assert np.array_equal(ro["anchors"], re["anchors"])


def test_fixed_mid_occ_parameter():
    db, contigs = cases.case_repeat_gene()
    flat = cases.flat_contigs(contigs)
    for mid in (3, 14, 15, 1000):
        ro = ol.OracleDB(*db.flat(), params=ol.default_params(mid_occ=mid)).map(*flat)
        re = el.EmulIndex(*db.flat(), params=el.default_params(mid_occ=mid)).map(*flat)
        assert ro["mid_occ"] == re["mid_occ"] == mid
        assert_same(re, ro["hits"], ro["cigar"])


def test_random_cases_against_oracle():
    from kaptive_b200 import synth

    db = cases.small_db(seed=5, n_extra=2)
    odb, edb = ol.OracleDB(*db.flat()), el.EmulIndex(*db.flat())
    for s in range(6):
        a = synth.make_assembly(db, s, seed=9000 + s, genome_len=120_000, mean_contigs=5 + 10 * (s % 3),
                                sub=(0, 0.1), indel=(0, 0.01), n_frac=2e-4, lowercase_frac=0.1 * (s % 2))
        ro = odb.map(*a.flat(), keep_stages=True)
        re = edb.map(*a.flat(), lane_bases=(256, 100, 33)[s % 3], keep_stages=True)
        assert_same(re, ro["hits"], ro["cigar"], ro["chains"])
        assert re["stage"]["mismatch"] == 0 and re["stage"]["ok"] > 0


def test_staged_alignment_equals_monolith_on_divergence_ladder():
    """kb_stage.cuh (plan, per-job DP, assemble) against kb_align1 chain by chain, 0-14 % substitutions / 0-4 % indels."""
    from kaptive_b200 import synth

    db = synth.make_db(n_loci=12, genes_per_locus=10, n_core=3, seed=11)
    edb = el.EmulIndex(*db.flat())
    ok = fb = 0
    for i in range(0, 28, 3):
        s = 0.005 * i
        a = synth.make_assembly(db, i % 12, seed=7000 + i, genome_len=200_000, mean_contigs=4, sub=(s, s + 0.005),
                                indel=(0.0015 * i, 0.0015 * i + 0.001))
        r = edb.map(*a.flat())
        assert r["stage"]["mismatch"] == 0
        ok, fb = ok + r["stage"]["ok"], fb + r["stage"]["fallback"]
    assert ok > 100 and ok > 5 * fb


def test_fast_sketch_on_low_complexity_sequence():
    """Homopolymers, short tandem repeats and N runs: the identical-k-mer paths of mm_sketch."""
    import numpy as np

    rng = np.random.default_rng(4)
    parts = [b"A" * 300, b"AT" * 200, b"ACG" * 150, b"N" * 40, b"ACGTTGCA" * 60, b"AAAAAAAAAAAAAAAC" * 30, b"N", b"GATTACA" * 90]
    for _ in range(10):
        parts.append(bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=int(rng.integers(1, 400)), p=[.24, .24, .24, .24, .04])))
        parts.append(bytes(rng.choice(np.frombuffer(b"AC", np.uint8), size=int(rng.integers(1, 200)))))
    seq = b"".join(parts)
    db = cases.small_db()
    contigs = [("lc", seq), ("lc2", seq[::-1]), ("h", b"C" * 5000)]
    flat = cases.flat_contigs(contigs)
    want = []
    for ci, (_, s) in enumerate(contigs):
        x, y = ol.sketch(s)
        want += [(int(a >> 8), ci, int(b)) for a, b in zip(x, y)]
    for lb in (-256, -17, 256):
        r = emul_index(db).map(*flat, lane_bases=lb, keep_stages=True)
        assert sorted(zip(r["mz_hash"].tolist(), r["mz_ctg"].tolist(), r["mz_pos"].tolist())) == sorted(want), lb
