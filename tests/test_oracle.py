"""CPU tier: pin the oracle (known answers, an independent pure-Python restatement of mm_sketch, the committed
golden vectors) so that 'matches the oracle' means something."""

import math

import numpy as np
import pytest

import cases
import oracle_lib as ol
from kaptive_b200 import synth

GOLD = np.load(cases.__file__.replace("cases.py", "golden/mapping_golden.npz"))


# ------------------------------------------------------------------ minimap2 hash / sketch known answers
def hash64_mask(key: int, mask: int) -> int:
    """minimap2 sketch.c hash64 in arbitrary-precision Python (the published invertible integer hash)."""
    M64 = (1 << 64) - 1
    key = (~key + (key << 21)) & M64 & mask
    key = key ^ key >> 24
    key = ((key + (key << 3)) + (key << 8)) & mask
    key = key ^ key >> 14
    key = ((key + (key << 2)) + (key << 4)) & mask
    key = key ^ key >> 28
    key = (key + (key << 31)) & mask
    return key


def py_sketch(seq: bytes, w: int = 10, k: int = 15):
    """Line-by-line Python port of minimap2's mm_sketch (sketch.c), used only here as a second opinion."""
    nt = {65: 0, 67: 1, 71: 2, 84: 3, 97: 0, 99: 1, 103: 2, 116: 3, 85: 3, 117: 3}
    MAX = (1 << 64) - 1
    shift1, mask = 2 * (k - 1), (1 << 2 * k) - 1
    kmer = [0, 0]
    buf = [(MAX, MAX)] * w
    mn = (MAX, MAX)
    out = []
    l = buf_pos = min_pos = 0
    for i, ch in enumerate(seq):
        c = nt.get(ch, 4)
        info = (MAX, MAX)
        if c < 4:
            kmer_span = min(l + 1, k)
            kmer[0] = (kmer[0] << 2 | c) & mask
            kmer[1] = (kmer[1] >> 2) | (3 ^ c) << shift1
            if kmer[0] == kmer[1]:
                continue
            z = 0 if kmer[0] < kmer[1] else 1
            l += 1
            if l >= k:
                info = (hash64_mask(kmer[z], mask) << 8 | kmer_span, i << 1 | z)
        else:
            l = 0
        buf[buf_pos] = info
        if l == w + k - 1 and mn[0] != MAX:
            for j in list(range(buf_pos + 1, w)) + list(range(0, buf_pos)):
                if mn[0] == buf[j][0] and buf[j][1] != mn[1]:
                    out.append(buf[j])
        if info[0] <= mn[0]:
            if l >= w + k and mn[0] != MAX:
                out.append(mn)
            mn, min_pos = info, buf_pos
        elif buf_pos == min_pos:
            if l >= w + k - 1 and mn[0] != MAX:
                out.append(mn)
            mn = (MAX, MAX)
            for j in list(range(buf_pos + 1, w)) + list(range(0, buf_pos + 1)):
                if mn[0] >= buf[j][0]:
                    mn, min_pos = buf[j], j
            if l >= w + k - 1 and mn[0] != MAX:
                for j in list(range(buf_pos + 1, w)) + list(range(0, buf_pos + 1)):
                    if mn[0] == buf[j][0] and mn[1] != buf[j][1]:
                        out.append(buf[j])
        buf_pos = (buf_pos + 1) % w
    if mn[0] != MAX:
        out.append(mn)
    return out


def test_hash32_equals_minimap2_hash64_on_30_bits():
    rng = np.random.default_rng(0)
    mask = (1 << 30) - 1
    for key in [0, 1, mask, 0x2AAAAAAA] + [int(x) for x in rng.integers(0, mask, size=200)]:
        assert ol.lib().kbo_hash32(key, mask) == hash64_mask(key, mask)
    # invertible hash: a bijection on 30 bits
    keys = rng.integers(0, mask, size=5000)
    hs = {ol.lib().kbo_hash32(int(x), mask) for x in set(int(x) for x in keys)}
    assert len(hs) == len(set(int(x) for x in keys))


@pytest.mark.parametrize("seq", [
    b"ACGT" * 40,
    b"A" * 100,
    b"ACGTTGCATGCATGCANNNNACGATCGATCGATCGATCGACTGACTAGCTAGCTAGCATCGATCGATCAGCTAGCTAGCTAGCATCGACNGT",
    b"acgtacgatcgatcgatcgatcgactagctagctagctagctagcatcgactagcatcgactacgactacgacatcgactacgatcagcatcgactacg",
    b"ACGTACGTAC",
    b"",
    b"ATATATATATATATATATATATATATATATATATATATATATATATATATATATATAT",
])
def test_sketch_matches_python_port(seq):
    x, y = ol.sketch(seq)
    ref = py_sketch(seq)
    assert [(int(a), int(b)) for a, b in zip(x, y)] == ref


def test_sketch_matches_python_port_random():
    rng = np.random.default_rng(3)
    for n in (30, 200, 1500):
        s = synth.random_dna(rng, n)
        s[rng.random(n) < 0.01] = ord("N")
        x, y = ol.sketch(s.tobytes())
        assert [(int(a), int(b)) for a, b in zip(x, y)] == py_sketch(s.tobytes())


def test_deterministic_math():
    L = ol.lib()
    for v in (2.0, 3.0, 7.5, 100.0, 12345.0):
        assert abs(L.kbo_log2_fast(v) - math.log2(v)) < 0.09  # minimap2's mg_log2 is a coarse approximation by design
    for v in (0.5, 1.0, 1.5, 2.0, 40.0, 1023.0, 1e6):
        got = L.kbo_logf(v)
        want = np.float32(math.log(v))
        assert abs(got - want) <= 2 * np.spacing(np.float32(abs(want)) if want else np.float32(1e-7))


# ------------------------------------------------------------------ alignment known answers
def _map_single(gene: bytes, target: bytes):
    db = ol.OracleDB(*cases.flat_contigs([("g", gene)]))
    return db.map(*cases.flat_contigs([("t", target)]))


def test_exact_gene_known_answer():
    rng = np.random.default_rng(11)
    g = synth.random_orf(rng, 300).tobytes()
    bg = synth.random_dna(rng, 3000).tobytes()
    r = _map_single(g, bg[:1500] + g + bg[1500:])
    assert len(r["hits"]) == 1
    h = r["hits"][0]
    assert (h["q_start"], h["q_end"], h["t_start"], h["t_end"], h["strand"]) == (0, 900, 1500, 2400, 1)
    assert h["matches"] == 900 and h["block_len"] == 900 and h["edit_distance"] == 0
    assert h["score"] == 2 * 900  # match score a = 2
    assert ol.cigar_string(r["cigar"]) == "900M"
    assert h["mapq"] == 60 and h["is_primary"] == 1


def test_reverse_strand_and_single_edits_known_answer():
    rng = np.random.default_rng(12)
    g = synth.random_orf(rng, 300)
    t = g.copy()
    t[450] = ord("A") if t[450] != ord("A") else ord("C")  # one substitution
    t = np.concatenate([t[:600], t[603:]])  # 3-base deletion in the target = 3I in the query's CIGAR
    bg = synth.random_dna(rng, 2000)
    target = np.concatenate([bg[:1000], synth.revcomp(t), bg[1000:]]).tobytes()
    r = _map_single(g.tobytes(), target)
    assert len(r["hits"]) == 1
    h = r["hits"][0]
    assert h["strand"] == -1 and (h["q_start"], h["q_end"]) == (0, 900)
    assert (h["t_start"], h["t_end"]) == (1000, 1897)
    assert h["matches"] == 896 and h["block_len"] == 900 and h["edit_distance"] == 4
    assert h["score"] == 2 * 896 - 4 - (4 + 2 * 3)  # 1 mismatch (b=4), one gap of 3: q + 3e
    cg = ol.cigar_string(r["cigar"])
    assert cg.count("I") == 1 and "3I" in cg and "D" not in cg


def test_no_hits_cases():
    db, contigs = cases.case_empty_assembly()
    r = ol.OracleDB(*db.flat()).map(*cases.flat_contigs(contigs))
    assert len(r["hits"]) == 0
    db, contigs = cases.case_no_locus()
    r = ol.OracleDB(*db.flat()).map(*cases.flat_contigs(contigs))
    assert len(r["hits"]) == 0


def test_zdrop_split_and_repeat_filter_are_exercised():
    db, contigs = cases.case_mosaic_gene()
    r = ol.OracleDB(*db.flat()).map(*cases.flat_contigs(contigs), keep_stages=True)
    assert len(r["hits"]) == 2 * len(r["chains"])  # every chain was split in two by the z-drop
    db, contigs = cases.case_repeat_gene()
    r = ol.OracleDB(*db.flat()).map(*cases.flat_contigs(contigs), keep_stages=True)
    assert r["mid_occ"] == 10
    g20 = r["hits"][r["hits"]["gene"] == 20]
    assert len(g20) >= 1 and g20["matches"].max() == len(db.genes[20])  # still found end to end without the repeat seeds


def test_is_interrupted_gene_is_reported_as_two_flanks():
    """Pins a DOCUMENTED deviation of mapping spec v1 from minimap2: without long-join / RMQ re-chaining a gene
    interrupted by a 2 kb insertion comes out as two hits, one per flank, where minimap2 (bw_long 20000) joins them into one hit with
    a 2 kb gap.  The flanks meet at the insertion point on the gene and lie 2 kb apart on the contig."""
    db, contigs = cases.case_is_insertion()
    r = ol.OracleDB(*db.flat()).map(*cases.flat_contigs(contigs))
    for gi in cases.IS_GENES:
        h = np.sort(r["hits"][(r["hits"]["gene"] == gi) & (r["hits"]["is_primary"] == 1)], order="q_start")
        cut = len(db.genes[gi]) * 45 // 100
        flanks = h[(h["q_end"] - h["q_start"]) > 200]
        assert len(flanks) == 2, (gi, h[["q_start", "q_end", "t_start", "t_end"]])
        a, b = flanks
        assert a["q_start"] == 0 and b["q_end"] == len(db.genes[gi]) and abs(int(a["q_end"]) - cut) <= 30 and abs(int(b["q_start"]) - cut) <= 30
        assert a["t_ctg"] == b["t_ctg"] and a["strand"] == b["strand"]
        gap = int(b["t_start"]) - int(a["t_end"]) if a["strand"] >= 0 else int(a["t_start"]) - int(b["t_end"])
        assert 1940 <= gap <= 2060, gap


# ------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("name", list(cases.CASES))
def test_oracle_reproduces_golden(name):
    db, contigs = cases.CASES[name]()
    r = ol.OracleDB(*db.flat()).map(*cases.flat_contigs(contigs), keep_stages=True)
    assert np.array_equal(r["hits"], GOLD[f"{name}/hits"])
    assert np.array_equal(r["cigar"], GOLD[f"{name}/cigar"])
    assert np.array_equal(r["chains"], GOLD[f"{name}/chains"])
    assert [r["mid_occ"], r["n_minimizers"], len(r["anchors"])] == list(GOLD[f"{name}/meta"])
