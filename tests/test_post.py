"""Post-mapping stages: the oracle restatement against vectors produced by the
reference's own numba kernels (tests/golden/post_golden.npz), and -- on the GPU tier -- the CUDA kernels against both."""

from pathlib import Path

import numpy as np
import pytest

import post_oracle_lib as po

G = np.load(Path(__file__).resolve().parent / "golden" / "post_golden.npz")
GOTOH_FIELDS = ("scores", "matches", "mismatches", "gaps", "q_starts", "q_ends", "t_starts", "t_ends")


def gotoh_golden():
    return np.stack([G[f"gotoh/{f}"] for f in GOTOH_FIELDS], axis=1).astype(np.int32)


def test_oracle_protein_align_matches_reference_kernel():
    res = po.protein_align(G["gotoh/q"], G["gotoh/q_len"], G["gotoh/t"], G["gotoh/t_len"])
    assert np.array_equal(res, gotoh_golden())
    # known answer observed in SURVEY.md section 8c
    assert list(res[-1][:4]) == [37, 11, 1, 1]
    tot = res[:, 1] + res[:, 2] + res[:, 3]
    pid = np.divide(res[:, 1] * 100.0, tot, out=np.zeros(len(res)), where=tot > 0)
    assert np.allclose(pid, G["gotoh/pidents"], atol=0.0)  # identical integers -> identical float64 identity


def test_oracle_extract_and_translate_match_reference_kernels():
    plen = G["ext/contig_len"]
    poff = po.offsets_of(plen)
    out, oo, ol_ = po.extract(G["ext/contigs"], poff, G["ext/ci"], G["ext/st"], G["ext/en"], G["ext/sd"])
    assert np.array_equal(ol_, G["ext/out_len"]) and np.array_equal(out, G["ext/out"])
    for key, stop in (("tr", True), ("tr2", False)):
        tr, _, tl = po.translate(out, oo, ol_, G["ext/frames"], stop)
        assert np.array_equal(tl, G[f"{key}/out_len"]) and np.array_equal(tr, G[f"{key}/out"])


def test_oracle_edge_cases():
    # empty inputs, frame beyond the sequence, unknown residues
    tr, _, tl = po.translate(np.frombuffer(b"ATGAAATAGCC", np.uint8), np.array([0, 3, 11]), np.array([11, 2, 0]), np.array([0, 2, 1]), True)
    assert bytes(tr) == b"MK" and list(tl) == [2, 0, 0]
    res = po.protein_align(np.frombuffer(b"MKTZZ", np.uint8), np.array([5, 0]), np.frombuffer(b"MKT", np.uint8), np.array([3, 0]))
    assert list(res[1]) == [0] * 8 and res[0][0] == 15 and res[0][1] == 3  # M, K, T on the BLOSUM62 diagonal: 5 + 5 + 5


@pytest.mark.gpu
def test_gpu_post_kernels_match_golden_and_oracle():
    from kaptive_b200 import post

    res = post.protein_align(G["gotoh/q"], G["gotoh/q_len"], G["gotoh/t"], G["gotoh/t_len"])
    assert np.array_equal(res, gotoh_golden())
    plen = G["ext/contig_len"]
    out, oo, ol_ = post.extract(G["ext/contigs"], po.offsets_of(plen), G["ext/ci"], G["ext/st"], G["ext/en"], G["ext/sd"])
    assert np.array_equal(ol_, G["ext/out_len"]) and np.array_equal(out, G["ext/out"])
    for key, stop in (("tr", True), ("tr2", False)):
        tr, _, tl = post.translate(out, oo, ol_, G["ext/frames"], stop)
        assert np.array_equal(tl, G[f"{key}/out_len"]) and np.array_equal(tr, G[f"{key}/out"])
    # larger random batch against the oracle (sizes the reference sees per batch of assemblies)
    rng = np.random.default_rng(9)
    aa = np.frombuffer(b"ARNDCQEGHILKMFPSTWYVX*", np.uint8)
    qs, ts = [], []
    for _ in range(3000):
        n = int(rng.integers(0, 500))
        t = aa[rng.integers(0, len(aa), size=n)]
        q = t.copy()
        m = rng.random(n) < rng.uniform(0, 0.4)
        q[m] = aa[rng.integers(0, len(aa), size=int(m.sum()))]
        if n > 20 and rng.random() < 0.5:
            c = int(rng.integers(1, n - 1))
            q = np.concatenate([q[:c], q[c + int(rng.integers(1, 40)):]])
        qs.append(q), ts.append(t)
    ql, tl = np.array([len(x) for x in qs], np.int32), np.array([len(x) for x in ts], np.int32)
    Q, T = np.concatenate(qs), np.concatenate(ts)
    assert np.array_equal(post.protein_align(Q, ql, T, tl), po.protein_align(Q, ql, T, tl))


# ---------------------------------------------------------------------------------------------- row a8: cull + cluster
A8 = np.load(Path(__file__).resolve().parent / "golden" / "a8_golden.npz")


def _a8_segments():
    off = A8["seg_off"]
    for s in range(len(off) - 1):
        yield s, slice(int(off[s]), int(off[s + 1]))


def test_oracle_cull_and_cluster_match_reference_kernels():
    """oracle/kb_post_oracle.c vs vectors of the reference's _cull_overlaps_kernel / _cluster_kernel (make_golden_a8.py);
    max_overlap_fraction and tolerance differ per segment, so the oracle is called one segment at a time."""
    for s, sl in _a8_segments():
        so = np.array([0, sl.stop - sl.start], np.int64)
        kept = po.cull_overlaps(A8["order_cull"][sl], A8["g1"][sl], A8["g2"][sl], A8["st"][sl], A8["en"][sl], A8["frac"][s], so)
        assert np.array_equal(kept, A8["kept"][sl]), s
        ids = po.cluster(A8["st"][sl], A8["en"][sl], A8["g1"][sl], A8["tol"][s], A8["order_cl"][sl], so)
        assert np.array_equal(ids, A8["cl"][sl]), s


def test_oracle_cull_batched_equals_per_segment():
    same = np.nonzero(A8["frac"] == 0.1)[0]
    off = A8["seg_off"]
    parts = [slice(int(off[s]), int(off[s + 1])) for s in same]
    cat = lambda k: np.concatenate([A8[k][p] for p in parts])  # noqa: E731
    so = np.concatenate([[0], np.cumsum([p.stop - p.start for p in parts])]).astype(np.int64)
    kept = po.cull_overlaps(cat("order_cull"), cat("g1"), cat("g2"), cat("st"), cat("en"), 0.1, so)
    assert np.array_equal(kept, cat("kept"))


def test_cull_order_is_the_reference_lexsort_with_wrapped_uint8_mapq():
    """kaptive_b200.post.cull_order vs the order the reference's own Alignments.cull_overlaps expression produced
    (core/alignment.py:675 negates a uint8 array: mapq 0 sorts first among hits that tie on score and matches; the golden
    set holds overlapping primary / secondary pairs in exactly that situation)."""
    from kaptive_b200 import post

    n_tied = 0
    for s, sl in _a8_segments():
        so = np.array([0, sl.stop - sl.start], np.int64)
        order = post.cull_order(A8["score"][sl], A8["matches"][sl], A8["mapq"][sl], so)
        assert np.array_equal(order, A8["order_cull"][sl]), s
        if sl.stop - sl.start > 8:
            pos = {int(v): i for i, v in enumerate(order)}
            assert pos[5] < pos[4] and pos[6] < pos[7]  # the mapq-0 hit of each tied pair is evaluated first
            n_tied += 1
    assert n_tied > 20


@pytest.mark.gpu
def test_gpu_cull_and_cluster_match_reference_kernels():
    """CUDA kernels through kaptive_b200.post (which also derives the evaluation order the way the reference does) against
    the reference's own outputs, batched over all segments that share a parameter value."""
    from kaptive_b200 import post

    off = A8["seg_off"]
    for frac in np.unique(A8["frac"]):
        segs = np.nonzero(A8["frac"] == frac)[0]
        parts = [slice(int(off[s]), int(off[s + 1])) for s in segs]
        cat = lambda k: np.concatenate([A8[k][p] for p in parts])  # noqa: E731
        so = np.concatenate([[0], np.cumsum([p.stop - p.start for p in parts])]).astype(np.int64)
        kept = post.cull_overlaps(cat("st"), cat("en"), cat("g1"), cat("g2"), cat("score"), cat("matches"), cat("mapq"), so, float(frac))
        assert np.array_equal(kept, cat("kept").astype(bool))
    for tol in np.unique(A8["tol"]):
        segs = np.nonzero(A8["tol"] == tol)[0]
        parts = [slice(int(off[s]), int(off[s + 1])) for s in segs]
        cat = lambda k: np.concatenate([A8[k][p] for p in parts])  # noqa: E731
        so = np.concatenate([[0], np.cumsum([p.stop - p.start for p in parts])]).astype(np.int64)
        ids = post.cluster_spatial(cat("st"), cat("en"), cat("g1"), so, int(tol))
        assert np.array_equal(ids, cat("cl"))
    # empty batch
    assert len(post.cull_overlaps([], [], [], [], [], [], [], [0], 0.1)) == 0
