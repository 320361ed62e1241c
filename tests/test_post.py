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
