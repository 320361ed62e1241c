#!/usr/bin/env python
"""Golden vectors for SURVEY.md section 8 row a8 (overlap cull + spatial clustering).  Run in the BUILD container only:

    PYTHONPATH=/root/reference/src NUMBA_CACHE_DIR=/tmp/numba python tests/golden/make_golden_a8.py

Inputs are random hit tables shaped like one assembly's alignments (10..600 hits on a few contigs / genes); outputs
come from the reference's own numba kernels, driven exactly as the reference drives them:
  cull    : order = np.lexsort((-mapq, -matches, -scores)), mapq uint8 so the negation wraps (core/alignment.py:466,669-675)
            _cull_overlaps_kernel(order, group1, group2, starts, ends, frac, n)   (core/interval.py:698-751)
  cluster : order = np.lexsort((ends, starts, groups))                 (core/interval.py:492)
            _cluster_kernel(starts, ends, groups, tolerance, order)    (core/interval.py:595-639)
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def main():
    from kaptive.core.interval import _cluster_kernel, _cull_overlaps_kernel

    rng = np.random.default_rng(8)
    out = {}
    segs = []
    for s in range(40):
        n = int(rng.integers(0, 12)) if s % 7 == 0 else int(rng.integers(10, 600))
        g1 = rng.integers(0, max(1, n // 20 + 1), size=n).astype(np.int32)      # gene / contig ids: many hits per group
        g2 = rng.integers(0, 2, size=n).astype(np.int32)
        st = rng.integers(0, 5000, size=n).astype(np.int32)
        ln = rng.integers(-2, 900, size=n).astype(np.int32)                      # a few empty / negative intervals
        en = (st + ln).astype(np.int32)
        score = rng.integers(50, 2500, size=n).astype(np.float64)
        score[rng.random(n) < 0.1] += 1e9                                        # priority boost of best-locus genes
        matches = rng.integers(20, 1200, size=n).astype(np.int32)
        mapq = rng.integers(0, 61, size=n).astype(np.uint8)
        if n > 4:  # exact ties in every key
            score[1], matches[1], mapq[1] = score[0], matches[0], mapq[0]
            score[3] = score[2]
        if n > 8:  # overlapping duplicate hits that tie on score and matches: a primary (mapq > 0) against a secondary (mapq 0)
            for a, b in ((4, 5), (6, 7)):
                g1[b], g2[b], st[b], en[b] = g1[a], g2[a], st[a] + 3, en[a] + 3
                score[b], matches[b] = score[a], matches[a]
            mapq[4], mapq[5], mapq[6], mapq[7] = 37, 0, 0, 12
        order_cull = np.lexsort((-mapq, -matches, -score)).astype(np.int32)  # uint8 negation, exactly as Alignments.cull_overlaps
        frac = float(rng.choice([0.1, 0.0, 0.5]))
        kept = _cull_overlaps_kernel(order_cull, g1, g2, st, en, frac, n) if n else np.zeros(0, bool)
        tol = int(rng.choice([0, 100, 30000]))
        order_cl = np.lexsort((en, st, g1)).astype(np.int32)
        cl = _cluster_kernel(st, en, g1, tol, order_cl) if n else np.zeros(0, np.int32)
        segs.append(dict(g1=g1, g2=g2, st=st, en=en, score=score, matches=matches, mapq=mapq, order_cull=order_cull, frac=frac,
                         kept=np.asarray(kept, dtype=np.uint8), tol=tol, order_cl=order_cl, cl=np.asarray(cl, np.int32)))
    for k in ("g1", "g2", "st", "en", "score", "matches", "mapq", "order_cull", "kept", "order_cl", "cl"):
        out[k] = np.concatenate([s[k] for s in segs])
    out["seg_off"] = np.concatenate([[0], np.cumsum([len(s["g1"]) for s in segs])]).astype(np.int64)
    out["frac"] = np.array([s["frac"] for s in segs], np.float64)
    out["tol"] = np.array([s["tol"] for s in segs], np.int32)
    np.savez_compressed(HERE / "a8_golden.npz", **out)
    print("segments", len(segs), "hits", len(out["g1"]), "kept", int(out["kept"].sum()), "clusters", int(sum(s["cl"].max() + 1 for s in segs if len(s["cl"]))))


if __name__ == "__main__":
    sys.exit(main())
