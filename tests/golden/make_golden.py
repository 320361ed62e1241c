#!/usr/bin/env python
"""Generates the committed golden vectors.  Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden.py

For every parity case in tests/cases.py it stores
  * mapping_golden.npz  : the oracle's hits / CIGARs / chains / anchor digest / mid_occ
  * kaptive_rows.json   : what the UNMODIFIED reference pipeline (kaptive.serotyping.Serotyper +
                          KaptiveRow.from_result, /root/reference/src) reports when its `rammappy` import
                          resolves to the oracle-backed shim (tests/ref_shim): the TSV row, best locus,
                          typeable flag.  Any mapper that reproduces the golden hits reproduces these rows.
  * post_golden.npz     : inputs/outputs of the reference's own numba kernels for the post-mapping stages
                          (_batched_banded_gotoh pairwise.py:395, extract seq.py:612, translate seq.py:671)
"""
import hashlib
import io
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import cases  # noqa: E402
import oracle_lib as ol  # noqa: E402
import ref_bridge as rb  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:32]


def main():
    assert rb.reference_available(), "the reference is needed to generate goldens"
    rb.import_reference("oracle")
    from kaptive.core.genome import GenomeAssembly
    from kaptive.serotyping import Serotyper
    from kaptive.serotyping.io import KaptiveRow

    npz, rows = {}, {}
    serotypers = {}
    for name, fn in cases.CASES.items():
        db, contigs = fn()
        odb = ol.OracleDB(*db.flat())
        r = odb.map(*cases.flat_contigs(contigs), keep_stages=True)
        npz[f"{name}/hits"] = r["hits"]
        npz[f"{name}/cigar"] = r["cigar"]
        npz[f"{name}/chains"] = r["chains"]
        npz[f"{name}/meta"] = np.array([r["mid_occ"], r["n_minimizers"], len(r["anchors"])], dtype=np.int64)
        rows[name] = {"anchors_sha": digest(r["anchors"]), "input_sha": digest(cases.flat_contigs(contigs)[0]),
                      "n_hits": int(len(r["hits"]))}
        if contigs and sum(len(s) for _, s in contigs) > 0:
            key = id(db)
            if key not in serotypers:
                serotypers[key] = Serotyper(rb.reference_database(db))
            g = GenomeAssembly.from_stream(io.BytesIO(cases.fasta_bytes([c for c in contigs if len(c[1]) > 0])), name)
            res = serotypers[key](g)
            row = KaptiveRow.from_result(res)
            rows[name].update({
                "kaptive_row": bytes(row).decode() if hasattr(row, "__bytes__") else str(row),
                "best_locus": res.best_locus_name, "typeable": bool(res.typeable),
                "percent_identity": round(float(res.percent_identity), 2), "percent_coverage": round(float(res.percent_coverage), 2),
            })
        print(name, rows[name].get("best_locus"), rows[name]["n_hits"])
    np.savez_compressed(HERE / "mapping_golden.npz", **npz)
    (HERE / "kaptive_rows.json").write_text(json.dumps(rows, indent=1, sort_keys=True))

    # ---- post-mapping kernels of the reference itself (bit-exact oracle exists in-tree)
    from kaptive.core.pairwise import PairwiseAligner
    from kaptive.core.seq import SeqRecord, Sequences

    rng = np.random.default_rng(5)
    aa = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
    qs, ts = [], []
    for i in range(48):
        n = int(rng.integers(5, 420))
        t = aa[rng.integers(0, 20, size=n)]
        q = t.copy()
        m = rng.random(n) < rng.uniform(0, 0.3)
        q[m] = aa[rng.integers(0, 20, size=int(m.sum()))]
        if i % 3 == 0 and n > 30:  # indels
            c = int(rng.integers(5, n - 10))
            q = np.concatenate([q[:c], q[c + int(rng.integers(1, 9)):]])
        if i % 5 == 0 and n > 30:
            c = int(rng.integers(5, n - 10))
            q = np.concatenate([q[:c], aa[rng.integers(0, 20, size=int(rng.integers(1, 30)))], q[c:]])
        if i % 7 == 0:
            q = np.concatenate([q, np.frombuffer(b"X*", dtype=np.uint8)])
        if i == 11:
            q = np.zeros(0, dtype=np.uint8)
        qs.append(q.tobytes()), ts.append(t.tobytes())
    qs.append(b"MKTLLILAVGALS"), ts.append(b"MKTLILAVGGLS")  # KAT observed in SURVEY.md section 8c: score 37, 11/1/1
    Q = Sequences.from_records([SeqRecord(seq=s, id=str(i)) for i, s in enumerate(qs)])
    T = Sequences.from_records([SeqRecord(seq=s, id=str(i)) for i, s in enumerate(ts)])
    pa = PairwiseAligner()(Q, T)
    post = {"gotoh/q": np.frombuffer(b"".join(qs), np.uint8), "gotoh/q_len": np.array([len(s) for s in qs], np.int32),
            "gotoh/t": np.frombuffer(b"".join(ts), np.uint8), "gotoh/t_len": np.array([len(s) for s in ts], np.int32)}
    for f in ("scores", "matches", "mismatches", "gaps", "q_starts", "q_ends", "t_starts", "t_ends"):
        post[f"gotoh/{f}"] = np.asarray(getattr(pa, f))
    post["gotoh/pidents"] = np.asarray(pa.pidents)
    # extract + translate
    contigs = [bytes(rng.choice(np.frombuffer(b"ACGTNacgtRY", np.uint8), size=int(n), p=[.23, .23, .23, .23, .02, .01, .01, .01, .01, .01, .01]))
               for n in (50, 1000, 3000, 7)]
    C = Sequences.from_records([SeqRecord(seq=s, id=f"c{i}") for i, s in enumerate(contigs)])
    n = 40
    ci = rng.integers(0, 3, size=n).astype(np.uint32)
    lens = np.array([len(contigs[c]) for c in ci])
    st = (rng.random(n) * (lens - 1)).astype(np.int32)
    en = np.minimum(lens, st + rng.integers(0, 700, size=n)).astype(np.int32)
    sd = rng.choice(np.array([1, -1], dtype=np.int8), size=n)
    ex = C.extract(ci, st, en, sd)
    frames = rng.integers(0, 3, size=n).astype(np.int8)
    tr = ex.translate(frames=frames, to_stop=True)
    tr2 = ex.translate(frames=frames, to_stop=False)
    post.update({"ext/contigs": np.frombuffer(b"".join(contigs), np.uint8), "ext/contig_len": np.array([len(c) for c in contigs], np.int32),
                 "ext/ci": ci, "ext/st": st, "ext/en": en, "ext/sd": sd, "ext/frames": frames,
                 "ext/out": np.asarray(ex.seqs), "ext/out_len": np.asarray(ex.lengths),
                 "tr/out": np.asarray(tr.seqs), "tr/out_len": np.asarray(tr.lengths),
                 "tr2/out": np.asarray(tr2.seqs), "tr2/out_len": np.asarray(tr2.lengths)})
    np.savez_compressed(HERE / "post_golden.npz", **post)
    print("gotoh KAT:", int(pa.scores[-1]), int(pa.matches[-1]), int(pa.mismatches[-1]), int(pa.gaps[-1]))


if __name__ == "__main__":
    main()
