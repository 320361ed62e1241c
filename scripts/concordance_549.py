#!/usr/bin/env python
"""BASELINE.json configs[4]: 549 synthetic assemblies over a depth-like fragmentation ladder (contig N50
from ~200 kb down to ~2 kb, the locus broken into 1-8+ pieces) vs ab_k- and ab_o-shaped databases (typed separately, as the reference
types one database per run).  The paper's subsampled read-set assemblies are not obtainable; the ladder stands in for them.

Reports, per database: hit-level identity GPU vs CPU oracle (every field + CIGAR, all 549 assemblies), call-level concordance
(best locus / typeable from type_many on both hit sets -- identical hits give identical calls by construction, checked anyway through
the scoring stage on the oracle's hits), and concordance with the ground truth (the embedded locus) per ladder step.

    python scripts/concordance_549.py [n_assemblies] > profiles/concordance_r2.json      (GPU box; ~5 min)
"""
import json
import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from kaptive_b200 import mapper, serotype, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 549
GENOME = 4_000_000
LADDER = [20, 40, 80, 160, 320, 640, 1280, 2000, 3000]  # mean contigs: N50 ~200 kb ... ~1.3 kb on 4 Mb
FIELDS = ("gene", "q_start", "q_end", "t_ctg", "t_len", "t_start", "t_end", "strand", "score", "matches", "block_len", "edit_distance", "mapq", "is_primary")
_G = {}


def _oracle(idx):
    import oracle_lib as ol

    odb = ol.OracleDB(*_G["flat_db"])
    out = []
    for i in idx:
        r = odb.map(*_G["asms"][i].flat())
        out.append((i, r["hits"], r["cigar"]))
    return out


def run_db(name, db, n_real, asms, cores):
    gi = mapper.GeneIndex(db.genes)
    tdb = serotype.TypingDB.from_synth(db)
    batch = mapper.AssemblyBatch.from_contigs([[s for _, s in a.contigs] for a in asms])
    t0 = time.perf_counter()
    res = gi.map(batch)
    typed = serotype.type_many(tdb, batch, res)
    gpu_s = time.perf_counter() - t0
    _G["flat_db"], _G["asms"] = db.flat(), asms
    shards = [list(range(i, len(asms), cores)) for i in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        parts = pool.map(_oracle, [s for s in shards if s])
    cpu_s = time.perf_counter() - t0
    ores = {i: (h, c) for part in parts for i, h, c in part}
    asm = res.hits["asm_id"]
    same_hits, n_hits = 0, 0
    o_parts = []
    for i in range(len(asms)):
        oh, oc = ores[i]
        lo, hi = np.searchsorted(asm, i, "left"), np.searchsorted(asm, i, "right")
        ok = hi - lo == len(oh) and all(np.array_equal(res.hits[f][lo:hi].astype(np.int64), oh[f].astype(np.int64)) for f in FIELDS)
        if ok and len(oh):
            g = np.concatenate([res.cigar_of(k) for k in range(lo, hi)])
            w = np.concatenate([oc[int(o) : int(o) + int(n)] for o, n in zip(oh["cigar_off"], oh["n_cigar"])])
            ok = np.array_equal(g, w)
        same_hits += bool(ok)
        n_hits += len(oh)
        o_parts.append({"asm_id": np.full(len(oh), i, np.int32), **{k: oh[k] for k in ("gene", "q_start", "q_end", "score")}})
    ohits = {k: np.concatenate([p[k] for p in o_parts]) for k in o_parts[0]}
    obest, oscore, _, _ = serotype.score_loci(tdb, ohits, len(asms))
    truth = np.array([a.locus for a in asms])
    per_step = {}
    for s, mc in enumerate(LADDER):
        sel = np.arange(len(asms)) % len(LADDER) == s
        per_step[str(mc)] = {"n": int(sel.sum()), "best_locus_correct": int((typed.best_locus[sel] == truth[sel]).sum()), "typeable": int(typed.typeable[sel].sum()),
                             "median_pieces": float(np.median(typed.n_pieces[sel])), "median_contigs": float(np.median([len(asms[i].contigs) for i in np.nonzero(sel)[0]]))}
    return {"db": name, "genes": len(db.genes), "loci": n_real, "assemblies": len(asms), "hits": int(n_hits),
            "assemblies_with_identical_hits_gpu_vs_oracle": int(same_hits), "best_locus_gpu_vs_oracle_hits": int((obest == typed.best_locus).sum()),
            "best_locus_score_identical": int((oscore == typed.best_locus_score).sum()), "best_locus_equals_embedded": int((typed.best_locus == truth).sum()),
            "typeable": int(typed.typeable.sum()), "by_mean_contigs": per_step, "gpu_map_and_type_s": gpu_s, "oracle_map_s": cpu_s, "oracle_cores": cores}


def main():
    cores = os.cpu_count() or 1
    out = {"config": f"{N} synthetic {GENOME / 1e6:g} Mb assemblies, fragmentation ladder (mean contigs {LADDER}), 0-6 % substitutions, 0-0.5 % indels", "results": []}
    for name, kw in (("ab_k-shaped", dict(n_loci=100, genes_per_locus=18, n_core=3, seed=41, prefix="KL")),
                     ("ab_o-shaped", dict(n_loci=15, genes_per_locus=9, n_core=2, n_extra=3, seed=42, prefix="OCL"))):
        db = synth.make_db(**kw)
        n_real = kw["n_loci"]
        asms = [synth.make_assembly(db, i % n_real, seed=5490 + i, genome_len=GENOME, mean_contigs=LADDER[i % len(LADDER)], sub=(0.0, 0.06), indel=(0.0, 0.005))
                for i in range(N)]
        out["results"].append(run_db(name, db, n_real, asms, cores))
        print(json.dumps(out["results"][-1]), file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
