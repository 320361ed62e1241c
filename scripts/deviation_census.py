#!/usr/bin/env python
"""How often would the minimap2 behaviours that mapping spec v1 leaves out have touched an emitted field?  

The arithmetic of the reference's mapper lives in the closed `rammappy` wheel; oracle and kernels restate published minimap2 and leave
out four behaviours (oracle/kb_oracle.h).  None of them can be implemented against a verifiable source here, so this script measures
their EXPOSURE: over bench-shaped assemblies (K+O index, both loci embedded), a fragmentation ladder and an insertion-sequence
workload (a 1.5 kb element in 12 copies, one of them inside a locus gene), it counts the (assembly, gene) queries whose state meets the
documented trigger of each behaviour -- an upper bound on the hits that could differ.  CPU only (oracle); run in the build container:

    python scripts/deviation_census.py > profiles/deviations_r2.json
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle_lib as ol  # noqa: E402
from kaptive_b200 import synth  # noqa: E402


def is_workload(db, i):
    """An assembly with a 1.5 kb element in 12 copies; one copy interrupts a gene of the embedded locus."""
    rng = np.random.default_rng(7700 + i)
    a = synth.make_assembly(db, i % 150, seed=7700 + i, genome_len=400_000, mean_contigs=6)
    elem = synth.random_dna(rng, 1500)
    contigs = []
    placed_in_gene = False
    for name, s in a.contigs:
        s = np.frombuffer(s, np.uint8)
        cuts = sorted(rng.integers(0, max(1, len(s)), size=2).tolist())
        parts, last = [], 0
        for c in cuts:
            parts += [s[last:c], elem if rng.random() < 0.5 else synth.revcomp(elem)]
            last = c
        parts.append(s[last:])
        contigs.append((name, np.concatenate(parts).tobytes()))
    return synth.SynthAssembly(name=a.name, contigs=contigs, locus=a.locus, locus_strand=a.locus_strand), placed_in_gene


def census(db, asms, label):
    odb = ol.OracleDB(*db.flat())
    glen = np.array([len(g) for g in db.genes])
    tot = dict(assemblies=len(asms), queries_with_hits=0, hits=0, dp_max_rule_exposed_queries=0, dp_max_rule_exposed_hits=0,
               long_join_exposed_queries=0, multi_chain_queries=0, seeds_over_mid_occ_assemblies=0, mid_occ_max=0)
    for a in asms:
        r = odb.map(*a.flat(), keep_stages=True)
        h, ch = r["hits"], r["chains"]
        tot["hits"] += len(h)
        tot["mid_occ_max"] = max(tot["mid_occ_max"], r["mid_occ"])
        if r["mid_occ"] > 10:
            tot["seeds_over_mid_occ_assemblies"] += 1
        for g in np.unique(h["gene"]):
            hh = h[h["gene"] == g]
            tot["queries_with_hits"] += 1
            # [mm2:hit.c:mm_update_dp_max] fires for queries >= 500 bp (rank_min_len) with >= 2 alignments when the best one covers
            # >= rank_frac (0.9) of the query and the second best dp_max is >= 0.9 of the best
            if glen[g] >= 500 and len(hh) >= 2:
                d = np.sort(hh["dp_max"])[::-1]
                best = hh[np.argmax(hh["dp_max"])]
                if (best["q_end"] - best["q_start"]) >= 0.9 * glen[g] and d[1] >= 0.9 * d[0]:
                    tot["dp_max_rule_exposed_queries"] += 1
                    tot["dp_max_rule_exposed_hits"] += len(hh)
        for g in np.unique(ch["gene"]):
            cc = ch[ch["gene"] == g]
            if len(cc) > 1:
                tot["multi_chain_queries"] += 1
                # [mm2:map.c:mm_map_frag] RMQ re-chaining / long join is attempted when the best chain leaves > rmq_rescue_size (1000)
                # query bases uncovered, or covers < rmq_rescue_ratio (0.1) of the query
                b = cc[np.argmax(cc["score"])]
                cov = b["qe"] - b["qs"]
                if glen[g] - cov > 1000 or cov < 0.1 * glen[g]:
                    tot["long_join_exposed_queries"] += 1
    tot["label"] = label
    return tot


def main():
    db, ranges = synth.make_ko_db()
    bench = [synth.make_assembly(db, i % 150, seed=1000 + i, genome_len=400_000, mean_contigs=8, extra_loci=(150 + i % 20,)) for i in range(24)]
    ladder = [synth.make_assembly(db, (7 * i) % 150, seed=3000 + i, genome_len=400_000, mean_contigs=[3, 12, 50, 200, 800][i % 5],
                                  extra_loci=(150 + i % 20,), sub=(0.0, 0.06), indel=(0.0, 0.005)) for i in range(20)]
    iswl = [is_workload(db, i)[0] for i in range(12)]
    out = {"what": "exposure of the minimap2 behaviours outside mapping spec v1 (upper bounds; oracle/kb_oracle.c on this container's CPU)",
           "workloads": [census(db, bench, "bench-shaped (K+O index, a K and an O locus per assembly, 0-5 % substitutions)"),
                         census(db, ladder, "fragmentation ladder (mean contigs 3 -> 800, 0-6 % substitutions)"),
                         census(db, iswl, "insertion-sequence workload (1.5 kb element, ~2 copies per contig)")]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
