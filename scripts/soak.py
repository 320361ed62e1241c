"""Parity soak (diagnostic, not part of the test suite): N random assemblies over a wide range of divergence, indel rate,
fragmentation and ambiguity, GPU (staged path) vs the CPU oracle, every hit field and CIGAR.  Usage: python scripts/soak.py [N] [seed]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import oracle_lib as ol
from kaptive_b200 import mapper, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 120
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
db = synth.make_db(n_loci=20, genes_per_locus=12, n_core=4, core_div=(0.02, 0.22), n_extra=3, seed=100 + seed)
gi = mapper.GeneIndex(db.genes)
odb = ol.OracleDB(*db.flat())
F = ("gene", "q_start", "q_end", "t_ctg", "t_len", "t_start", "t_end", "strand", "score", "matches", "block_len", "edit_distance", "mapq", "is_primary")
asms = []
for i in range(N):
    s = float(rng.choice([0.0, 0.01, 0.03, 0.06, 0.1, 0.15]))
    ind = float(rng.choice([0.0, 0.002, 0.01, 0.03]))
    asms.append(synth.make_assembly(db, int(rng.integers(0, 20)), seed=50_000 + 1000 * seed + i, genome_len=int(rng.choice([60_000, 250_000, 600_000])),
                                    mean_contigs=float(rng.choice([1, 5, 40, 300])), sub=(s, s + 0.01), indel=(ind, ind + 0.002),
                                    n_frac=float(rng.choice([0, 1e-4, 2e-3])), extra_loci=tuple(int(x) for x in rng.integers(0, 20, size=int(rng.integers(0, 3)))),
                                    lowercase_frac=float(rng.choice([0, 0.3]))))
t0 = time.perf_counter()
res = gi.map_contigs([[c for _, c in a.contigs] for a in asms])
t1 = time.perf_counter()
bad = tot = 0
for ai, a in enumerate(asms):
    ro = odb.map(*a.flat())
    idx = np.nonzero(res.hits["asm_id"] == ai)[0]
    ok = len(idx) == len(ro["hits"])
    if ok:
        for f in F:
            ok = ok and np.array_equal(res.hits[f][idx].astype(np.int64), ro["hits"][f].astype(np.int64))
        for k, i in enumerate(idx):
            h = ro["hits"][k]
            ok = ok and np.array_equal(res.cigar_of(i), ro["cigar"][h["cigar_off"] : h["cigar_off"] + h["n_cigar"]])
    tot += len(ro["hits"])
    if not ok:
        bad += 1
        print("MISMATCH assembly", ai, a.name, len(idx), len(ro["hits"]))
print(f"soak seed {seed}: {N} assemblies, {tot} hits, {bad} mismatching assemblies; GPU {t1 - t0:.2f} s; counters {res.counters}")
sys.exit(1 if bad else 0)
