"""Seeded assemblies for the typing tests (type_many vs the unmodified reference Serotyper): a combined K+O-shaped database and
assemblies over a divergence / fragmentation ladder (BASELINE.json configs[2] and [4] in miniature): clean, mutated, fragmented
down to sub-kilobase contigs, locus split over many pieces, a second locus of the same class nearby, no locus at all."""

from __future__ import annotations

import numpy as np

from kaptive_b200 import synth


def make_db():
    k = synth.make_db(n_loci=14, genes_per_locus=10, n_core=3, seed=31, prefix="KL")
    o = synth.make_db(n_loci=5, genes_per_locus=6, n_core=2, n_extra=4, seed=32, prefix="OL")
    return synth.combine(k, o), 14, 5


def make_assemblies(db, n_k: int, n_o: int, n: int = 48):
    ladder = [3, 6, 12, 25, 50, 100, 200, 400]
    out = []
    for i in range(n):
        kl, ol = i % n_k, n_k + i % n_o
        sub = (0.0, 0.02) if i % 4 else (0.04, 0.12)
        if i % 11 == 10:  # no locus of the database at all
            rng = np.random.default_rng(9100 + i)
            g = synth.random_dna(rng, 120_000, 0.57).tobytes()
            out.append(synth.SynthAssembly(name=f"asm_none_{i}", contigs=[("contig_1", g[:70_000]), ("contig_2", g[70_000:])], locus=-1, locus_strand=1))
            continue
        extra = (ol,) if i % 3 else (ol, (kl + 5) % n_k)
        out.append(synth.make_assembly(db, kl, seed=9100 + i, genome_len=300_000, mean_contigs=ladder[i % 8], extra_loci=extra, sub=sub,
                                       indel=(0.0, 0.004)))
    return out
