"""Seeded parity cases shared by the golden generator and the CPU / GPU parity tests.

Each case returns ``(db, contigs)``: a :class:`kaptive_b200.synth.SynthDB` and a list of ``(name, bytes)``
contigs of ONE assembly.  They cover the regular path plus the edge cases the domain has: empty and ragged
inputs (no contigs, contigs shorter than a k-mer / a window), ambiguous bases, lower case, both strands,
fragmentation through genes, repeats above the occurrence cut-off, z-drop splits, long indels.
"""

from __future__ import annotations

import functools

import numpy as np

from kaptive_b200 import synth


@functools.lru_cache(maxsize=8)
def small_db(seed: int = 1, n_extra: int = 1) -> synth.SynthDB:
    return synth.make_db(n_loci=10, genes_per_locus=8, n_core=2, seed=seed, n_extra=n_extra)


def _asm(db, locus, seed, **kw):
    kw.setdefault("genome_len", 250_000)
    kw.setdefault("mean_contigs", 10)
    return synth.make_assembly(db, locus, seed=seed, **kw).contigs


def _custom(seed: int, pieces: list[np.ndarray], spacer: int = 20_000, cuts: int = 0):
    """Random background with the given pieces inserted `spacer` apart; optional random contig breakpoints."""
    rng = np.random.default_rng(seed)
    parts = [synth.random_dna(rng, spacer, 0.57)]
    for p in pieces:
        parts.append(p)
        parts.append(synth.random_dna(rng, spacer, 0.57))
    g = np.concatenate(parts)
    if cuts:
        bps = np.unique(rng.integers(1, len(g), size=cuts))
        bounds = np.concatenate([[0], bps, [len(g)]])
    else:
        bounds = np.array([0, len(g)])
    return [(f"c{i + 1}", g[bounds[i] : bounds[i + 1]].tobytes()) for i in range(len(bounds) - 1)]


def _gene(db, i) -> np.ndarray:
    return np.frombuffer(db.genes[i], dtype=np.uint8)


def case_exact():
    db = small_db()
    return db, _asm(db, 3, 1000, mean_contigs=1, sub=(0, 0), indel=(0, 0), n_frac=0)


def case_mutated(k: int):
    db = small_db()
    return db, _asm(db, k % 10, 2000 + k, genome_len=300_000, mean_contigs=12)


def case_clean_typeable():
    db = small_db()
    return db, _asm(db, 6, 2100, sub=(0.01, 0.02), indel=(0, 0), mean_contigs=4)


def case_fragmented():
    db = small_db()
    return db, _asm(db, 2, 2200, genome_len=200_000, mean_contigs=70)


def case_lowercase():
    db = small_db()
    return db, _asm(db, 4, 2300, lowercase_frac=0.3)


def case_n_rich():
    db = small_db()
    return db, _asm(db, 5, 2400, n_frac=0.003)


def case_two_loci():
    db = small_db()
    return db, _asm(db, 1, 2500, extra_loci=(7,))


def case_divergent():
    db = small_db()
    return db, _asm(db, 8, 2600, sub=(0.11, 0.12), indel=(0.001, 0.002))


def case_repeat_gene():
    """A 70-base fragment of a gene present 15 times in a 400 kb assembly: few enough distinct repeated
    minimizers that the census keeps mid_occ at 10, so those seeds are dropped as repeats (rep_len > 0)."""
    db = small_db()
    g = _gene(db, 20)
    frag = g[300:370]
    pieces = [frag.copy() for _ in range(14)]
    pieces.append(np.frombuffer(db.loci[2], dtype=np.uint8))
    return db, _custom(2700, pieces, spacer=25_000)


def case_mosaic_gene():
    """Genes whose middle ~55% is replaced by random sequence: the gap fill z-drops and the region is split in two."""
    db = small_db()
    out = []
    rng = np.random.default_rng(78)
    n = 0
    for gi in range(len(db.genes)):
        g = _gene(db, gi).copy()
        if len(g) < 1250 or db.gene_locus[gi] > 9:
            continue
        g[330:-330] = synth.random_dna(rng, len(g) - 660)
        out.append(g)
        n += 1
        if n == 4:
            break
    return db, _custom(2800, out)


def case_big_indel():
    """40 bp deletion + 25 bp insertion + 70 bp deletion inside genes: long-gap seed filters and long gaps in the DP."""
    db = small_db()
    rng = np.random.default_rng(79)
    out = []
    for gi in (19, 28, 35, 41):
        g = _gene(db, gi)
        g = np.concatenate([g[:200], g[240:500], synth.random_dna(rng, 25), g[500:700], g[770:]])
        out.append(synth.mutate(rng, g, 0.01))
    return db, _custom(2900, out, cuts=2)


IS_GENES = (19, 28, 35)


def case_is_insertion():
    """Three genes each interrupted by a 2 kb insertion (an IS element inside a locus gene, a common Kaptive input).  minimap2 with its
    default long-join (bw_long 20000) reports ONE hit per gene with a 2 kb deletion-side gap; mapping spec v1 has no long-join and reports
    the two flanks as TWO hits.  The golden vectors pin that documented deviation so that it cannot change silently."""
    db = small_db()
    rng = np.random.default_rng(81)
    out = []
    for gi in IS_GENES:
        g = _gene(db, gi)
        cut = len(g) * 45 // 100
        out.append(np.concatenate([g[:cut], synth.random_dna(rng, 2000, 0.5), g[cut:]]))
    return db, _custom(3100, out)


def case_tiny_contigs():
    db = small_db()
    rng = np.random.default_rng(80)
    contigs = [("e0", b""), ("t5", synth.random_dna(rng, 5).tobytes()), ("t14", synth.random_dna(rng, 14).tobytes()),
               ("t15", synth.random_dna(rng, 15).tobytes()), ("t24", synth.random_dna(rng, 24).tobytes()),
               ("t25", synth.random_dna(rng, 25).tobytes()), ("allN", b"N" * 300)]
    contigs += _asm(db, 9, 3000, genome_len=120_000, mean_contigs=3)
    contigs += [("g_only", db.genes[5]), ("e1", b""), ("tail", db.genes[12][:100])]
    return db, contigs


def case_empty_assembly():
    return small_db(), []


def case_no_locus():
    db = small_db()
    rng = np.random.default_rng(81)
    return db, [("bg", synth.random_dna(rng, 100_000, 0.57).tobytes())]


def case_long_contig_boundaries():
    """Contig lengths around the 256-base lane and 8192-base chunk boundaries of the scan kernel."""
    db = small_db()
    rng = np.random.default_rng(82)
    locus = np.frombuffer(db.loci[0], dtype=np.uint8)
    contigs = []
    for i, n in enumerate((255, 256, 257, 8191, 8192, 8193, 16384, 16385)):
        s = synth.random_dna(rng, n, 0.5)
        contigs.append((f"b{i}", s.tobytes()))
    contigs.append(("locus", np.concatenate([synth.random_dna(rng, 8192 - 300), locus, synth.random_dna(rng, 777)]).tobytes()))
    return db, contigs


def case_long_tail():
    """A 4.2 kb gene of which only the first 2.3 kb are in the assembly (the rest of the contig is unrelated sequence): no seed beyond,
    so the right end extension is a 1.9 k x 3.8 k rectangle (7.2 M cells) -- over the 4 M cells the staged DP kernels keep scratch for,
    under minimap2's max_sw_mat of 100 M: the chain has to go through the full-size kernel, and the extension it finds (a few bases
    into the unrelated sequence) must be the oracle's, not "skipped" (max_sw_cells)."""
    rng = np.random.default_rng(77)
    gene = synth.random_orf(rng, 1400)
    db = small_db()
    db = synth.SynthDB(genes=db.genes + [gene.tobytes()], gene_locus=np.append(db.gene_locus, db.gene_locus.max() + 1).astype(np.int32),
                       gene_pos=np.append(db.gene_pos, 1).astype(np.int32), gene_start=np.append(db.gene_start, 0).astype(np.int32),
                       gene_end=np.append(db.gene_end, len(gene)).astype(np.int32), gene_strand=np.append(db.gene_strand, 1).astype(np.int8),
                       extra=np.append(db.extra, False), loci=db.loci + [gene.tobytes()], locus_names=db.locus_names + ["KLlong"],
                       gene_names=db.gene_names + ["KLlong_01_glong"])
    head = synth.mutate(rng, gene[:2300], 0.01)
    return db, _custom(78, [head], spacer=9000)


CASES = {
    "exact": case_exact,
    "mutated0": lambda: case_mutated(0),
    "mutated1": lambda: case_mutated(1),
    "mutated2": lambda: case_mutated(2),
    "clean_typeable": case_clean_typeable,
    "fragmented": case_fragmented,
    "lowercase": case_lowercase,
    "n_rich": case_n_rich,
    "two_loci": case_two_loci,
    "divergent": case_divergent,
    "repeat_gene": case_repeat_gene,
    "mosaic_gene": case_mosaic_gene,
    "big_indel": case_big_indel,
    "is_insertion": case_is_insertion,
    "tiny_contigs": case_tiny_contigs,
    "empty_assembly": case_empty_assembly,
    "no_locus": case_no_locus,
    "boundaries": case_long_contig_boundaries,
    "long_tail": lambda: case_long_tail(),
}


def flat_contigs(contigs):
    lengths = np.array([len(s) for _, s in contigs], dtype=np.int32)
    offsets = np.zeros(len(contigs), dtype=np.int64)
    if len(contigs) > 1:
        np.cumsum(lengths[:-1].astype(np.int64), out=offsets[1:])
    seqs = np.frombuffer(b"".join(s for _, s in contigs), dtype=np.uint8).copy() if contigs else np.zeros(0, np.uint8)
    return seqs, offsets, lengths


def fasta_bytes(contigs) -> bytes:
    out = []
    for n, s in contigs:
        out.append(b">" + n.encode() + b" some description\n")
        out.extend(s[i : i + 70] + b"\n" for i in range(0, len(s), 70))
    return b"".join(out)
