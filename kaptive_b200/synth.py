"""Seeded synthetic K/O-locus databases and draft assemblies.

The real ``kpsc_k`` / ``kpsc_o`` / ``ab_k`` / ``ab_o`` databases are downloaded by the
reference at first use (``/root/reference/src/kaptive/db/manager.py:63-73``) and cannot be
fetched here, so every test and benchmark input is generated from seeds:

* a database of ``n_loci`` loci x ``genes_per_locus`` stop-free ORFs (~1 kb), of which
  ``n_core`` gene families are shared by every locus at a per-locus divergence (the
  galF / wzi / gnd / ugd pattern of real K loci), plus optional "extra" genes;
* an assembly: random background (GC 0.57) with one DB locus embedded after
  substitutions and small indels, random strand, cut into contigs at random breakpoints
  (breakpoints may fall inside the locus), a sprinkling of ``N``.

Nothing here touches the GPU; arrays are plain numpy.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = ord("N")
for _a, _b in zip(b"ACGTacgtNn", b"TGCAtgcaNn"):
    _COMP[_a] = _b
_STOPS = {b"TAA", b"TAG", b"TGA"}


def revcomp(seq: np.ndarray) -> np.ndarray:
    return _COMP[seq[::-1]]


def random_dna(rng: np.random.Generator, n: int, gc: float = 0.5) -> np.ndarray:
    """n random bases as uint8 ASCII with the given GC fraction."""
    p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    return _ACGT[rng.choice(4, size=n, p=p)]


def _remove_stops(orf: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    """Replace in-frame stop codons (except the last codon) so the ORF translates end to end."""
    n_codon = len(orf) // 3
    cod = orf[: n_codon * 3].reshape(n_codon, 3)
    is_stop = (cod[:, 0] == ord("T")) & (
        ((cod[:, 1] == ord("A")) & ((cod[:, 2] == ord("A")) | (cod[:, 2] == ord("G"))))
        | ((cod[:, 1] == ord("G")) & (cod[:, 2] == ord("A")))
    )
    idx = np.nonzero(is_stop)[0]
    cod[idx, 0] = _ACGT[rng.integers(0, 3, size=len(idx))]  # A, C or G: never a stop
    return cod.reshape(-1)


def random_orf(rng: np.random.Generator, n_codons: int, gc: float = 0.5) -> np.ndarray:
    body = _remove_stops(random_dna(rng, n_codons * 3, gc), rng)
    body[:3] = np.frombuffer(b"ATG", dtype=np.uint8)
    body[-3:] = np.frombuffer(b"TAA", dtype=np.uint8)
    return body


def mutate(
    rng: np.random.Generator, seq: np.ndarray, sub: float, indel: float = 0.0, keep_orf: bool = False
) -> np.ndarray:
    """Point substitutions at rate ``sub`` and 1-3 bp indels at rate ``indel`` (per base)."""
    s = seq.copy()
    n = len(s)
    if sub > 0:
        pos = np.nonzero(rng.random(n) < sub)[0]
        if keep_orf:  # leave start/stop codon alone
            pos = pos[(pos >= 3) & (pos < n - 3)]
        # substitute with a different base
        cur = np.searchsorted(_ACGT, s[pos])
        s[pos] = _ACGT[(cur + rng.integers(1, 4, size=len(pos))) % 4]
    if indel > 0:
        pos = np.nonzero(rng.random(n) < indel)[0]
        if len(pos):
            out = []
            last = 0
            for p in pos:
                out.append(s[last:p])
                ln = int(rng.integers(1, 4))
                if rng.random() < 0.5:
                    out.append(random_dna(rng, ln))
                    last = p
                else:
                    last = min(n, p + ln)
            out.append(s[last:])
            s = np.concatenate(out)
    if keep_orf:
        s = s[: len(s) // 3 * 3].copy()
        s = _remove_stops(s, rng)
        s[:3] = np.frombuffer(b"ATG", dtype=np.uint8)
        s[-3:] = np.frombuffer(b"TAA", dtype=np.uint8)
    return s


@dataclass
class SynthDB:
    """Gene queries + locus layout, mirroring the fields of ``kaptive.db.Database`` the hot path reads
    (``db.genes`` seqs/offsets/lengths, ``gene_locus_indices``, ``extra_genes``;
    reference ``src/kaptive/db/core.py:82-98``)."""

    genes: list[bytes]
    gene_locus: np.ndarray  # int32, locus index per gene (extra genes: index of their pseudo-locus)
    gene_pos: np.ndarray  # int32, 1-based position within locus
    gene_start: np.ndarray  # int32, start within locus sequence
    gene_end: np.ndarray
    gene_strand: np.ndarray  # int8 +1/-1
    extra: np.ndarray  # bool
    loci: list[bytes]
    locus_names: list[str]
    gene_names: list[str] = field(default_factory=list)

    def flat(self) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Concatenated gene bytes, int64 offsets, int32 lengths (the C-ABI input layout)."""
        lengths = np.array([len(g) for g in self.genes], dtype=np.int32)
        offsets = np.zeros(len(self.genes), dtype=np.int64)
        if len(lengths) > 1:
            np.cumsum(lengths[:-1], out=offsets[1:])
        seqs = np.frombuffer(b"".join(self.genes), dtype=np.uint8).copy() if self.genes else np.zeros(0, np.uint8)
        return seqs, offsets, lengths


def make_db(
    n_loci: int = 150,
    genes_per_locus: int = 20,
    n_core: int = 4,
    core_div: tuple[float, float] = (0.03, 0.18),
    gene_codons: tuple[int, int] = (200, 500),
    n_extra: int = 0,
    seed: int = 1,
    prefix: str = "KL",
) -> SynthDB:
    rng = np.random.default_rng(seed)
    core_anc = [random_orf(rng, int(rng.integers(*gene_codons))) for _ in range(n_core)]
    genes, g_locus, g_pos, g_start, g_end, g_strand, names = [], [], [], [], [], [], []
    loci = []
    for li in range(n_loci):
        parts = [random_dna(rng, int(rng.integers(100, 300)))]
        off = len(parts[0])
        # core genes sit at the ends of the locus as in real K loci (galF..wzc at 5', gnd/ugd at 3')
        n_head = (n_core + 1) // 2
        order = list(range(genes_per_locus))
        for gi in order:
            if gi < n_head:
                anc = core_anc[gi]
                orf = mutate(rng, anc, float(rng.uniform(*core_div)), 0.0, keep_orf=True)
            elif gi >= genes_per_locus - (n_core - n_head):
                anc = core_anc[n_head + gi - (genes_per_locus - (n_core - n_head))]
                orf = mutate(rng, anc, float(rng.uniform(*core_div)), 0.0, keep_orf=True)
            else:
                orf = random_orf(rng, int(rng.integers(*gene_codons)))
            strand = 1 if (gi < genes_per_locus - 2 or rng.random() < 0.7) else -1
            placed = orf if strand == 1 else revcomp(orf)
            genes.append(orf.tobytes())
            g_locus.append(li)
            g_pos.append(gi + 1)
            g_start.append(off)
            g_end.append(off + len(placed))
            g_strand.append(strand)
            names.append(f"{prefix}{li + 1}_{gi + 1:02}_g{gi + 1}")
            parts.append(placed)
            off += len(placed)
            sp = random_dna(rng, int(rng.integers(20, 120)))
            parts.append(sp)
            off += len(sp)
        loci.append(np.concatenate(parts).tobytes())
    extra = [False] * len(genes)
    locus_names = [f"{prefix}{i + 1}" for i in range(n_loci)]
    for ei in range(n_extra):
        orf = random_orf(rng, int(rng.integers(*gene_codons)))
        genes.append(orf.tobytes())
        g_locus.append(n_loci + ei)
        g_pos.append(1)
        g_start.append(0)
        g_end.append(len(orf))
        g_strand.append(1)
        names.append(f"Extra_genes_x{ei + 1}_01_x{ei + 1}")
        extra.append(True)
        loci.append(orf.tobytes())
        locus_names.append(f"Extra_genes_x{ei + 1}")
    return SynthDB(
        genes=genes,
        gene_locus=np.array(g_locus, dtype=np.int32),
        gene_pos=np.array(g_pos, dtype=np.int32),
        gene_start=np.array(g_start, dtype=np.int32),
        gene_end=np.array(g_end, dtype=np.int32),
        gene_strand=np.array(g_strand, dtype=np.int8),
        extra=np.array(extra, dtype=bool),
        loci=loci,
        locus_names=locus_names,
        gene_names=names,
    )


def combine(*dbs: SynthDB) -> SynthDB:
    """Several gene sets in ONE index, loci renumbered consecutively (BASELINE.json configs[2]: kpsc_k + kpsc_o combined)."""
    genes, loci, names, gnames = [], [], [], []
    cols = {k: [] for k in ("gene_locus", "gene_pos", "gene_start", "gene_end", "gene_strand", "extra")}
    for d in dbs:
        base = len(loci)
        genes += d.genes
        loci += d.loci
        names += d.locus_names
        gnames += d.gene_names
        cols["gene_locus"].append(d.gene_locus + base)
        for k in ("gene_pos", "gene_start", "gene_end", "gene_strand", "extra"):
            cols[k].append(getattr(d, k))
    c = {k: np.concatenate(v) for k, v in cols.items()}
    return SynthDB(genes=genes, gene_locus=c["gene_locus"].astype(np.int32), gene_pos=c["gene_pos"], gene_start=c["gene_start"],
                   gene_end=c["gene_end"], gene_strand=c["gene_strand"], extra=c["extra"], loci=loci, locus_names=names, gene_names=gnames)


def make_ko_db(k_loci: int = 150, k_genes: int = 20, k_core: int = 4, o_loci: int = 20, o_genes: int = 10, o_core: int = 2,
               o_extra: int = 15, seed: int = 1) -> tuple[SynthDB, tuple[tuple[int, int], ...]]:
    """kpsc_k-shaped + kpsc_o-shaped gene sets in one index: 150 x 20 K genes with 4 core families,
    20 x 10 O genes with 2 core families + 15 "Extra genes" records.  Returns the database and, per locus class, the half-open range
    of locus indices an assembly draws its embedded locus from ((0, 150) for K, (150, 170) for O)."""
    k = make_db(n_loci=k_loci, genes_per_locus=k_genes, n_core=k_core, seed=seed, prefix="KL")
    o = make_db(n_loci=o_loci, genes_per_locus=o_genes, n_core=o_core, n_extra=o_extra, seed=seed + 1, prefix="OL")
    return combine(k, o), ((0, k_loci), (k_loci, k_loci + o_loci))


@dataclass
class SynthAssembly:
    name: str
    contigs: list[tuple[str, bytes]]
    locus: int
    locus_strand: int

    def flat(self) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        lengths = np.array([len(s) for _, s in self.contigs], dtype=np.int32)
        offsets = np.zeros(len(lengths), dtype=np.int64)
        if len(lengths) > 1:
            np.cumsum(lengths[:-1].astype(np.int64), out=offsets[1:])
        seqs = np.frombuffer(b"".join(s for _, s in self.contigs), dtype=np.uint8).copy()
        return seqs, offsets, lengths

    def fasta(self) -> bytes:
        out = []
        for n, s in self.contigs:
            out.append(b">" + n.encode() + b"\n")
            out.extend(s[i : i + 80] + b"\n" for i in range(0, len(s), 80))
        return b"".join(out)


def make_assembly(
    db: SynthDB,
    locus: int,
    seed: int,
    genome_len: int = 5_000_000,
    mean_contigs: float = 80.0,
    sub: tuple[float, float] = (0.0, 0.05),
    indel: tuple[float, float] = (0.0, 0.005),
    gc: float = 0.57,
    n_frac: float = 1e-4,
    extra_loci: tuple[int, ...] = (),
    lowercase_frac: float = 0.0,
) -> SynthAssembly:
    """One draft assembly with DB locus ``locus`` (and ``extra_loci``) embedded."""
    rng = np.random.default_rng(seed)
    pieces = []
    for li in (locus, *extra_loci):
        ls = np.frombuffer(db.loci[li], dtype=np.uint8)
        ls = mutate(rng, ls, float(rng.uniform(*sub)), float(rng.uniform(*indel)))
        strand = 1 if rng.random() < 0.5 else -1
        if strand < 0:
            ls = revcomp(ls)
        pieces.append((ls, strand))
    bg_len = max(1000, genome_len - sum(len(p) for p, _ in pieces))
    bg = random_dna(rng, bg_len, gc)
    cuts = np.sort(rng.integers(0, bg_len, size=len(pieces)))
    parts, last = [], 0
    for (ls, _), c in zip(pieces, cuts):
        parts.append(bg[last:c])
        parts.append(ls)
        last = c
    parts.append(bg[last:])
    genome = np.concatenate(parts)
    if n_frac > 0:
        npos = np.nonzero(rng.random(len(genome)) < n_frac)[0]
        genome[npos] = ord("N")
    if lowercase_frac > 0:
        lo = np.nonzero(rng.random(len(genome)) < lowercase_frac)[0]
        genome[lo] |= 0x20
    n_ctg = max(1, int(rng.poisson(mean_contigs)))
    bps = np.unique(rng.integers(1, len(genome), size=n_ctg - 1)) if n_ctg > 1 else np.zeros(0, dtype=np.int64)
    bounds = np.concatenate([[0], bps, [len(genome)]]).astype(np.int64)
    contigs = []
    for i in range(len(bounds) - 1):
        s = genome[bounds[i] : bounds[i + 1]]
        if rng.random() < 0.5:
            s = revcomp(s)
        contigs.append((f"contig_{i + 1}", s.tobytes()))
    return SynthAssembly(name=f"asm_{seed}", contigs=contigs, locus=locus, locus_strand=pieces[0][1])
