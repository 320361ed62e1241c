"""TEST INFRASTRUCTURE: glue between this repo's synthetic data and the *unmodified* reference code under
/root/reference/src (available only in the build container, never on the GPU box).  Used by
tests/golden/make_golden.py to produce the committed golden vectors, and by the opt-in CPU tests that run
the reference's own Serotyper on top of a ``rammappy`` shim."""

from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np

REF_SRC = Path("/root/reference/src")
ROOT = Path(__file__).resolve().parent.parent


def reference_available() -> bool:
    return (REF_SRC / "kaptive" / "serotyping" / "core.py").exists()


def import_reference(shim: str = "oracle"):
    """Import the reference package with a ``rammappy`` module in front: 'oracle' (tests/ref_shim) or 'gpu'."""
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_kb")
    shim_dir = ROOT / ("tests/ref_shim" if shim == "oracle" else "kaptive_b200/shim")
    for p in (str(REF_SRC), str(shim_dir)):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    # a `rammappy` cached from the other shim directory (test_abi imports the product shim) would win over sys.path: drop it, and the
    # reference modules that bound it at import time
    stale = [m for m in sys.modules if (m == "rammappy" or m.startswith("rammappy.")) and
             not str(getattr(sys.modules[m], "__file__", "") or "").startswith(str(shim_dir))]
    if stale:
        for m in stale + [m for m in sys.modules if m == "kaptive" or m.startswith("kaptive.")]:
            sys.modules.pop(m, None)
    import kaptive  # noqa: F401

    return kaptive


def reference_database(db, id_threshold: float = 82.5):
    """Hand-built kaptive.db.Database from a SynthDB (pattern of /root/reference/tests/test_db.py:59-98)."""
    from kaptive.core.interval import Interval, Intervals, Strand
    from kaptive.core.kmers import FracMinHashIndex
    from kaptive.core.seq import SeqRecord, Sequences
    from kaptive.db import Database
    from kaptive.db.models import DatabaseMetadata, Phenotypes

    meta = DatabaseMetadata(
        name="Synthetic K", keyword="synth_k", genbank="synth.gbk", organism="Synthetica", taxon=1, antigen="K",
        pathway="Wzx", version="1.0.0", id_threshold=id_threshold, doi=[], owner="o", repo="r", branch="main",
        contact={}, phenotype_logic={}, antigenic_units={},
    )  # fmt: skip
    loci = Sequences.from_records([SeqRecord(seq=s, id=n) for s, n in zip(db.loci, db.locus_names)])
    genes = Sequences.from_records([SeqRecord(seq=s, id=n) for s, n in zip(db.genes, db.gene_names)])
    translations = genes.translate()
    n_loci = len(db.loci)
    offs = np.zeros(n_loci, dtype=np.uint32)
    lens = np.zeros(n_loci, dtype=np.uint32)
    for li in range(n_loci):
        idx = np.nonzero(db.gene_locus == li)[0]
        offs[li], lens[li] = idx[0], len(idx)
    intervals = Intervals.from_intervals(
        [Interval(int(s), int(e), strand=Strand.FORWARD if st > 0 else Strand.REVERSE) for s, e, st in zip(db.gene_start, db.gene_end, db.gene_strand)]
    )
    clusters = tuple(n.split("_")[-1] for n in db.gene_names)
    ckeys = tuple(dict.fromkeys(clusters))
    cmap = {k: i for i, k in enumerate(ckeys)}
    return Database(
        metadata=meta, loci=loci, serotypes=tuple(db.locus_names), locus_gene_offsets=offs, locus_gene_lengths=lens,
        gene_intervals=intervals, genes=genes, translations=translations, extra_genes=db.extra.copy(),
        gene_locus_indices=db.gene_locus.astype(np.uint16), cluster_keys=ckeys,
        gene_cluster_ids=np.array([cmap[c] for c in clusters], dtype=np.uint16), description_keys=("hypothetical protein",),
        gene_description_ids=np.zeros(len(db.genes), dtype=np.uint16), gene_positions=db.gene_pos.astype(np.uint16),
        phenotypes=Phenotypes.empty(), loci_sketches=FracMinHashIndex.build(loci, sort_by_hash=False),
    )  # fmt: skip
