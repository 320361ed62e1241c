"""Host-side throughput of the batch FASTA ingest (GB/s of FASTA bytes) for 1, 4, all cores -- a diagnostic, not the bench."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from kaptive_b200 import ingest, synth
rng = np.random.default_rng(1)
n, L = 48, 5_000_000
files = []
for i in range(n):
    s = synth.random_dna(rng, L, 0.57).tobytes()
    cuts = sorted(rng.integers(1, L, size=79).tolist())
    parts, last = [], 0
    for k, c in enumerate(cuts + [L]):
        ctg = s[last:c]; last = c
        parts.append(b">contig_%d\n" % (k + 1))
        parts.append(b"\n".join(ctg[j : j + 80] for j in range(0, len(ctg), 80)) + b"\n")
    files.append(b"".join(parts))
tot = sum(len(f) for f in files)
for th in (1, 4, os.cpu_count() or 1):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); b = ingest.ingest_fasta(files, threads=th); best = min(best, time.perf_counter() - t0)
    print(f"threads {th:3d}: {tot / best / 1e9:6.2f} GB/s FASTA -> {n / best:7.0f} assemblies/s  ({len(b.contig_len)} contigs)")
