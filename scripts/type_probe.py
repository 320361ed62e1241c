"""Where type_many spends its time (GPU box): python scripts/type_probe.py [n_asm]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from kaptive_b200 import _lib, mapper, serotype, synth, workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
db, ranges = synth.make_ko_db()
gi = mapper.GeneIndex(db.genes)
wl = workload.make_device_workload(db, n, 5_000_000, seed=1000, locus_ranges=ranges)
batch = mapper.AssemblyBatch(wl.ascii.data_ptr(), wl.contig_off, wl.contig_len, wl.asm_contig_start)
res = gi.map(batch)
tdb = serotype.TypingDB.from_synth(db)
for it in range(3):
    t0 = time.perf_counter(); best = serotype.score_loci(tdb, res.hits, n); t1 = time.perf_counter()
    typed = serotype.type_many(tdb, batch, res); t2 = time.perf_counter()
    tt = np.zeros(4); _lib.load().kb_type_debug_times(_lib.ptr(tt))
    print(f"score_loci {1e3*(t1-t0):.1f} ms; type_many {1e3*(t2-t1):.1f} ms; inside call: pass1 {1e3*tt[0]:.1f} jobs {1e3*tt[1]:.1f} device {1e3*tt[2]:.1f} pass2 {1e3*tt[3]:.1f}; hits {len(res)} kept {len(typed.gene_hits['gene'])}")
