"""Timing probe for the host-buffer path: batch_create (H2D + pack), map, fetch -- not a benchmark, a diagnostic."""
import ctypes as C, sys, time, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from kaptive_b200 import mapper, synth, workload
from kaptive_b200._lib import check, load, ptr
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
db = synth.make_db(n_loci=150, genes_per_locus=20, n_core=4, seed=1)
gi = mapper.GeneIndex(db.genes, device=0)
wl = workload.make_device_workload(db, n, 5_000_000, seed=1000, device="cuda:0", first_index=0)
host = torch.empty(n * 5_000_000, dtype=torch.uint8).pin_memory(); host.copy_(wl.ascii[: n * 5_000_000]); torch.cuda.synchronize()
L = load()
off, ln, acs = wl.contig_off, wl.contig_len, wl.asm_contig_start
for rep in range(4):
    t0 = time.perf_counter()
    b = mapper.AssemblyBatch(host.data_ptr(), off, ln, acs, device=0)
    t1 = time.perf_counter()
    r = gi.map(b, fetch=True)
    t2 = time.perf_counter()
    print(f"n={n} batch_create {1e3*(t1-t0):.1f} ms ({n*5e6/(t1-t0)/1e9:.1f} GB/s)  map+fetch {1e3*(t2-t1):.1f} ms  stages {r.stage_ms}")
plans = sys.argv[2].split("/") if len(sys.argv) > 2 else ["500", "150,425", "500", "150,425"]
for plan in plans:  # slab plans "a,b,c" (the last size repeats); a leading "t3:" uses three host threads
    thr = "2"
    if plan.startswith("t3:"): thr, plan = "3", plan[3:]
    os.environ["KAPTIVE_B200_SLAB_PLAN"] = plan
    os.environ["KAPTIVE_B200_SLAB_THREADS"] = thr
    cap = 1024 * n
    h, arrays = mapper.alloc_hits(cap); cig = np.zeros(cap * 16, dtype=np.uint32)
    ts = []
    for rep in range(6):
        nh, ncg = C.c_int64(0), C.c_int64(0)
        t0 = time.perf_counter()
        check(L.kb_map_assemblies(gi._h, C.c_void_p(host.data_ptr()), ptr(off), ptr(ln), ptr(acs), n, C.byref(h), C.byref(nh), ptr(cig), len(cig), C.byref(ncg)))
        ts.append(1e3 * (time.perf_counter() - t0))
    print(f"plan={plan} threads={thr}: " + " ".join(f"{t:.0f}" for t in ts) + f" ms; best {n / min(ts) * 1e3:.0f} asm/s, median {n / sorted(ts)[len(ts) // 2] * 1e3:.0f} asm/s, hits {nh.value}")
