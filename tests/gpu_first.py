"""First GPU bring-up script: map small synthetic assemblies through the C-ABI and diff against the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ["KAPTIVE_B200_KEEP_STAGES"] = "1"
os.environ["KAPTIVE_B200_FORCE_CENSUS"] = "1"
import numpy as np
import oracle_lib as ol
from kaptive_b200 import synth, mapper

db = synth.make_db(n_loci=10, genes_per_locus=8, n_core=2, seed=1)
odb = ol.OracleDB(*db.flat())
gi = mapper.GeneIndex(db.genes)
print("index: genes", gi.n_genes, "minimizers", gi.n_minimizers)
asms = [synth.make_assembly(db, s % 10, seed=3000 + s, genome_len=300000, mean_contigs=12) for s in range(6)]
t = time.time()
res = gi.map_contigs([[c for _, c in a.contigs] for a in asms])
print("gpu map: %.3fs" % (time.time() - t), res.stage_ms, res.counters, "mid_occ", res.mid_occ)
ok = True
for ai, a in enumerate(asms):
    ro = odb.map(*a.flat(), keep_stages=True)
    sel = res.hits["asm_id"] == ai
    n = int(sel.sum())
    if res.mid_occ[ai] != ro["mid_occ"]:
        print("asm", ai, "mid_occ differs", res.mid_occ[ai], ro["mid_occ"]); ok = False
    ga = res.anchors[res.anchors[:, 0] == ai][:, 1:] if res.anchors is not None else np.zeros((0, 6), np.int32)
    oa = np.stack([ro["anchors"][k] for k in ("gene", "rev", "rid", "tpos", "qpos", "flags")], axis=1) if len(ro["anchors"]) else np.zeros((0, 6), np.int32)
    if ga.shape != oa.shape or not np.array_equal(ga, oa):
        print("asm", ai, "ANCHORS differ", ga.shape, oa.shape); ok = False
    gc = res.chains[res.chains[:, 0] == ai][:, 1:] if res.chains is not None else np.zeros((0, 9), np.int32)
    oc = np.stack([ro["chains"][k] for k in ("gene", "score", "cnt", "rev", "rid", "rs", "re", "qs", "qe")], axis=1) if len(ro["chains"]) else np.zeros((0, 9), np.int32)
    if gc.shape != oc.shape or not np.array_equal(gc, oc):
        print("asm", ai, "CHAINS differ", gc.shape, oc.shape); ok = False
    oh = ro["hits"]
    if n != len(oh):
        print("asm", ai, "hit count differs", n, len(oh)); ok = False; continue
    idx = np.nonzero(sel)[0]
    for k, i in enumerate(idx):
        for f, of in (("gene", "gene"), ("q_start", "q_start"), ("q_end", "q_end"), ("t_ctg", "t_ctg"), ("t_len", "t_len"), ("t_start", "t_start"),
                      ("t_end", "t_end"), ("strand", "strand"), ("score", "score"), ("matches", "matches"), ("block_len", "block_len"),
                      ("edit_distance", "edit_distance"), ("mapq", "mapq"), ("is_primary", "is_primary")):
            if int(res.hits[f][i]) != int(oh[of][k]):
                print("asm", ai, "hit", k, f, int(res.hits[f][i]), int(oh[of][k])); ok = False
        c1 = res.cigar_of(i); c2 = ro["cigar"][oh["cigar_off"][k]: oh["cigar_off"][k] + oh["n_cigar"][k]]
        if not np.array_equal(c1, c2):
            print("asm", ai, "hit", k, "cigar differs", ol.cigar_string(c1)[:80], ol.cigar_string(c2)[:80]); ok = False
    print("asm", ai, "hits", n, "anchors", len(ga), "chains", len(gc))
print("PARITY", "OK" if ok else "FAIL")
sys.exit(0 if ok else 1)
