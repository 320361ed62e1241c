"""GPU tier (pytest -m gpu): the CUDA path, called through the C-ABI, against the committed golden vectors, the
oracle, and size-independent properties at full assembly size."""

import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

import cases
import oracle_lib as ol

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
GOLD = np.load(ROOT / "tests/golden/mapping_golden.npz")
ROWS = json.loads((ROOT / "tests/golden/kaptive_rows.json").read_text())
FIELDS = ("gene", "q_start", "q_end", "t_ctg", "t_len", "t_start", "t_end", "strand", "score", "matches", "block_len",
          "edit_distance", "mapq", "is_primary")


OWN_DB = ("long_tail",)  # cases with a gene set of their own: mapped by their own test, not in the shared batch


@pytest.fixture(scope="module")
def env():
    os.environ["KAPTIVE_B200_KEEP_STAGES"] = "1"
    os.environ["KAPTIVE_B200_FORCE_CENSUS"] = "1"
    from kaptive_b200 import mapper

    names = [n for n in cases.CASES if n not in OWN_DB]
    built = {n: cases.CASES[n]() for n in names}
    db = built[names[0]][0]
    assert all(b[0] is db for b in built.values())
    gi = mapper.GeneIndex(db.genes)
    res = gi.map_contigs([[s for _, s in built[n][1]] for n in names])
    yield {"mapper": mapper, "names": names, "built": built, "db": db, "gi": gi, "res": res}
    os.environ.pop("KAPTIVE_B200_KEEP_STAGES", None)
    os.environ.pop("KAPTIVE_B200_FORCE_CENSUS", None)


def check_against(res, ai, hits, cigar):
    idx = np.nonzero(res.hits["asm_id"] == ai)[0]
    assert len(idx) == len(hits)
    for f in FIELDS:
        assert np.array_equal(res.hits[f][idx].astype(np.int64), hits[f].astype(np.int64)), f
    for k, i in enumerate(idx):
        h = hits[k]
        assert np.array_equal(res.cigar_of(i), cigar[h["cigar_off"] : h["cigar_off"] + h["n_cigar"]])


@pytest.mark.parametrize("name", [n for n in cases.CASES if n not in OWN_DB])
def test_gpu_matches_golden(env, name):
    ai = env["names"].index(name)
    check_against(env["res"], ai, GOLD[f"{name}/hits"], GOLD[f"{name}/cigar"])
    mid, n_mz, n_anchor = (int(x) for x in GOLD[f"{name}/meta"])
    assert int(env["res"].mid_occ[ai]) == mid
    chains = env["res"].chains
    got = chains[chains[:, 0] == ai][:, 1:] if chains is not None else np.zeros((0, 9), np.int32)
    want = GOLD[f"{name}/chains"]
    want = np.stack([want[k] for k in want.dtype.names], axis=1) if len(want) else np.zeros((0, 9), np.int32)
    assert np.array_equal(got, want)
    anchors = env["res"].anchors
    assert (0 if anchors is None else int((anchors[:, 0] == ai).sum())) == n_anchor


def test_minimizer_count_matches_golden(env):
    assert env["res"].counters["minimizers"] == sum(int(GOLD[f"{n}/meta"][1]) for n in env["names"])


def test_batch_composition_does_not_change_results(env):
    """Mapping an assembly alone == mapping it inside a batch (assemblies are independent units)."""
    for name in ("mutated1", "mosaic_gene", "tiny_contigs"):
        contigs = env["built"][name][1]
        r1 = env["gi"].map_contigs([[s for _, s in contigs]])
        check_against(r1, 0, GOLD[f"{name}/hits"], GOLD[f"{name}/cigar"])


def test_map_into_caller_arrays_equals_fresh_arrays(env):
    contigs = env["built"]["mutated1"][1]
    b = env["mapper"].AssemblyBatch.from_contigs([[s for _, s in contigs]])
    h, arrays = env["mapper"].alloc_hits(4096)
    out = (h, arrays, np.zeros(1 << 16, dtype=np.uint32))
    r0, r1 = env["gi"].map(b), env["gi"].map(b, out=out)
    assert len(r0) == len(r1) > 0
    for k in r0.hits:
        assert np.array_equal(r0.hits[k], r1.hits[k]), k
    assert np.array_equal(r0.cigar, r1.cigar)
    tiny = (env["mapper"].alloc_hits(1)[0], {}, np.zeros(1, dtype=np.uint32))
    with pytest.raises(Exception):
        env["gi"].map(b, out=tiny)


def test_scan_minimizers_equal_sequential_sketch(env):
    name = "boundaries"
    contigs = env["built"][name][1]
    b = env["mapper"].AssemblyBatch.from_contigs([[s for _, s in contigs]])
    h, c, p = env["gi"].scan_minimizers(b, 0, 200_000)
    want = []
    for ci, (_, s) in enumerate(contigs):
        x, y = ol.sketch(s)
        want += [(int(a >> 8), ci, int(bb)) for a, bb in zip(x, y)]
    assert sorted(zip(h.tolist(), c.tolist(), p.tolist())) == sorted(want)


def test_scan_minimizers_adversarial_sequences(env):
    """Tandem repeats, two-letter sequence, inverted repeats, N runs, contigs around the w + k thresholds and across chunk
    boundaries: the position-parallel sketch must emit mm_sketch's multiset, identical-k-mer rules included."""
    sys.path.insert(0, str(ROOT / "scripts"))
    from proto_sketch_parallel import rand_seq

    rng = np.random.default_rng(5)
    lens = [0, 1, 14, 15, 23, 24, 25, 26, 31, 32, 33, 40, 63, 64, 65, 100, 300, 700, 8191, 8192, 8193, 8215, 8217, 20000, 70000]
    contigs = [rand_seq(rng, n, t % 5) for t, n in enumerate(lens * 3)]
    contigs += [b"A" * 9000, b"AC" * 4600, (b"ACGTTGCA" * 40 + b"N") * 30, b"N" * 50 + rand_seq(rng, 500, 0) + b"N" * 50]
    b = env["mapper"].AssemblyBatch.from_contigs([[c for c in contigs if len(c)]])
    h, c, p = env["gi"].scan_minimizers(b, 0, 2_000_000)
    want = []
    for ci, s in enumerate(c for c in contigs if len(c)):
        x, y = ol.sketch(s)
        want += [(int(a >> 8), ci, int(bb)) for a, bb in zip(x, y)]
    assert len(h) == len(want)
    assert sorted(zip(h.tolist(), c.tolist(), p.tolist())) == sorted(want)


def test_host_buffer_entry_point_equals_batch_path(env):
    from kaptive_b200 import _lib

    name = "mutated2"
    seqs, off, ln = cases.flat_contigs(env["built"][name][1])
    acs = np.array([0, len(ln)], dtype=np.int32)
    h, arrays = env["mapper"].alloc_hits(4096)
    cig = np.zeros(1 << 16, dtype=np.uint32)
    nh, nc = C.c_int64(0), C.c_int64(0)
    L = _lib.load()
    _lib.check(L.kb_map_assemblies(env["gi"]._h, _lib.ptr(seqs), _lib.ptr(off), _lib.ptr(ln), _lib.ptr(acs), 1, C.byref(h), C.byref(nh),
                                   _lib.ptr(cig), len(cig), C.byref(nc)))
    g = GOLD[f"{name}/hits"]
    assert nh.value == len(g)
    for f in FIELDS:
        assert np.array_equal(arrays[f][: nh.value].astype(np.int64), g[f].astype(np.int64)), f


def test_index_image_roundtrip(env):
    img = env["gi"].serialize()
    gi2 = env["mapper"].GeneIndex.deserialize(img)
    assert (gi2.n_genes, gi2.n_minimizers) == (env["gi"].n_genes, env["gi"].n_minimizers)
    name = "two_loci"
    r = gi2.map_contigs([[s for _, s in env["built"][name][1]]])
    check_against(r, 0, GOLD[f"{name}/hits"], GOLD[f"{name}/cigar"])


def test_capacity_and_argument_errors(env):
    from kaptive_b200 import _lib

    L = _lib.load()
    seqs, off, ln = cases.flat_contigs(env["built"]["exact"][1])
    acs = np.array([0, len(ln)], dtype=np.int32)
    h, _ = env["mapper"].alloc_hits(1)
    nh, nc = C.c_int64(0), C.c_int64(0)
    rc = L.kb_map_assemblies(env["gi"]._h, _lib.ptr(seqs), _lib.ptr(off), _lib.ptr(ln), _lib.ptr(acs), 1, C.byref(h), C.byref(nh), None, 0, C.byref(nc))
    assert rc == -4 and b"smaller" in L.kb_last_error()
    assert nh.value == len(GOLD["exact/hits"])  # the size is still reported so the caller can retry


def test_rammappy_shim_drop_in(env):
    """The module the reference imports (serotyping/core.py:15-16): Index.build -> Aligner -> map_batch -> hit iterators."""
    sys.path.insert(0, str(ROOT / "kaptive_b200" / "shim"))
    import rammappy
    from rammappy.align import Aligner

    name = "clean_typeable"
    db, contigs = env["built"][name]
    recs = rammappy.fasta.parse_fasta_bytes(cases.fasta_bytes(contigs))
    assert [(n, s) for n, s in recs] == [(n, s) for n, s in contigs]
    index = rammappy.Index.build([(n.encode(), s) for n, s in recs])
    aligner = Aligner(index=index, preset=None, do_cigar=True, do_cs=False, do_md=False)
    opts = aligner.options
    opts.filtering.best_n = 50000
    opts.filtering.pri_ratio = 0.0
    aligner.options = opts
    queries = [(str(i).encode(), g) for i, g in enumerate(db.genes)]
    its = aligner.map_batch(queries)
    assert len(its) == len(queries)
    g = GOLD[f"{name}/hits"]
    k = 0
    for qi, it in enumerate(its):
        for h in it:
            w = g[k]
            assert w["gene"] == qi
            assert h.target_name == contigs[w["t_ctg"]][0].encode()
            assert (h.query_start, h.query_end, h.target_len, h.target_start, h.target_end) == (w["q_start"], w["q_end"], w["t_len"], w["t_start"], w["t_end"])
            assert ("Forward" in repr(h.strand)) == (w["strand"] > 0)
            assert (h.block_len, h.matches, h.edit_distance, h.score, h.mapq, h.is_primary) == (w["block_len"], w["matches"], w["edit_distance"], w["score"], w["mapq"], bool(w["is_primary"]))
            cg = GOLD[f"{name}/cigar"][w["cigar_off"] : w["cigar_off"] + w["n_cigar"]]
            assert h.cigar == ol.cigar_string(cg).encode()
            k += 1
    assert k == len(g)
    # identical hits => the unmodified reference pipeline reports exactly this row (generated in the build container)
    assert ROWS[name]["best_locus"] == "KL7" and ROWS[name]["typeable"] is True


def test_full_size_assemblies_properties_and_oracle(env):
    """BASELINE-size inputs (5 Mb assemblies, 150 x 20 gene db): oracle equality on two assemblies and
    size-independent properties on all of them."""
    from kaptive_b200 import synth

    db = synth.make_db(n_loci=150, genes_per_locus=20, n_core=4, seed=1)
    gi = env["mapper"].GeneIndex(db.genes)
    asms = [synth.make_assembly(db, (7 * i) % 150, seed=1000 + i, genome_len=5_000_000) for i in range(6)]
    res = gi.map_contigs([[s for _, s in a.contigs] for a in asms])
    res2 = gi.map_contigs([[s for _, s in a.contigs] for a in asms])
    for f in FIELDS:  # idempotence / determinism
        assert np.array_equal(res.hits[f], res2.hits[f])
    odb = ol.OracleDB(*db.flat())
    for ai in (0, 3):
        ro = odb.map(*asms[ai].flat())
        check_against(res, ai, ro["hits"], ro["cigar"])
    h = res.hits
    lens = np.array([len(g) for g in db.genes])
    assert np.all(h["q_start"] >= 0) and np.all(h["q_end"] <= lens[h["gene"]]) and np.all(h["q_start"] < h["q_end"])
    assert np.all(h["t_start"] >= 0) and np.all(h["t_end"] <= h["t_len"]) and np.all(h["matches"] <= h["block_len"])
    assert np.all(h["mapq"] <= 60)
    for ai, a in enumerate(asms):  # hits arrive ordered by gene within an assembly
        sel = h["asm_id"] == ai
        assert np.all(np.diff(h["gene"][sel]) >= 0)
        # every non-core gene of the embedded locus is found, almost end to end
        mine = np.nonzero(db.gene_locus == a.locus)[0]
        for g in mine[2:-2]:
            cov = (h["q_end"] - h["q_start"])[sel & (h["gene"] == g)]
            assert len(cov) and cov.sum() >= 0.6 * lens[g]
    # CIGAR consistency: query / target spans
    for i in range(0, len(res), 37):
        cg = res.cigar_of(i)
        ops, ln = cg & 0xF, cg >> 4
        assert int(ln[(ops == 0) | (ops == 1)].sum()) == int(h["q_end"][i] - h["q_start"][i])
        assert int(ln[(ops == 0) | (ops == 2)].sum()) == int(h["t_end"][i] - h["t_start"][i])


def test_divergence_ladder_equals_oracle(env):
    """Loci embedded at 0-14 % substitutions and 0-4 % indels: exercises the certified band pass of the gap fills on
    both sides of its certificate (accepted, and rejected -> full DP) plus z-drop splits; every hit equals the oracle's."""
    from kaptive_b200 import synth

    db = synth.make_db(n_loci=12, genes_per_locus=10, n_core=3, seed=11)
    gi = env["mapper"].GeneIndex(db.genes)
    asms = []
    for i in range(28):
        s = 0.005 * i
        asms.append(synth.make_assembly(db, i % 12, seed=7000 + i, genome_len=200_000, mean_contigs=4, sub=(s, s + 0.005),
                                        indel=(0.0015 * i, 0.0015 * i + 0.001)))
    res = gi.map_contigs([[s for _, s in a.contigs] for a in asms])
    odb = ol.OracleDB(*db.flat())
    total = 0
    for ai, a in enumerate(asms):
        ro = odb.map(*a.flat())
        check_against(res, ai, ro["hits"], ro["cigar"])
        total += len(ro["hits"])
    assert total > 200


def test_host_buffer_entry_point_slabs_equal_batch_path(env, monkeypatch):
    """kb_map_assemblies cuts large inputs into slabs that overlap copies and kernels; hits must come back in assembly
    order with the same values as one batch (slab size forced down to 2 assemblies so that 7 assemblies make 4 slabs)."""
    from kaptive_b200 import _lib

    monkeypatch.setenv("KAPTIVE_B200_SLAB", "2")
    names = env["names"][:7]
    seqs, off, ln, acs = [], [], [], [0]
    base = 0
    for n in names:
        s, o, l = cases.flat_contigs(env["built"][n][1])
        seqs.append(s), off.append(o + base), ln.append(l)
        base += len(s)
        acs.append(acs[-1] + len(l))
    seqs, off, ln = np.concatenate(seqs), np.concatenate(off).astype(np.int64), np.concatenate(ln).astype(np.int32)
    acs = np.array(acs, dtype=np.int32)
    h, arrays = env["mapper"].alloc_hits(1 << 14)
    cig = np.zeros(1 << 18, dtype=np.uint32)
    nh, nc = C.c_int64(0), C.c_int64(0)
    L = _lib.load()
    _lib.check(L.kb_map_assemblies(env["gi"]._h, _lib.ptr(seqs), _lib.ptr(off), _lib.ptr(ln), _lib.ptr(acs), len(names), C.byref(h),
                                   C.byref(nh), _lib.ptr(cig), len(cig), C.byref(nc)))
    k = 0
    for ai, n in enumerate(names):
        g, gc = GOLD[f"{n}/hits"], GOLD[f"{n}/cigar"]
        sel = slice(k, k + len(g))
        assert np.all(arrays["asm_id"][sel] == ai)
        for f in FIELDS:
            assert np.array_equal(arrays[f][sel].astype(np.int64), g[f].astype(np.int64)), (n, f)
        for j in range(len(g)):
            co, ncg = int(arrays["cigar_off"][k + j]), int(arrays["n_cigar"][k + j])
            assert np.array_equal(cig[co : co + ncg], gc[g["cigar_off"][j] : g["cigar_off"][j] + g["n_cigar"][j]])
        k += len(g)
    assert k == nh.value


def test_combined_k_and_o_database_and_fragmentation_ladder_equal_oracle(env):
    """BASELINE.json configs[2] and configs[4] as parity cases: a combined K+O-shaped gene set (150 x 20 genes would be
    the bench size; here 14 x 10 K genes + 5 x 6 O genes + 4 extra genes in ONE index), assemblies carrying both a K and an O
    locus, and a depth-like fragmentation ladder (mean contig count 3 -> 400 on a 300 kb genome, i.e. contig N50 from
    ~100 kb down to < 1 kb, the locus broken into 1-8+ pieces).  Every hit equals the oracle's, so any call the unmodified
    typing logic derives from the hits (best locus, typeable) is concordant by construction."""
    from kaptive_b200 import synth

    k = synth.make_db(n_loci=14, genes_per_locus=10, n_core=3, seed=21, prefix="KL")
    o = synth.make_db(n_loci=5, genes_per_locus=6, n_core=2, n_extra=4, seed=22, prefix="OL")
    genes = k.genes + o.genes
    loci = k.loci + o.loci
    db = synth.SynthDB(genes=genes, gene_locus=np.concatenate([k.gene_locus, o.gene_locus + len(k.loci)]),
                       gene_pos=np.concatenate([k.gene_pos, o.gene_pos]), gene_start=np.concatenate([k.gene_start, o.gene_start]),
                       gene_end=np.concatenate([k.gene_end, o.gene_end]), gene_strand=np.concatenate([k.gene_strand, o.gene_strand]),
                       extra=np.concatenate([k.extra, o.extra]), loci=loci, locus_names=k.locus_names + o.locus_names)
    gi = env["mapper"].GeneIndex(db.genes)
    asms = []
    ladder = [3, 6, 12, 25, 50, 100, 200, 400]
    for i, mc in enumerate(ladder * 2):
        kl, ol_ = i % 14, 14 + i % 5
        asms.append(synth.make_assembly(db, kl, seed=8100 + i, genome_len=300_000, mean_contigs=mc, extra_loci=(ol_,),
                                        sub=(0.0, 0.04), indel=(0.0, 0.004)))
    res = gi.map_contigs([[s for _, s in a.contigs] for a in asms])
    odb = ol.OracleDB(*db.flat())
    n_k = n_o = 0
    for ai, a in enumerate(asms):
        ro = odb.map(*a.flat())
        check_against(res, ai, ro["hits"], ro["cigar"])
        g = ro["hits"]["gene"]
        n_k += int((g < len(k.genes)).sum())
        n_o += int((g >= len(k.genes)).sum())
    assert n_k > 100 and n_o > 40


def test_batch_fasta_ingest_feeds_the_mapper(env):
    """FASTA bytes -> kb_fasta_ingest_* (host threads) -> kb_batch_create -> kb_map_batch gives the golden hits."""
    names = ["fragmented", "lowercase", "n_rich", "empty_assembly", "two_loci"]
    files = [cases.fasta_bytes(env["built"][n][1]) for n in names]
    batch = env["mapper"].AssemblyBatch.from_fasta(files, threads=4)
    assert [len(x) for x in batch.contig_names] == [sum(1 for _ in env["built"][n][1]) for n in names]
    res = env["gi"].map(batch, fetch=True)
    for ai, n in enumerate(names):
        check_against(res, ai, GOLD[f"{n}/hits"], GOLD[f"{n}/cigar"])


def test_packed_ingest_from_fasta_gz_feeds_the_mapper(env, tmp_path, monkeypatch):
    """FASTA(.gz) files -> ingest.read_fasta_files (the reference's openers) -> kb_fasta_ingest_pack (2 bit + N mask on the host)
    -> kb_batch_create_packed / kb_map_assemblies_packed (slabs of 2 assemblies) -> the golden hits: the host packer and the device
    pack kernel are interchangeable."""
    import gzip

    from kaptive_b200 import ingest

    names = ["fragmented", "lowercase", "n_rich", "empty_assembly", "two_loci", "tiny_contigs", "boundaries"]
    paths = []
    for k, n in enumerate(names):
        data = cases.fasta_bytes(env["built"][n][1])
        p = tmp_path / (f"{n}.fasta.gz" if k % 2 == 0 else f"{n}.fna")
        p.write_bytes(gzip.compress(data) if k % 2 == 0 else data)
        paths.append(p)
    blobs, ids = ingest.read_fasta_files(paths, threads=4)
    assert ids == names
    pb = ingest.ingest_fasta_packed(blobs, threads=4)
    res = env["gi"].map(env["mapper"].AssemblyBatch.from_packed(pb), fetch=True)
    for ai, n in enumerate(names):
        check_against(res, ai, GOLD[f"{n}/hits"], GOLD[f"{n}/cigar"])
    monkeypatch.setenv("KAPTIVE_B200_SLAB", "2")
    res2 = env["gi"].map_packed(pb)
    for ai, n in enumerate(names):
        check_against(res2, ai, GOLD[f"{n}/hits"], GOLD[f"{n}/cigar"])


def test_dp_over_the_fast_limit_goes_through_the_full_size_kernel(env):
    """long_tail: a 1.9 k x 3.8 k end extension (7.2 M cells) is over what the staged kernels keep scratch for (4 M) and under
    minimap2's max_sw_mat (100 M): the chain is handed to the full-size kernel and the hit equals the oracle's golden."""
    db, contigs = cases.CASES["long_tail"]()
    gi = env["mapper"].GeneIndex(db.genes)
    res = gi.map_contigs([[s for _, s in contigs]])
    check_against(res, 0, GOLD["long_tail/hits"], GOLD["long_tail/cigar"])
    assert res.counters["slow_chains"] >= 1
    g = len(db.genes) - 1
    h = res.hits
    assert int(h["q_end"][h["gene"] == g][0]) == 2300  # with the extension skipped (a 4 M limit) it would stop at 2291


def test_index_sidecar_round_trip(env, tmp_path):
    """kaptive_b200.sidecar: the gene index cached beside the compiled database (db/manager.py:539-558 caches <kw>.pkl the same way)
    maps exactly like a freshly built one; a stale key (other genes) is rejected and rebuilt."""
    from kaptive_b200 import sidecar

    db = env["db"]
    pkl = tmp_path / "synth_k.pkl"
    gi1, cached1 = sidecar.index_for(pkl, db.genes)
    gi2, cached2 = sidecar.index_for(pkl, db.genes)
    assert (cached1, cached2) == (False, True) and sidecar.sidecar_path(pkl).exists()
    n = "fragmented"
    for gi in (gi1, gi2):
        res = gi.map_contigs([[s for _, s in env["built"][n][1]]])
        check_against(res, 0, GOLD[f"{n}/hits"], GOLD[f"{n}/cigar"])
    assert sidecar.load(sidecar.sidecar_path(pkl), db.genes[:-1]) is None
    sidecar.sidecar_path(pkl).write_bytes(b"garbage")
    assert sidecar.index_for(pkl, db.genes)[1] is False


def test_one_warp_per_chain_kernel_alone_gives_the_same_hits(env, monkeypatch):
    """KAPTIVE_B200_STAGED=0 sends every chain through kb_align_kernel (all of mm_align1 in one warp), the kernel that
    otherwise only serves the chains the staged path hands back: it has to stay bit-identical."""
    monkeypatch.setenv("KAPTIVE_B200_STAGED", "0")
    names = ["mutated1", "divergent", "big_indel", "mosaic_gene", "boundaries"]
    res = env["gi"].map_contigs([[s for _, s in env["built"][n][1]] for n in names])
    assert res.counters["slow_chains"] == 0  # the staged path did not run at all
    for ai, n in enumerate(names):
        check_against(res, ai, GOLD[f"{n}/hits"], GOLD[f"{n}/cigar"])


def test_batched_census_equals_the_per_assembly_census(monkeypatch):
    """Assemblies that carry 14 copies of a database gene take the occurrence census (mm_idx_cal_max_occ).  The batched form (one scan,
    two sorts and a run-length encode per group of assemblies) gives every assembly the mid_occ of the per-assembly form, groups of
    every size included, and the hits do not change."""
    import torch

    from kaptive_b200 import mapper, synth, workload

    db = synth.make_db(n_loci=6, genes_per_locus=6, n_core=1, seed=5)
    wl = workload.make_device_workload(db, 9, 300_000, mean_contigs=6, seed=77, device="cuda:0", repeats=14, repeat_seq=db.genes[3])
    gi = mapper.GeneIndex(db.genes, device=0)
    batch = mapper.AssemblyBatch(wl.ascii.data_ptr(), wl.contig_off, wl.contig_len, wl.asm_contig_start, device=0)
    results = {}
    for name, env in (("serial", {"KAPTIVE_B200_CENSUS_SERIAL": "1"}), ("batched", {}), ("small groups", {"KAPTIVE_B200_CENSUS_GROUP": "700000"})):
        for k in ("KAPTIVE_B200_CENSUS_SERIAL", "KAPTIVE_B200_CENSUS_GROUP"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        r = gi.map(batch)
        results[name] = (np.array(r.mid_occ), {k: np.array(v) for k, v in r.hits.items()}, np.array(r.cigar))
    torch.cuda.synchronize()
    ref = results["serial"]
    assert len(ref[1]["gene"]) > 0 and (ref[0] >= 10).all()
    for name in ("batched", "small groups"):
        got = results[name]
        assert np.array_equal(got[0], ref[0]), (name, got[0], ref[0])
        assert all(np.array_equal(got[1][k], ref[1][k]) for k in ref[1]) and np.array_equal(got[2], ref[2]), name


def test_batched_census_with_only_some_assemblies_flagged(monkeypatch):
    """Only assemblies with repeats need the census: the batched form groups runs of flagged assemblies (bridging gaps of up to four
    unflagged ones, which are scanned but not evaluated) and must leave every other assembly at the floor."""
    from kaptive_b200 import mapper, synth, workload

    db = synth.make_db(n_loci=6, genes_per_locus=6, n_core=1, seed=5)
    n = 16
    rep = workload.make_device_workload(db, n, 200_000, mean_contigs=4, seed=91, device="cuda:0", repeats=14, repeat_seq=db.genes[3])
    plain = workload.make_device_workload(db, n, 200_000, mean_contigs=4, seed=91, device="cuda:0")
    flagged = {0, 1, 4, 5, 14}  # a run, a gap of two (bridged), a gap of eight (a new group)
    asms = []
    for a in range(n):
        seq, off, ln = (rep if a in flagged else plain).host_assembly(a)
        asms.append([seq[o : o + k].tobytes() for o, k in zip(off, ln)])
    gi = mapper.GeneIndex(db.genes, device=0)
    batch = mapper.AssemblyBatch.from_contigs(asms, device=0)
    out = {}
    for name, env in (("serial", {"KAPTIVE_B200_CENSUS_SERIAL": "1"}), ("batched", {})):
        monkeypatch.delenv("KAPTIVE_B200_CENSUS_SERIAL", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        r = gi.map(batch)
        out[name] = (np.array(r.mid_occ), {k: np.array(v) for k, v in r.hits.items()}, np.array(r.cigar))
    a, b = out["serial"], out["batched"]
    assert np.array_equal(a[0], b[0]), (a[0], b[0])
    assert all(np.array_equal(a[1][k], b[1][k]) for k in a[1]) and np.array_equal(a[2], b[2])
    assert len(a[1]["gene"]) > 0
