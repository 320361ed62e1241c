// kb_host_emul.cpp -- TEST INFRASTRUCTURE: runs the product's device logic
// (kaptive_b200/csrc/*.cuh, the KB_HD functions the CUDA kernels are made of) sequentially
// on the host, so that the CPU-only test tier can compare it with the oracle before any GPU
// time is spent.  It is NOT a fallback: nothing in kaptive_b200/ links or loads it.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../kaptive_b200/csrc/kb_host.h"
#include "../../kaptive_b200/csrc/kb_scan.cuh"
#include "../../kaptive_b200/csrc/kb_final.cuh"
#include "../../kaptive_b200/csrc/kb_stage.cuh"

// same helpers the scan kernel defines (kb_scan.cu is a .cu file; restated here for the host build)
static inline bool ht_lookup(const uint64_t *ht, uint32_t ht_mask, uint32_t hash, uint32_t *start, uint32_t *count)
{
    uint32_t slot = hash & ht_mask;
    for (;;) {
        uint64_t e = ht[slot];
        if (e == 0) return false;
        if ((uint32_t)(e >> KB_HT_KEY_SHIFT) == hash) {
            *start = (uint32_t)(e >> KB_HT_START_SHIFT) & KB_HT_START_MASK;
            *count = (uint32_t)e & KB_HT_COUNT_MASK;
            return true;
        }
        slot = (slot + 1) & ht_mask;
    }
}

struct EmuHit {
    int32_t gene, q_start, q_end, t_ctg, t_len, t_start, t_end, strand, score, matches, block_len, edit_distance, mapq,
        is_primary, dp_max, chain_score, chain_cnt, cigar_off, n_cigar;
};
struct EmuAnchor { int32_t gene, rev, rid, tpos, qpos, flags; };
struct EmuChain { int32_t gene, score, cnt, rev, rid, rs, re, qs, qe; };

struct EmuResult {
    std::vector<EmuHit> hits;
    std::vector<uint32_t> cigar;
    std::vector<EmuAnchor> anchors;
    std::vector<EmuChain> chains;
    std::vector<uint32_t> mz_hash, mz_pos;
    std::vector<int32_t> mz_ctg;
    int32_t mid_occ = 0;
    int64_t n_minimizers = 0;
    int64_t stage_ok = 0, stage_fallback = 0, stage_mismatch = 0;  // staged path (kb_stage.cuh) vs kb_align1, per chain
};

struct PackedFetch {
    const uint32_t *seq2, *nmask;
    int64_t soff;
    int operator()(int i) const { return kb_fetch_base(seq2, nmask, soff + i); }
};

extern "C" {

void *kbe_index_create(const uint8_t *seqs, const int64_t *off, const int32_t *len, int32_t n, const kb_params_t *p)
{
    KbHostIndex *h = new KbHostIndex();
    std::string err = h->build(seqs, off, len, n, *p);
    if (!err.empty()) {
        fprintf(stderr, "kbe_index_create: %s\n", err.c_str());
        delete h;
        return nullptr;
    }
    // exercise the serializer on the way
    std::vector<uint8_t> img((size_t)h->serialized_size());
    h->serialize(img.data());
    KbHostIndex *h2 = new KbHostIndex();
    err = h2->deserialize(img.data(), (int64_t)img.size());
    delete h;
    if (!err.empty()) {
        fprintf(stderr, "kbe_index_create: %s\n", err.c_str());
        delete h2;
        return nullptr;
    }
    return h2;
}
void kbe_index_destroy(void *h) { delete (KbHostIndex *)h; }

// one assembly; lane_bases lets tests vary the slice size to stress the look-back logic
EmuResult *kbe_map_assembly(void *idx_, const uint8_t *ascii, const int64_t *ctg_off, const int32_t *ctg_len, int32_t n_ctg,
                            int32_t lane_bases, int32_t keep_stages)
{
    KbHostIndex &H = *(KbHostIndex *)idx_;
    KbIndexView ix = H.host_view();
    const kb_params_t &P = H.p;
    EmuResult *R = new EmuResult();
    int32_t acs[2] = {0, n_ctg};
    KbHostBatchLayout L;
    std::string err = L.build(ctg_off, ctg_len, acs, 1);
    if (!err.empty()) { fprintf(stderr, "layout: %s\n", err.c_str()); return R; }
    std::vector<uint32_t> seq2, nmask;
    kb_pack_host(ascii, ctg_off, L, seq2, nmask);
    KbBatchView bt;
    bt.n_asm = 1, bt.n_ctg = n_ctg, bt.n_chunks = (int64_t)L.chunk_ctg.size(), bt.total_bases = L.total_bases;
    bt.seq2 = seq2.data(), bt.nmask = nmask.data(), bt.ctg_soff = L.ctg_soff.data(), bt.ctg_len = L.ctg_len.data();
    bt.ctg_asm = L.ctg_asm.data(), bt.ctg_vstart = L.ctg_vstart.data(), bt.asm_ctg_start = L.asm_ctg_start.data();
    bt.chunk_ctg = L.chunk_ctg.data(), bt.chunk_start = L.chunk_start.data();

    // ---- scan: every slice of lane_bases bases independently, exactly as a GPU lane would
    std::vector<uint64_t> akey;
    std::vector<uint32_t> aval;
    std::vector<uint16_t> occ((size_t)ix.n_entries, 0);
    std::vector<uint32_t> all_hash;
    const bool fast = lane_bases < 0;  // negative slice size: run the branch-free sketch the scan kernel uses
    if (fast) lane_bases = -lane_bases;
    if (lane_bases == 0) lane_bases = KB_LANE_BASES;
    for (int c = 0; c < n_ctg; ++c) {
        PackedFetch fetch{bt.seq2, bt.nmask, bt.ctg_soff[c]};
        for (int s = 0; s < ctg_len[c]; s += lane_bases) {
            int e = std::min(ctg_len[c], s + lane_bases);
            auto emit = [&](uint32_t hx, uint32_t hy) {
                all_hash.push_back(hx);
                if (keep_stages) R->mz_hash.push_back(hx), R->mz_pos.push_back(hy), R->mz_ctg.push_back(c);
                uint32_t st, cnt;
                if (!ht_lookup(ix.ht, ix.ht_mask, hx, &st, &cnt)) return;
                for (uint32_t j = 0; j < cnt; ++j) {
                    const KbEntry &en = ix.ent[st + j];
                    int rev = (int)(en.qpos_z & 1u) != (int)(hy & 1u);
                    uint64_t key = ((uint64_t)0 << KB_KEY_ASM_SHIFT) | ((uint64_t)en.gene << KB_KEY_GENE_SHIFT) |
                                   ((uint64_t)rev << KB_KEY_REV_SHIFT) | (uint64_t)(bt.ctg_vstart[c] + (int)(hy >> 1));
                    akey.push_back(key), aval.push_back(st + j);
                    if (occ[st + j] < 0xffff) ++occ[st + j];
                }
            };
            if (fast) kb_fast_slice<10, 15>(ctg_len[c], s, e, fetch, emit);
            else kb_sketch_slice<10, 15>(ctg_len[c], s, e, fetch, emit);
        }
    }
    R->n_minimizers = (int64_t)all_hash.size();
    // ---- census (minimap2 mm_idx_cal_max_occ on the indexed assembly)
    int32_t mid = P.mid_occ;
    if (mid <= 0) {
        std::sort(all_hash.begin(), all_hash.end());
        std::vector<uint32_t> cnts;
        for (size_t st = 0, i = 1; i <= all_hash.size(); ++i)
            if (i == all_hash.size() || all_hash[i] != all_hash[st]) cnts.push_back((uint32_t)(i - st)), st = i;
        if (cnts.empty()) mid = INT32_MAX;
        else {
            std::sort(cnts.begin(), cnts.end());
            mid = (int32_t)(cnts[(uint32_t)((1. - P.mid_occ_frac) * cnts.size())] + 1);
        }
        if (mid < P.min_mid_occ) mid = P.min_mid_occ;
        if (P.max_mid_occ > P.min_mid_occ && mid > P.max_mid_occ) mid = P.max_mid_occ;
    }
    R->mid_occ = mid;
    // ---- sort by key only (what the radix sort does; payload order among equal keys is arbitrary: shuffle-proofed by fix-up)
    int64_t na = (int64_t)akey.size();
    std::vector<int64_t> ord((size_t)na);
    for (int64_t i = 0; i < na; ++i) ord[(size_t)i] = i;
    std::sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b) { return akey[(size_t)a] != akey[(size_t)b] ? akey[(size_t)a] < akey[(size_t)b] : aval[(size_t)a] > aval[(size_t)b]; });
    std::vector<uint64_t> skey((size_t)na);
    std::vector<uint32_t> sval((size_t)na);
    for (int64_t i = 0; i < na; ++i) skey[(size_t)i] = akey[(size_t)ord[(size_t)i]], sval[(size_t)i] = aval[(size_t)ord[(size_t)i]];
    // ---- groups
    std::vector<int64_t> gstart;
    for (int64_t i = 0; i < na; ++i)
        if (i == 0 || (skey[(size_t)i] >> KB_KEY_GENE_SHIFT) != (skey[(size_t)i - 1] >> KB_KEY_GENE_SHIFT)) gstart.push_back(i);
    gstart.push_back(na);
    int64_t ng = (int64_t)gstart.size() - 1;
    // ---- chain
    std::vector<uint32_t> wx((size_t)na + 1);
    std::vector<int32_t> wy((size_t)na + 1), wf((size_t)na + 1), wp((size_t)na + 1), wv((size_t)na + 1), wt((size_t)na + 1);
    std::vector<uint64_t> wz((size_t)na + 1), wu((size_t)na + 1), cx((size_t)na + 1), cy((size_t)na + 1);
    KbChainWork W{wx.data(), wy.data(), wf.data(), wp.data(), wv.data(), wt.data(), wz.data(), wu.data()};
    std::vector<KbGroupInfo> ginfo((size_t)ng + 1);
    std::vector<KbChainRec> chains((size_t)na / 3 + 16);
    unsigned long long n_chains = 0;
    for (int64_t g = 0; g < ng; ++g) {
        kb_chain_group(ix, bt, skey.data(), sval.data(), gstart[(size_t)g], gstart[(size_t)g + 1], occ.data(), &mid, W, cx.data(),
                       cy.data(), &ginfo[(size_t)g], chains.data(), &n_chains, (int64_t)chains.size(), (int32_t)g);
        if (keep_stages) {
            const KbGroupInfo &gi = ginfo[(size_t)g];
            for (int32_t i = 0; i < gi.n_seed; ++i) {
                uint32_t xv = wx[(size_t)(gstart[(size_t)g] + i)];
                int32_t yv = wy[(size_t)(gstart[(size_t)g] + i)];
                int32_t vpos = (int32_t)(xv & KB_VPOS_MASK);
                int32_t c = kb_vpos_to_ctg(bt, 0, vpos);
                R->anchors.push_back(EmuAnchor{gi.gene, (int32_t)(xv >> KB_KEY_REV_SHIFT), c, vpos - bt.ctg_vstart[c], yv & 0x3fffffff, (yv >> 30) & 1});
            }
        }
    }
    // NB: the anchor dump above is taken after chaining reused x/y?  x/y are not modified by chaining, only read.
    if (keep_stages)
        for (int64_t g = 0; g < ng; ++g) {
            const KbGroupInfo &gi = ginfo[(size_t)g];
            for (int32_t i = 0; i < gi.n_chains; ++i) {
                const KbChainRec &c = chains[(size_t)(gi.chain_base + i)];
                R->chains.push_back(EmuChain{gi.gene, c.score, c.cnt, c.rev, c.rid, c.rs, c.re, c.qs, c.qe});
            }
        }
    // ---- align: one "warp" of NL=1 lanes per chain, z-drop splits handled in place
    std::vector<uint8_t> scratch(kb_align_scratch_bytes(P.max_sw_cells));
    KbAlignScratch S = kb_align_scratch_at(scratch.data(), P.max_sw_cells);
    std::vector<KbRawHit> raw;
    std::vector<uint32_t> pool;
    int64_t cells = 0;
    for (unsigned long long ci = 0; ci < n_chains; ++ci) {
        const KbChainRec &c = chains[(size_t)ci];
        const KbGroupInfo &gi = ginfo[(size_t)c.group];
        KbReg r;
        memset(&r, 0, sizeof(r));
        int reg_idx = (int)((int64_t)ci - gi.chain_base);
        r.as = c.as, r.cnt = c.cnt, r.score = c.score, r.score0 = c.score0, r.mlen = c.mlen, r.blen = c.blen, r.parent = c.parent, r.id = reg_idx;
        r.hash = c.hash, r.rev = c.rev, r.rid = c.rid, r.rs = c.rs, r.re = c.re, r.qs = c.qs, r.qe = c.qe;
        int32_t subsc = c.subsc, n_sub = c.n_sub;
        // staged path (plan -> one DP per job -> assemble), checked against kb_align1 below
        KbReg rst = r;
        std::vector<uint32_t> cig_st;
        int staged = -1;
        {
            KbPlan pl;
            std::vector<KbJob> jobs(KB_JOBS_PER_CHAIN_MAX);
            std::vector<int32_t> K((size_t)gi.n_a + 8);
            KbJobCount count;
            KbJobWrite write{jobs.data(), (int32_t)ci};
            int nj = kb_stage_plan(ix, bt, gi.asm_id, gi.gene, r.as, r.cnt, r.mlen, gi.n_a, cx.data() + gi.a_base, cy.data() + gi.a_base,
                                   K.data(), pl, count);
            if (nj >= 0)
                nj = kb_stage_plan(ix, bt, gi.asm_id, gi.gene, r.as, r.cnt, r.mlen, gi.n_a, cx.data() + gi.a_base, cy.data() + gi.a_base,
                                   K.data(), pl, write);
            if (nj >= 0) {
                std::vector<uint32_t> jobcig;
                const uint8_t *gq = c.rev ? ix.gseq_rev : ix.gseq_fwd;
                for (int k = 0; k < nj; ++k) {
                    KbJob &J = jobs[(size_t)k];
                    std::vector<uint8_t> q((size_t)J.qlen), t((size_t)J.tlen);
                    const bool back = J.kind == KB_JOB_LEFT;
                    for (int x = 0; x < J.qlen; ++x) q[(size_t)x] = back ? gq[J.qbase + J.qoff - 1 - x] : gq[J.qbase + J.qoff + x];
                    for (int x = 0; x < J.tlen; ++x)
                        t[(size_t)x] = (uint8_t)kb_fetch_base(bt.seq2, bt.nmask, back ? J.tpos - 1 - x : J.tpos + x);
                    KbEz ez;
                    int64_t dummy = 0;
                    kb_extd2<1>(kb_dp_const(P), 0, J.qlen, q.data(), J.tlen, t.data(), J.w, J.zdrop, J.flag, ez, S, &dummy);
                    J.score = ez.score, J.max = ez.max, J.max_t = ez.max_t, J.max_q = ez.max_q, J.zdropped = ez.zdropped, J.n_cigar = ez.n_cigar;
                    J.cigar_off = (int64_t)jobcig.size(), J.state = 1;
                    for (int x = 0; x < ez.n_cigar; ++x) jobcig.push_back(S.ezcig[x]);
                }
                cig_st.assign(jobcig.size() + 8, 0);
                jobcig.push_back(0);
                staged = kb_stage_assemble(ix, bt, gi.gene, pl, jobs.data(), jobcig.data(), rst, cig_st.data());
            }
        }
        for (int split = 0;; ++split) {
            KbReg r2;
            memset(&r2, 0, sizeof(r2));
            int e = kb_align1<1>(ix, bt, 0, gi.asm_id, gi.gene, r, r2, gi.n_a, cx.data() + gi.a_base, cy.data() + gi.a_base, S, &cells);
            KbRawHit h;
            memset(&h, 0, sizeof(h));
            h.group = c.group, h.reg_idx = reg_idx, h.split_idx = split;
            h.cnt = r.cnt, h.score = r.score, h.score0 = r.score0, h.hash = r.hash;
            h.rev = r.rev, h.rid = r.rid, h.rs = r.rs, h.re = r.re, h.qs = r.qs, h.qe = r.qe;
            h.has_p = r.has_p, h.dp_score = r.dp_score, h.dp_max = r.dp_max, h.dp_max2 = 0, h.n_ambi = r.n_ambi, h.mlen = r.mlen, h.blen = r.blen;
            h.parent = r.parent, h.subsc = subsc, h.n_sub = n_sub, h.n_cigar = e ? 0 : r.n_cigar, h.err = e;
            h.cigar_off = (int64_t)pool.size();
            for (int i = 0; i < h.n_cigar; ++i) pool.push_back(S.cigar[i]);
            if (split == 0) {
                if (staged != 0) ++R->stage_fallback;
                else {
                    bool same = e == 0 && r2.cnt == 0 && rst.cnt == r.cnt && rst.score == r.score && rst.rs == r.rs && rst.re == r.re &&
                                rst.qs == r.qs && rst.qe == r.qe && rst.has_p == r.has_p && rst.dp_score == r.dp_score &&
                                rst.dp_max == r.dp_max && rst.n_ambi == r.n_ambi && rst.mlen == r.mlen && rst.blen == r.blen &&
                                rst.n_cigar == r.n_cigar;
                    for (int i = 0; same && i < r.n_cigar; ++i) same = cig_st[(size_t)i] == S.cigar[i];
                    if (same) ++R->stage_ok;
                    else ++R->stage_mismatch;
                }
            }
            raw.push_back(h);
            if (e == 0 && r2.cnt > 0) r = r2;
            else break;
        }
    }
    // ---- finalize per group (raw is already in (group, reg, split) order)
    std::vector<int32_t> fw(raw.size() + 1);
    std::vector<uint64_t> fcov(raw.size() + 1);
    for (size_t st = 0; st < raw.size();) {
        size_t en = st;
        while (en < raw.size() && raw[en].group == raw[st].group) ++en;
        const KbGroupInfo &gi = ginfo[(size_t)raw[st].group];
        KbHitView *hv = static_cast<KbHitView *>(&raw[st]);
        int kept = kb_finalize_group(P, hv, (int)(en - st), gi.rep_len, fw.data(), fcov.data());
        for (int i = 0; i < kept; ++i) {
            const KbHitView &h = hv[i];
            EmuHit o;
            o.gene = gi.gene, o.q_start = h.qs, o.q_end = h.qe, o.t_ctg = h.rid, o.t_len = ctg_len[h.rid], o.t_start = h.rs, o.t_end = h.re;
            o.strand = h.rev ? -1 : 1, o.score = h.dp_score, o.matches = h.mlen, o.block_len = h.blen;
            o.edit_distance = h.blen - h.mlen + h.n_ambi, o.mapq = h.mapq, o.is_primary = h.parent == i, o.dp_max = h.dp_max;
            o.chain_score = h.score0, o.chain_cnt = h.cnt, o.cigar_off = (int32_t)R->cigar.size(), o.n_cigar = h.n_cigar;
            for (int k = 0; k < h.n_cigar; ++k) R->cigar.push_back(pool[(size_t)(h.cigar_off + k)]);
            R->hits.push_back(o);
        }
        st = en;
    }
    return R;
}

int64_t kbe_result_counts(EmuResult *r, int64_t *out)
{
    out[0] = (int64_t)r->hits.size(), out[1] = (int64_t)r->cigar.size(), out[2] = (int64_t)r->anchors.size();
    out[3] = (int64_t)r->chains.size(), out[4] = r->mid_occ, out[5] = r->n_minimizers, out[6] = (int64_t)r->mz_hash.size();
    out[7] = r->stage_mismatch << 42 | r->stage_fallback << 21 | r->stage_ok;
    return 0;
}
void kbe_result_fetch(EmuResult *r, EmuHit *hits, uint32_t *cigar, EmuAnchor *anchors, EmuChain *chains, uint32_t *mzh, int32_t *mzc, uint32_t *mzp)
{
    if (hits && !r->hits.empty()) memcpy(hits, r->hits.data(), r->hits.size() * sizeof(EmuHit));
    if (cigar && !r->cigar.empty()) memcpy(cigar, r->cigar.data(), r->cigar.size() * 4);
    if (anchors && !r->anchors.empty()) memcpy(anchors, r->anchors.data(), r->anchors.size() * sizeof(EmuAnchor));
    if (chains && !r->chains.empty()) memcpy(chains, r->chains.data(), r->chains.size() * sizeof(EmuChain));
    if (mzh && !r->mz_hash.empty()) memcpy(mzh, r->mz_hash.data(), r->mz_hash.size() * 4);
    if (mzc && !r->mz_ctg.empty()) memcpy(mzc, r->mz_ctg.data(), r->mz_ctg.size() * 4);
    if (mzp && !r->mz_pos.empty()) memcpy(mzp, r->mz_pos.data(), r->mz_pos.size() * 4);
}
void kbe_result_free(EmuResult *r) { delete r; }
void kbe_params_default(kb_params_t *p);
}  // extern "C"
