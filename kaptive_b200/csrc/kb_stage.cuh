// kb_stage.cuh -- the staged form of the base-level alignment step (minimap2 mm_align1; reference call site
// src/kaptive/serotyping/core.py:154).
//
// kb_align1 (kb_align.cuh) walks one chain from its left extension over the gap fills to the right extension,
// one DP after the other, inside one warp.  Every one of those DP problems is fixed by the chain's anchors alone --
// only a z-drop inside a gap fill changes what comes next -- so the same work can be laid out as
//
//   plan      one thread per chain: trimmed anchors, seed filters, windows -> a list of DP jobs
//   DP        one warp per job, homogeneous kernels: certified band pass (gap fills), row-stripe wavefront (the rest)
//   assemble  one thread per chain: z-drop test of every fill, CIGAR concatenation, mm_update_extra -> raw hit
//
// with small kernels that stay inside the instruction cache and need a fraction of the monolith's registers.
// A chain whose plan or assembly meets anything outside the common case (a z-drop in a fill, a DP over a size
// limit, a zero-length segment, a capacity overflow) is handed, untouched, to kb_align1: the staged path never
// has to reproduce the rare branches, it only has to recognise them.  Everything here is KB_HD: tests/host_emul
// runs plan -> kb_extd2<1> per job -> assemble on the CPU and checks it against kb_align1 chain by chain.
#pragma once
#include "kb_align.cuh"

#define KB_JOB_LEFT 1
#define KB_JOB_FILL 2
#define KB_JOB_RIGHT 3
#define KB_JOBS_PER_CHAIN_MAX 48

struct KbJob {
    int32_t chain;   // index of the chain (KbChainRec) the job belongs to
    int32_t kind;    // KB_JOB_*
    int32_t qoff;    // query segment starts at qseq0[qoff]; a left extension reads qseq0[qoff - 1 - x]
    int32_t qlen, tlen;
    int32_t w, zdrop, flag;
    int64_t tpos;    // target segment starts at packed base tpos (storage offset); left extension: tpos - 1 - x
    int64_t qbase;   // offset of qseq0 in gseq_fwd / gseq_rev
    int32_t qrev;    // 1: gseq_rev
    // results
    int32_t score, max, max_t, max_q, zdropped, n_cigar;
    int32_t state;   // 0 = pending, 1 = done
    int64_t cigar_off;
};

struct KbPlan {
    int32_t ok;        // 1: staged path; 0: kb_align1 does the chain
    int32_t n_jobs;
    int64_t job_base;
    int32_t has_left, has_right;
    int32_t rs, qs, re, qe;  // first anchor (after the left extension starts) and last anchor, half-k shifted
    int32_t rev, qlen;
    int64_t soff;
};

// sequence views shared by the DP kernels and the assembly: plain bytes (gene sequences, nt4 codes) and the
// 2-bit + mask packing of the contigs, both readable forwards or backwards
struct KbByteSeq {
    const uint8_t *p;
    KB_HD int operator[](int x) const { return p[x]; }
    KB_HD KbByteSeq operator+(int d) const { return KbByteSeq{p + d}; }
};
struct KbPackSeq {  // remembers the last 32-base group it touched: the assembly code reads bases in sequence
    const uint32_t *seq2, *nmask;
    int64_t pos;
    mutable int64_t grp = -1;
    mutable uint32_t w0 = 0, w1 = 0, mw = 0;
    KB_HD int operator[](int x) const
    {
        const int64_t b = pos + x, g = b >> 5;
        if (g != grp) grp = g, w0 = seq2[2 * g], w1 = seq2[2 * g + 1], mw = nmask[g];
        const int r = (int)(b & 31);
        if ((mw >> r) & 1u) return 4;
        return (int)(((r < 16 ? w0 : w1) >> (2 * (r & 15))) & 3u);
    }
    KB_HD KbPackSeq operator+(int d) const { return KbPackSeq{seq2, nmask, pos + d}; }
};

// Plan one chain.  `K` is int32 scratch for the seed filters (>= cnt entries).  Every DP job is handed to
// sink(job) -> bool (false = no room); chain / job_base are the caller's business.  Returns the number of jobs, or -1
// when the chain has to take the kb_align1 path.  The seed-filter flags written into ay[] are the ones kb_align1
// would write (idempotent), so a chain can be planned twice (count, then emit) and still be handed to kb_align1.
template <class Sink>
KB_HD int kb_stage_plan(const KbIndexView &ix, const KbBatchView &bt, int asm_id, int gene, int r_as, int r_cnt, int r_mlen, int n_a,
                        const uint64_t *ax, uint64_t *ay, int32_t *K, KbPlan &pl, Sink &sink)
{
    const kb_params_t &P = ix.p;
    const int hk = P.k >> 1;
    int32_t as1, cnt1;
    pl.ok = 0, pl.n_jobs = 0;
    if (r_cnt == 0) return -1;
    const int bw = P.ext_bw;
    int bw_long = (int)(20000 * 1.5 + 1.);
    if (bw_long < bw) bw_long = bw;
    kb_fix_bad_ends(r_as, r_cnt, r_mlen, ax, ay, P.bw, P.min_chain_score * 2, &as1, &cnt1);
    kb_filter_bad_seeds(as1, cnt1, ax, ay, 10, 40, P.max_gap >> 1, 10, K);
    kb_filter_bad_seeds_alt(as1, cnt1, ax, ay, 30, P.max_gap >> 1, K);
    KbWin W;
    kb_align_window(ix, bt, asm_id, gene, r_as, r_cnt, as1, cnt1, n_a, ax, ay, W);
    if (W.re0 - W.rs0 > KB_TFULL_MAX || W.re0 <= W.rs0) return -1;
    const int64_t qbase = ix.gene_seq_off[gene];
    int nj = 0;
    auto add = [&](int kind, int qoff, int qlen, int64_t tpos, int tlen, int w, int zdrop, int flag) -> bool {
        if (nj >= KB_JOBS_PER_CHAIN_MAX) return false;
        const bool track = !(flag & KB_EZ_GLOBAL_NO_ZDROP);
        if (!kb_rows_eligible(P.max_sw_cells, qlen, tlen, w, track)) return false;
        KbJob j;
        j.chain = -1, j.kind = kind, j.qoff = qoff, j.qlen = qlen, j.tlen = tlen, j.w = w, j.zdrop = zdrop, j.flag = flag;
        j.tpos = tpos, j.qbase = qbase, j.qrev = W.rev;
        j.score = KB_NEG_INF, j.max = 0, j.max_t = j.max_q = -1, j.zdropped = 0, j.n_cigar = 0, j.state = 0, j.cigar_off = 0;
        if (!sink(nj, j)) return false;
        ++nj;
        return true;
    };
    int32_t rs = W.rs, qs = W.qs;
    pl.rs = rs, pl.qs = qs, pl.re = W.re, pl.qe = W.qe, pl.rev = W.rev, pl.qlen = W.qlen, pl.soff = W.soff;
    pl.has_left = (qs > 0 && rs > 0) ? 1 : 0;
    if (pl.has_left) {
        if (!add(KB_JOB_LEFT, qs, qs - W.qs0, W.soff + rs, rs - W.rs0, bw, P.zdrop, KB_EZ_EXTZ_ONLY | KB_EZ_RIGHT | KB_EZ_REV_CIGAR)) return -1;
    }
    for (int i = 1; i < cnt1; ++i) {  // gap filling
        if ((ay[as1 + i] & (KB_SEED_IGNORE | KB_SEED_TANDEM)) && i != cnt1 - 1) continue;
        const int32_t re = (int32_t)ax[as1 + i] - hk, qe = (int32_t)ay[as1 + i] - hk;
        if (i == cnt1 - 1 || (ay[as1 + i] & KB_SEED_LONG_JOIN) || (qe - qs >= P.min_ksw_len && re - rs >= P.min_ksw_len)) {
            int bw1 = bw_long;
            if (ay[as1 + i] & KB_SEED_LONG_JOIN) bw1 = qe - qs > re - rs ? qe - qs : re - rs;
            if (!add(KB_JOB_FILL, qs, qe - qs, W.soff + rs, re - rs, bw1, -1, KB_EZ_GLOBAL_NO_ZDROP)) return -1;
            rs = re, qs = qe;
        }
    }
    pl.has_right = (W.qe < W.qe0 && W.re < W.re0) ? 1 : 0;
    if (pl.has_right) {
        if (!add(KB_JOB_RIGHT, W.qe, W.qe0 - W.qe, W.soff + W.re, W.re0 - W.re, bw, P.zdrop, KB_EZ_EXTZ_ONLY)) return -1;
    }
    pl.ok = 1, pl.n_jobs = nj;
    return nj;
}
struct KbJobCount {  // sink that only counts
    KB_HD bool operator()(int, const KbJob &) { return true; }
};
struct KbJobWrite {  // sink that stores job k at out[k]
    KbJob *out;
    int32_t chain;
    KB_HD bool operator()(int k, const KbJob &j)
    {
        out[k] = j;
        out[k].chain = chain;
        return true;
    }
};

// Assemble one planned chain from its finished jobs.  `cig` is scratch for the chain's CIGAR (>= the sum of the jobs'
// n_cigar, at most KB_CIG_MAX are used), `jobcig` the pool the DP kernels wrote the per-job CIGARs to.  Returns 0
// when r (and cig[0 .. r.n_cigar)) is exactly what kb_align1 produces for the chain, or -1 when the chain has to be
// redone by kb_align1 (z-drop in a fill, DP error).
KB_HD int kb_stage_assemble(const KbIndexView &ix, const KbBatchView &bt, int gene, const KbPlan &pl, const KbJob *jobs,
                            const uint32_t *jobcig, KbReg &r, uint32_t *cig, bool enabled = true)
{
    // Single exit on purpose: on the GPU one thread per chain runs this and every lane of a warp -- also the ones whose
    // chain has already failed (rc != 0) or that have nothing to do (!enabled) -- walks down to the re-convergence point
    // inside kb_update_extra<true>, so that the long per-column loop there is executed by the whole warp in step.
    const kb_params_t &P = ix.p;
    const uint8_t *qseq0 = (pl.rev ? ix.gseq_rev : ix.gseq_fwd) + ix.gene_seq_off[gene];
    int jb = 0, rc = enabled ? 0 : -1;
    int32_t rs1 = pl.rs, qs1 = pl.qs, re1 = pl.re, qe1 = pl.qe;
    r.has_p = 0, r.dp_score = 0, r.dp_max = 0, r.n_ambi = 0, r.n_cigar = 0;
    if (rc == 0 && pl.has_left) {
        const KbJob &J = jobs[jb++];
        if (J.state != 1 || J.n_cigar < 0) rc = -1;
        else {
            if (J.n_cigar > 0) {
                kb_append_cigar(r, cig, J.n_cigar, jobcig + J.cigar_off);
                r.has_p = 1;
                r.dp_score += J.max;
            }
            rs1 = pl.rs - (J.max_t + 1);
            qs1 = pl.qs - (J.max_q + 1);
        }
    }
    const int n_fill = enabled ? pl.n_jobs - pl.has_left - pl.has_right : 0;
    for (int f = 0; f < n_fill && rc == 0; ++f) {
        const KbJob &J = jobs[jb++];
        if (J.state != 1 || J.n_cigar < 0 || J.zdropped) {
            rc = -1;
            break;
        }
        const KbByteSeq qseq{qseq0 + J.qoff};
        const KbPackSeq tseq{bt.seq2, bt.nmask, J.tpos};
        // mm_test_zdrop walks the alignment base by base; it cannot fire when everything the walk can lose stays within
        // the threshold: its drop is at most the sum of the negative steps, i.e. at most
        // sum over M columns of (a - column score) + sum over gaps of (q + e len)
        //   = a * M_total - (score + sum of dual-affine gap costs) + sum of single-affine gap costs.
        {
            const uint32_t *cg = jobcig + J.cigar_off;
            int64_t m_tot = 0, g1 = 0, g2 = 0;
            for (int k = 0; k < J.n_cigar; ++k) {
                const int64_t len = cg[k] >> 4;
                if ((cg[k] & 0xf) == 0) m_tot += len;
                else {
                    const int64_t c1 = P.q + P.e * len, c2 = P.q2 + P.e2 * len;
                    g1 += c1, g2 += c1 < c2 ? c1 : c2;
                }
            }
            if (P.a * m_tot - ((int64_t)J.score + g2) + g1 > P.zdrop)
                if (kb_test_zdrop(P, qseq, tseq, J.n_cigar, jobcig + J.cigar_off) != 0) {
                    rc = -1;
                    break;
                }
        }
        if (J.n_cigar > 0) {
            kb_append_cigar(r, cig, J.n_cigar, jobcig + J.cigar_off);
            r.has_p = 1;
        }
        r.dp_score += J.score;
    }
    if (rc == 0 && pl.has_right) {
        const KbJob &J = jobs[jb++];
        if (J.state != 1 || J.n_cigar < 0) rc = -1;
        else {
            if (J.n_cigar > 0) {
                kb_append_cigar(r, cig, J.n_cigar, jobcig + J.cigar_off);
                r.has_p = 1;
                r.dp_score += J.max;
            }
            re1 = pl.re + (J.max_t + 1);
            qe1 = pl.qe + (J.max_q + 1);
        }
    }
    if (rc == 0 && r.n_cigar > KB_CIG_MAX) rc = -1;
    if (rc != 0) r.n_cigar = 0, r.has_p = 0;  // the walk below then has nothing to do, but is still entered by this lane
    r.rs = rs1, r.re = re1;
    if (pl.rev) r.qs = pl.qlen - qe1, r.qe = pl.qlen - qs1;
    else r.qs = qs1, r.qe = qe1;
    kb_update_extra<true>(P, r, cig, KbByteSeq{qseq0 + qs1}, KbPackSeq{bt.seq2, bt.nmask, pl.soff + rs1});
    return rc;
}
