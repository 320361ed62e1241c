// kb_chain.cuh -- per-query seed filtering, chaining DP, backtracking and chain bookkeeping.
//
// One "group" = all anchors of one (assembly, gene) pair, i.e. one query of the
// reference's Aligner.map_batch call (src/kaptive/serotyping/core.py:154).  The
// routines restate minimap2's collect_seed_hits / mg_lchain_dp / mg_chain_backtrack /
// compact_a / mm_gen_regs / mm_set_parent for that query.  They are plain sequential
// code (KB_HD) executed by one GPU thread per group: a batch holds 10^5..10^6 groups,
// which is where the parallelism comes from.
#pragma once
#include "kb_common.cuh"

struct KbChainWork {  // all arrays are indexed by global (sorted) anchor index
    uint32_t *x;      // rev << 27 | vpos, filtered anchors compacted to the front of the group's range
    int32_t *y;       // strand-oriented query position | tandem << 30
    int32_t *f, *p, *v, *t;
    uint64_t *z, *u;
};

struct KbGroupInfo {
    int32_t asm_id, gene;
    int64_t a_base;      // start of the group's range in cx/cy
    int32_t n_a;         // chain anchors kept (compacted)
    int32_t n_chains;
    int64_t chain_base;  // first KbChainRec of the group
    int32_t rep_len;
    int32_t n_seed;      // anchors after the occurrence filters
};

struct KbChainRec {
    int32_t group;
    int32_t as, cnt, score, score0;
    uint32_t hash;
    int32_t rev, rid, rs, re, qs, qe;
    int32_t parent, subsc, n_sub, mlen, blen;
    int32_t pad;
};

#define KB_PARENT_UNSET (-1)
#define KB_PARENT_TMP_PRI (-2)

// ascending sort of 64-bit keys: insertion sort for short arrays, heapsort otherwise
KB_HD void kb_sort_u64(uint64_t *a, int64_t n)
{
    if (n < 24) {
        for (int64_t i = 1; i < n; ++i) {
            uint64_t t = a[i];
            int64_t j = i - 1;
            for (; j >= 0 && a[j] > t; --j) a[j + 1] = a[j];
            a[j + 1] = t;
        }
        return;
    }
    for (int64_t s = n / 2 - 1; s >= 0; --s) {
        int64_t i = s;
        uint64_t t = a[i];
        for (;;) {
            int64_t c = 2 * i + 1;
            if (c >= n) break;
            if (c + 1 < n && a[c + 1] > a[c]) ++c;
            if (a[c] <= t) break;
            a[i] = a[c], i = c;
        }
        a[i] = t;
    }
    for (int64_t e = n - 1; e > 0; --e) {
        uint64_t t = a[e];
        a[e] = a[0];
        int64_t i = 0;
        for (;;) {
            int64_t c = 2 * i + 1;
            if (c >= e) break;
            if (c + 1 < e && a[c + 1] > a[c]) ++c;
            if (a[c] <= t) break;
            a[i] = a[c], i = c;
        }
        a[i] = t;
    }
}

// minimap2 lchain.c comput_sc (one segment, not cDNA); x = rev<<27|vpos, y = qpos
KB_HD int32_t kb_comput_sc(uint32_t xi, int32_t yi, uint32_t xj, int32_t yj, int32_t max_dist_x, int32_t max_dist_y,
                           int32_t bw, float chn_pen_gap, float chn_pen_skip, int32_t q_span)
{
    int32_t dq = yi - yj, dr, dd, dg, sc;
    if (dq <= 0 || dq > max_dist_x) return INT32_MIN;
    dr = (int32_t)(xi - xj);
    if (dr == 0 || dq > max_dist_y) return INT32_MIN;
    dd = dr > dq ? dr - dq : dq - dr;
    if (dd > bw) return INT32_MIN;
    dg = dr < dq ? dr : dq;
    sc = q_span < dg ? q_span : dg;
    if (dd || dg > q_span) {
        float lin_pen = kb_fadd(kb_fmul(chn_pen_gap, (float)dd), kb_fmul(chn_pen_skip, (float)dg));
        float log_pen = dd >= 1 ? kb_log2_fast((float)(dd + 1)) : 0.0f;
        sc -= (int)kb_fadd(lin_pen, kb_fmul(.5f, log_pen));
    }
    return sc;
}

// vpos (virtual position within the assembly) -> contig index (global) ; contigs of an assembly have increasing vstart
KB_HD int32_t kb_vpos_to_ctg(const KbBatchView &bt, int32_t asm_id, int32_t vpos)
{
    int32_t lo = bt.asm_ctg_start[asm_id], hi = bt.asm_ctg_start[asm_id + 1] - 1;
    while (lo < hi) {
        int32_t mid = (lo + hi + 1) >> 1;
        if (bt.ctg_vstart[mid] <= vpos) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

// minimap2 hit.c mm_reg_set_coor; cx/cy are minimap2-format anchors (x: rev<<63|rid<<32|rpos, y: flags|span<<32|qpos)
KB_HD void kb_reg_set_coor(KbChainRec &r, int32_t qlen, const uint64_t *cx, const uint64_t *cy)
{
    int32_t k = r.as, q_span = (int32_t)(cy[k] >> 32 & 0xff);
    r.rev = (int32_t)(cx[k] >> 63);
    r.rid = (int32_t)(cx[k] << 1 >> 33);
    r.rs = (int32_t)cx[k] + 1 > q_span ? (int32_t)cx[k] + 1 - q_span : 0;
    r.re = (int32_t)cx[k + r.cnt - 1] + 1;
    if (!r.rev) {
        r.qs = (int32_t)cy[k] + 1 - q_span;
        r.qe = (int32_t)cy[k + r.cnt - 1] + 1;
    } else {
        r.qs = qlen - ((int32_t)cy[k + r.cnt - 1] + 1);
        r.qe = qlen - ((int32_t)cy[k] + 1 - q_span);
    }
}

// minimap2 hit.c mm_cal_fuzzy_len
KB_HD void kb_cal_fuzzy_len(KbChainRec &r, const uint64_t *cx, const uint64_t *cy)
{
    r.mlen = r.blen = 0;
    if (r.cnt <= 0) return;
    r.mlen = r.blen = (int32_t)(cy[r.as] >> 32 & 0xff);
    for (int i = r.as + 1; i < r.as + r.cnt; ++i) {
        int span = (int)(cy[i] >> 32 & 0xff);
        int tl = (int32_t)cx[i] - (int32_t)cx[i - 1];
        int ql = (int32_t)cy[i] - (int32_t)cy[i - 1];
        r.blen += tl > ql ? tl : ql;
        r.mlen += tl > span && ql > span ? span : tl < ql ? tl : ql;
    }
}

// Generic view of what minimap2's mm_set_parent touches, so the same routine serves chains (before
// alignment) and hits (after): R must expose qs,qe,parent,subsc,n_sub,score,cnt,rid,rs,re and hp()/dpmax()/dpmax2()/set_dpmax2().
template <class R>
KB_HD void kb_set_parent(float mask_level, int mask_len, int n, R *r, int sub_diff, int32_t *w, uint64_t *cov)
{
    if (n <= 0) return;
    w[0] = 0, r[0].parent = 0;
    int k = 1;
    for (int i = 1; i < n; ++i) {
        R &ri = r[i];
        int si = ri.qs, ei = ri.qe, n_cov = 0, uncov_len = 0, j;
        for (j = 0; j < k; ++j) {
            R &rp = r[w[j]];
            int sj = rp.qs, ej = rp.qe;
            if (ej <= si || sj >= ei) continue;
            if (sj < si) sj = si;
            if (ej > ei) ej = ei;
            cov[n_cov++] = (uint64_t)sj << 32 | (uint32_t)ej;
        }
        if (n_cov > 0) {
            int x = si;
            kb_sort_u64(cov, n_cov);
            for (int jj = 0; jj < n_cov; ++jj) {
                if ((int)(cov[jj] >> 32) > x) uncov_len += (int)(cov[jj] >> 32) - x;
                x = (int32_t)cov[jj] > x ? (int32_t)cov[jj] : x;
            }
            if (ei > x) uncov_len += ei - x;
            for (j = 0; j < k; ++j) {
                R &rp = r[w[j]];
                int sj = rp.qs, ej = rp.qe, mn, mx, ol;
                if (ej <= si || sj >= ei) continue;
                mn = ej - sj < ei - si ? ej - sj : ei - si;
                mx = ej - sj > ei - si ? ej - sj : ei - si;
                ol = si < sj ? (ei < sj ? 0 : ei < ej ? ei - sj : ej - sj) : (ej < si ? 0 : ej < ei ? ej - si : ei - si);
                if (kb_fsub(kb_fdiv((float)ol, (float)mn), kb_fdiv((float)uncov_len, (float)mx)) > mask_level && uncov_len <= mask_len) {
                    int cnt_sub = 0, sci = ri.score;
                    ri.parent = rp.parent;
                    rp.subsc = rp.subsc > sci ? rp.subsc : sci;
                    if (ri.cnt >= rp.cnt) cnt_sub = 1;
                    if (rp.hp() && ri.hp() && (rp.rid != ri.rid || rp.rs != ri.rs || rp.re != ri.re || ol != mn)) {
                        sci = ri.dpmax();
                        if (rp.dpmax2() < sci) rp.set_dpmax2(sci);
                        if (rp.dpmax() - ri.dpmax() <= sub_diff) cnt_sub = 1;
                    }
                    if (cnt_sub) ++rp.n_sub;
                    break;
                }
            }
        } else j = k;
        if (j == k) w[k++] = i, ri.parent = i, ri.n_sub = 0;
    }
}

struct KbChainRegView : KbChainRec {
    KB_HD bool hp() const { return false; }
    KB_HD int dpmax() const { return 0; }
    KB_HD int dpmax2() const { return 0; }
    KB_HD void set_dpmax2(int) {}
};

// The whole per-query chaining stage.  Returns the number of chains (regs) written.
KB_HD int kb_chain_group(const KbIndexView &ix, const KbBatchView &bt, const uint64_t *akey, const uint32_t *aval,
                         int64_t gs, int64_t ge, const uint16_t *occ, const int32_t *mid_occ_arr, KbChainWork W,
                         uint64_t *cx_, uint64_t *cy_, KbGroupInfo *gi, KbChainRec *chains, unsigned long long *chain_counter,
                         int64_t chain_cap, int32_t group_id, const int32_t *occ_skip = nullptr)
{
    // occ_skip[asm] != 0: no gene minimizer occurs more often in this assembly than the mid_occ in force, so the per-anchor
    // look-up in the (large, randomly addressed) occurrence table cannot filter anything and is skipped
    const kb_params_t &P = ix.p;
    const int32_t asm_id = (int32_t)(akey[gs] >> KB_KEY_ASM_SHIFT);
    const int32_t gene = (int32_t)(akey[gs] >> KB_KEY_GENE_SHIFT) & (KB_MAX_GENES - 1);
    const int32_t qlen = ix.gene_len[gene], nmin = ix.gene_nmin[gene];
    const int32_t mid = mid_occ_arr[asm_id];
    const int32_t K = P.k;
    const bool qflt = nmin > mid && P.q_occ_frac > 0.0f && mid > 0;
    const bool skip_occ = occ_skip && occ_skip[asm_id];
    uint32_t *x = W.x + gs;
    int32_t *y = W.y + gs, *f = W.f + gs, *p = W.p + gs, *v = W.v + gs, *t = W.t + gs;
    uint64_t *z = W.z + gs, *u = W.u + gs, *cx = cx_ + gs, *cy = cy_ + gs;
    int64_t n = 0, n_rep = 0;

    gi->asm_id = asm_id, gi->gene = gene, gi->a_base = gs, gi->n_a = 0, gi->n_chains = 0, gi->chain_base = 0, gi->rep_len = 0, gi->n_seed = 0;

    // ---- seed filters: mm_seed_mz_flt (query side) and mid_occ (indexed side), collect_seed_hits coordinates
    for (int64_t i = gs; i < ge; ++i) {
        uint32_t e = aval[i];
        KbEntry en = ix.ent[e];
        if (qflt && en.qocc > mid && (float)en.qocc > kb_fmul((float)nmin, P.q_occ_frac)) continue;
        const int32_t occ_n = skip_occ ? 0 : occ[(int64_t)asm_id * ix.n_entries + e];
        if (occ_n > mid) {
            t[n_rep++] = (int32_t)(en.qpos_z >> 1);  // t[] is free until chaining starts; n_rep <= i - gs - n
            continue;
        }
        uint32_t xv = (uint32_t)akey[i] & ((1u << (KB_KEY_REV_SHIFT + 1)) - 1);
        int32_t qp = (int32_t)(en.qpos_z >> 1);
        if (xv >> KB_KEY_REV_SHIFT) qp = qlen - (qp + 1 - K) - 1;
        if (en.mi_flags >> 31) qp |= 1 << 30;
        x[n] = xv, y[n] = qp;
        ++n;
    }
    // t[] and x/y share index space only through gs: t entries written so far sit at [0, n_rep), x/y at [0, n); both < ge-gs
    if (n_rep > 0) {  // rep_len = length of the union of the query spans of the filtered seeds
        for (int64_t i = 0; i < n_rep; ++i) z[i] = (uint64_t)(uint32_t)t[i];
        kb_sort_u64(z, n_rep);
        int rep_len = 0, rep_st = 0, rep_en = 0;
        for (int64_t i = 0; i < n_rep; ++i) {
            if (i > 0 && z[i] == z[i - 1]) continue;
            int en = (int)z[i] + 1, st = en - K;
            if (st > rep_en) rep_len += rep_en - rep_st, rep_st = st, rep_en = en;
            else rep_en = en;
        }
        rep_len += rep_en - rep_st;
        gi->rep_len = rep_len;
    }
    gi->n_seed = (int32_t)n;
    if (n < P.min_cnt) return 0;
    // anchors with identical (strand, target position): order by query position (the radix sort saw only the key)
    for (int64_t i = 1; i < n; ++i) {
        if (x[i] != x[i - 1]) continue;
        uint32_t xt = x[i];
        int32_t yt = y[i];
        int64_t j = i - 1;
        for (; j >= 0 && x[j] == xt && (y[j] & 0x3fffffff) > (yt & 0x3fffffff); --j) y[j + 1] = y[j];
        y[j + 1] = yt;
    }

    // ---- mg_lchain_dp
    int32_t max_dist_x = P.max_gap, max_dist_y = P.max_gap, bw = P.bw;
    if (max_dist_x < bw) max_dist_x = bw;
    if (max_dist_y < bw) max_dist_y = bw;
    const float chn_pen_gap = (float)(P.chain_gap_scale * 0.01 * P.k), chn_pen_skip = 0.0f;
    int64_t st = 0, max_ii = -1;
    for (int64_t i = 0; i < n; ++i) t[i] = 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t max_j = -1, end_j, j;
        int32_t max_f = K, n_skip = 0;
        const uint32_t xi = x[i];
        const int32_t yi = y[i] & 0x3fffffff;
        while (st < i && ((xi >> KB_KEY_REV_SHIFT) != (x[st] >> KB_KEY_REV_SHIFT) || xi > x[st] + (uint32_t)max_dist_x)) ++st;
        if (i - st > P.max_chain_iter) st = i - P.max_chain_iter;
        for (j = i - 1; j >= st; --j) {
            int32_t sc = kb_comput_sc(xi, yi, x[j], y[j] & 0x3fffffff, max_dist_x, max_dist_y, bw, chn_pen_gap, chn_pen_skip, K);
            if (sc == INT32_MIN) continue;
            sc += f[j];
            if (sc > max_f) {
                max_f = sc, max_j = j;
                if (n_skip > 0) --n_skip;
            } else if (t[j] == (int32_t)i) {
                if (++n_skip > P.max_chain_skip) break;
            }
            if (p[j] >= 0) t[p[j]] = (int32_t)i;
        }
        end_j = j;
        if (max_ii < 0 || xi - x[max_ii] > (uint32_t)max_dist_x) {
            int32_t mx = INT32_MIN;
            max_ii = -1;
            for (j = i - 1; j >= st; --j)
                if (mx < f[j]) mx = f[j], max_ii = j;
        }
        if (max_ii >= 0 && max_ii < end_j) {
            int32_t tmp = kb_comput_sc(xi, yi, x[max_ii], y[max_ii] & 0x3fffffff, max_dist_x, max_dist_y, bw, chn_pen_gap, chn_pen_skip, K);
            if (tmp != INT32_MIN && max_f < tmp + f[max_ii]) max_f = tmp + f[max_ii], max_j = max_ii;
        }
        f[i] = max_f, p[i] = (int32_t)max_j;
        v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f;
        if (max_ii < 0 || (xi - x[max_ii] <= (uint32_t)max_dist_x && f[max_ii] < f[i])) max_ii = i;
    }

    // ---- mg_chain_backtrack (single pass; z ties ordered by anchor index)
    const int32_t min_sc = P.min_chain_score, min_cnt = P.min_cnt, max_drop = bw;
    int64_t n_z = 0;
    for (int64_t i = 0; i < n; ++i)
        if (f[i] >= min_sc) z[n_z++] = (uint64_t)(uint32_t)f[i] << 32 | (uint32_t)i;
    if (n_z == 0) return 0;
    kb_sort_u64(z, n_z);
    for (int64_t i = 0; i < n; ++i) t[i] = 0;
    int64_t n_v = 0;
    int32_t n_u = 0;
    for (int64_t k = n_z - 1; k >= 0; --k) {
        const int32_t zi = (int32_t)(uint32_t)z[k], zf = (int32_t)(z[k] >> 32);
        if (t[zi] != 0) continue;
        // mg_chain_bk_end
        int64_t i = zi, end_i = -1, max_i = i;
        {
            int32_t max_s = 0;
            do {
                int32_t s;
                t[i] = 2;
                end_i = i = p[i];
                s = i < 0 ? zf : zf - f[i];
                if (s > max_s) max_s = s, max_i = i;
                else if (max_s - s > max_drop) break;
            } while (i >= 0 && t[i] == 0);
            for (i = zi; i >= 0 && i != end_i; i = p[i]) t[i] = 0;
        }
        end_i = max_i;
        int64_t n_v0 = n_v;
        for (i = zi; i != end_i; i = p[i]) v[n_v++] = (int32_t)i, t[i] = 1;
        int32_t sc = i < 0 ? zf : zf - f[i];
        if (sc >= min_sc && n_v > n_v0 && n_v - n_v0 >= min_cnt) u[n_u++] = (uint64_t)(uint32_t)sc << 32 | (uint64_t)(n_v - n_v0);
        else n_v = n_v0;
    }
    if (n_u == 0) return 0;

    // ---- compact_a: chains ordered by the target position of their first anchor (ties: backtrack order)
    // t[] is free again: t[i] = start of chain i in v[]
    {
        int32_t k0 = 0;
        for (int32_t i = 0; i < n_u; ++i) t[i] = k0, k0 += (int32_t)u[i];
    }
    for (int32_t i = 0; i < n_u; ++i) {
        int32_t first = v[t[i] + (int32_t)(uint32_t)u[i] - 1];  // chains are stored end-first in v[]
        z[i] = (uint64_t)x[first] << 32 | (uint32_t)i;           // key x; ties by k (monotone in i)
    }
    kb_sort_u64(z, n_u);
    int64_t ko = 0;
    for (int32_t i = 0; i < n_u; ++i) {
        int32_t src = (int32_t)(uint32_t)z[i], ni = (int32_t)(uint32_t)u[src], k0 = t[src];
        for (int32_t j = 0; j < ni; ++j) {
            int32_t a = v[k0 + (ni - j - 1)];
            uint32_t xv = x[a];
            int32_t vpos = (int32_t)(xv & KB_VPOS_MASK);
            int32_t c = kb_vpos_to_ctg(bt, asm_id, vpos);
            uint64_t rid = (uint64_t)(c - bt.asm_ctg_start[asm_id]);
            cx[ko] = (uint64_t)(xv >> KB_KEY_REV_SHIFT) << 63 | rid << 32 | (uint32_t)(vpos - bt.ctg_vstart[c]);
            cy[ko] = (uint64_t)K << 32 | (uint32_t)(y[a] & 0x3fffffff) | ((y[a] >> 30 & 1) ? KB_SEED_TANDEM : 0ULL);
            ++ko;
        }
        z[i] = u[src];  // z[i] (the sort key) is consumed: reuse the slot for u in compact order (score<<32|cnt)
    }
    for (int32_t i = 0; i < n_u; ++i) u[i] = z[i];

    // ---- mm_gen_regs: sort by (score, cnt ^ h) descending; equal keys: later chain first
    {
        int32_t k = 0;
        for (int32_t i = 0; i < n_u; ++i) {
            uint32_t h = (uint32_t)kb_hash64_full((kb_hash64_full(cx[k]) + kb_hash64_full(cy[k])) ^ ix.gene_hash[gene]);
            z[i] = u[i] ^ h;
            t[i] = k;
            k += (int32_t)(uint32_t)u[i];
        }
    }
    // descending order with "larger index first" on ties == ascending sort of (key, idx) read backwards;
    // idx does not fit beside a 64-bit key, so sort indices with an explicit comparison (n_u is small)
    for (int32_t i = 0; i < n_u; ++i) v[i] = i;
    for (int32_t i = 1; i < n_u; ++i) {
        int32_t vi = v[i], j = i - 1;
        for (; j >= 0 && (z[v[j]] < z[vi] || (z[v[j]] == z[vi] && v[j] < vi)); --j) v[j + 1] = v[j];
        v[j + 1] = vi;
    }
    unsigned long long base;
#ifdef __CUDA_ARCH__
    base = atomicAdd(chain_counter, (unsigned long long)n_u);
#else
    base = *chain_counter, *chain_counter += (unsigned long long)n_u;
#endif
    gi->n_a = (int32_t)ko, gi->n_chains = n_u, gi->chain_base = (int64_t)base;
    if ((int64_t)base + n_u > chain_cap) return n_u;  // overflow: caller re-runs with a larger buffer
    KbChainRegView *regs = static_cast<KbChainRegView *>(chains + base);
    for (int32_t i = 0; i < n_u; ++i) {
        KbChainRegView &r = regs[i];
        int32_t src = v[i];
        r.group = group_id;
        r.parent = KB_PARENT_UNSET;
        r.score = r.score0 = (int32_t)(z[src] >> 32);
        r.hash = (uint32_t)z[src];
        r.cnt = (int32_t)(uint32_t)u[src];
        r.as = t[src];
        r.subsc = 0, r.n_sub = 0, r.pad = 0;
        kb_reg_set_coor(r, qlen, cx, cy);
        kb_cal_fuzzy_len(r, cx, cy);
    }
    // ---- mm_set_parent (chain level); pri_ratio = 0 makes mm_select_sub a no-op (serotyping/core.py:151)
    kb_set_parent(P.mask_level, P.mask_len, n_u, regs, P.a * 2 + P.b, f, W.z + gs + n_u);
    return n_u;
}
