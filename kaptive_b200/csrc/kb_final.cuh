// kb_final.cuh -- per-query post-alignment stage: minimap2's mm_filter_regs, mm_hit_sort,
// second mm_set_parent and mm_set_mapq over the aligned regions of one (assembly, gene)
// query.  The order it produces is the order in which the reference receives the hits of
// a query from Aligner.map_batch (src/kaptive/core/alignment.py:409-446), which matters
// because the reference's np.lexsort / np.unique tie-breaks are stable in that order
// (src/kaptive/serotyping/core.py:174,241).
#pragma once
#include "kb_align.cuh"

struct KbHitView : KbRawHit {
    KB_HD bool hp() const { return has_p != 0; }
    KB_HD int dpmax() const { return dp_max; }
    KB_HD int dpmax2() const { return dp_max2; }
    KB_HD void set_dpmax2(int v) { dp_max2 = v; }
};

// minimap2 hit.c mm_set_mapq2 (long reads)
KB_HD void kb_set_mapq(int n_regs, KbHitView *regs, int min_chain_sc, int match_sc, int rep_len)
{
    const float q_coef = 40.0f;
    int64_t sum_sc = 0;
    if (n_regs == 0) return;
    for (int i = 0; i < n_regs; ++i)
        if (regs[i].parent == i) sum_sc += regs[i].score;
    const float uniq_ratio = kb_fdiv((float)sum_sc, (float)(sum_sc + rep_len));
    for (int i = 0; i < n_regs; ++i) {
        KbHitView &r = regs[i];
        if (r.parent == i) {
            int mapq, subsc;
            float pen_s1 = kb_fmul(r.score > 100 ? 1.0f : kb_fmul(0.01f, (float)r.score), uniq_ratio);
            float pen_cm = r.cnt > 10 ? 1.0f : kb_fmul(0.1f, (float)r.cnt);
            pen_cm = pen_s1 < pen_cm ? pen_s1 : pen_cm;
            subsc = r.subsc > min_chain_sc ? r.subsc : min_chain_sc;
            if (r.has_p && r.dp_max2 > 0 && r.dp_max > 0) {
                float identity = kb_fdiv((float)r.mlen, (float)r.blen);
                float x = kb_fmul((float)r.dp_max2, (float)subsc);
                x = kb_fdiv(x, (float)r.dp_max), x = kb_fdiv(x, (float)r.score0);
                float lg = kb_logf(kb_fdiv((float)r.dp_max, (float)match_sc));
                float t = kb_fmul(identity, pen_cm);
                t = kb_fmul(t, q_coef), t = kb_fmul(t, kb_fsub(1.0f, kb_fmul(x, x))), t = kb_fmul(t, lg);
                mapq = (int)t;
                t = kb_fmul(6.02f, identity), t = kb_fmul(t, identity), t = kb_fmul(t, (float)(r.dp_max - r.dp_max2));
                t = kb_fdiv(t, (float)match_sc), t = kb_fadd(t, .499f);
                int mapq_alt = (int)t;
                mapq = mapq < mapq_alt ? mapq : mapq_alt;
            } else {
                float x = kb_fdiv((float)subsc, (float)r.score0), t;
                if (r.has_p) {
                    float identity = kb_fdiv((float)r.mlen, (float)r.blen);
                    t = kb_fmul(identity, pen_cm), t = kb_fmul(t, q_coef), t = kb_fmul(t, kb_fsub(1.0f, x));
                    t = kb_fmul(t, kb_logf(kb_fdiv((float)r.dp_max, (float)match_sc)));
                } else {
                    t = kb_fmul(pen_cm, q_coef), t = kb_fmul(t, kb_fsub(1.0f, x)), t = kb_fmul(t, kb_logf((float)r.score));
                }
                mapq = (int)t;
            }
            mapq -= (int)kb_fadd(kb_fmul(4.343f, kb_logf((float)(r.n_sub + 1))), .499f);
            mapq = mapq > 0 ? mapq : 0;
            r.mapq = mapq < 60 ? mapq : 60;
            if (r.has_p && r.dp_max > r.dp_max2 && r.mapq == 0) r.mapq = 1;
        } else r.mapq = 0;
    }
}

// hits[0..n): the raw hits of one query in minimap2 regs[] order. Returns the number kept;
// kept hits are moved to the front in final order. w: >= n int32; cov: >= n uint64 (scratch).
KB_HD int kb_finalize_group(const kb_params_t &P, KbHitView *hits, int n, int rep_len, int32_t *w, uint64_t *cov)
{
    int k = 0;
    for (int i = 0; i < n; ++i) {  // mm_filter_regs (+ regions without a CIGAR, + failed regions)
        KbHitView &r = hits[i];
        int flt = 0;
        if (r.cnt < P.min_cnt) flt = 1;
        if (r.err || !r.has_p || r.n_cigar == 0) flt = 1;
        else if (r.mlen < P.min_chain_score) flt = 1;
        else if (r.dp_max < P.min_dp_max) flt = 1;
        if (!flt) {
            if (k < i) hits[k] = hits[i];
            ++k;
        }
    }
    n = k;
    if (n > 1) {  // mm_hit_sort: (dp_max, hash) descending, equal keys: later region first
        for (int i = 1; i < n; ++i) {
            KbHitView t = hits[i];
            uint64_t kt = (uint64_t)(uint32_t)t.dp_max << 32 | t.hash;
            int j = i - 1;
            // stable descending insertion == ascending stable sort read backwards needs "later first" on ties:
            // element i (later) goes in front of equal keys
            for (; j >= 0; --j) {
                uint64_t kj = (uint64_t)(uint32_t)hits[j].dp_max << 32 | hits[j].hash;
                if (kj > kt) break;
                hits[j + 1] = hits[j];
            }
            hits[j + 1] = t;
        }
    }
    if (n > 0) {
        // subsc / n_sub / dp_max2 are carried over from the chain-level pass: minimap2 does not reset them here
        kb_set_parent(P.mask_level, P.mask_len, n, hits, P.a * 2 + P.b, w, cov);
        kb_set_mapq(n, hits, P.min_chain_score, P.a, rep_len);
    }
    return n;
}
