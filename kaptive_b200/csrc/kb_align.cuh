// kb_align.cuh -- base-level extension of one chain: the "banded gapped extension
// (+CIGAR)" step of the reference's map_batch call (src/kaptive/serotyping/core.py:154;
// hit fields consumed at src/kaptive/core/alignment.py:414-446).
//
// Restates minimap2's mm_align1 (fix_bad_ends, bad-seed filters, left extension, gap
// filling with z-drop test and split, right extension, mm_update_extra) around a
// dual-affine DP (ksw_extd2 recurrences and tie rules) that is computed by NL
// cooperating lanes, one anti-diagonal at a time: lanes stride over the cells of the
// anti-diagonal, the previous two anti-diagonals live in per-warp scratch, the
// anti-diagonal maximum needed for z-drop is a warp reduction.  NL = 32 on the GPU
// (one warp per chain), NL = 1 in the host emulation used by the CPU-only tests.
#pragma once
#include "kb_chain.cuh"

#define KB_DP_MAXLEN 16384      // longest query/target side of one DP problem
#define KB_CIG_MAX 8192         // CIGAR operations per hit
#define KB_TFULL_MAX 65536      // target window of one chain
#define KB_EZ_EXTZ_ONLY 0x1
#define KB_EZ_RIGHT 0x2
#define KB_EZ_REV_CIGAR 0x4
#define KB_EZ_GLOBAL_NO_ZDROP 0x8
#define KB_RING_WORDS 520       // 512 anti-diagonals + 8 alias words (a lane's 8 slots never wrap inside one step)

struct KbAlignScratch {
    int32_t *dp;      // 11 * KB_DP_MAXLEN
    uint8_t *tb;      // max_sw_cells + 64
    int32_t *off;     // 3 * 2 * KB_DP_MAXLEN : off, off_end, ppos
    uint8_t *qbuf;    // KB_DP_MAXLEN (reversed query for the left extension)
    uint8_t *tbuf;    // KB_DP_MAXLEN (reversed target for the left extension)
    uint8_t *tfull;   // KB_TFULL_MAX target codes of [rs0, re0)
    uint32_t *cigar;  // KB_CIG_MAX, CIGAR of the hit being built
    uint32_t *ezcig;  // KB_CIG_MAX, CIGAR of the last DP
    uint32_t *wmax;   // device only: per-warp ring of KB_RING_WORDS words in shared memory (per-anti-diagonal maxima of kb_rows)
};

KB_HD size_t kb_align_scratch_bytes(int64_t max_sw_cells)
{
    size_t b = 0;
    b += (size_t)11 * KB_DP_MAXLEN * 4;
    b += ((size_t)max_sw_cells + 64 + 15) & ~(size_t)15;
    b += (size_t)6 * KB_DP_MAXLEN * 4;
    b += (size_t)2 * KB_DP_MAXLEN;
    b += KB_TFULL_MAX;
    b += (size_t)2 * KB_CIG_MAX * 4;
    return (b + 255) & ~(size_t)255;
}

KB_HD KbAlignScratch kb_align_scratch_at(uint8_t *base, int64_t max_sw_cells)
{
    KbAlignScratch S;
    uint8_t *p = base;
    S.dp = (int32_t *)p, p += (size_t)11 * KB_DP_MAXLEN * 4;
    S.off = (int32_t *)p, p += (size_t)6 * KB_DP_MAXLEN * 4;
    S.cigar = (uint32_t *)p, p += (size_t)KB_CIG_MAX * 4;
    S.ezcig = (uint32_t *)p, p += (size_t)KB_CIG_MAX * 4;
    S.qbuf = p, p += KB_DP_MAXLEN;
    S.tbuf = p, p += KB_DP_MAXLEN;
    S.tfull = p, p += KB_TFULL_MAX;
    S.tb = p;
    S.wmax = nullptr;
    (void)max_sw_cells;
    return S;
}

struct KbEz {
    int32_t max, max_q, max_t, score, zdropped, n_cigar;
};

// raw hit = one aligned region, before the per-query filter / sort / MAPQ stage
struct KbRawHit {
    int32_t group, reg_idx, split_idx;  // position in minimap2's regs[] order
    int32_t cnt, score, score0;         // chain-level (anchors, chain score after split, original chain score)
    uint32_t hash;
    int32_t rev, rid, rs, re, qs, qe;
    int32_t has_p, dp_score, dp_max, dp_max2, n_ambi, mlen, blen;
    int32_t parent, subsc, n_sub, mapq;
    int32_t n_cigar;
    int64_t cigar_off;
    int32_t err, pad;
};

// ---------------------------------------------------------------- warp helpers
template <int NL>
KB_HD void kb_sync()
{
#ifdef __CUDA_ARCH__
    if (NL > 1) __syncwarp();
#endif
}
template <int NL>
KB_HD int32_t kb_bcast(int32_t v)
{
#ifdef __CUDA_ARCH__
    if (NL > 1) return __shfl_sync(0xffffffffu, v, 0);
#endif
    return v;
}
// (max H, lowest t among equals) over the lanes
template <int NL>
KB_HD void kb_reduce_max(int32_t &h, int32_t &t)
{
#ifdef __CUDA_ARCH__
    if (NL > 1) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            int32_t oh = __shfl_xor_sync(0xffffffffu, h, d);
            int32_t ot = __shfl_xor_sync(0xffffffffu, t, d);
            if (oh > h || (oh == h && ot < t)) h = oh, t = ot;
        }
    }
#endif
}

KB_HD int32_t kb_gapcost2(const kb_params_t &P, int l)
{
    int32_t c1 = P.q + P.e * l, c2 = P.q2 + P.e2 * l;
    return c1 < c2 ? c1 : c2;
}

// the handful of scoring constants the DP inner loops need, passed by value so they live in registers
struct KbDpConst {
    int32_t a, b, q, e, q2, e2, sc_ambi, max_sw_cells;
};
KB_HD KbDpConst kb_dp_const(const kb_params_t &P)
{
    KbDpConst c;
    c.a = P.a, c.b = P.b, c.q = P.q, c.e = P.e, c.q2 = P.q2, c.e2 = P.e2, c.sc_ambi = P.sc_ambi, c.max_sw_cells = P.max_sw_cells;
    return c;
}
KB_HD int32_t kb_gapcost2(const KbDpConst &P, int l)
{
    int32_t c1 = P.q + P.e * l, c2 = P.q2 + P.e2 * l;
    return c1 < c2 ? c1 : c2;
}
KB_HD int kb_sub_score(const KbDpConst &P, int ct, int cq)
{
    return (ct > 3 || cq > 3) ? -P.sc_ambi : (ct == cq ? P.a : -P.b);
}

KB_HD int kb_sub_score(const kb_params_t &P, int ct, int cq)
{
    return (ct > 3 || cq > 3) ? -P.sc_ambi : (ct == cq ? P.a : -P.b);
}

// ksw_backtrack (is_rot) over the traceback bytes in S.tb; sequential, lane 0; result broadcast to the warp.
template <int NL>
KB_HD void kb_backtrack(int lane, int qlen, int tlen, int flag, KbEz &ez, const KbAlignScratch &S)
{
    const int32_t *off = S.off, *off_end = S.off + 2 * KB_DP_MAXLEN, *ppos = S.off + 4 * KB_DP_MAXLEN;
    const uint8_t *p = S.tb;
    int n_cigar = 0;
    if (lane == 0) {
        int i0 = -1, j0 = -1;
        uint32_t *cg = S.ezcig;
        if (!ez.zdropped && !(flag & KB_EZ_EXTZ_ONLY)) i0 = tlen - 1, j0 = qlen - 1;
        else if (ez.max_t >= 0 && ez.max_q >= 0) i0 = ez.max_t, j0 = ez.max_q;
        if (i0 >= 0 && j0 >= 0) {
            int i = i0, j = j0, state = 0;
            auto push = [&](uint32_t op, int len) {
                if (n_cigar == 0 || op != (cg[n_cigar - 1] & 0xf)) {
                    if (n_cigar < KB_CIG_MAX) cg[n_cigar] = (uint32_t)len << 4 | op;
                    ++n_cigar;
                } else if (n_cigar <= KB_CIG_MAX) cg[n_cigar - 1] += (uint32_t)len << 4;
            };
            while (i >= 0 && j >= 0) {
                int force_state = -1, r = i + j;
                uint32_t tmp;
                if (i < off[r]) force_state = 2;
                if (i > off_end[r]) force_state = 1;
                tmp = force_state < 0 ? p[ppos[r] + i - off[r]] : 0;
                if (state == 0) state = tmp & 7;
                else if (!(tmp >> (state + 2) & 1)) state = 0;
                if (state == 0) state = tmp & 7;
                if (force_state >= 0) state = force_state;
                if (state == 0) push(0, 1), --i, --j;
                else if (state == 1 || state == 3) push(2, 1), --i;
                else push(1, 1), --j;
            }
            if (i >= 0) push(2, i + 1);
            if (j >= 0) push(1, j + 1);
            if (n_cigar > KB_CIG_MAX) n_cigar = -1;  // overflow: reported as an error on the hit
            else if (!(flag & KB_EZ_REV_CIGAR))
                for (int a = 0; a < n_cigar >> 1; ++a) {
                    uint32_t tmp = cg[a];
                    cg[a] = cg[n_cigar - 1 - a], cg[n_cigar - 1 - a] = tmp;
                }
        }
    }
    kb_sync<NL>();
    ez.n_cigar = kb_bcast<NL>(n_cigar);
}

// Dual-affine DP, anti-diagonal order (ksw_extd2 recurrences; see oracle/kb_oracle.c:extd2 for the spec).
template <int NL>
KB_HD void kb_extd2(const KbDpConst P, int lane, int qlen, const uint8_t *qs, int tlen, const uint8_t *ts, int w,
                    int zdrop, int flag, KbEz &ez, const KbAlignScratch S, int64_t *cell_counter)
{
    const int rb = (flag & KB_EZ_RIGHT) ? 1 : 0;
    const int q = P.q, e = P.e, q2 = P.q2, e2 = P.e2;
    ez.max = 0, ez.max_q = ez.max_t = -1, ez.score = KB_NEG_INF, ez.zdropped = 0, ez.n_cigar = 0;
    if (qlen <= 0 || tlen <= 0) return;
    if ((int64_t)qlen * tlen > P.max_sw_cells || qlen > KB_DP_MAXLEN || tlen > KB_DP_MAXLEN) {
        ez.zdropped = 1;
        return;
    }
    const int n_diag = qlen + tlen - 1;
    int32_t *H[3] = {S.dp, S.dp + KB_DP_MAXLEN, S.dp + 2 * KB_DP_MAXLEN};
    int32_t *E1[2] = {S.dp + 3 * KB_DP_MAXLEN, S.dp + 4 * KB_DP_MAXLEN};
    int32_t *E2[2] = {S.dp + 5 * KB_DP_MAXLEN, S.dp + 6 * KB_DP_MAXLEN};
    int32_t *F1[2] = {S.dp + 7 * KB_DP_MAXLEN, S.dp + 8 * KB_DP_MAXLEN};
    int32_t *F2[2] = {S.dp + 9 * KB_DP_MAXLEN, S.dp + 10 * KB_DP_MAXLEN};
    int32_t *off = S.off, *off_end = S.off + 2 * KB_DP_MAXLEN, *ppos = S.off + 4 * KB_DP_MAXLEN;
    uint8_t *p = S.tb;
    int64_t tb_n = 0;
    int last_st = 0, last_en = -1, last2_st = 0, last2_en = -1;

    for (int r = 0; r < n_diag; ++r) {
        int st = 0, en = tlen - 1;
        if (st < r - qlen + 1) st = r - qlen + 1;
        if (en > r) en = r;
        if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
        if (en > (r + w) >> 1) en = (r + w) >> 1;
        if (st > en) {
            ez.zdropped = 1;
            break;
        }
        int32_t *Hc = H[r % 3], *H1 = H[(r + 2) % 3], *Hd = H[(r + 1) % 3];
        int32_t *e1c = E1[r & 1], *e1p = E1[(r & 1) ^ 1], *e2c = E2[r & 1], *e2p = E2[(r & 1) ^ 1];
        int32_t *f1c = F1[r & 1], *f1p = F1[(r & 1) ^ 1], *f2c = F2[r & 1], *f2p = F2[(r & 1) ^ 1];
        if (lane == 0) off[r] = st, off_end[r] = en, ppos[r] = (int32_t)tb_n;
        uint8_t *pr = p + tb_n - st;
        tb_n += en - st + 1;
        int32_t max_H = INT32_MIN, max_t = 0x7fffffff;
        for (int t = st + lane; t <= en; t += NL) {
            const int j = r - t;
            int32_t h_up, h_left, h_diag, a1, a2, b1, b2, z, hq, hq2;
            uint8_t d;
            if (t == 0) h_up = -kb_gapcost2(P, j + 1), a1 = a2 = KB_NEG_INF;
            else if (t - 1 >= last_st && t - 1 <= last_en) h_up = H1[t - 1], a1 = e1p[t - 1], a2 = e2p[t - 1];
            else h_up = KB_NEG_INF, a1 = a2 = KB_NEG_INF;
            if (j == 0) h_left = -kb_gapcost2(P, t + 1), b1 = b2 = KB_NEG_INF;
            else if (t >= last_st && t <= last_en) h_left = H1[t], b1 = f1p[t], b2 = f2p[t];
            else h_left = KB_NEG_INF, b1 = b2 = KB_NEG_INF;
            if (t == 0 && j == 0) h_diag = 0;
            else if (t == 0) h_diag = -kb_gapcost2(P, j);
            else if (j == 0) h_diag = -kb_gapcost2(P, t);
            else if (t - 1 >= last2_st && t - 1 <= last2_en) h_diag = Hd[t - 1];
            else h_diag = KB_NEG_INF;
            a1 = (h_up - q > a1 ? h_up - q : a1) - e;
            a2 = (h_up - q2 > a2 ? h_up - q2 : a2) - e2;
            b1 = (h_left - q > b1 ? h_left - q : b1) - e;
            b2 = (h_left - q2 > b2 ? h_left - q2 : b2) - e2;
            z = h_diag + kb_sub_score(P, ts[t], qs[j]);
            d = 0;  // (x >= y) == (x + 1 > y): rb = 1 gives the gap-preferring tie rule of KB_EZ_RIGHT
            if (a1 + rb > z) d = 1, z = a1;
            if (b1 + rb > z) d = 2, z = b1;
            if (a2 + rb > z) d = 3, z = a2;
            if (b2 + rb > z) d = 4, z = b2;
            hq = z - q - rb, hq2 = z - q2 - rb;
            if (a1 > hq) d |= 0x08;
            if (b1 > hq) d |= 0x10;
            if (a2 > hq2) d |= 0x20;
            if (b2 > hq2) d |= 0x40;
            Hc[t] = z, e1c[t] = a1, e2c[t] = a2, f1c[t] = b1, f2c[t] = b2;
            pr[t] = d;
            if (z > max_H) max_H = z, max_t = t;  // lanes visit t in increasing order: first maximum = lowest t
        }
        if (!(flag & KB_EZ_GLOBAL_NO_ZDROP)) {
            kb_reduce_max<NL>(max_H, max_t);
            if (max_H > ez.max) {
                ez.max = max_H, ez.max_t = max_t, ez.max_q = r - max_t;
            } else if (max_t >= ez.max_t && r - max_t >= ez.max_q) {
                int tl = max_t - ez.max_t, ql = (r - max_t) - ez.max_q, l;
                l = tl > ql ? tl - ql : ql - tl;
                if (zdrop >= 0 && ez.max - max_H > zdrop + l * e2) {
                    ez.zdropped = 1;
                    break;
                }
            }
        }
        kb_sync<NL>();
        if (r == n_diag - 1 && en == tlen - 1) ez.score = Hc[tlen - 1];
        last2_st = last_st, last2_en = last_en;
        last_st = st, last_en = en;
    }
    kb_sync<NL>();
    if (cell_counter && lane == 0) *cell_counter += tb_n;
    kb_backtrack<NL>(lane, qlen, tlen, flag, ez, S);
}

// limits of the register-resident DP kernels (kb_rows); `track` = the per-anti-diagonal maximum is needed
KB_HD bool kb_rows_eligible(int64_t max_sw_cells, int qlen, int tlen, int w, bool track)
{
    if (qlen <= 0 || tlen <= 0 || qlen > KB_DP_MAXLEN || tlen > KB_DP_MAXLEN) return false;
    const int64_t tiles = (tlen + 255) / 256;
    const int dlen = tlen > qlen ? tlen - qlen : qlen - tlen;
    if ((int64_t)qlen * tlen > max_sw_cells || tiles * (qlen + 31) * 256 > max_sw_cells) return false;
    if (w < 2 || (!track && dlen >= w)) return false;
    if (track && (qlen > 4096 || tlen > 4096)) return false;
    return true;
}
// the certified band pass is worth trying (kb_global_band re-checks its own margin)
KB_HD bool kb_band_eligible(int64_t max_sw_cells, int qlen, int tlen, int w, int flag)
{
    const int width = qlen < tlen ? qlen : tlen;
    const int dlen = tlen > qlen ? tlen - qlen : qlen - tlen;
    return width > 32 && w >= qlen + tlen && (flag & KB_EZ_GLOBAL_NO_ZDROP) && !(flag & KB_EZ_RIGHT) && (63 - dlen) / 2 >= 8 &&
           (int64_t)32 * (qlen + tlen + 8) <= max_sw_cells && (int64_t)qlen * tlen <= max_sw_cells;
}

#ifndef KB_DP_STAT
#define KB_DP_STAT(kind, path, cells) ((void)0)
#define KB_DP_STAT_RAW(slot, v) ((void)0)
#endif
#ifdef __CUDACC__
#define kb_backtrack_lane0(lane, qlen, tlen, flag, ez, S) kb_backtrack<32>(lane, qlen, tlen, flag, ez, S)
#include "kb_align_reg.cuh"
#include "kb_align_reg16.cuh"
#include "kb_band16.cuh"
// device-side choice between the register-resident DP and the scratch-memory DP (identical results)
static __device__ __noinline__ void kb_dp_device(const KbDpConst P, int lane, int qlen, const uint8_t *qs, int tlen, const uint8_t *ts,
                                                 int w, int zdrop, int flag, KbEz &ez, const KbAlignScratch S, int64_t *cell_counter)
{
    const int kind = (flag & KB_EZ_GLOBAL_NO_ZDROP) ? 0 : ((flag & KB_EZ_EXTZ_ONLY) ? 1 : 2);
    (void)kind;
    const KbPtrSeq sq{qs}, st{ts};
    if (kb_band_eligible(P.max_sw_cells, qlen, tlen, w, flag)) {
        const int ok = kb_global_band(P, lane, qlen, sq, tlen, st, flag, ez, S, cell_counter);
        if (lane == 0) KB_DP_STAT(kind, ok ? 0 : 1, (int64_t)32 * (qlen + tlen + 1));
        if (ok) return;
    }
    // row-stripe wavefront for every rectangle within its limits (keys of the z-drop tracker hold 12 bits of t)
    const bool track = !(flag & KB_EZ_GLOBAL_NO_ZDROP);
    if (kb_rows_eligible(P.max_sw_cells, qlen, tlen, w, track)) {
        if (lane == 0) KB_DP_STAT(kind, tlen > 256 ? 3 : 2, (int64_t)qlen * tlen);
        if (track) kb_rows<true>(P, lane, qlen, sq, tlen, st, w, zdrop, flag, ez, S, cell_counter);
        else kb_rows<false>(P, lane, qlen, sq, tlen, st, w, zdrop, flag, ez, S, cell_counter);
    } else {
        if (lane == 0) KB_DP_STAT(kind, 4, (int64_t)qlen * tlen);
        kb_extd2<32>(P, lane, qlen, qs, tlen, ts, w, zdrop, flag, ez, S, cell_counter);
    }
}
#endif

template <int NL>
KB_HD void kb_dp(const kb_params_t &PP, int lane, int qlen, const uint8_t *qs, int tlen, const uint8_t *ts, int w, int zdrop, int flag,
                 KbEz &ez, const KbAlignScratch &S, int64_t *cell_counter)
{
    const KbDpConst P = kb_dp_const(PP);
#ifdef __CUDA_ARCH__
    if (NL == 32) {
        kb_dp_device(P, lane, qlen, qs, tlen, ts, w, zdrop, flag, ez, S, cell_counter);
        return;
    }
#endif
    kb_extd2<NL>(P, lane, qlen, qs, tlen, ts, w, zdrop, flag, ez, S, cell_counter);
}

// ---------------------------------------------------------------- mm_align1 pieces (all lanes run them redundantly)

// minimap2 align.c mm_fix_bad_ends
KB_HD void kb_fix_bad_ends(int r_as, int r_cnt, int r_mlen, const uint64_t *ax, const uint64_t *ay, int bw, int min_match,
                           int32_t *as, int32_t *cnt)
{
    int32_t i, l, m;
    *as = r_as, *cnt = r_cnt;
    if (r_cnt < 3) return;
    m = l = (int32_t)(ay[r_as] >> 32 & 0xff);
    for (i = r_as + 1; i < r_as + r_cnt - 1; ++i) {
        int32_t lq, lr, mn, mx, q_span = (int32_t)(ay[i] >> 32 & 0xff);
        if (ay[i] & KB_SEED_LONG_JOIN) break;
        lr = (int32_t)ax[i] - (int32_t)ax[i - 1];
        lq = (int32_t)ay[i] - (int32_t)ay[i - 1];
        mn = lr < lq ? lr : lq, mx = lr > lq ? lr : lq;
        if (mx - mn > l >> 1) *as = i;
        l += mn;
        m += mn < q_span ? mn : q_span;
        if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r_mlen >> 1) break;
    }
    *cnt = r_as + r_cnt - *as;
    m = l = (int32_t)(ay[r_as + r_cnt - 1] >> 32 & 0xff);
    for (i = r_as + r_cnt - 2; i > *as; --i) {
        int32_t lq, lr, mn, mx, q_span = (int32_t)(ay[i + 1] >> 32 & 0xff);
        if (ay[i + 1] & KB_SEED_LONG_JOIN) break;
        lr = (int32_t)ax[i + 1] - (int32_t)ax[i];
        lq = (int32_t)ay[i + 1] - (int32_t)ay[i];
        mn = lr < lq ? lr : lq, mx = lr > lq ? lr : lq;
        if (mx - mn > l >> 1) *cnt = i + 1 - *as;
        l += mn;
        m += mn < q_span ? mn : q_span;
        if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r_mlen >> 1) break;
    }
}

KB_HD int kb_anchor_gap(const uint64_t *ax, const uint64_t *ay, int i)
{
    return ((int32_t)ay[i] - (int32_t)ay[i - 1]) - ((int32_t)ax[i] - (int32_t)ax[i - 1]);
}

// index (relative to as1) of the k-th gap longer than min_gap, or -1; replaces collect_long_gaps' K[] array
KB_HD int kb_long_gap_at(const uint64_t *ax, const uint64_t *ay, int as1, int cnt1, int min_gap, int k)
{
    int n = 0;
    for (int i = 1; i < cnt1; ++i) {
        int gap = kb_anchor_gap(ax, ay, as1 + i);
        if (gap < -min_gap || gap > min_gap) {
            if (n == k) return i;
            ++n;
        }
    }
    return -1;
}
KB_HD int kb_count_long_gaps(const uint64_t *ax, const uint64_t *ay, int as1, int cnt1, int min_gap)
{
    int n = 0;
    for (int i = 1; i < cnt1; ++i) {
        int gap = kb_anchor_gap(ax, ay, as1 + i);
        if (gap < -min_gap || gap > min_gap) ++n;
    }
    return n;
}

// minimap2 align.c mm_filter_bad_seeds; K[] lives in scratch (int32, >= cnt1 entries)
KB_HD void kb_filter_bad_seeds(int as1, int cnt1, const uint64_t *ax, uint64_t *ay, int min_gap, int diff_thres, int max_ext_len,
                               int max_ext_cnt, int32_t *K)
{
    int n = 0;
    for (int i = 1; i < cnt1; ++i) {
        int gap = kb_anchor_gap(ax, ay, as1 + i);
        if (gap < -min_gap || gap > min_gap) K[n++] = i;
    }
    if (n <= 1) return;
    int mx = 0, max_st = -1, max_en = -1;
    for (int k = 0;; ++k) {
        int gap, l, n_ins = 0, n_del = 0, qs, rs, max_diff = 0, max_diff_l = -1, i;
        if (k == n || k >= max_en) {
            if (max_en > 0)
                for (i = K[max_st]; i < K[max_en]; ++i) ay[as1 + i] |= KB_SEED_IGNORE;
            mx = 0, max_st = max_en = -1;
            if (k == n) break;
        }
        i = K[k];
        gap = kb_anchor_gap(ax, ay, as1 + i);
        if (gap > 0) n_ins += gap;
        else n_del += -gap;
        qs = (int32_t)ay[as1 + i - 1];
        rs = (int32_t)ax[as1 + i - 1];
        for (l = k + 1; l < n && l <= k + max_ext_cnt; ++l) {
            int j = K[l], diff;
            if ((int32_t)ay[as1 + j] - qs > max_ext_len || (int32_t)ax[as1 + j] - rs > max_ext_len) break;
            gap = kb_anchor_gap(ax, ay, as1 + j);
            if (gap > 0) n_ins += gap;
            else n_del += -gap;
            diff = n_ins + n_del - (n_ins > n_del ? n_ins - n_del : n_del - n_ins);
            if (max_diff < diff) max_diff = diff, max_diff_l = l;
        }
        if (max_diff > diff_thres && max_diff > mx) mx = max_diff, max_st = k, max_en = max_diff_l;
    }
}

// minimap2 align.c mm_filter_bad_seeds_alt
KB_HD void kb_filter_bad_seeds_alt(int as1, int cnt1, const uint64_t *ax, uint64_t *ay, int min_gap, int max_ext, int32_t *K)
{
    int n = 0;
    for (int i = 1; i < cnt1; ++i) {
        int gap = kb_anchor_gap(ax, ay, as1 + i);
        if (gap < -min_gap || gap > min_gap) K[n++] = i;
    }
    if (n <= 1) return;
    for (int k = 0; k < n;) {
        int i = K[k], l;
        int gap1 = kb_anchor_gap(ax, ay, as1 + i);
        int re1 = (int32_t)ax[as1 + i], qe1 = (int32_t)ay[as1 + i];
        gap1 = gap1 > 0 ? gap1 : -gap1;
        for (l = k + 1; l < n; ++l) {
            int j = K[l], gap2, q_span_pre, rs2, qs2, m;
            if ((int32_t)ay[as1 + j] - qe1 > max_ext || (int32_t)ax[as1 + j] - re1 > max_ext) break;
            gap2 = kb_anchor_gap(ax, ay, as1 + j);
            q_span_pre = (int)(ay[as1 + j - 1] >> 32 & 0xff);
            rs2 = (int32_t)ax[as1 + j - 1] + q_span_pre;
            qs2 = (int32_t)ay[as1 + j - 1] + q_span_pre;
            m = rs2 - re1 < qs2 - qe1 ? rs2 - re1 : qs2 - qe1;
            gap2 = gap2 > 0 ? gap2 : -gap2;
            if (m > gap1 + gap2) break;
            re1 = (int32_t)ax[as1 + j], qe1 = (int32_t)ay[as1 + j];
            gap1 = gap2;
        }
        if (l > k + 1) {
            int end = K[l - 1];
            for (int j = K[k]; j < end; ++j) ay[as1 + j] |= KB_SEED_IGNORE;
            ay[as1 + end] |= KB_SEED_LONG_JOIN;
        }
        k = l;
    }
}

// minimap2 align.c mm_test_zdrop without the inversion test
template <class SQ, class ST>
KB_HD int kb_test_zdrop(const kb_params_t &P, SQ qseq, ST tseq, int n_cigar, const uint32_t *cigar)
{
    int32_t score = 0, mx = INT32_MIN, max_i = -1, max_j = -1, i = 0, j = 0, max_zdrop = 0;
    for (int k = 0; k < n_cigar; ++k) {
        uint32_t op = cigar[k] & 0xf, len = cigar[k] >> 4;
        if (op == 0) {
            for (uint32_t l = 0; l < len; ++l) {
                score += kb_sub_score(P, tseq[i + l], qseq[j + l]);
                if (score < mx) {
                    int li = i + (int)l - max_i, lj = j + (int)l - max_j;
                    int diff = li > lj ? li - lj : lj - li;
                    int z = mx - score - diff * P.e;
                    if (z > max_zdrop) max_zdrop = z;
                } else mx = score, max_i = i + (int)l, max_j = j + (int)l;
            }
            i += len, j += len;
        } else {
            score -= P.q + P.e * (int)len;
            if (op == 1) j += len;
            else i += len;
            if (score < mx) {
                int li = i - max_i, lj = j - max_j;
                int diff = li > lj ? li - lj : lj - li;
                int z = mx - score - diff * P.e;
                if (z > max_zdrop) max_zdrop = z;
            } else mx = score, max_i = i, max_j = j;
        }
    }
    return max_zdrop > P.zdrop ? 1 : 0;
}

// working copy of one region while it is aligned (minimap2 mm_reg1_t + mm_extra_t)
struct KbReg {
    int32_t as, cnt, score, score0, mlen, blen, parent, id;
    uint32_t hash;
    int32_t rev, rid, rs, re, qs, qe;
    int32_t has_p, dp_score, dp_max, n_ambi, n_cigar;
};

// minimap2 align.c mm_append_cigar (lane 0 only)
KB_HD void kb_append_cigar(KbReg &r, uint32_t *cigar, int n_cigar, const uint32_t *src)
{
    if (n_cigar <= 0) return;
    r.has_p = 1;
    if (r.n_cigar > 0 && (cigar[r.n_cigar - 1] & 0xf) == (src[0] & 0xf)) {
        cigar[r.n_cigar - 1] += (src[0] >> 4) << 4;
        for (int i = 1; i < n_cigar; ++i)
            if (r.n_cigar + i - 1 < KB_CIG_MAX) cigar[r.n_cigar + i - 1] = src[i];
        r.n_cigar += n_cigar - 1;
    } else {
        for (int i = 0; i < n_cigar; ++i)
            if (r.n_cigar + i < KB_CIG_MAX) cigar[r.n_cigar + i] = src[i];
        r.n_cigar += n_cigar;
    }
}

// minimap2 align.c mm_fix_cigar + mm_update_extra (lane 0 only)
// FLAT = true walks the alignment as ONE loop over its columns (a small state machine over the CIGAR) instead of nested
// per-operation loops: same arithmetic in the same order, but when one GPU thread per chain runs it (kb_assemble_kernel) the
// lanes of a warp stay in step instead of serialising each other's inner loops.
template <bool FLAT = false, class SQ, class ST>
KB_HD void kb_update_extra(const kb_params_t &P, KbReg &r, uint32_t *cigar, SQ qseq, ST tseq)
{
    int32_t toff = 0, qoff = 0, to_shrink = 0, qshift = 0, tshift = 0;
    int n = r.n_cigar, k;
    if (n > 1) {
        for (k = 0; k < n; ++k) {
            uint32_t op = cigar[k] & 0xf, len = cigar[k] >> 4;
            if (len == 0) to_shrink = 1;
            if (op == 0) toff += len, qoff += len;
            else if (op == 1 || op == 2) {
                if (k > 0 && k < n - 1 && (cigar[k - 1] & 0xf) == 0 && (cigar[k + 1] & 0xf) == 0) {
                    int l, prev_len = (int)(cigar[k - 1] >> 4);
                    if (op == 1) {
                        for (l = 0; l < prev_len; ++l)
                            if (qseq[qoff - 1 - l] != qseq[qoff + (int)len - 1 - l]) break;
                    } else {
                        for (l = 0; l < prev_len; ++l)
                            if (tseq[toff - 1 - l] != tseq[toff + (int)len - 1 - l]) break;
                    }
                    if (l > 0) cigar[k - 1] -= (uint32_t)l << 4, cigar[k + 1] += (uint32_t)l << 4, qoff -= l, toff -= l;
                    if (l == prev_len) to_shrink = 1;
                }
                if (op == 2) toff += len;
                else qoff += len;
            }
        }
        for (k = 0; k < n - 2; ++k) {
            if ((cigar[k] & 0xf) > 0 && (cigar[k] & 0xf) + (cigar[k + 1] & 0xf) == 3) {
                int l;
                uint32_t s[3] = {0, 0, 0};
                for (l = k; l < n; ++l) {
                    uint32_t op = cigar[l] & 0xf;
                    if (op == 1 || op == 2) s[op] += cigar[l] >> 4;
                    else break;
                }
                if (s[1] > 0 && s[2] > 0 && l - k > 2) {
                    cigar[k] = s[1] << 4 | 1;
                    cigar[k + 1] = s[2] << 4 | 2;
                    for (k += 2; k < l; ++k) cigar[k] &= 0xf;
                    to_shrink = 1;
                }
                k = l;
            }
        }
        if (to_shrink) {
            int l = 0;
            for (k = 0; k < n; ++k)
                if (cigar[k] >> 4 != 0) cigar[l++] = cigar[k];
            n = l;
            for (k = l = 0; k < n; ++k)
                if (k == n - 1 || (cigar[k] & 0xf) != (cigar[k + 1] & 0xf)) cigar[l++] = cigar[k];
                else cigar[k + 1] += cigar[k] >> 4 << 4;
            n = l;
        }
        if (n > 0 && ((cigar[0] & 0xf) == 1 || (cigar[0] & 0xf) == 2)) {
            int32_t l = (int32_t)(cigar[0] >> 4);
            if ((cigar[0] & 0xf) == 1) {
                if (r.rev) r.qe -= l;
                else r.qs += l;
                qshift = l;
            } else r.rs += l, tshift = l;
            --n;
            for (k = 0; k < n; ++k) cigar[k] = cigar[k + 1];
        }
        r.n_cigar = n;
    }
    qseq = qseq + qshift, tseq = tseq + tshift;
    toff = qoff = 0;
    double s = 0.0, mx = 0.0;
    r.blen = r.mlen = 0, r.n_ambi = 0;
    if (FLAT) {
        KB_WARP_RECONVERGE();  // FLAT is the one-thread-per-chain caller: every lane of the warp gets here (kb_stage_assemble)
        int32_t blen = 0, mlen = 0, n_ambi = 0;
        uint32_t op = 3, rem = 0;
        k = -1;
        for (;;) {
            if (rem == 0) {
                if (++k >= r.n_cigar) break;
                op = cigar[k] & 0xf, rem = cigar[k] >> 4;
                if (op == 1 || op == 2) {  // the gap's penalty does not depend on its bases: applied when the gap is entered
                    double pen = kb_dmul((double)P.e, (double)kb_log2_fast((float)(1.0 + rem)));
                    pen = kb_dadd((double)P.q, pen);
                    s = kb_dadd(s, -pen);
                    if (s < 0) s = 0;
                } else if (op != 0) rem = 0;
                continue;
            }
            if (op == 0) {
                const int cq = qseq[qoff], ct = tseq[toff];
                const bool amb = ct > 3 || cq > 3, diff = !amb && ct != cq;
                s = kb_dadd(s, amb ? (double)-P.sc_ambi : (diff ? (double)-P.b : (double)P.a));
                if (s < 0) s = 0;
                else mx = mx > s ? mx : s;
                n_ambi += amb, blen += !amb, mlen += !(amb || diff);
                ++qoff, ++toff;
            } else if (op == 1) {
                const bool amb = qseq[qoff] > 3;
                n_ambi += amb, blen += !amb, ++qoff;
            } else {
                const bool amb = tseq[toff] > 3;
                n_ambi += amb, blen += !amb, ++toff;
            }
            --rem;
        }
        r.blen = blen, r.mlen = mlen, r.n_ambi = n_ambi;
        r.dp_max = (int32_t)kb_dadd(mx, .499);
        return;
    }
    for (k = 0; k < r.n_cigar; ++k) {
        uint32_t op = cigar[k] & 0xf, len = cigar[k] >> 4;
        if (op == 0) {
            int n_ambi = 0, n_diff = 0;
            for (uint32_t l = 0; l < len; ++l) {
                int cq = qseq[qoff + l], ct = tseq[toff + l];
                if (ct > 3 || cq > 3) ++n_ambi, s = kb_dadd(s, (double)-P.sc_ambi);
                else if (ct != cq) ++n_diff, s = kb_dadd(s, (double)-P.b);
                else s = kb_dadd(s, (double)P.a);
                if (s < 0) s = 0;
                else mx = mx > s ? mx : s;
            }
            r.blen += len - n_ambi, r.mlen += len - (n_ambi + n_diff), r.n_ambi += n_ambi;
            toff += len, qoff += len;
        } else if (op == 1 || op == 2) {
            int n_ambi = 0;
            if (op == 1) {
                for (uint32_t l = 0; l < len; ++l)
                    if (qseq[qoff + l] > 3) ++n_ambi;
            } else {
                for (uint32_t l = 0; l < len; ++l)
                    if (tseq[toff + l] > 3) ++n_ambi;
            }
            r.blen += len - n_ambi, r.n_ambi += n_ambi;
            double pen = kb_dmul((double)P.e, (double)kb_log2_fast((float)(1.0 + len)));
            pen = kb_dadd((double)P.q, pen);
            s = kb_dadd(s, -pen);
            if (s < 0) s = 0;
            if (op == 1) qoff += len;
            else toff += len;
        }
    }
    r.dp_max = (int32_t)kb_dadd(mx, .499);
}

// minimap2 hit.c mm_split_reg
KB_HD void kb_split_reg(KbReg &r, KbReg &r2, int n, int qlen, const uint64_t *ax, const uint64_t *ay)
{
    if (n <= 0 || n >= r.cnt) return;
    r2 = r;
    r2.id = -1;
    r2.has_p = 0, r2.n_cigar = 0, r2.dp_score = r2.dp_max = r2.n_ambi = 0;
    r2.cnt = r.cnt - n;
    r2.score = (int32_t)kb_fadd(kb_fmul((float)r.score, kb_fdiv((float)r2.cnt, (float)r.cnt)), .499f);
    r2.as = r.as + n;
    if (r.parent == r.id) r2.parent = KB_PARENT_TMP_PRI;
    KbChainRec t;
    t.as = r2.as, t.cnt = r2.cnt;
    kb_reg_set_coor(t, qlen, ax, ay);
    r2.rev = t.rev, r2.rid = t.rid, r2.rs = t.rs, r2.re = t.re, r2.qs = t.qs, r2.qe = t.qe;
    r.cnt -= r2.cnt;
    r.score -= r2.score;
    t.as = r.as, t.cnt = r.cnt;
    kb_reg_set_coor(t, qlen, ax, ay);
    r.rev = t.rev, r.rid = t.rid, r.rs = t.rs, r.re = t.re, r.qs = t.qs, r.qe = t.qe;
}

// first half of minimap2 align.c mm_align1: trimmed anchor range, first / last anchor and the query / target windows
// the two end extensions may use.  Pure function of the chain and its (already seed-filtered) anchors.
struct KbWin {
    int32_t as1, cnt1, rs, qs, re, qe, rs0, qs0, re0, qe0;
    int32_t rev, rid, ctg, tlen_full, qlen;
    int64_t soff;
};
KB_HD void kb_align_window(const KbIndexView &ix, const KbBatchView &bt, int asm_id, int gene, int r_as, int r_cnt, int as1, int cnt1,
                           int n_a, const uint64_t *ax, const uint64_t *ay, KbWin &W)
{
    const kb_params_t &P = ix.p;
    const int qlen = ix.gene_len[gene];
    const int32_t rid = (int32_t)(ax[r_as] << 1 >> 33), rev = (int32_t)(ax[r_as] >> 63);
    const int ctg = bt.asm_ctg_start[asm_id] + rid;
    const int32_t tlen_full = bt.ctg_len[ctg];
    const int hk = P.k >> 1;
    int32_t i, l, rs0, re0, qs0, qe0, rs, re, qs, qe, rs1, qs1, re1, qe1;
    rs = (int32_t)ax[as1] - hk, qs = (int32_t)ay[as1] - hk;
    re = (int32_t)ax[as1 + cnt1 - 1] - hk, qe = (int32_t)ay[as1 + cnt1 - 1] - hk;

    rs0 = (int32_t)ax[r_as] + 1 - (int32_t)(ay[r_as] >> 32 & 0xff);
    qs0 = (int32_t)ay[r_as] + 1 - (int32_t)(ay[r_as] >> 32 & 0xff);
    if (rs0 < 0) rs0 = 0;
    rs1 = qs1 = 0;
    for (i = r_as - 1, l = 0; i >= 0 && ax[i] >> 32 == ax[r_as] >> 32; --i) {
        int32_t x = (int32_t)ax[i] + 1 - (int32_t)(ay[i] >> 32 & 0xff);
        int32_t y = (int32_t)ay[i] + 1 - (int32_t)(ay[i] >> 32 & 0xff);
        if (x < rs0 && y < qs0) {
            if (++l > P.min_cnt) {
                l = rs0 - x > qs0 - y ? rs0 - x : qs0 - y;
                rs1 = rs0 - l, qs1 = qs0 - l;
                if (rs1 < 0) rs1 = 0;
                break;
            }
        }
    }
    if (qs > 0 && rs > 0) {
        l = qs < P.max_gap ? qs : P.max_gap;
        qs1 = qs1 > qs - l ? qs1 : qs - l;
        qs0 = qs0 < qs1 ? qs0 : qs1;
        l += l * P.a > P.q ? (l * P.a - P.q) / P.e : 0;
        l = l < P.max_gap ? l : P.max_gap;
        l = l < rs ? l : rs;
        rs1 = rs1 > rs - l ? rs1 : rs - l;
        rs0 = rs0 < rs1 ? rs0 : rs1;
        rs0 = rs0 < rs ? rs0 : rs;
    } else rs0 = rs, qs0 = qs;
    re0 = (int32_t)ax[r_as + r_cnt - 1] + 1;
    qe0 = (int32_t)ay[r_as + r_cnt - 1] + 1;
    re1 = tlen_full, qe1 = qlen;
    for (i = r_as + r_cnt, l = 0; i < n_a && ax[i] >> 32 == ax[r_as] >> 32; ++i) {
        int32_t x = (int32_t)ax[i] + 1, y = (int32_t)ay[i] + 1;
        if (x > re0 && y > qe0) {
            if (++l > P.min_cnt) {
                l = x - re0 > y - qe0 ? x - re0 : y - qe0;
                re1 = re0 + l, qe1 = qe0 + l;
                break;
            }
        }
    }
    if (qe < qlen && re < tlen_full) {
        l = qlen - qe < P.max_gap ? qlen - qe : P.max_gap;
        qe1 = qe1 < qe + l ? qe1 : qe + l;
        qe0 = qe0 > qe1 ? qe0 : qe1;
        l += l * P.a > P.q ? (l * P.a - P.q) / P.e : 0;
        l = l < P.max_gap ? l : P.max_gap;
        l = l < tlen_full - re ? l : tlen_full - re;
        re1 = re1 < re + l ? re1 : re + l;
        re0 = re0 > re1 ? re0 : re1;
    } else re0 = re, qe0 = qe;
    W.as1 = as1, W.cnt1 = cnt1, W.rs = rs, W.qs = qs, W.re = re, W.qe = qe, W.rs0 = rs0, W.qs0 = qs0, W.re0 = re0, W.qe0 = qe0;
    W.rev = rev, W.rid = rid, W.ctg = ctg, W.tlen_full = tlen_full, W.qlen = qlen, W.soff = bt.ctg_soff[ctg];
}

// minimap2 align.c mm_align1 (long-read path).  ax/ay: the query's compacted anchors (n_a of them), minimap2 format.
// Returns an error code (0 = ok); r2.cnt > 0 when the region was split by a z-drop.
template <int NL>
KB_HD int kb_align1(const KbIndexView &ix, const KbBatchView &bt, int lane, int asm_id, int gene, KbReg &r, KbReg &r2,
                    int n_a, const uint64_t *ax, uint64_t *ay, const KbAlignScratch &S, int64_t *cell_counter)
{
    const kb_params_t &P = ix.p;
    const int hk = P.k >> 1;
    int32_t as1, cnt1, i, l, bw, bw_long, dropped = 0, rs0, re0, qs0, qe0, rs, re, qs, qe, rs1, qs1, re1, qe1;
    KbEz ez;

    r2.cnt = 0;
    if (r.cnt == 0) return 0;
    bw = P.ext_bw;
    bw_long = (int)(20000 * 1.5 + 1.);
    if (bw_long < bw) bw_long = bw;

    kb_fix_bad_ends(r.as, r.cnt, r.mlen, ax, ay, P.bw, P.min_chain_score * 2, &as1, &cnt1);
    if (lane == 0) {  // the seed filters set flags in ay[]: one writer, then everyone reads
        kb_filter_bad_seeds(as1, cnt1, ax, ay, 10, 40, P.max_gap >> 1, 10, S.off);
        kb_filter_bad_seeds_alt(as1, cnt1, ax, ay, 30, P.max_gap >> 1, S.off);
    }
    kb_sync<NL>();
    KbWin W;
    kb_align_window(ix, bt, asm_id, gene, r.as, r.cnt, as1, cnt1, n_a, ax, ay, W);
    const int qlen = W.qlen, rev = W.rev;
    const int32_t tlen_full = W.tlen_full;
    (void)tlen_full, (void)l;
    const int64_t soff = W.soff;
    const uint8_t *qseq0 = (rev ? ix.gseq_rev : ix.gseq_fwd) + ix.gene_seq_off[gene];
    rs = W.rs, qs = W.qs, re = W.re, qe = W.qe, rs0 = W.rs0, qs0 = W.qs0, re0 = W.re0, qe0 = W.qe0;

    if (re0 - rs0 > KB_TFULL_MAX || re0 <= rs0) return 2;
    // target window [rs0, re0) -> codes, cooperatively
    for (int x = lane; x < re0 - rs0; x += NL) S.tfull[x] = (uint8_t)kb_fetch_base(bt.seq2, bt.nmask, soff + rs0 + x);
    kb_sync<NL>();
    const uint8_t *tfull = S.tfull - rs0;  // tfull[pos] for pos in [rs0, re0)

    if (qs > 0 && rs > 0) {  // left extension on reversed sequences
        int ql = qs - qs0, tl = rs - rs0;
        if (ql <= KB_DP_MAXLEN && tl <= KB_DP_MAXLEN) {
            for (int x = lane; x < ql; x += NL) S.qbuf[x] = qseq0[qs - 1 - x];
            for (int x = lane; x < tl; x += NL) S.tbuf[x] = tfull[rs - 1 - x];
        }
        kb_sync<NL>();
        kb_dp<NL>(P, lane, ql, S.qbuf, tl, S.tbuf, bw, P.zdrop, KB_EZ_EXTZ_ONLY | KB_EZ_RIGHT | KB_EZ_REV_CIGAR, ez, S, cell_counter);
        if (ez.n_cigar < 0) return 3;
        if (ez.n_cigar > 0) {
            if (lane == 0) kb_append_cigar(r, S.cigar, ez.n_cigar, S.ezcig);
            r.has_p = 1;
            r.n_cigar = kb_bcast<NL>(r.n_cigar);
            r.dp_score += ez.max;
        }
        rs1 = rs - (ez.max_t + 1);
        qs1 = qs - (ez.max_q + 1);
    } else rs1 = rs, qs1 = qs;
    re1 = rs, qe1 = qs;

    for (i = 1; i < cnt1; ++i) {  // gap filling
        if ((ay[as1 + i] & (KB_SEED_IGNORE | KB_SEED_TANDEM)) && i != cnt1 - 1) continue;
        re = (int32_t)ax[as1 + i] - hk, qe = (int32_t)ay[as1 + i] - hk;
        re1 = re, qe1 = qe;
        if (i == cnt1 - 1 || (ay[as1 + i] & KB_SEED_LONG_JOIN) || (qe - qs >= P.min_ksw_len && re - rs >= P.min_ksw_len)) {
            int j, bw1 = bw_long, zdrop_code;
            const uint8_t *tseq = tfull + rs, *qseq = qseq0 + qs;
            if (ay[as1 + i] & KB_SEED_LONG_JOIN) bw1 = qe - qs > re - rs ? qe - qs : re - rs;
            kb_dp<NL>(P, lane, qe - qs, qseq, re - rs, tseq, bw1, -1, KB_EZ_GLOBAL_NO_ZDROP, ez, S, cell_counter);
            if (ez.n_cigar < 0) return 3;
            zdrop_code = ez.zdropped ? 1 : kb_test_zdrop(P, qseq, tseq, ez.n_cigar, S.ezcig);
            if (zdrop_code != 0) {
                kb_dp<NL>(P, lane, qe - qs, qseq, re - rs, tseq, bw1, P.zdrop, 0, ez, S, cell_counter);
                if (ez.n_cigar < 0) return 3;
            }
            if (ez.n_cigar > 0) {
                if (lane == 0) kb_append_cigar(r, S.cigar, ez.n_cigar, S.ezcig);
                r.has_p = 1;
                r.n_cigar = kb_bcast<NL>(r.n_cigar);
            }
            if (ez.zdropped) {
                r.has_p = 1;
                for (j = i - 1; j >= 0; --j)
                    if ((int32_t)ax[as1 + j] <= rs + ez.max_t) break;
                dropped = 1;
                if (j < 0) j = 0;
                r.dp_score += ez.max;
                re1 = rs + (ez.max_t + 1);
                qe1 = qs + (ez.max_q + 1);
                if (cnt1 - (j + 1) >= P.min_cnt) kb_split_reg(r, r2, as1 + j + 1 - r.as, qlen, ax, ay);
                break;
            } else r.dp_score += ez.score;
            rs = re, qs = qe;
        }
    }

    if (!dropped && qe < qe0 && re < re0) {  // right extension
        kb_dp<NL>(P, lane, qe0 - qe, qseq0 + qe, re0 - re, tfull + re, bw, P.zdrop, KB_EZ_EXTZ_ONLY, ez, S, cell_counter);
        if (ez.n_cigar < 0) return 3;
        if (ez.n_cigar > 0) {
            if (lane == 0) kb_append_cigar(r, S.cigar, ez.n_cigar, S.ezcig);
            r.has_p = 1;
            r.n_cigar = kb_bcast<NL>(r.n_cigar);
            r.dp_score += ez.max;
        }
        re1 = re + (ez.max_t + 1);
        qe1 = qe + (ez.max_q + 1);
    }
    if (r.n_cigar > KB_CIG_MAX) return 3;

    r.rs = rs1, r.re = re1;
    if (rev) r.qs = qlen - qe1, r.qe = qlen - qs1;
    else r.qs = qs1, r.qe = qe1;

    if (r.has_p) {
        kb_sync<NL>();
        if (lane == 0) kb_update_extra(P, r, S.cigar, qseq0 + qs1, tfull + rs1);
        kb_sync<NL>();
        r.qs = kb_bcast<NL>(r.qs), r.qe = kb_bcast<NL>(r.qe), r.rs = kb_bcast<NL>(r.rs);
        r.n_cigar = kb_bcast<NL>(r.n_cigar), r.blen = kb_bcast<NL>(r.blen), r.mlen = kb_bcast<NL>(r.mlen);
        r.n_ambi = kb_bcast<NL>(r.n_ambi), r.dp_max = kb_bcast<NL>(r.dp_max);
    }
    return 0;
}
