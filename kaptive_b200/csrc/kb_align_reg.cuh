// kb_align_reg.cuh -- register-resident variants of the anti-diagonal dual-affine DP (device only).
//
// Same recurrences, tie rules, per-anti-diagonal maximum / z-drop test and traceback bytes as
// kb_extd2 in kb_align.cuh (the spec is oracle/kb_oracle.c:extd2), but the previous two
// anti-diagonals never touch memory: target column t lives in lane (t & 31), register slot
// ((t >> 5) % M).  Along an anti-diagonal the "left" neighbour (t, j-1) is the lane's own slot, the
// "up" neighbour (t-1, j) is the previous lane (or lane 31 of the previous slot) and arrives by warp
// shuffle, and the diagonal neighbour (t-1, j-1) is the "up" value received one step earlier.
//
//   kb_extd2_reg        one pass, circular slots: needs min(qlen, tlen) <= 32 * (M - 1); tracks the
//                       per-anti-diagonal maximum (extension / z-drop mode) when TRACK
//   kb_extd2_reg_tiled  global alignment of any width: the target is cut into tiles of 32 * (M - 1)
//                       columns, a tile is swept over all query rows, and its last column (H, E1, E2
//                       per row) is spilled to per-warp scratch for the next tile's first column
//
// Both need a band that never binds (w >= qlen + tlen); then every anti-diagonal is a full slice of
// the rectangle and the traceback offset of a cell has a closed form (kb_ppos_rect) instead of the
// off[]/ppos[] arrays.  kb_dp_device dispatches; everything else falls back to kb_extd2.
#pragma once
#ifdef __CUDACC__

// All DP operands live in global memory; say so, otherwise loads through pointers that crossed a call are generic.
__device__ __forceinline__ int kb_ld_u8(const uint8_t *p)
{
    unsigned v;
    asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(__cvta_generic_to_global(p)));
    return (int)v;
}
__device__ __forceinline__ void kb_st_u8(uint8_t *p, int v)
{
    asm volatile("st.global.u8 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t kb_ld_s32(const int32_t *p)
{
    int32_t v;
    asm volatile("ld.global.s32 %0, [%1];" : "=r"(v) : "l"(__cvta_generic_to_global(p)));
    return v;
}
__device__ __forceinline__ void kb_st_s32(int32_t *p, int32_t v)
{
    asm volatile("st.global.s32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(v) : "memory");
}

// number of cells on anti-diagonals 0..r-1 of a qlen x tlen rectangle
__device__ __forceinline__ int kb_ppos_rect(int r, int qlen, int tlen)
{
    const int a = qlen < tlen ? qlen : tlen, b = qlen < tlen ? tlen : qlen;
    if (r <= a) return r * (r + 1) / 2;
    int base = a * (a + 1) / 2;
    if (r <= b) return base + (r - a) * a;
    base += (b - a) * a;
    const int k = r - b;  // anti-diagonals b .. r-1 have lengths a-1, a-2, ...
    return base + k * (a - 1) - k * (k - 1) / 2;
}

// one DP cell; returns H and updates the affine states and the traceback byte.
// rb = 1 selects the gap-preferring tie rule of KB_EZ_RIGHT: (x >= y) == (x + 1 > y) for integers.
__device__ __forceinline__ int32_t kb_cell(const KbDpConst &P, int rb, int32_t h_up, int32_t a1, int32_t a2, int32_t h_left,
                                           int32_t b1, int32_t b2, int32_t h_diag, int ct, int cq, int32_t &oa1, int32_t &oa2,
                                           int32_t &ob1, int32_t &ob2, int &d)
{
    a1 = max(h_up - P.q, a1) - P.e;
    a2 = max(h_up - P.q2, a2) - P.e2;
    b1 = max(h_left - P.q, b1) - P.e;
    b2 = max(h_left - P.q2, b2) - P.e2;
    int32_t z = h_diag + ((ct > 3 || cq > 3) ? -P.sc_ambi : (ct == cq ? P.a : -P.b));
    d = 0;
    if (a1 + rb > z) d = 1, z = a1;
    if (b1 + rb > z) d = 2, z = b1;
    if (a2 + rb > z) d = 3, z = a2;
    if (b2 + rb > z) d = 4, z = b2;
    const int32_t hq = z - P.q - rb, hq2 = z - P.q2 - rb;
    d |= (a1 > hq ? 0x08 : 0) | (b1 > hq ? 0x10 : 0) | (a2 > hq2 ? 0x20 : 0) | (b2 > hq2 ? 0x40 : 0);
    oa1 = a1, oa2 = a2, ob1 = b1, ob2 = b2;
    return z;
}

// ksw_backtrack over a full rectangle (no band): tb byte of cell (i, j) at kb_ppos_rect(i + j) + i - st(i + j).
// Runs of diagonal moves are resolved 32 at a time: every lane fetches the byte of (i - k, j - k) and a ballot
// finds how far the H-state diagonal run goes, so the dependent-load chain is paid once per run, not per base.
__device__ __forceinline__ void kb_backtrack_rect(int lane, int qlen, int tlen, int flag, KbEz &ez, const KbAlignScratch &S)
{
    const uint8_t *p = S.tb;
    uint32_t *cg = S.ezcig;
    int n_cigar = 0;
    int i = -1, j = -1;
    if (!ez.zdropped && !(flag & KB_EZ_EXTZ_ONLY)) i = tlen - 1, j = qlen - 1;
    else if (ez.max_t >= 0 && ez.max_q >= 0) i = ez.max_t, j = ez.max_q;
    int state = 0;
    uint32_t last = 0xffffffffu;  // last pushed op (lane 0 keeps the CIGAR, every lane tracks the counters)
    auto push = [&](uint32_t op, int len) {
        if (n_cigar == 0 || op != last) {
            if (lane == 0 && n_cigar < KB_CIG_MAX) cg[n_cigar] = (uint32_t)len << 4 | op;
            ++n_cigar, last = op;
        } else if (lane == 0 && n_cigar <= KB_CIG_MAX) cg[n_cigar - 1] += (uint32_t)len << 4;
    };
    while (i >= 0 && j >= 0) {
        const int ik = i - lane, jk = j - lane;  // the cell `lane` steps up the diagonal
        const bool in = ik >= 0 && jk >= 0;
        uint32_t tmp = 0xff;
        if (in) {
            const int r = ik + jk, st = r - qlen + 1 > 0 ? r - qlen + 1 : 0;
            tmp = (uint32_t)kb_ld_u8(p + kb_ppos_rect(r, qlen, tlen) + ik - st);
        }
        const uint32_t t0 = __shfl_sync(0xffffffffu, tmp, 0);
        // resolve the current cell exactly as ksw_backtrack does
        if (state == 0) state = t0 & 7;
        else if (!(t0 >> (state + 2) & 1)) state = 0;
        if (state == 0) state = t0 & 7;
        const unsigned run = __ballot_sync(0xffffffffu, in && (tmp & 7) == 0);
        if (state == 0) {
            // H state, diagonal move: also take the following cells whose own maximum is the diagonal
            int n = run == 0xffffffffu ? 32 : __ffs(~run) - 1;
            if (n < 1) n = 1;
            push(0, n), i -= n, j -= n;
        } else if (state == 1 || state == 3) push(2, 1), --i;
        else push(1, 1), --j;
    }
    if (i >= 0) push(2, i + 1);
    if (j >= 0) push(1, j + 1);
    __syncwarp();
    if (n_cigar > KB_CIG_MAX) n_cigar = -1;
    else if (!(flag & KB_EZ_REV_CIGAR) && lane == 0)
        for (int a = 0; a < n_cigar >> 1; ++a) {
            uint32_t t = cg[a];
            cg[a] = cg[n_cigar - 1 - a], cg[n_cigar - 1 - a] = t;
        }
    __syncwarp();
    ez.n_cigar = n_cigar;
}

// The register-resident DP.  tiled = false: one pass with circular slots (needs min(qlen, tlen) <= 224), optional
// per-anti-diagonal maximum / z-drop tracking.  tiled = true: global alignment of any width by 224-column tiles.
// One function (runtime flags, one copy of the unrolled slot loop) on purpose: the align kernel is instruction-cache
// sensitive, warps sit in different DP problems at any time.
static __device__ __noinline__ void kb_extd2_reg8(const KbDpConst P, int lane, int qlen, const uint8_t *__restrict__ qs, int tlen,
                                           const uint8_t *__restrict__ ts, int zdrop, int flag, bool tiled, KbEz &ez,
                                           const KbAlignScratch S, int64_t *cell_counter)
{
    constexpr int M = 8, TW = 32 * (M - 1);
    const int rb = (flag & KB_EZ_RIGHT) ? 1 : 0;
    const bool track = !(flag & KB_EZ_GLOBAL_NO_ZDROP);
    uint8_t *p = S.tb;
    int32_t *edge = S.dp;  // edge[parity][3][KB_DP_MAXLEN]: H, E1, E2 of the previous tile's last column, per query row
    KbEz z_;
    z_.max = 0, z_.max_q = z_.max_t = -1, z_.score = KB_NEG_INF, z_.zdropped = 0, z_.n_cigar = 0;
    int32_t last_h = 0;
    const int tstep = tiled ? TW : tlen;
    int tile = 0;
    for (int T0 = 0; T0 < tlen && !z_.zdropped; T0 += tstep, ++tile) {
        const int T1 = T0 + tstep < tlen ? T0 + tstep : tlen;  // columns [T0, T1); T0 is a multiple of 32
        const int32_t *ein = edge + (size_t)((tile & 1) ^ 1) * 3 * KB_DP_MAXLEN;
        int32_t *eout = edge + (size_t)(tile & 1) * 3 * KB_DP_MAXLEN;
        int32_t H1[M], HD[M], E1r[M], E2r[M], F1r[M], F2r[M];
#pragma unroll
        for (int m = 0; m < M; ++m) H1[m] = HD[m] = E1r[m] = E2r[m] = F1r[m] = F2r[m] = KB_NEG_INF;
        const int r_end = T1 - 1 + qlen - 1;
        int tbo = kb_ppos_rect(T0, qlen, tlen);  // traceback offset of anti-diagonal r, kept incrementally
        for (int r = T0; r <= r_end; ++r) {
            const int gst = r - qlen + 1 > 0 ? r - qlen + 1 : 0;  // the whole anti-diagonal is [gst, gen]
            const int gen = r < tlen - 1 ? r : tlen - 1;
            const int st = gst > T0 ? gst : T0;
            const int en = gen < T1 - 1 ? gen : T1 - 1;
            const int b_lo = st >> 5, b_hi = en >> 5;
            uint8_t *pr = p + tbo - gst;
            tbo += gen - gst + 1;
            int32_t max_H = INT32_MIN, max_t = 0x7fffffff;
            // slot M-1 is overwritten before slot 0 needs it as its wrap-around source: keep a copy
            const int32_t oH = H1[M - 1], oE1 = E1r[M - 1], oE2 = E2r[M - 1];
            const int bT0 = T0 >> 5;
            const bool new_col = en == r;  // the cell (t = r, j = 0) enters on this anti-diagonal (it sits in block b_hi)
#pragma unroll
            for (int m = M - 1; m >= 0; --m) {
                // the block of 32 columns in [b_lo, b_hi] that maps to slot m, if any
                const int b = b_lo + ((m - b_lo) & (M - 1));
                if (b > b_hi) continue;  // warp-uniform
                const int32_t sH = m > 0 ? H1[m > 0 ? m - 1 : 0] : oH;
                const int32_t sE1 = m > 0 ? E1r[m > 0 ? m - 1 : 0] : oE1;
                const int32_t sE2 = m > 0 ? E2r[m > 0 ? m - 1 : 0] : oE2;
                int32_t upH = __shfl_up_sync(0xffffffffu, H1[m], 1), wH = __shfl_sync(0xffffffffu, sH, 31);
                int32_t upE1 = __shfl_up_sync(0xffffffffu, E1r[m], 1), wE1 = __shfl_sync(0xffffffffu, sE1, 31);
                int32_t upE2 = __shfl_up_sync(0xffffffffu, E2r[m], 1), wE2 = __shfl_sync(0xffffffffu, sE2, 31);
                // Column t-1 of lane 0 is lane 31 of the previous block -- or, at the left edge of the rectangle / of the
                // tile, the virtual column -1 / the previous tile's spilled last column (warp-uniform branches).
                if (b == bT0) {
                    if (T0 == 0) wH = -kb_gapcost2(P, r + 1), wE1 = wE2 = KB_NEG_INF;  // (t = -1, j = r)
                    else if (lane == 0) {
                        const int j = r - T0;
                        wH = kb_ld_s32(ein + j), wE1 = kb_ld_s32(ein + KB_DP_MAXLEN + j), wE2 = kb_ld_s32(ein + 2 * KB_DP_MAXLEN + j);
                    }
                }
                if (lane == 0) upH = wH, upE1 = wE1, upE2 = wE2;
                const int t = (b << 5) + lane;
                if (new_col && b == b_hi && t == r) {  // virtual row -1 for the column that starts now
                    H1[m] = -kb_gapcost2(P, t + 1), F1r[m] = F2r[m] = KB_NEG_INF;
                    HD[m] = t == 0 ? 0 : -kb_gapcost2(P, t);
                }
                if (t >= st && t <= en) {
                    const int j = r - t;
                    int d;
                    const int32_t z = kb_cell(P, rb, upH, upE1, upE2, H1[m], F1r[m], F2r[m], HD[m], kb_ld_u8(ts + t), kb_ld_u8(qs + j), E1r[m],
                                              E2r[m], F1r[m], F2r[m], d);
                    HD[m] = upH;  // H(t-1, j): the diagonal neighbour of (t, j+1) on the next anti-diagonal
                    H1[m] = z;
                    kb_st_u8(pr + t, d);
                    if (z > max_H || (z == max_H && t < max_t)) max_H = z, max_t = t;
                    if (t == T1 - 1) {  // last column: spill it for the next tile; the last one ends with H(tlen-1, qlen-1)
                        if (tiled) kb_st_s32(eout + j, z), kb_st_s32(eout + KB_DP_MAXLEN + j, E1r[m]), kb_st_s32(eout + 2 * KB_DP_MAXLEN + j, E2r[m]);
                        last_h = z;
                    }
                }
            }
            if (track) {  // never together with tiled
                kb_reduce_max<32>(max_H, max_t);
                if (max_H > z_.max) {
                    z_.max = max_H, z_.max_t = max_t, z_.max_q = r - max_t;
                } else if (max_t >= z_.max_t && r - max_t >= z_.max_q) {
                    int tl = max_t - z_.max_t, ql = (r - max_t) - z_.max_q, l;
                    l = tl > ql ? tl - ql : ql - tl;
                    if (zdrop >= 0 && z_.max - max_H > zdrop + l * P.e2) {
                        z_.zdropped = 1;
                        break;
                    }
                }
            }
        }
        __syncwarp();  // edge column visible to the lane that owns the next tile's first column
    }
    // H(tlen-1, qlen-1) is the last value the owner of column tlen-1 produced
    if (!z_.zdropped) z_.score = __shfl_sync(0xffffffffu, last_h, (tlen - 1) & 31);
    if (cell_counter && lane == 0) *cell_counter += (int64_t)qlen * tlen;
    kb_backtrack_rect(lane, qlen, tlen, flag, z_, S);
    ez = z_;
}

// ---------------------------------------------------------------------------------------------------------------
// Certified band pass for the global gap fills (flag KB_EZ_GLOBAL_NO_ZDROP, band of the spec never binding).
//
// The 32 lanes hold 32 consecutive cells of one anti-diagonal, i.e. the 64 diagonals d = t - j in [dlo, dlo + 63]
// placed symmetrically around the two corners of the rectangle (d = 0 and d = tlen - qlen).  Everything outside is
// -inf.  Lane l works on t' = T(r') + l with T(r') = ceil((r' + dlo) / 2) in coordinates shifted by one (index 0 is
// the virtual boundary row / column), so T advances on every other anti-diagonal: on "A" steps the (t-1, j)
// neighbour comes from lane l-1 and (t, j-1) is the lane's own previous cell, on "B" steps (t, j-1) comes from lane
// l+1 and (t-1, j) is the lane's own; (t-1, j-1) is always the lane's own cell of two steps ago.  Six live registers
// per lane, three shuffles per anti-diagonal, one traceback word (4 anti-diagonals x 1 B) stored every fourth step.
//
// Exactness: a path that leaves the band must reach diagonal dlo - 1 or dlo + 64, which costs at least one gap of
// that many target bases and one of the matching number of query bases (a single gap is the cheapest way to spend a
// gap length because the dual-affine cost is sub-additive), and leaves at most tlen - D columns that can score +a.
// If the band score is strictly above that bound every optimal path of the full DP lies inside the band; on such a
// path the full DP's candidates that attain the maximum have identical values in the band DP and the others can
// only be lower, so every traceback decision (including ties) is the same.  tests/proto_band.py is the executable
// statement of this argument (band pass vs full DP on random and tandem inputs); the GPU parity tests cover it end
// to end.  Returns 1 when certified (ez complete, CIGAR in S.ezcig), 0 when the caller has to run the full DP.
#define KB_BAND_MIN_MARGIN 8
template <bool EDGE>
__device__ __forceinline__ void kb_band_step(const KbDpConst &P, int lane, bool stepB, int tp, int jp, int ct, int cq, int32_t &H1,
                                             int32_t &H2, int32_t &E1, int32_t &E2, int32_t &F1, int32_t &F2, uint32_t &acc)
{
    int32_t uH, uE1, uE2, lH, lF1, lF2;
    if (!stepB) {
        uH = __shfl_up_sync(0xffffffffu, H1, 1), uE1 = __shfl_up_sync(0xffffffffu, E1, 1), uE2 = __shfl_up_sync(0xffffffffu, E2, 1);
        if (lane == 0) uH = uE1 = uE2 = KB_NEG_INF;
        lH = H1, lF1 = F1, lF2 = F2;
    } else {
        lH = __shfl_down_sync(0xffffffffu, H1, 1), lF1 = __shfl_down_sync(0xffffffffu, F1, 1), lF2 = __shfl_down_sync(0xffffffffu, F2, 1);
        if (lane == 31) lH = lF1 = lF2 = KB_NEG_INF;
        uH = H1, uE1 = E1, uE2 = E2;
    }
    int d;
    int32_t z = kb_cell(P, 0, uH, uE1, uE2, lH, lF1, lF2, H2, ct, cq, E1, E2, F1, F2, d);
    if (EDGE) {
        if (tp <= 0 || jp <= 0) {  // virtual row / column (and cells before them, which no valid cell ever reads)
            const int m = tp > jp ? tp : jp;
            z = (tp < 0 || jp < 0) ? KB_NEG_INF : (m == 0 ? 0 : -kb_gapcost2(P, m));
            E1 = E2 = F1 = F2 = KB_NEG_INF, d = 0;
        }
    }
    H2 = H1, H1 = z;
    acc = (acc >> 8) | ((uint32_t)d << 24);
}

static __device__ __noinline__ int kb_global_band(const KbDpConst P, int lane, int qlen, const uint8_t *__restrict__ qs, int tlen,
                                                  const uint8_t *__restrict__ ts, int flag, KbEz &ez, const KbAlignScratch S,
                                                  int64_t *cell_counter)
{
    const int d1 = tlen - qlen, lo_d = d1 < 0 ? d1 : 0, hi_d = d1 > 0 ? d1 : 0;
    const int margin = (63 - (hi_d - lo_d)) >> 1;
    const int r_end = tlen + qlen;  // last anti-diagonal in shifted coordinates: the cell (tlen, qlen)
    if (margin < KB_BAND_MIN_MARGIN || (int64_t)32 * (r_end + 8) > P.max_sw_cells) return 0;
    const int dlo = lo_d - margin, dhi = dlo + 63;
    uint32_t *tbw = reinterpret_cast<uint32_t *>(S.tb);
    int32_t H1 = KB_NEG_INF, H2 = KB_NEG_INF, E1 = KB_NEG_INF, E2 = KB_NEG_INF, F1 = KB_NEG_INF, F2 = KB_NEG_INF;
    uint32_t acc = 0;
    int tp = ((dlo + 1) >> 1) + lane, jp = -tp;  // anti-diagonal 0
    int ct = 4, cq = 4;
    // anti-diagonals that can hold boundary cells: t' = 0 while r' <= -dlo, j' = 0 while r' <= dhi
    int r_edge = (-dlo > dhi ? -dlo : dhi) + 1;
    if ((r_edge + dlo) & 1) ++r_edge;  // the fast loop starts on an A step
    if (r_edge > r_end + 1) r_edge = r_end + 1;
    int rp = 0;
    for (; rp < r_edge; ++rp) {
        const bool stepB = (rp + dlo) & 1;
        if (rp > 0) {
            if (stepB) ++tp;
            else ++jp;
        }
        ct = kb_ld_u8(ts + (tp < 1 ? 0 : (tp > tlen ? tlen : tp) - 1));
        cq = kb_ld_u8(qs + (jp < 1 ? 0 : (jp > qlen ? qlen : jp) - 1));
        kb_band_step<true>(P, lane, stepB, tp, jp, ct, cq, H1, H2, E1, E2, F1, F2, acc);
        if ((rp & 3) == 3) tbw[(rp >> 2) * 32 + lane] = acc;
    }
    // interior: every in-range cell has real neighbours; cells past the far edges compute garbage nobody reads
    for (; rp + 1 <= r_end; rp += 2) {
        ++jp;
        cq = kb_ld_u8(qs + (jp > qlen ? qlen : jp) - 1);
        kb_band_step<false>(P, lane, false, tp, jp, ct, cq, H1, H2, E1, E2, F1, F2, acc);
        if ((rp & 3) == 3) tbw[(rp >> 2) * 32 + lane] = acc;
        ++tp;
        ct = kb_ld_u8(ts + (tp > tlen ? tlen : tp) - 1);
        kb_band_step<false>(P, lane, true, tp, jp, ct, cq, H1, H2, E1, E2, F1, F2, acc);
        if (((rp + 1) & 3) == 3) tbw[((rp + 1) >> 2) * 32 + lane] = acc;
    }
    if (rp == r_end) {
        ++jp;
        cq = kb_ld_u8(qs + (jp > qlen ? qlen : jp) - 1);
        kb_band_step<false>(P, lane, false, tp, jp, ct, cq, H1, H2, E1, E2, F1, F2, acc);
        if ((rp & 3) == 3) tbw[(rp >> 2) * 32 + lane] = acc;
    }
    if ((r_end & 3) != 3) tbw[(r_end >> 2) * 32 + lane] = acc >> (8 * (3 - (r_end & 3)));
    if (cell_counter && lane == 0) *cell_counter += (int64_t)32 * (r_end + 1);
    const int score = __shfl_sync(0xffffffffu, H1, (tlen - ((r_end + dlo + 1) >> 1)) & 31);
    {  // certificate
        const int D_hi = dhi + 1, I_hi = D_hi - d1, I_lo = 1 - dlo, D_lo = I_lo + d1;
        int b_hi = (tlen - D_hi < 0 || qlen - I_hi < 0) ? KB_NEG_INF : P.a * (tlen - D_hi) - kb_gapcost2(P, D_hi) - kb_gapcost2(P, I_hi);
        int b_lo = (tlen - D_lo < 0 || qlen - I_lo < 0) ? KB_NEG_INF : P.a * (tlen - D_lo) - kb_gapcost2(P, D_lo) - kb_gapcost2(P, I_lo);
        if (!(score > (b_hi > b_lo ? b_hi : b_lo))) return 0;
    }
    __syncwarp();
    // ksw_backtrack from the far corner; diagonal runs 32 cells at a time (a diagonal keeps its lane index)
    uint32_t *cg = S.ezcig;
    int n_cigar = 0, i = tlen - 1, j = qlen - 1, state = 0;
    uint32_t last = 0xffffffffu;
    auto push = [&](uint32_t op, int len) {
        if (n_cigar == 0 || op != last) {
            if (lane == 0 && n_cigar < KB_CIG_MAX) cg[n_cigar] = (uint32_t)len << 4 | op;
            ++n_cigar, last = op;
        } else if (lane == 0 && n_cigar <= KB_CIG_MAX) cg[n_cigar - 1] += (uint32_t)len << 4;
    };
    while (i >= 0 && j >= 0) {
        const int l = i + 1 - ((i + j + 2 + dlo + 1) >> 1);
        if ((unsigned)l > 31u) return 0;  // cannot happen once certified
        const int ik = i - lane, jk = j - lane;
        const bool in = ik >= 0 && jk >= 0;
        uint32_t tmp = 0xff;
        if (in) {
            const int r2 = ik + jk + 2;
            tmp = (tbw[(r2 >> 2) * 32 + l] >> ((r2 & 3) * 8)) & 0xffu;
        }
        const uint32_t t0 = __shfl_sync(0xffffffffu, tmp, 0);
        if (state == 0) state = t0 & 7;
        else if (!(t0 >> (state + 2) & 1)) state = 0;
        if (state == 0) state = t0 & 7;
        const unsigned run = __ballot_sync(0xffffffffu, in && (tmp & 7) == 0);
        if (state == 0) {
            int n = run == 0xffffffffu ? 32 : __ffs(~run) - 1;
            if (n < 1) n = 1;
            push(0, n), i -= n, j -= n;
        } else if (state == 1 || state == 3) push(2, 1), --i;
        else push(1, 1), --j;
    }
    if (i >= 0) push(2, i + 1);
    if (j >= 0) push(1, j + 1);
    __syncwarp();
    if (n_cigar > KB_CIG_MAX) n_cigar = -1;
    else if (!(flag & KB_EZ_REV_CIGAR) && lane == 0)
        for (int a = 0; a < n_cigar >> 1; ++a) {
            uint32_t t = cg[a];
            cg[a] = cg[n_cigar - 1 - a], cg[n_cigar - 1 - a] = t;
        }
    __syncwarp();
    ez.max = 0, ez.max_q = ez.max_t = -1, ez.zdropped = 0;
    ez.score = score, ez.n_cigar = n_cigar;
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// Row-stripe wavefront: the DP of a whole rectangle (any band half-width w of the spec, |d| = |t - j| <= w).
//
// The target is cut into tiles of 32 * K columns; inside a tile lane l owns the K consecutive columns
// t0 .. t0 + K - 1 (t0 = T0 + l * K) and at step s computes row j = s - l of them, left to right.  Along the row
// the (t-1, j) neighbour and its two E states are the values the lane has just produced (for its first column:
// what lane l-1 produced one step earlier, by shuffle), (t, j-1) and its F states are the column's own registers,
// (t-1, j-1) is the left column's previous H.  Per lane 3 * K state registers, per step 3 shuffles for 32 * K
// cells, and only 31 ramp steps per tile.  The last column of a tile (H, E1, E2 per row) is spilled for lane 0 of
// the next tile.  Traceback bytes go to tb[tile][s][column-in-tile]: 32 * K contiguous bytes per step.
//
// Extension / z-drop mode (TRACK): ksw2 evaluates, per anti-diagonal r, the maximum H and the lowest t attaining
// it.  Every cell folds key = (H + 2^19) << 12 | (4095 - t) into a per-warp shared-memory ring (red.shared.max,
// conflict free: lanes sit 7 anti-diagonals apart), the ring is drained into rmax[r] every 256 steps, and the
// sequential z-drop rule runs over rmax[] after the last tile.  Cells beyond a z-drop are computed in vain but
// never influence the result (a traceback only moves towards smaller r).
template <bool MASK, bool TRACK, int K>
__device__ __forceinline__ void kb_rows_body(const KbDpConst &P, int rb, int w, int d0, int nval, unsigned ring, int rbase, int32_t ckey,
                                             int cq, int32_t &hu, int32_t &e1, int32_t &e2, int32_t &hd, int32_t (&Hc)[K],
                                             int32_t (&F1)[K], int32_t (&F2)[K], const int (&tc)[K], uint32_t (&tbw)[(K + 3) / 4])
{
#pragma unroll
    for (int m = 0; m < K; ++m) {
        int d;
        const int32_t hl = Hc[m];
        int32_t z = kb_cell(P, rb, hu, e1, e2, hl, F1[m], F2[m], hd, tc[m], cq, e1, e2, F1[m], F2[m], d);
        bool ok = true;
        if (MASK) {
            ok = (unsigned)(d0 + m + w) <= (unsigned)(2 * w);
            if (!ok) z = e1 = e2 = F1[m] = F2[m] = KB_NEG_INF, d = 0;
        }
        hd = hl, Hc[m] = z, hu = z;
        if ((m & 3) == 0) tbw[m >> 2] = (uint32_t)d;
        else tbw[m >> 2] |= (uint32_t)d << (8 * (m & 3));
        if (TRACK) {
            if (m < nval && ok) {
                const uint32_t key = (uint32_t)(z * 4096 + (ckey - m));
                asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(ring + (((unsigned)(rbase + m) & 511u) << 2)), "r"(key) : "memory");
            }
        }
    }
}

#define KB_ROWS_KEY_BIAS (int32_t)(0x80000000u + 4095u)
template <int K, bool TRACK>
static __device__ __noinline__ void kb_rows(const KbDpConst P, int lane, int qlen, const uint8_t *__restrict__ qs, int tlen,
                                            const uint8_t *__restrict__ ts, int w, int zdrop, int flag, KbEz &ez, const KbAlignScratch S,
                                            int64_t *cell_counter)
{
    constexpr int TW = 32 * K;
    const int rb = (flag & KB_EZ_RIGHT) ? 1 : 0;
    const int ntile = (tlen + TW - 1) / TW, nstep = qlen + 31, n_diag = qlen + tlen - 1;
    const size_t tile_bytes = (size_t)nstep * TW;
    uint8_t *tb = S.tb;
    int32_t *edge = S.dp;                                   // [parity][3][KB_DP_MAXLEN]
    uint32_t *rmax = reinterpret_cast<uint32_t *>(S.off);   // per anti-diagonal key (TRACK)
    const unsigned ring = (unsigned)__cvta_generic_to_shared(S.wmax);
    const bool banded = tlen - 1 > w || qlen - 1 > w;
    if (TRACK) {
        for (int r = lane; r < n_diag; r += 32) rmax[r] = 0;
        __syncwarp();
    }
    auto drain = [&](int r_lo) {  // fold ring entries of anti-diagonals [r_lo, r_lo + 512) into rmax[]
        __syncwarp();
        for (int r = r_lo + lane; r < r_lo + 512; r += 32) {
            if (r < 0 || r >= n_diag) continue;
            const uint32_t v = S.wmax[r & 511];
            if (v) {
                S.wmax[r & 511] = 0;
                if (v > rmax[r]) rmax[r] = v;
            }
        }
        __syncwarp();
    };
    int32_t score = KB_NEG_INF;
    for (int tile = 0; tile < ntile; ++tile) {
        const int T0 = tile * TW, t0 = T0 + lane * K;
        const int32_t *ein = edge + (size_t)((tile & 1) ^ 1) * 3 * KB_DP_MAXLEN;
        int32_t *eout = edge + (size_t)(tile & 1) * 3 * KB_DP_MAXLEN;
        const bool spill = tile + 1 < ntile;  // then the tile is full width and its last column is lane 31's last
        int32_t Hc[K], F1[K], F2[K];
        int tc[K];
#pragma unroll
        for (int m = 0; m < K; ++m) {
            const int t = t0 + m;
            Hc[m] = -kb_gapcost2(P, t + 1), F1[m] = F2[m] = KB_NEG_INF;  // virtual row j = -1
            tc[m] = t < tlen ? kb_ld_u8(ts + t) : 4;
        }
        int nval = tlen - t0;
        nval = nval < 0 ? 0 : (nval > K ? K : nval);
        // what the lane offers to lane + 1: (H, E1, E2) of its last column in the row it has just finished
        int32_t oh = -kb_gapcost2(P, t0 + K), oe1 = KB_NEG_INF, oe2 = KB_NEG_INF;
        int32_t dg = T0 == 0 ? 0 : -kb_gapcost2(P, T0);  // lane 0: H(T0 - 1, -1); other lanes: set by the first shuffle
        uint8_t *tbt = tb + (size_t)tile * tile_bytes + lane * K;
        for (int s = 0; s < nstep; ++s) {
            int32_t uh = __shfl_up_sync(0xffffffffu, oh, 1), ue1 = __shfl_up_sync(0xffffffffu, oe1, 1), ue2 = __shfl_up_sync(0xffffffffu, oe2, 1);
            const int j = s - lane;
            const bool act = (unsigned)j < (unsigned)qlen;
            if (lane == 0 && act) {
                if (T0 == 0) uh = -kb_gapcost2(P, j + 1), ue1 = ue2 = KB_NEG_INF;
                else uh = kb_ld_s32(ein + j), ue1 = kb_ld_s32(ein + KB_DP_MAXLEN + j), ue2 = kb_ld_s32(ein + 2 * KB_DP_MAXLEN + j);
            }
            const int d0 = t0 - j;
            const bool edge_lane = banded && act && (d0 < -w || d0 + K - 1 > w);
            const bool any_edge = banded && __any_sync(0xffffffffu, edge_lane);
            if (act) {
                const int cq = kb_ld_u8(qs + j);
                int32_t hu = uh, e1 = ue1, e2 = ue2, hd = dg;
                uint32_t tbw[(K + 3) / 4];
                if (any_edge) kb_rows_body<true, TRACK, K>(P, rb, w, d0, nval, ring, t0 + j, KB_ROWS_KEY_BIAS - t0, cq, hu, e1, e2, hd, Hc, F1, F2, tc, tbw);
                else kb_rows_body<false, TRACK, K>(P, rb, w, d0, nval, ring, t0 + j, KB_ROWS_KEY_BIAS - t0, cq, hu, e1, e2, hd, Hc, F1, F2, tc, tbw);
                oh = hu, oe1 = e1, oe2 = e2;
                uint32_t *dst = reinterpret_cast<uint32_t *>(tbt + (size_t)s * TW);
#pragma unroll
                for (int x = 0; x < (K + 3) / 4; ++x) asm volatile("st.global.u32 [%0], %1;" ::"l"(__cvta_generic_to_global(dst + x)), "r"(tbw[x]) : "memory");
                if (spill && lane == 31) kb_st_s32(eout + j, oh), kb_st_s32(eout + KB_DP_MAXLEN + j, oe1), kb_st_s32(eout + 2 * KB_DP_MAXLEN + j, oe2);
            }
            dg = uh;
            if (TRACK && (s & 255) == 255) drain(T0 + s - 287);
        }
        if (TRACK) drain(T0 + nstep - 288), drain(T0 + nstep + 224);
        if (!spill) {  // H(tlen - 1, qlen - 1): the last row of column tlen - 1
            const int c = tlen - 1 - T0, ms = c % K;
            int32_t hv = Hc[0];
#pragma unroll
            for (int m = 1; m < K; ++m)
                if (m == ms) hv = Hc[m];
            score = __shfl_sync(0xffffffffu, hv, c / K);
        }
        __syncwarp();  // spilled column visible to lane 0 of the next tile
    }
    KbEz z_;
    z_.max = 0, z_.max_q = z_.max_t = -1, z_.score = KB_NEG_INF, z_.zdropped = 0, z_.n_cigar = 0;
    if (TRACK) {  // [mm2:ksw2.h:ksw_apply_zdrop] over the per-anti-diagonal maxima, in order
        int zd = 0, mx = 0, mt = -1, mq = -1;
        if (lane == 0) {
            for (int r = 0; r < n_diag; ++r) {
                const uint32_t key = rmax[r];
                if (key == 0) {  // empty anti-diagonal (band excludes it): the spec stops here
                    zd = 1;
                    break;
                }
                const int32_t max_H = (int32_t)(key >> 12) - (1 << 19), max_t = 4095 - (int32_t)(key & 4095u);
                if (max_H > mx) mx = max_H, mt = max_t, mq = r - max_t;
                else if (max_t >= mt && r - max_t >= mq) {
                    const int tl = max_t - mt, ql = (r - max_t) - mq, l = tl > ql ? tl - ql : ql - tl;
                    if (zdrop >= 0 && mx - max_H > zdrop + l * P.e2) {
                        zd = 1;
                        break;
                    }
                }
            }
        }
        z_.zdropped = __shfl_sync(0xffffffffu, zd, 0), z_.max = __shfl_sync(0xffffffffu, mx, 0);
        z_.max_t = __shfl_sync(0xffffffffu, mt, 0), z_.max_q = __shfl_sync(0xffffffffu, mq, 0);
    }
    if (!z_.zdropped) z_.score = score;
    if (cell_counter && lane == 0) *cell_counter += (int64_t)qlen * tlen;
    // ksw_backtrack; diagonal runs are resolved 32 cells at a time
    uint32_t *cg = S.ezcig;
    int n_cigar = 0, i = -1, j = -1, state = 0;
    if (!z_.zdropped && !(flag & KB_EZ_EXTZ_ONLY)) i = tlen - 1, j = qlen - 1;
    else if (z_.max_t >= 0 && z_.max_q >= 0) i = z_.max_t, j = z_.max_q;
    uint32_t last = 0xffffffffu;
    auto push = [&](uint32_t op, int len) {
        if (n_cigar == 0 || op != last) {
            if (lane == 0 && n_cigar < KB_CIG_MAX) cg[n_cigar] = (uint32_t)len << 4 | op;
            ++n_cigar, last = op;
        } else if (lane == 0 && n_cigar <= KB_CIG_MAX) cg[n_cigar - 1] += (uint32_t)len << 4;
    };
    while (i >= 0 && j >= 0) {
        const int ik = i - lane, jk = j - lane;
        const bool in = ik >= 0 && jk >= 0;
        uint32_t tmp = 0xff;
        if (in) {
            const int c = ik & (TW - 1);
            tmp = (uint32_t)kb_ld_u8(tb + (size_t)(ik / TW) * tile_bytes + (size_t)(jk + c / K) * TW + c);
        }
        const uint32_t t0v = __shfl_sync(0xffffffffu, tmp, 0);
        if (state == 0) state = t0v & 7;
        else if (!(t0v >> (state + 2) & 1)) state = 0;
        if (state == 0) state = t0v & 7;
        const unsigned run = __ballot_sync(0xffffffffu, in && (tmp & 7) == 0);
        if (state == 0) {
            int n = run == 0xffffffffu ? 32 : __ffs(~run) - 1;
            if (n < 1) n = 1;
            push(0, n), i -= n, j -= n;
        } else if (state == 1 || state == 3) push(2, 1), --i;
        else push(1, 1), --j;
    }
    if (i >= 0) push(2, i + 1);
    if (j >= 0) push(1, j + 1);
    __syncwarp();
    if (n_cigar > KB_CIG_MAX) n_cigar = -1;
    else if (!(flag & KB_EZ_REV_CIGAR) && lane == 0)
        for (int a = 0; a < n_cigar >> 1; ++a) {
            uint32_t t = cg[a];
            cg[a] = cg[n_cigar - 1 - a], cg[n_cigar - 1 - a] = t;
        }
    __syncwarp();
    z_.n_cigar = n_cigar;
    ez = z_;
}

#endif  // __CUDACC__
