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

#endif  // __CUDACC__
