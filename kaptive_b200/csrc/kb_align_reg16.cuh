// kb_align_reg16.cuh -- the row-stripe wavefront of kb_align_reg.cuh with TWO cells per instruction (device only).
//
// Blackwell's DPX instructions work on packed signed halfwords (VIADDMNMX.S16x2, VIMNMX3.S16x2, VIMNMX.S16x2 with two
// predicate outputs, VIADD.16x2); kb_rows16 keeps every DP value of kb_rows as an x8-domain score with its priority tag
// in 16 bits and lets one warp act as 64 virtual lanes: the low halfword of every register belongs to virtual lane
// `lane`, the high halfword to virtual lane `lane + 32`.  A tile is 64 stripes of up to 8 columns (512 columns); virtual
// lane v computes row s - v of its stripe at step s, so the high half simply runs 32 steps behind the low half, and
// lane 0's high half takes its left neighbour from lane 31's low half (a rotating shuffle).  Recurrences, tie rules,
// per-anti-diagonal maxima / z-drop rule and traceback bytes are those of kb_rows (the spec is oracle/kb_oracle.c:extd2).
//
// Range.  A score of the rectangle lies in [-(gap(qlen) + gap(tlen) + q2 + e2 + q + e), a * min(qlen, tlen)]; times 8 plus
// the tag this must fit a signed halfword, which kb_rows16_eligible checks (genes are ~1 kb: with the default scoring
// any rectangle with qlen + tlen <= 3900 qualifies).  No value is ever a "minus infinity" sentinel: the states that
// kb_rows initialises with KB_NEG8 start here at "gap opened from the boundary cell", which the first real cell
// reproduces anyway (same value, same tag), so nothing can wrap.  Only unbanded rectangles are taken (|t - j| <= w for
// every cell); the banded ones and the few over the range stay with kb_rows.
//
// Substitution scores: the QUERY base of a row gives a word of four score bytes (against target A, C, G, T; all equal for
// an ambiguous query base), one word per half, built once per step; every column keeps a byte-permute selector made
// from its two target bases, so one PRMT yields both halves' sign-extended scores.  A tile that holds an ambiguous TARGET
// base (0.01 % of the bases) runs a variant of the row that patches those columns.
#pragma once
#ifdef __CUDACC__

#define KB_R16_RING 1024                      // anti-diagonals in the per-warp ring of kb_rows16 (TRACK)
#define KB_R16_RING_WORDS (KB_R16_RING + 8)   // + 8 alias words, folded back by drain()
#define KB_R16_MIN_TLEN 33                    // below this the 32-lane kernel's shorter ramp wins

__device__ __forceinline__ uint32_t kb_add2(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("add.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t kb_pack2(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
__device__ __forceinline__ int kb_lo16(uint32_t v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int kb_hi16(uint32_t v) { return (int)v >> 16; }

KB_HD bool kb_rows16_eligible(const KbDpConst &P, int qlen, int tlen, int w)
{
    if (qlen <= 0 || tlen < KB_R16_MIN_TLEN || qlen > 4000 || tlen > 4000) return false;
    if (tlen - 1 > w || qlen - 1 > w) return false;  // a band of the spec would bind: kb_rows masks it
    const int64_t tiles = (tlen + 511) / 512;
    if (tiles * (qlen + 63) * 512 > P.max_sw_cells || (int64_t)qlen * tlen > P.max_sw_cells) return false;
    const int lo = kb_gapcost2(P, qlen + 1) + kb_gapcost2(P, tlen + 1) + P.q + P.e + P.q2 + P.e2 + P.b + P.sc_ambi + 8;
    const int hi = P.a * (qlen < tlen ? qlen : tlen) + 8;
    return lo < 4000 && hi < 4000 && P.a <= 15 && P.b <= 15 && P.sc_ambi <= 15;
}

struct KbC16 {                      // kb_c8's constants, the same value in both halves
    uint32_t oe1, oe2, of1, of2;    // open + first extension with the state's tag
    uint32_t nx1, nx2;              // minus the extension cost
    uint32_t th1, th2;              // flag thresholds relative to H8, as ">=": -8 q + (rb ? 0 : 8)
    int rb;
};
__device__ __forceinline__ KbC16 kb_c16(const KbDpConst &P, int rb)
{
    KbC16 c;
    const int tE1 = rb ? 4 : 6, tF1 = 5, tE2 = rb ? 6 : 4, tF2 = rb ? 7 : 3;
    auto both = [](int v) { return kb_pack2(v, v); };
    c.rb = rb;
    c.oe1 = both(-8 * (P.q + P.e) + tE1), c.of1 = both(-8 * (P.q + P.e) + tF1);
    c.oe2 = both(-8 * (P.q2 + P.e2) + tE2), c.of2 = both(-8 * (P.q2 + P.e2) + tF2);
    c.nx1 = both(-8 * P.e), c.nx2 = both(-8 * P.e2);
    // kb_cell8 tests "state > H8 + (-8 q + (rb ? -1 : 7))"; the packed compare answers ">=", so the threshold is one higher
    c.th1 = both(-8 * P.q + (rb ? 0 : 8)), c.th2 = both(-8 * P.q2 + (rb ? 0 : 8));
    return c;
}
// score word of a query base: byte k = 8 * score(k, cq) + diagonal tag against target k = A, C, G, T
__device__ __forceinline__ uint32_t kb_qrow16(const KbDpConst &P, int tag_d, int cq)
{
    const uint32_t mm = (uint32_t)(-8 * P.b + tag_d) & 0xffu, ma = (uint32_t)(8 * P.a + tag_d) & 0xffu;
    if (cq > 3) return ((uint32_t)(-8 * P.sc_ambi + tag_d) & 0xffu) * 0x01010101u;
    return (mm * 0x01010101u) ^ ((mm ^ ma) << (8 * cq));
}
// byte-permute selector of a column: low half = sign-extended byte ct_lo of the first operand, high half = byte ct_hi of the second
__device__ __forceinline__ uint32_t kb_sel16(int ct_lo, int ct_hi)
{
    const uint32_t a = (uint32_t)(ct_lo & 3), b = (uint32_t)(ct_hi & 3) + 4;
    return a | (a | 8u) << 4 | b << 8 | (b | 8u) << 12;
}

// One packed cell = kb_cell8 on both halves.  d: low three bits of each half = tag of the winning candidate, bits 3..6 = flags.
__device__ __forceinline__ uint32_t kb_cell16(const KbC16 &c, uint32_t hu, uint32_t &e1, uint32_t &e2, uint32_t hl, uint32_t &f1, uint32_t &f2,
                                              uint32_t hd, uint32_t s, uint32_t &d)
{
    f1 = __viaddmax_s16x2(hl, c.of1, kb_add2(f1, c.nx1));
    f2 = __viaddmax_s16x2(hl, c.of2, kb_add2(f2, c.nx2));
    const uint32_t pre = __vmaxs2(__viaddmax_s16x2(hd, s, f1), f2);
    e1 = __viaddmax_s16x2(hu, c.oe1, kb_add2(e1, c.nx1));
    e2 = __viaddmax_s16x2(hu, c.oe2, kb_add2(e2, c.nx2));
    const uint32_t zk = __vimax3_s16x2(pre, e1, e2);
    const uint32_t z = zk & 0xfff8fff8u;
    const uint32_t t1 = kb_add2(z, c.th1), t2 = kb_add2(z, c.th2);
    d = zk & 0x00070007u;
    bool ph, pl;
    (void)__vibmax_s16x2(e1, t1, &ph, &pl);
    if (pl) d += 8u;
    if (ph) d += 8u << 16;
    (void)__vibmax_s16x2(f1, t1, &ph, &pl);
    if (pl) d += 16u;
    if (ph) d += 16u << 16;
    (void)__vibmax_s16x2(e2, t2, &ph, &pl);
    if (pl) d += 32u;
    if (ph) d += 32u << 16;
    (void)__vibmax_s16x2(f2, t2, &ph, &pl);
    if (pl) d += 64u;
    if (ph) d += 64u << 16;
    return z;
}

template <bool TN, bool TRACK, int M>
__device__ __forceinline__ void kb_rows16_cell(const KbC16 &c, uint32_t qlo, uint32_t qhi, const uint32_t (&sel)[8], uint32_t nbits, uint32_t sNw,
                                               uint32_t &hu, uint32_t &e1, uint32_t &e2, uint32_t &hd, uint32_t (&Hc)[8], uint32_t (&F1)[8],
                                               uint32_t (&F2)[8], uint32_t (&tb)[4], unsigned ring_lo, unsigned ring_hi, int32_t ckey_lo,
                                               int32_t ckey_hi, int nvm_lo, int nvm_hi)
{
    uint32_t s = (uint32_t)kb_prmt(qlo, qhi, sel[M]);
    if (TN) {  // ambiguous target base in this column of either half: the score is -sc_ambi whatever the query base
        const uint32_t m = ((nbits >> M) & 1u ? 0x0000ffffu : 0u) | ((nbits >> (8 + M)) & 1u ? 0xffff0000u : 0u);
        s = (s & ~m) | (sNw & m);
    }
    uint32_t d;
    const uint32_t hl = Hc[M];
    const uint32_t z = kb_cell16(c, hu, e1, e2, hl, F1[M], F2[M], hd, s, d);
    hd = hl, Hc[M] = z, hu = z;
    tb[M >> 1] += d << (8 * (M & 1));  // d < 128 per half: two columns share a halfword
    if (TRACK) {
        // key = (H + 2^19) << 12 | (4095 - t), as in kb_rows; slot M is M words past the half's slot-0 anti-diagonal
        const int zl = kb_lo16(z), zh = kb_hi16(z);
        asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 k;\n\t"
                     "setp.lt.s32 p, %3, %4;\n\tmad.lo.s32 k, %0, 512, %1;\n\t"
                     "@p red.shared.max.u32 [%2+%5], k;\n\t}" ::"r"(zl), "r"(ckey_lo - M), "r"(ring_lo), "n"(M), "r"(nvm_lo), "n"(4 * M)
                     : "memory");
        asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 k;\n\t"
                     "setp.lt.s32 p, %3, %4;\n\tmad.lo.s32 k, %0, 512, %1;\n\t"
                     "@p red.shared.max.u32 [%2+%5], k;\n\t}" ::"r"(zh), "r"(ckey_hi - M), "r"(ring_hi), "n"(M), "r"(nvm_hi), "n"(4 * M)
                     : "memory");
    }
}
template <bool TN, bool TRACK>
__device__ __forceinline__ void kb_rows16_body(const KbC16 &c, int kact, uint32_t qlo, uint32_t qhi, const uint32_t (&sel)[8], uint32_t nbits,
                                               uint32_t sNw, uint32_t &hu, uint32_t &e1, uint32_t &e2, uint32_t &hd, uint32_t (&Hc)[8],
                                               uint32_t (&F1)[8], uint32_t (&F2)[8], uint32_t (&tb)[4], unsigned ring_lo, unsigned ring_hi,
                                               int32_t ckey_lo, int32_t ckey_hi, int nvm_lo, int nvm_hi)
{
    tb[0] = tb[1] = tb[2] = tb[3] = 0;
#define KB_RC(M) \
    kb_rows16_cell<TN, TRACK, M>(c, qlo, qhi, sel, nbits, sNw, hu, e1, e2, hd, Hc, F1, F2, tb, ring_lo, ring_hi, ckey_lo, ckey_hi, nvm_lo, nvm_hi)
    switch (kact) {  // a stripe of kact < 8 columns lives in the LAST kact slots
    case 8: KB_RC(0);
    case 7: KB_RC(1);
    case 6: KB_RC(2);
    case 5: KB_RC(3);
    case 4: KB_RC(4);
    case 3: KB_RC(5);
    case 2: KB_RC(6);
    default: KB_RC(7);
    }
#undef KB_RC
}

// [mm2:ksw2.h:ksw_apply_zdrop] over the per-anti-diagonal keys rmax[0, n_diag), in order: see kb_rows for the derivation.
static __device__ __forceinline__ void kb_zdrop_scan(const KbDpConst &P, int lane, const uint32_t *rmax, int n_diag, int zdrop, KbEz &z_)
{
    const int chunk = (n_diag + 31) >> 5, lo = lane * chunk, hi = lo + chunk < n_diag ? lo + chunk : n_diag;
    int lmx = INT32_MIN, lr = -1, lt = 0;
    for (int r = lo; r < hi; ++r) {
        const uint32_t key = rmax[r];
        if (key == 0) continue;
        const int32_t h = (int32_t)(key >> 12) - (1 << 19);
        if (h > lmx) lmx = h, lr = r, lt = 4095 - (int32_t)(key & 4095u);
    }
    int smx = lmx, sr = lr, st = lt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int omx = __shfl_up_sync(0xffffffffu, smx, d), orr = __shfl_up_sync(0xffffffffu, sr, d), ot = __shfl_up_sync(0xffffffffu, st, d);
        if (lane >= d && omx >= smx) smx = omx, sr = orr, st = ot;  // the earlier chunk wins ties: first occurrence
    }
    int pmx = __shfl_up_sync(0xffffffffu, smx, 1), pr = __shfl_up_sync(0xffffffffu, sr, 1), pt = __shfl_up_sync(0xffffffffu, st, 1);
    if (lane == 0) pmx = INT32_MIN;
    int mx = 0, mt = -1, mq = -1;
    if (pmx > 0) mx = pmx, mt = pt, mq = pr - pt;
    int stop = INT32_MAX, bmx = 0, bmt = -1, bmq = -1;
    for (int r = lo; r < hi; ++r) {
        const uint32_t key = rmax[r];
        bool brk = key == 0;  // empty anti-diagonal (band excludes it): the spec stops here
        if (!brk) {
            const int32_t max_H = (int32_t)(key >> 12) - (1 << 19), max_t = 4095 - (int32_t)(key & 4095u);
            if (max_H > mx) mx = max_H, mt = max_t, mq = r - max_t;
            else if (max_t >= mt && r - max_t >= mq) {
                const int tl = max_t - mt, ql = (r - max_t) - mq, l = tl > ql ? tl - ql : ql - tl;
                brk = zdrop >= 0 && mx - max_H > zdrop + l * P.e2;
                if (brk) KB_DP_STAT_RAW(30, 1), KB_DP_STAT_RAW(31, (int64_t)r * 1000 / n_diag);
            }
        }
        if (brk) {
            stop = r, bmx = mx, bmt = mt, bmq = mq;
            break;
        }
    }
    const int first = __reduce_min_sync(0xffffffffu, stop);
    int owner = 31;  // no stop: the state after the last chunk
    if (first != INT32_MAX) owner = __ffs(__ballot_sync(0xffffffffu, stop == first)) - 1, mx = bmx, mt = bmt, mq = bmq;
    z_.zdropped = first != INT32_MAX;
    z_.max = __shfl_sync(0xffffffffu, mx, owner), z_.max_t = __shfl_sync(0xffffffffu, mt, owner), z_.max_q = __shfl_sync(0xffffffffu, mq, owner);
}

// S.wmax must point at KB_R16_RING_WORDS zeroed words of shared memory when TRACK.
template <bool TRACK, class SQ, class ST>
static __device__ __noinline__ void kb_rows16(const KbDpConst P, int lane, int qlen, const SQ qs, int tlen, const ST ts, int zdrop, int flag,
                                              KbEz &ez, const KbAlignScratch S, int64_t *cell_counter)
{
    const int rb = (flag & KB_EZ_RIGHT) ? 1 : 0;
    const KbC16 c = kb_c16(P, rb);
    const int tag_d = rb ? 3 : 7;
    const int ntile = (tlen + 511) >> 9, nstep = qlen + 63, n_diag = qlen + tlen - 1;
    const int klast = (tlen - ((ntile - 1) << 9) + 63) >> 6;  // columns per virtual lane in the last tile
    const size_t tile_bytes = (size_t)nstep * 512;
    uint8_t *tb = S.tb;
    int32_t *edge = S.dp;                                  // [parity][3][KB_DP_MAXLEN]
    uint32_t *rmax = reinterpret_cast<uint32_t *>(S.off);  // per anti-diagonal key (TRACK)
    const unsigned ring = TRACK ? (unsigned)__cvta_generic_to_shared(S.wmax) : 0u;
    const uint32_t sNw = kb_pack2(-8 * P.sc_ambi + tag_d, -8 * P.sc_ambi + tag_d);
    const int of1 = kb_lo16(c.of1), of2 = kb_lo16(c.of2), oe1 = kb_lo16(c.oe1), oe2 = kb_lo16(c.oe2);
    if (TRACK) {
        for (int r = lane; r < n_diag; r += 32) rmax[r] = 0;
        __syncwarp();
    }
    auto drain = [&](int r_lo) {  // fold the ring entries of anti-diagonals [r_lo, r_lo + KB_R16_RING) into rmax[]
        __syncwarp();
        if (lane < 8) {  // alias words first: word RING + x stands for word x
            const uint32_t a = S.wmax[KB_R16_RING + lane];
            if (a) {
                S.wmax[KB_R16_RING + lane] = 0;
                if (a > S.wmax[lane]) S.wmax[lane] = a;
            }
        }
        __syncwarp();
        for (int r = r_lo + lane; r < r_lo + KB_R16_RING; r += 32) {
            if (r < 0 || r >= n_diag) continue;
            const uint32_t v = S.wmax[r & (KB_R16_RING - 1)];
            if (v) {
                S.wmax[r & (KB_R16_RING - 1)] = 0;
                if (v > rmax[r]) rmax[r] = v;
            }
        }
        __syncwarp();
    };
    int32_t score = KB_NEG_INF;
    for (int tile = 0; tile < ntile; ++tile) {
        const bool spill = tile + 1 < ntile;  // then the tile is full width: its last column is slot 7 of lane 31's high half
        const int kact = spill ? 8 : klast, koff = 8 - kact;
        const int T0 = tile << 9;
        const int t0_lo = T0 + lane * kact, t0_hi = t0_lo + 32 * kact;     // first column of the two stripes
        const int t0s_lo = t0_lo - koff, t0s_hi = t0_hi - koff;            // slot m holds column t0s + m (m >= koff)
        const int32_t *ein = edge + (size_t)((tile & 1) ^ 1) * 3 * KB_DP_MAXLEN;
        int32_t *eout = edge + (size_t)(tile & 1) * 3 * KB_DP_MAXLEN;
        uint32_t Hc[8], F1[8], F2[8], sel[8];
        uint32_t nbits = 0;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int tl = t0s_lo + m, th = t0s_hi + m;
            const int cl = (m >= koff && tl < tlen) ? ts(tl) : 0, ch = (m >= koff && th < tlen) ? ts(th) : 0;
            sel[m] = kb_sel16(cl, ch);
            if (cl > 3) nbits |= 1u << m;
            if (ch > 3) nbits |= 1u << (8 + m);
            // virtual row j = -1 of the low half (the high half is initialised when its first row comes up)
            const int h0 = -8 * kb_gapcost2(P, tl + 1);
            Hc[m] = kb_pack2(h0, 0), F1[m] = kb_pack2(h0 + of1, 0), F2[m] = kb_pack2(h0 + of2, 0);
        }
        const bool tile_has_n = __any_sync(0xffffffffu, nbits != 0);
        int nval_lo = tlen - t0_lo, nval_hi = tlen - t0_hi;
        nval_lo = nval_lo < 0 ? 0 : (nval_lo > kact ? kact : nval_lo), nval_hi = nval_hi < 0 ? 0 : (nval_hi > kact ? kact : nval_hi);
        const int nvm_lo = nval_lo > 0 ? koff + nval_lo : 0, nvm_hi = nval_hi > 0 ? koff + nval_hi : 0;  // slots below nvm hold real columns
        // what a virtual lane offers to the next one: (H, E1, E2) of its last column in the row it has just finished
        uint32_t oh = kb_pack2(-8 * kb_gapcost2(P, t0_lo + kact), -8 * kb_gapcost2(P, t0_hi + kact)), oe1p = 0, oe2p = 0;
        uint32_t dg = kb_pack2(T0 == 0 ? 0 : -8 * kb_gapcost2(P, T0), 0);  // lane 0 low: H(T0 - 1, -1); everything else: set by the shuffles
        uint8_t *tbt = tb + (size_t)tile * tile_bytes + lane * 8;
        // the step at which the cell (tlen - 1, qlen - 1) is finished, and who owns it (last tile only)
        const int cc_fin = tlen - 1 - T0, v_fin = cc_fin / kact, ms_fin = koff + cc_fin % kact, s_fin = qlen - 1 + v_fin;
        int cq_lo_next = qs(lane == 0 ? 0 : qlen - 1), cq_hi_next = qs(qlen - 1);
        for (int s = 0; s < nstep; ++s) {
            const int cq_lo = cq_lo_next, cq_hi = cq_hi_next;
            const int j_lo = s - lane, j_hi = j_lo - 32;
            {
                int jn = j_lo + 1;
                jn = jn < 0 ? 0 : (jn >= qlen ? qlen - 1 : jn);
                cq_lo_next = qs(jn);
                jn = j_hi + 1;
                jn = jn < 0 ? 0 : (jn >= qlen ? qlen - 1 : jn);
                cq_hi_next = qs(jn);
            }
            const int src = (lane + 31) & 31;
            uint32_t uh = __shfl_sync(0xffffffffu, oh, src), ue1 = __shfl_sync(0xffffffffu, oe1p, src), ue2 = __shfl_sync(0xffffffffu, oe2p, src);
            const bool act_lo = (unsigned)j_lo < (unsigned)qlen, act_hi = (unsigned)j_hi < (unsigned)qlen;
            if (lane == 0) {  // low half: the rectangle's left edge or the previous tile's last column; high half: lane 31's low half
                int bh = 0, b1 = 0, b2 = 0;
                if (act_lo) {
                    if (T0 == 0) bh = -8 * kb_gapcost2(P, j_lo + 1), b1 = bh + oe1, b2 = bh + oe2;
                    else bh = kb_ld_s32(ein + j_lo), b1 = kb_ld_s32(ein + KB_DP_MAXLEN + j_lo), b2 = kb_ld_s32(ein + 2 * KB_DP_MAXLEN + j_lo);
                }
                uh = (uint32_t)kb_prmt((uint32_t)bh, uh, 0x5410), ue1 = (uint32_t)kb_prmt((uint32_t)b1, ue1, 0x5410);
                ue2 = (uint32_t)kb_prmt((uint32_t)b2, ue2, 0x5410);
            }
            if (j_hi == 0) {  // the high half's stripe starts now: virtual row j = -1
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const int h0 = -8 * kb_gapcost2(P, t0s_hi + m + 1);
                    Hc[m] = (uint32_t)kb_prmt(Hc[m], (uint32_t)h0, 0x5410);
                    F1[m] = (uint32_t)kb_prmt(F1[m], (uint32_t)(h0 + of1), 0x5410);
                    F2[m] = (uint32_t)kb_prmt(F2[m], (uint32_t)(h0 + of2), 0x5410);
                }
            }
            if (act_lo || act_hi) {
                const uint32_t qlo = kb_qrow16(P, tag_d, cq_lo), qhi = kb_qrow16(P, tag_d, cq_hi);
                uint32_t hu = uh, e1 = ue1, e2 = ue2, hd = dg;
                uint32_t tbw[4];
                const unsigned rs_lo = ring + (((unsigned)(t0s_lo + j_lo) & (KB_R16_RING - 1)) << 2);
                const unsigned rs_hi = ring + (((unsigned)(t0s_hi + j_hi) & (KB_R16_RING - 1)) << 2);
                const int nv_lo = act_lo ? nvm_lo : 0, nv_hi = act_hi ? nvm_hi : 0;
                if (tile_has_n)
                    kb_rows16_body<true, TRACK>(c, kact, qlo, qhi, sel, nbits, sNw, hu, e1, e2, hd, Hc, F1, F2, tbw, rs_lo, rs_hi,
                                                KB_ROWS_KEY_BIAS - t0s_lo, KB_ROWS_KEY_BIAS - t0s_hi, nv_lo, nv_hi);
                else
                    kb_rows16_body<false, TRACK>(c, kact, qlo, qhi, sel, nbits, sNw, hu, e1, e2, hd, Hc, F1, F2, tbw, rs_lo, rs_hi,
                                                 KB_ROWS_KEY_BIAS - t0s_lo, KB_ROWS_KEY_BIAS - t0s_hi, nv_lo, nv_hi);
                // a half that is not on a real row keeps its offer (the row -1 value its neighbour needs as a diagonal)
                const uint32_t keep = act_lo ? (act_hi ? 0x3210u : 0x7610u) : 0x3254u;
                oh = (uint32_t)kb_prmt(hu, oh, keep), oe1p = (uint32_t)kb_prmt(e1, oe1p, keep), oe2p = (uint32_t)kb_prmt(e2, oe2p, keep);
                // a virtual lane's slot is 8 bytes wide whatever kact is: low half at [lane * 8], high half 256 bytes on
                if (act_lo) {
                    const uint32_t w0 = (uint32_t)kb_prmt(tbw[0], tbw[1], 0x5410), w1 = (uint32_t)kb_prmt(tbw[2], tbw[3], 0x5410);
                    asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(__cvta_generic_to_global(tbt)), "r"(w0), "r"(w1) : "memory");
                }
                if (act_hi) {
                    const uint32_t w0 = (uint32_t)kb_prmt(tbw[0], tbw[1], 0x7632), w1 = (uint32_t)kb_prmt(tbw[2], tbw[3], 0x7632);
                    asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(__cvta_generic_to_global(tbt + 256)), "r"(w0), "r"(w1) : "memory");
                    if (spill && lane == 31)
                        kb_st_s32(eout + j_hi, kb_hi16(oh)), kb_st_s32(eout + KB_DP_MAXLEN + j_hi, kb_hi16(oe1p)),
                            kb_st_s32(eout + 2 * KB_DP_MAXLEN + j_hi, kb_hi16(oe2p));
                }
            }
            dg = uh;
            tbt += 512;
            if (!spill && s == s_fin) {  // H(tlen - 1, qlen - 1) has just been computed by virtual lane v_fin
                uint32_t hv = Hc[0];
#pragma unroll
                for (int m = 1; m < 8; ++m)
                    if (m == ms_fin) hv = Hc[m];
                hv = __shfl_sync(0xffffffffu, hv, v_fin & 31);
                score = (v_fin >> 5 ? kb_hi16(hv) : kb_lo16(hv)) >> 3;
            }
            if (TRACK && (s & 511) == 511) drain(T0 + s - 543);
        }
        if (TRACK) drain(T0 + nstep - 544), drain(T0 + nstep + 480);
        __syncwarp();  // spilled column visible to lane 0 of the next tile
    }
    KbEz z_;
    z_.max = 0, z_.max_q = z_.max_t = -1, z_.score = KB_NEG_INF, z_.zdropped = 0, z_.n_cigar = 0;
    if (TRACK) kb_zdrop_scan(P, lane, rmax, n_diag, zdrop, z_);
    if (!z_.zdropped) z_.score = score;
    if (cell_counter && lane == 0) *cell_counter += (int64_t)qlen * tlen;
    int i = -1, j = -1;
    if (!z_.zdropped && !(flag & KB_EZ_EXTZ_ONLY)) i = tlen - 1, j = qlen - 1;
    else if (z_.max_t >= 0 && z_.max_q >= 0) i = z_.max_t, j = z_.max_q;
    z_.n_cigar = kb_backtrack_warp(lane, i, j, rb, flag, S.ezcig, [&](int ii, int jj) -> uint32_t {
        const int tile = ii >> 9, cc = ii & 511, kk = tile + 1 < ntile ? 8 : klast;
        const int v = cc / kk;
        return (uint32_t)kb_ld_u8(tb + (size_t)tile * tile_bytes + (size_t)(jj + v) * 512 + v * 8 + (8 - kk) + (cc - v * kk));
    });
    ez = z_;
}

#endif  // __CUDACC__
