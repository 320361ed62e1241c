// kb_align_reg16.cuh -- the row-stripe wavefront of kb_align_reg.cuh with TWO cells per instruction (device only).
//
// Blackwell's DPX instructions work on packed signed halfwords (VIADDMNMX.S16x2, VIMNMX3.S16x2, VIMNMX.S16x2 with two
// predicate outputs, VIADD.16x2).  kb_rows16 keeps every DP value of kb_rows as an x8-domain score with its priority tag in
// 16 bits and runs TWO DP JOBS in one warp: the low halfword of every register belongs to job A, the high halfword to job B.
// Geometry, step schedule and neighbour shuffles are exactly those of kb_rows (tiles of 256 columns, lane = a stripe of up
// to 8 columns, row s - lane at step s), so a pair costs one pass of kb_rows over the larger of the two rectangles;
// the kernel sorts its jobs by size and pairs neighbours, which keeps that maximum close to both.  Recurrences, tie rules,
// per-anti-diagonal maxima / z-drop rule and traceback bytes are those of kb_rows (the spec is oracle/kb_oracle.c:extd2);
// the two jobs may differ in every parameter (lengths, band, z-drop, left / right tie rule).
// (A first version made one job 64 virtual lanes wide instead: correct, but the 63-step ramp and the 64-column rounding ate
// the gain -- 109 thread instructions per useful cell pair; profiles/r2_summary.md.)
//
// Range.  A score of the rectangle lies in [-(gap(qlen) + gap(tlen) + q2 + e2 + q + e), a * min(qlen, tlen)]; times 8 plus
// the tag this must fit a signed halfword, which kb_rows16_eligible checks (genes are ~1 kb: with the default scoring
// any rectangle with qlen + tlen <= 3600 qualifies).  No value is ever a "minus infinity" sentinel: the states that
// kb_rows initialises with KB_NEG8 start here at "gap opened from the boundary cell", which the first real cell
// reproduces anyway (same value, same tag), so nothing can wrap.  Cells the band of the spec excludes (|t - j| > w) are set to
// KB_NEG16, a value below every real one of an eligible rectangle that is re-imposed on every such cell, so it cannot decay
// towards the wrap either; rectangles over the range stay with kb_rows.
//
// Substitution scores: the QUERY base of a row gives a word of four score bytes (against target A, C, G, T; all equal for
// an ambiguous query base), one word per job, built once per step; every column keeps a byte-permute selector made
// from the two jobs' target bases, so one PRMT yields both halves' sign-extended scores.  A tile that holds an ambiguous
// TARGET base (0.01 % of the bases) runs a variant of the row that patches those columns.
#pragma once
#ifdef __CUDACC__

#define KB_R16_RING_WORDS (2 * KB_RING_WORDS)  // one ring of kb_rows' size per job of the pair
// stripe width of the last tile: rounded up to an even number of columns per lane, so that the row code exists in four widths instead of
// eight (the copies share the instruction cache of an SM; the extra column of an odd width costs less than the misses)
#ifndef KB_R16_EVEN
#define KB_R16_EVEN 1
#endif
#define KB_R16_KACT(k) (KB_R16_EVEN ? (((k) + 1) & ~1) : (k))
// per-warp shared memory of kb_rows16: the two rings, a 5-entry table of query score words per job, and the two query segments
// as nt4 bytes in DP order with 32 bytes of padding on either side (a lane reads row s - lane for -31 <= s - lane < qlen + 31)
#define KB_R16_QMAX 2048
#define KB_R16_LUT_OFF KB_R16_RING_WORDS
#define KB_R16_SQ_OFF (KB_R16_RING_WORDS + 16)
#define KB_R16_SQ_WORDS ((KB_R16_QMAX + 64) / 4)
#define KB_R16_SMEM_WORDS (KB_R16_SQ_OFF + 2 * KB_R16_SQ_WORDS)

__device__ __forceinline__ uint32_t kb_add2(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("add.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t kb_pack2(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
__device__ __forceinline__ int kb_lo16(uint32_t v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int kb_hi16(uint32_t v) { return (int)v >> 16; }

KB_HD bool kb_rows16_eligible(const KbDpConst &P, int qlen, int tlen, int w, bool track)
{
    if (qlen <= 0 || tlen <= 0 || qlen > KB_R16_QMAX || tlen > 4000) return false;
    const int dlen = tlen > qlen ? tlen - qlen : qlen - tlen;
    if (w < 2 || (!track && dlen >= w)) return false;  // as kb_rows_eligible: a global alignment whose end cell the band excludes
    // each job of a pair gets half of the warp's traceback scratch; a pair's tiles and steps are those of its larger member
    // (kb_rows16_pair_fits); two jobs that do not fit together are run one after the other, each with the scratch to itself
    const int64_t tiles = (tlen + 255) / 256;
    if (tiles * (qlen + 31) * 256 > P.max_sw_cells || (int64_t)qlen * tlen > P.max_sw_cells) return false;
    const int lo = kb_gapcost2(P, qlen + 1) + kb_gapcost2(P, tlen + 1) + P.q + P.e + P.q2 + P.e2 + P.b + P.sc_ambi + 8;
    const int hi = P.a * (qlen < tlen ? qlen : tlen) + 8;
    return lo < 3700 && hi < 4000 && P.a <= 15 && P.b <= 15 && P.sc_ambi <= 15;  // lo: real values stay above KB_NEG16 + gap costs
}

struct KbC16 {                      // kb_c8's constants: low half for job A (tie rule rbA), high half for job B (rbB)
    uint32_t oe1, oe2, of1, of2;    // open + first extension with the state's tag
    uint32_t nx1, nx2;              // minus the extension cost
    uint32_t th1, th2;              // flag thresholds relative to H8, as ">=": -8 q + (rb ? 0 : 8)
};
__device__ __forceinline__ KbC16 kb_c16(const KbDpConst &P, int rbA, int rbB)
{
    KbC16 c;
    auto tE1 = [](int rb) { return rb ? 4 : 6; };
    auto tE2 = [](int rb) { return rb ? 6 : 4; };
    auto tF2 = [](int rb) { return rb ? 7 : 3; };
    const int tF1 = 5;
    c.oe1 = kb_pack2(-8 * (P.q + P.e) + tE1(rbA), -8 * (P.q + P.e) + tE1(rbB)), c.of1 = kb_pack2(-8 * (P.q + P.e) + tF1, -8 * (P.q + P.e) + tF1);
    c.oe2 = kb_pack2(-8 * (P.q2 + P.e2) + tE2(rbA), -8 * (P.q2 + P.e2) + tE2(rbB));
    c.of2 = kb_pack2(-8 * (P.q2 + P.e2) + tF2(rbA), -8 * (P.q2 + P.e2) + tF2(rbB));
    c.nx1 = kb_pack2(-8 * P.e, -8 * P.e), c.nx2 = kb_pack2(-8 * P.e2, -8 * P.e2);
    // kb_cell8 tests "state > H8 + (-8 q + (rb ? -1 : 7))"; the packed compare answers ">=", so the threshold is one higher
    c.th1 = kb_pack2(-8 * P.q + (rbA ? 0 : 8), -8 * P.q + (rbB ? 0 : 8)), c.th2 = kb_pack2(-8 * P.q2 + (rbA ? 0 : 8), -8 * P.q2 + (rbB ? 0 : 8));
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

// value of a cell the band of the spec excludes: far below every real value of an eligible rectangle, far above -32768 - gap costs
#define KB_NEG16 (-30000)

// One row of a job pair: the first kact of 8 cells, straight line with an exit after every cell (slot m holds column
// t0 + m); ONE copy of the code serves every stripe width, which keeps the hot loop of the kernel inside the instruction cache
// (a copy per width ran 2.4 x slower).  SLOW adds what few rows need: ambiguous target bases (nbits), cells outside a band.
template <bool SLOW, bool TRACK, int KACT>
__device__ __forceinline__ void kb_rows16_row(const KbC16 &c, uint32_t qlo, uint32_t qhi, const uint32_t (&sel)[8], uint32_t nbits,
                                              uint32_t sNw, uint32_t &hu, uint32_t &e1, uint32_t &e2, uint32_t hd, uint32_t (&Hc)[8],
                                              uint32_t (&F1)[8], uint32_t (&F2)[8], uint32_t (&tb)[4], unsigned ring_lo, unsigned ring_hi,
                                              uint32_t tinv, int nv_lo, int nv_hi, int w_lo, int w_hi, int d0)
{
    tb[0] = tb[1] = tb[2] = tb[3] = 0;
#pragma unroll
    for (int M = 0; M < KACT; ++M) {
        uint32_t s = (uint32_t)kb_prmt(qlo, qhi, sel[M]);
        if (SLOW) {  // ambiguous target base in this column of either job: the score is -sc_ambi whatever the query base
            const uint32_t m = ((nbits >> M) & 1u ? 0x0000ffffu : 0u) | ((nbits >> (8 + M)) & 1u ? 0xffff0000u : 0u);
            s = (s & ~m) | (sNw & m);
        }
        uint32_t d;
        const uint32_t hl = Hc[M];
        uint32_t z = kb_cell16(c, hu, e1, e2, hl, F1[M], F2[M], hd, s, d);
        bool okl = true, okh = true;
        if (SLOW) {
            okl = (unsigned)(d0 + M + w_lo) <= (unsigned)(2 * w_lo), okh = (unsigned)(d0 + M + w_hi) <= (unsigned)(2 * w_hi);
            const uint32_t m = (okl ? 0u : 0x0000ffffu) | (okh ? 0u : 0xffff0000u);
            if (m) {
                const uint32_t neg = kb_pack2(KB_NEG16, KB_NEG16) & m;
                z = (z & ~m) | neg, e1 = (e1 & ~m) | neg, e2 = (e2 & ~m) | neg, F1[M] = (F1[M] & ~m) | neg, F2[M] = (F2[M] & ~m) | neg, d &= ~m;
            }
        }
        hd = hl, Hc[M] = z, hu = z;
        tb[M >> 1] += d << (8 * (M & 1));  // d < 128 per half: two columns share a halfword
        if (TRACK) {
            // key = (H8 ^ 0x8000) << 16 | (65535 - t): unsigned order = (H, lowest t); slot M is M words past slot 0's anti-diagonal in the
            // job's own ring; a key of 0 (job not on a real row, padded column, cell outside the band) leaves the ring unchanged
            const uint32_t zx = z ^ 0x80008000u, tm = tinv - M;
            uint32_t kl = (uint32_t)kb_prmt(zx, tm, 0x1054), kh = (uint32_t)kb_prmt(zx, tm, 0x3254);
            kl = (M < nv_lo && okl) ? kl : 0u, kh = (M < nv_hi && okh) ? kh : 0u;  // okl / okh are constants in the fast form
            asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(ring_lo + 4 * M), "r"(kl) : "memory");
            asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(ring_hi + 4 * M), "r"(kh) : "memory");
        }
    }
}
// dispatch on the stripe width: the fast form has a straight-line copy per width (no exits, no register shuffling at joins), the
// rarely used slow form one copy with an exit after every cell
template <bool SLOW, bool TRACK>
__device__ __forceinline__ void kb_rows16_rowk(const KbC16 &c, int kact, uint32_t qlo, uint32_t qhi, const uint32_t (&sel)[8], uint32_t nbits,
                                               uint32_t sNw, uint32_t &hu, uint32_t &e1, uint32_t &e2, uint32_t hd, uint32_t (&Hc)[8],
                                               uint32_t (&F1)[8], uint32_t (&F2)[8], uint32_t (&tb)[4], unsigned ring_lo, unsigned ring_hi,
                                               uint32_t tinv, int nv_lo, int nv_hi, int w_lo, int w_hi, int d0)
{
#define KB_RK(K) \
    kb_rows16_row<SLOW, TRACK, K>(c, qlo, qhi, sel, nbits, sNw, hu, e1, e2, hd, Hc, F1, F2, tb, ring_lo, ring_hi, tinv, nv_lo, nv_hi, w_lo, w_hi, d0)
#if KB_R16_EVEN
    switch (kact) {
    case 8: KB_RK(8); break;
    case 6: KB_RK(6); break;
    case 4: KB_RK(4); break;
    default: KB_RK(2); break;
    }
#else
    switch (kact) {
    case 8: KB_RK(8); break;
    case 7: KB_RK(7); break;
    case 6: KB_RK(6); break;
    case 5: KB_RK(5); break;
    case 4: KB_RK(4); break;
    case 3: KB_RK(3); break;
    case 2: KB_RK(2); break;
    default: KB_RK(1); break;
    }
#endif
#undef KB_RK
}

// [mm2:ksw2.h:ksw_apply_zdrop] over the per-anti-diagonal keys rmax[0, n_diag), in order: see kb_rows for the derivation.
// KEY16: keys of kb_rows16, (H8 ^ 0x8000) << 16 | (65535 - t); otherwise keys of kb_rows, (H + 2^19) << 12 | (4095 - t).
template <bool KEY16>
__device__ __forceinline__ void kb_key_decode(uint32_t key, int32_t &h, int32_t &t)
{
    if (KEY16) h = (int32_t)(short)((key >> 16) ^ 0x8000u) >> 3, t = 65535 - (int32_t)(key & 0xffffu);
    else h = (int32_t)(key >> 12) - (1 << 19), t = 4095 - (int32_t)(key & 4095u);
}
template <bool KEY16>
static __device__ __forceinline__ void kb_zdrop_scan(const KbDpConst &P, int lane, const uint32_t *rmax, int n_diag, int zdrop, KbEz &z_)
{
    const int chunk = (n_diag + 31) >> 5, lo = lane * chunk, hi = lo + chunk < n_diag ? lo + chunk : n_diag;
    int lmx = INT32_MIN, lr = -1, lt = 0;
    for (int r = lo; r < hi; ++r) {
        const uint32_t key = rmax[r];
        if (key == 0) continue;
        int32_t h, t;
        kb_key_decode<KEY16>(key, h, t);
        if (h > lmx) lmx = h, lr = r, lt = t;
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
            int32_t max_H, max_t;
            kb_key_decode<KEY16>(key, max_H, max_t);
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

// columns per lane of tile `tile` for a pair whose targets are tlenA / tlenB long (0: absent): the wider need of the two
__device__ __forceinline__ int kb_pair_kact(int tlenA, int tlenB, int tile)
{
    const int tl = tlenA > tlenB ? tlenA : tlenB, ntile = (tl + 255) >> 8;
    return tile + 1 < ntile ? 8 : KB_R16_KACT((tl - ((ntile - 1) << 8) + 31) >> 5);
}
KB_HD bool kb_rows16_pair_fits(const KbDpConst &P, int qlenA, int tlenA, int qlenB, int tlenB)
{
    const int ql = qlenA > qlenB ? qlenA : qlenB, tl = tlenA > tlenB ? tlenA : tlenB;
    return (int64_t)((tl + 255) >> 8) * (ql + 31) * 256 <= P.max_sw_cells / 2;
}

// One job of a pair, as the DP pass sees it.  An absent job B has qlen = tlen = 0.
template <class SQ, class ST>
struct KbPairJob {
    int qlen, tlen, w, zdrop, flag;
    SQ qs;
    ST ts;
};
struct KbPairOut {  // what the DP pass leaves for kb_rows16_finish
    int32_t scoreA, scoreB;
    int nstep;
};

// The DP pass of a pair: traceback bytes of job A in S.tb[0, max_sw_cells / 2), of job B in the second half; per-anti-diagonal
// keys in S.off (A) and S.off + 2 * KB_DP_MAXLEN (B) when TRACK.  S.wmax: KB_R16_RING_WORDS zeroed words of shared memory.
template <bool TRACK, class SQ, class ST>
static __device__ __noinline__ void kb_rows16_dp(const KbDpConst P, int lane, const KbPairJob<SQ, ST> A, const KbPairJob<SQ, ST> B, KbPairOut &out,
                                                 const KbAlignScratch S)
{
    const int rbA = (A.flag & KB_EZ_RIGHT) ? 1 : 0, rbB = (B.flag & KB_EZ_RIGHT) ? 1 : 0;
    const KbC16 c = kb_c16(P, rbA, rbB);
    const int tagA = rbA ? 3 : 7, tagB = rbB ? 3 : 7;
    const int qmax = A.qlen > B.qlen ? A.qlen : B.qlen, tmax = A.tlen > B.tlen ? A.tlen : B.tlen;
    const int ntile = (tmax + 255) >> 8, nstep = qmax + 31;
    const int ntA = (A.tlen + 255) >> 8, ntB = (B.tlen + 255) >> 8;
    const int ndA = A.qlen + A.tlen - 1, ndB = B.qlen + B.tlen - 1;
    const bool bandedA = A.tlen - 1 > A.w || A.qlen - 1 > A.w, bandedB = B.qlen > 0 && (B.tlen - 1 > B.w || B.qlen - 1 > B.w);
    const size_t tile_bytes = (size_t)nstep * 256;
    uint8_t *tbA = S.tb, *tbB = S.tb + (P.max_sw_cells >> 1);
    uint32_t *edge = reinterpret_cast<uint32_t *>(S.dp);   // [parity][3][KB_DP_MAXLEN], packed like the registers
    uint32_t *rmaxA = reinterpret_cast<uint32_t *>(S.off), *rmaxB = rmaxA + 2 * KB_DP_MAXLEN;
    uint32_t *wmA = S.wmax, *wmB = S.wmax + KB_RING_WORDS;
    const unsigned ringA = TRACK ? (unsigned)__cvta_generic_to_shared(wmA) : 0u, ringB = ringA + 4 * KB_RING_WORDS;
    const uint32_t sNw = kb_pack2(-8 * P.sc_ambi + tagA, -8 * P.sc_ambi + tagB);
    const int of1 = kb_lo16(c.of1);
    const uint32_t dOF1 = c.of1, dOF2 = c.of2, dOE1 = c.oe1, dOE2 = c.oe2;
    out.scoreA = out.scoreB = KB_NEG_INF, out.nstep = nstep;
    (void)of1;
    // the two query segments as bytes in shared memory (row j at byte 32 + j), and each job's table of score words by query base
    const unsigned sm_base = (unsigned)__cvta_generic_to_shared(S.wmax);
    const unsigned sqA = sm_base + 4 * KB_R16_SQ_OFF + 32, sqB = sqA + 4 * KB_R16_SQ_WORDS, lutA = sm_base + 4 * KB_R16_LUT_OFF, lutB = lutA + 32;
    {
        uint8_t *ba = reinterpret_cast<uint8_t *>(S.wmax + KB_R16_SQ_OFF) + 32, *bb = ba + 4 * KB_R16_SQ_WORDS;
        for (int x = lane; x < A.qlen; x += 32) ba[x] = (uint8_t)A.qs(x);
        for (int x = lane; x < B.qlen; x += 32) bb[x] = (uint8_t)B.qs(x);
        if (lane < 5) S.wmax[KB_R16_LUT_OFF + lane] = kb_qrow16(P, tagA, lane), S.wmax[KB_R16_LUT_OFF + 8 + lane] = kb_qrow16(P, tagB, lane);
        __syncwarp();
    }
    auto lds8 = [](unsigned a) {
        unsigned v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
        return v;
    };
    auto lds32 = [](unsigned a) {
        unsigned v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
        return v;
    };
    if (TRACK) {
        for (int r = lane; r < ndA; r += 32) rmaxA[r] = 0;
        for (int r = lane; r < ndB; r += 32) rmaxB[r] = 0;
        __syncwarp();
    }
    auto drain1 = [&](uint32_t *wm, uint32_t *rmax, int n_diag, int r_lo) {  // as kb_rows: ring entries [r_lo, r_lo + 512) into rmax[]
        if (lane < 8) {  // alias words first: word 512 + x stands for word x
            const uint32_t a = wm[512 + lane];
            if (a) {
                wm[512 + lane] = 0;
                if (a > wm[lane]) wm[lane] = a;
            }
        }
        __syncwarp();
        for (int r = r_lo + lane; r < r_lo + 512; r += 32) {
            if (r < 0 || r >= n_diag) continue;
            const uint32_t v = wm[r & 511];
            if (v) {
                wm[r & 511] = 0;
                if (v > rmax[r]) rmax[r] = v;
            }
        }
    };
    auto drain = [&](int r_lo) {
        __syncwarp();
        drain1(wmA, rmaxA, ndA, r_lo);
        drain1(wmB, rmaxB, ndB, r_lo);
        __syncwarp();
    };
    for (int tile = 0; tile < ntile; ++tile) {
        const bool spill = tile + 1 < ntile;  // then the tile is full width and its last column is lane 31's slot 7
        const int kact = spill ? 8 : KB_R16_KACT((tmax - ((ntile - 1) << 8) + 31) >> 5);
        const int T0 = tile << 8, t0 = T0 + lane * kact;  // slot m holds column t0 + m
        const bool onA = tile < ntA, onB = tile < ntB;      // a job whose target ends in an earlier tile sits this one out
        const uint32_t *ein = edge + (size_t)((tile & 1) ^ 1) * 3 * KB_DP_MAXLEN;
        uint32_t *eout = edge + (size_t)(tile & 1) * 3 * KB_DP_MAXLEN;
        uint32_t Hc[8], F1[8], F2[8], sel[8];
        uint32_t nbits = 0;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int t = t0 + m;
            const int ca = (m < kact && t < A.tlen) ? A.ts(t) : 0, cb = (m < kact && t < B.tlen) ? B.ts(t) : 0;
            sel[m] = kb_sel16(ca, cb);
            if (ca > 3) nbits |= 1u << m;
            if (cb > 3) nbits |= 1u << (8 + m);
            const int h0 = -8 * kb_gapcost2(P, t + 1);  // virtual row j = -1
            Hc[m] = kb_pack2(h0, h0), F1[m] = kb_add2(Hc[m], dOF1), F2[m] = kb_add2(Hc[m], dOF2);
        }
        const bool tile_slow = __any_sync(0xffffffffu, nbits != 0);  // an ambiguous target base: every row of the tile takes the slow form
        int nvA = A.tlen - t0, nvB = B.tlen - t0;  // real columns of the stripe in either job
        nvA = nvA < 0 ? 0 : (nvA > kact ? kact : nvA), nvB = nvB < 0 ? 0 : (nvB > kact ? kact : nvB);
        // what the lane offers to lane + 1: (H, E1, E2) of its last column in the row it has just finished
        const int hoff = -8 * kb_gapcost2(P, t0 + kact);
        uint32_t oh = kb_pack2(hoff, hoff), oe1p = 0, oe2p = 0;
        const int dg0 = T0 == 0 ? 0 : -8 * kb_gapcost2(P, T0);
        uint32_t dg = kb_pack2(dg0, dg0);  // lane 0: H(T0 - 1, -1); other lanes: set by the first shuffle
        uint8_t *tbt = tbA + (size_t)tile * tile_bytes + lane * 8;
        const size_t b_off = (size_t)(P.max_sw_cells >> 1);
        // the step at which the cell (tlen - 1, qlen - 1) of a job is finished, its lane and slot (the job's last tile only)
        const int ccA = A.tlen - 1 - T0, ccB = B.tlen - 1 - T0;
        const int sfA = (tile == ntA - 1) ? A.qlen - 1 + ccA / kact : -1, sfB = (tile == ntB - 1) ? B.qlen - 1 + ccB / kact : -1;
        // score words of the row a lane is on, fetched one step ahead (bytes outside [0, qlen) are never used: the job is not active there)
        uint32_t qlo_next = lds32(lutA + 4 * (lds8(sqA - lane) & 7u)), qhi_next = lds32(lutB + 4 * (lds8(sqB - lane) & 7u));
        for (int s = 0; s < nstep; ++s) {
            const uint32_t qlo = qlo_next, qhi = qhi_next;
            const int j = s - lane;
            qlo_next = lds32(lutA + 4 * (lds8(sqA + j + 1) & 7u)), qhi_next = lds32(lutB + 4 * (lds8(sqB + j + 1) & 7u));
            uint32_t uh = __shfl_up_sync(0xffffffffu, oh, 1), ue1 = __shfl_up_sync(0xffffffffu, oe1p, 1), ue2 = __shfl_up_sync(0xffffffffu, oe2p, 1);
            const bool actA = onA && (unsigned)j < (unsigned)A.qlen, actB = onB && (unsigned)j < (unsigned)B.qlen;
            if (lane == 0 && (actA || actB)) {  // the rectangle's left edge, or the previous tile's last column
                if (T0 == 0) {
                    const int bh = -8 * kb_gapcost2(P, j + 1);
                    uh = kb_pack2(bh, bh), ue1 = kb_add2(uh, dOE1), ue2 = kb_add2(uh, dOE2);
                } else uh = kb_ld_u32(ein + j), ue1 = kb_ld_u32(ein + KB_DP_MAXLEN + j), ue2 = kb_ld_u32(ein + 2 * KB_DP_MAXLEN + j);
            }
            // band of the spec: a lane whose stripe touches |t - j| > w in this row (either job)
            const int d0 = t0 - j;  // diagonal of slot 0
            bool slow = tile_slow;
            if (bandedA || bandedB) {
                const bool edge_lane = (bandedA && actA && (d0 < -A.w || d0 + kact - 1 > A.w)) || (bandedB && actB && (d0 < -B.w || d0 + kact - 1 > B.w));
                slow = slow || __any_sync(0xffffffffu, edge_lane);
            }
            if (actA || actB) {
                uint32_t hu = uh, e1 = ue1, e2 = ue2;
                uint32_t tbw[4];
                const unsigned ro = ((unsigned)(t0 + j) & 511u) << 2;
                int nv_lo = actA ? nvA : 0, nv_hi = actB ? nvB : 0;
                asm volatile("" : "+r"(nv_lo), "+r"(nv_hi));  // two plain registers: otherwise every cell re-derives them from their conditions
                if (slow)
                    kb_rows16_rowk<true, TRACK>(c, kact, qlo, qhi, sel, nbits, sNw, hu, e1, e2, dg, Hc, F1, F2, tbw, ringA + ro, ringB + ro,
                                               0xffffu - (uint32_t)t0, nv_lo, nv_hi, A.w, B.w, d0);
                else
                    kb_rows16_rowk<false, TRACK>(c, kact, qlo, qhi, sel, nbits, sNw, hu, e1, e2, dg, Hc, F1, F2, tbw, ringA + ro, ringB + ro,
                                                0xffffu - (uint32_t)t0, nv_lo, nv_hi, A.w, B.w, d0);
                oh = hu, oe1p = e1, oe2p = e2;
                // a lane's slot is 8 bytes wide whatever kact is; job A's bytes are the low halves, job B's the high halves
                if (actA) {
                    const uint32_t w0 = (uint32_t)kb_prmt(tbw[0], tbw[1], 0x5410), w1 = (uint32_t)kb_prmt(tbw[2], tbw[3], 0x5410);
                    asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(__cvta_generic_to_global(tbt)), "r"(w0), "r"(w1) : "memory");
                }
                if (actB) {
                    const uint32_t w0 = (uint32_t)kb_prmt(tbw[0], tbw[1], 0x7632), w1 = (uint32_t)kb_prmt(tbw[2], tbw[3], 0x7632);
                    asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(__cvta_generic_to_global(tbt + b_off)), "r"(w0), "r"(w1) : "memory");
                }
                if (spill && lane == 31) kb_st_u32(eout + j, oh), kb_st_u32(eout + KB_DP_MAXLEN + j, oe1p), kb_st_u32(eout + 2 * KB_DP_MAXLEN + j, oe2p);
            }
            dg = uh;
            tbt += 256;
            if (s == sfA || s == sfB) {  // H(tlen - 1, qlen - 1) of a job has just been computed
                const int cc = s == sfA ? ccA : ccB, ms = cc % kact;
                uint32_t hv = Hc[0];
#pragma unroll
                for (int m = 1; m < 8; ++m)
                    if (m == ms) hv = Hc[m];
                hv = __shfl_sync(0xffffffffu, hv, cc / kact);
                if (s == sfA) out.scoreA = kb_lo16(hv) >> 3;
                if (s == sfB) {  // (both can finish in the same step)
                    const int msb = ccB % kact;
                    uint32_t hb = Hc[0];
#pragma unroll
                    for (int m = 1; m < 8; ++m)
                        if (m == msb) hb = Hc[m];
                    hb = __shfl_sync(0xffffffffu, hb, ccB / kact);
                    out.scoreB = kb_hi16(hb) >> 3;
                }
            }
            if (TRACK && (s & 255) == 255) drain(T0 + s - 287);
        }
        if (TRACK) drain(T0 + nstep - 288), drain(T0 + nstep + 224);
        __syncwarp();  // spilled column visible to lane 0 of the next tile
    }
}

// z-drop rule + traceback of ONE job of the pair (which = 0: A, 1: B) into S.ezcig / ez.
template <bool TRACK, class SQ, class ST>
static __device__ __noinline__ void kb_rows16_finish(const KbDpConst P, int lane, const KbPairJob<SQ, ST> J, int which, int tlen_other,
                                                     const KbPairOut out, KbEz &ez, const KbAlignScratch S, int64_t *cell_counter)
{
    const int rb = (J.flag & KB_EZ_RIGHT) ? 1 : 0;
    const int n_diag = J.qlen + J.tlen - 1;
    const uint8_t *tb = S.tb + (which ? (size_t)(P.max_sw_cells >> 1) : 0);
    const uint32_t *rmax = reinterpret_cast<const uint32_t *>(S.off) + (which ? 2 * KB_DP_MAXLEN : 0);
    const size_t tile_bytes = (size_t)out.nstep * 256;
    const int tlA = which ? tlen_other : J.tlen, tlB = which ? J.tlen : tlen_other;
    KbEz z_;
    z_.max = 0, z_.max_q = z_.max_t = -1, z_.score = KB_NEG_INF, z_.zdropped = 0, z_.n_cigar = 0;
    if (TRACK && !(J.flag & KB_EZ_GLOBAL_NO_ZDROP)) kb_zdrop_scan<true>(P, lane, rmax, n_diag, J.zdrop, z_);
    if (!z_.zdropped) z_.score = which ? out.scoreB : out.scoreA;
    if (cell_counter && lane == 0) *cell_counter += (int64_t)J.qlen * J.tlen;
    int i = -1, j = -1;
    if (!z_.zdropped && !(J.flag & KB_EZ_EXTZ_ONLY)) i = J.tlen - 1, j = J.qlen - 1;
    else if (z_.max_t >= 0 && z_.max_q >= 0) i = z_.max_t, j = z_.max_q;
    z_.n_cigar = kb_backtrack_warp(lane, i, j, rb, J.flag, S.ezcig, [&](int ii, int jj) -> uint32_t {
        const int tile = ii >> 8, cc = ii & 255, kk = kb_pair_kact(tlA, tlB, tile);
        const int l = cc / kk;
        return (uint32_t)kb_ld_u8(tb + (size_t)tile * tile_bytes + (size_t)(jj + l) * 256 + l * 8 + (cc - l * kk));
    });
    ez = z_;
}

#endif  // __CUDACC__
