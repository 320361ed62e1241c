// kb_align_reg.cuh -- register-resident forms of the dual-affine DP (device only).
//
// Same recurrences, tie rules, per-anti-diagonal maximum / z-drop rule and traceback as kb_extd2 in
// kb_align.cuh (the spec is oracle/kb_oracle.c:extd2); no DP value ever touches memory.
//
//   kb_rows         the DP of a whole rectangle for any band half-width of the spec: column tiles, lane = a
//                   stripe of up to 8 columns, one row per step, skewed by one step per lane
//   kb_global_band  certified 64-diagonal band pass for the global gap fills: a sliding window of 32 cells per
//                   anti-diagonal; accepted only when a bound proves no optimal path can leave the band
//
// Cell arithmetic ("x8 domain").  Every score is kept multiplied by 8 and the five candidates of
// H = max(diag, E1, F1, E2, F2) carry their priority in the low three bits, so the spec's ordered, tie-aware
// selection `if (a1 > z) d = 1, z = a1; ...` (or its gap-preferring twin under KB_EZ_RIGHT) collapses into plain
// integer maxima: equal scores are separated by the tag, unequal scores differ by at least 8.  The tags ride
// along for free: E1' = max(H - q - e, E1 - e) becomes max(H8 + (-8(q+e) + tag), E1k - 8e) because subtracting a
// multiple of 8 keeps the tag of E1k.  H8 = max & ~7 is the clean score handed to the neighbours and the low
// three bits of the maximum are the traceback state.  The "gap continues" flags compare the tagged gap states
// with one threshold per gap cost pair: 8 a1 > 8 (z - q)  <=>  a1k > H8 - 8q + 7 for any tag in 0..6 (strict
// rule), 8 a1 >= 8 (z - q)  <=>  a1k > H8 - 8q - 1 (KB_EZ_RIGHT).  The substitution score (with the diagonal's
// tag) is one byte permute: per target column a word holds the four scores against A, C, G, T, the selector is
// derived from the query base once per row, byte 4 (ambiguous query base) comes from a constant.
#pragma once
#ifdef __CUDACC__

#define KB_NEG8 (-0x30000000)

// All DP operands live in global memory; say so, otherwise loads through pointers that crossed a call are generic.
__device__ __forceinline__ int kb_ld_u8(const uint8_t *p)
{
    unsigned v;
    asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(__cvta_generic_to_global(p)));
    return (int)v;
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
__device__ __forceinline__ void kb_st_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.global.u32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t kb_ld_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(__cvta_generic_to_global(p)));
    return v;
}
__device__ __forceinline__ int32_t kb_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    int32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// sequence accessors of the DP kernels: element x of a query / target segment
struct KbPtrSeq {  // plain nt4 bytes, forwards (the monolith's staged buffers)
    const uint8_t *p;
    __device__ __forceinline__ int operator()(int x) const { return kb_ld_u8(p + x); }
};
struct KbDirBytes {  // nt4 bytes, forwards (dir = 1) or backwards (dir = -1) from p
    const uint8_t *p;
    int dir;
    __device__ __forceinline__ int operator()(int x) const { return kb_ld_u8(p + dir * x); }
};
struct KbDirPack {  // 2-bit + mask packed contig bases, forwards or backwards from pos
    const uint32_t *seq2, *nmask;
    int64_t pos;
    int dir;
    __device__ __forceinline__ int operator()(int x) const
    {
        const int64_t b = pos + (int64_t)dir * x;
        if ((kb_ld_u32(nmask + (b >> 5)) >> (b & 31)) & 1u) return 4;
        return (int)((kb_ld_u32(seq2 + (b >> 4)) >> (2 * (b & 15))) & 3u);
    }
};

// sequential reader on top of an accessor: for the packed contigs it keeps the 32-base group it touched last (two sequence
// words + one mask word), so a lane that walks along its diagonal pays three loads per 32 bases instead of two per base
template <class S>
struct KbSeqReader {
    const S s;
    __device__ __forceinline__ explicit KbSeqReader(const S &a) : s(a) {}
    __device__ __forceinline__ int operator()(int x) { return s(x); }
};
template <>
struct KbSeqReader<KbDirPack> {
    const KbDirPack s;
    int64_t grp = -1;
    uint32_t w0 = 0, w1 = 0, mw = 0;
    __device__ __forceinline__ explicit KbSeqReader(const KbDirPack &a) : s(a) {}
    __device__ __forceinline__ int operator()(int x)
    {
        const int64_t b = s.pos + (int64_t)s.dir * x, g = b >> 5;
        if (g != grp) grp = g, w0 = kb_ld_u32(s.seq2 + 2 * g), w1 = kb_ld_u32(s.seq2 + 2 * g + 1), mw = kb_ld_u32(s.nmask + g);
        const int r = (int)(b & 31);
        if ((mw >> r) & 1u) return 4;
        return (int)(((r < 16 ? w0 : w1) >> (2 * (r & 15))) & 3u);
    }
};

// constants of the x8 domain for one DP call
struct KbC8 {
    int32_t oe1, oe2, of1, of2;  // open + first extension, with the tag of the state: -8 (q + e) + tag
    int32_t x1, x2;              // extension: 8 e, 8 e2
    int32_t th1, th2;            // flag thresholds relative to H8: -8 q + (rb ? -1 : 7)
    uint32_t sN;                 // byte 0: score against an ambiguous query base (with the diagonal tag)
    int32_t tag_d;               // tag of the diagonal candidate
    int32_t rb;
};
__device__ __forceinline__ KbC8 kb_c8(const KbDpConst &P, int rb)
{
    KbC8 c;
    const int tE1 = rb ? 4 : 6, tF1 = 5, tE2 = rb ? 6 : 4, tF2 = rb ? 7 : 3;
    c.tag_d = rb ? 3 : 7, c.rb = rb;
    c.oe1 = -8 * (P.q + P.e) + tE1, c.of1 = -8 * (P.q + P.e) + tF1;
    c.oe2 = -8 * (P.q2 + P.e2) + tE2, c.of2 = -8 * (P.q2 + P.e2) + tF2;
    c.x1 = 8 * P.e, c.x2 = 8 * P.e2;
    c.th1 = -8 * P.q + (rb ? -1 : 7), c.th2 = -8 * P.q2 + (rb ? -1 : 7);
    c.sN = (uint32_t)(-8 * P.sc_ambi + c.tag_d) & 0xffu;
    // keep the constants in registers: left alone, ptxas re-derives them from the parameter bank in every cell (two
    // extra LDC per cell); a value that went through a shuffle is opaque to it
    c.x1 = __shfl_sync(0xffffffffu, c.x1, 0), c.x2 = __shfl_sync(0xffffffffu, c.x2, 0);
    return c;
}
// initial (never winning) values of the tagged gap states
__device__ __forceinline__ int32_t kb_neg_e1(const KbC8 &c) { return KB_NEG8 + (c.rb ? 4 : 6); }
__device__ __forceinline__ int32_t kb_neg_e2(const KbC8 &c) { return KB_NEG8 + (c.rb ? 6 : 4); }
__device__ __forceinline__ int32_t kb_neg_f1(const KbC8 &c) { return KB_NEG8 + 5; }
__device__ __forceinline__ int32_t kb_neg_f2(const KbC8 &c) { return KB_NEG8 + (c.rb ? 7 : 3); }
// score word of a target base: byte k = 8 * score(ct, k) + diagonal tag for k = A, C, G, T
__device__ __forceinline__ uint32_t kb_score_row(const KbDpConst &P, const KbC8 &c, int ct)
{
    const uint32_t mm = (uint32_t)(-8 * P.b + c.tag_d) & 0xffu, ma = (uint32_t)(8 * P.a + c.tag_d) & 0xffu;
    if (ct > 3) return c.sN * 0x01010101u;
    return (mm * 0x01010101u) ^ ((mm ^ ma) << (8 * ct));
}
// byte-permute selector of a query base: byte cq sign-extended to 32 bits (cq = 4 selects byte 0 of the second operand)
__device__ __forceinline__ uint32_t kb_score_sel(int cq) { return (uint32_t)(cq > 4 ? 4 : cq) * 0x1111u + 0x8880u; }
// traceback state (0 = diagonal, 1 = E1, 2 = F1, 3 = E2, 4 = F2) of a stored byte
__device__ __forceinline__ int kb_tb_state(uint32_t byte, int rb) { return rb ? (int)(byte & 7) - 3 : 7 - (int)(byte & 7); }

// One cell.  In: H8 of (t-1, j) with its tagged E states, H8 of (t, j-1) with its tagged F states, H8 of (t-1, j-1).
// Out: H8 (returned), the cell's tagged E / F states (in place), the traceback byte.
__device__ __forceinline__ int32_t kb_cell8(const KbC8 &c, int32_t hu, int32_t &e1, int32_t &e2, int32_t hl, int32_t &f1, int32_t &f2,
                                            int32_t hd, uint32_t srow, uint32_t sel, uint32_t &d)
{
    // the chain that links a row's cells runs hu -> (e1, e2) -> zk -> H8: everything else is computed off it
    f1 = max(hl + c.of1, f1 - c.x1);
    f2 = max(hl + c.of2, f2 - c.x2);
    const int32_t s = kb_prmt(srow, c.sN, sel);
    const int32_t pre = max(max(hd + s, f1), f2);
    const int32_t x1 = e1 - c.x1, x2 = e2 - c.x2;
    e1 = max(hu + c.oe1, x1);
    e2 = max(hu + c.oe2, x2);
    const int32_t zk = max(max(pre, e1), e2);
    const int32_t z = zk & ~7;
    const int32_t t1 = z + c.th1, t2 = z + c.th2;
    d = (uint32_t)(zk & 7);
    asm("{\n\t.reg .pred p1, p2, p3, p4;\n\t"
        "setp.gt.s32 p1, %1, %5;\n\tsetp.gt.s32 p2, %2, %5;\n\tsetp.gt.s32 p3, %3, %6;\n\tsetp.gt.s32 p4, %4, %6;\n\t"
        "@p1 add.u32 %0, %0, 8;\n\t@p2 add.u32 %0, %0, 16;\n\t@p3 add.u32 %0, %0, 32;\n\t@p4 add.u32 %0, %0, 64;\n\t}"
        : "+r"(d)
        : "r"(e1), "r"(f1), "r"(e2), "r"(f2), "r"(t1), "r"(t2));
    return z;
}

// ksw_backtrack shared by both kernels: `byte_at(i, j)` returns the traceback byte of a cell.  Runs of diagonal
// moves are resolved 32 at a time: every lane fetches the byte of (i - k, j - k) and a ballot finds how far the
// H-state run goes, so the dependent-load chain is paid once per run, not once per base.
template <class ByteAt>
__device__ __forceinline__ int kb_backtrack_warp(int lane, int i, int j, int rb, int flag, uint32_t *cg, ByteAt byte_at)
{
    int n_cigar = 0, state = 0;
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
        uint32_t tmp = 0xffu;
        int st = 1;
        if (in) tmp = byte_at(ik, jk), st = kb_tb_state(tmp, rb);
        const uint32_t t0 = __shfl_sync(0xffffffffu, tmp, 0);
        const int s0 = __shfl_sync(0xffffffffu, st, 0);
        if (state == 0) state = s0;
        else if (!(t0 >> (state + 2) & 1)) state = 0;
        if (state == 0) state = s0;
        const unsigned run = __ballot_sync(0xffffffffu, in && st == 0);
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
    return n_cigar;
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
// any path leaving the band [dlo, dhi] scores at most this (see "Exactness" above)
__device__ __forceinline__ int kb_band_bound64(const KbDpConst &P, int qlen, int tlen, int dlo, int dhi)
{
    const int d1 = tlen - qlen;
    const int D_hi = dhi + 1, I_hi = D_hi - d1, I_lo = 1 - dlo, D_lo = I_lo + d1;
    const int b_hi = (tlen - D_hi < 0 || qlen - I_hi < 0) ? KB_NEG_INF : P.a * (tlen - D_hi) - kb_gapcost2(P, D_hi) - kb_gapcost2(P, I_hi);
    const int b_lo = (tlen - D_lo < 0 || qlen - I_lo < 0) ? KB_NEG_INF : P.a * (tlen - D_lo) - kb_gapcost2(P, D_lo) - kb_gapcost2(P, I_lo);
    return b_hi > b_lo ? b_hi : b_lo;
}
template <bool EDGE>
__device__ __forceinline__ void kb_band_step(const KbDpConst &P, const KbC8 &c, int lane, bool stepB, int tp, int jp, uint32_t srow,
                                             uint32_t sel, int32_t &H1, int32_t &H2, int32_t &E1, int32_t &E2, int32_t &F1, int32_t &F2,
                                             uint32_t &acc)
{
    int32_t uH, lH;
    if (!stepB) {
        uH = __shfl_up_sync(0xffffffffu, H1, 1), E1 = __shfl_up_sync(0xffffffffu, E1, 1), E2 = __shfl_up_sync(0xffffffffu, E2, 1);
        if (lane == 0) uH = E1 = E2 = KB_NEG8;
        lH = H1;
    } else {
        lH = __shfl_down_sync(0xffffffffu, H1, 1), F1 = __shfl_down_sync(0xffffffffu, F1, 1), F2 = __shfl_down_sync(0xffffffffu, F2, 1);
        if (lane == 31) lH = F1 = F2 = KB_NEG8;
        uH = H1;
    }
    uint32_t d;
    int32_t z = kb_cell8(c, uH, E1, E2, lH, F1, F2, H2, srow, sel, d);
    if (EDGE) {
        if (tp <= 0 || jp <= 0) {  // virtual row / column (and cells before them, which no valid cell ever reads)
            const int m = tp > jp ? tp : jp;
            z = (tp < 0 || jp < 0) ? KB_NEG8 : (m == 0 ? 0 : -8 * kb_gapcost2(P, m));
            E1 = E2 = F1 = F2 = KB_NEG8, d = 0;
        }
    }
    H2 = H1, H1 = z;
    acc = (acc >> 8) | (d << 24);
}

template <class SQ, class ST>
static __device__ __noinline__ int kb_global_band(const KbDpConst P, int lane, int qlen, const SQ qs, int tlen, const ST ts, int flag, KbEz &ez,
                                                  const KbAlignScratch S, int64_t *cell_counter)
{
    const int d1 = tlen - qlen, lo_d = d1 < 0 ? d1 : 0, hi_d = d1 > 0 ? d1 : 0;
    const int margin = (63 - (hi_d - lo_d)) >> 1;
    const int r_end = tlen + qlen;  // last anti-diagonal in shifted coordinates: the cell (tlen, qlen)
    ez.score = KB_NEG_INF;
    if (margin < KB_BAND_MIN_MARGIN || (int64_t)32 * (r_end + 8) > P.max_sw_cells) return 0;
    const KbC8 c = kb_c8(P, 0);
    const int dlo = lo_d - margin, dhi = dlo + 63;
    KbSeqReader<ST> tsr(ts);
    uint32_t *tbw = reinterpret_cast<uint32_t *>(S.tb);
    int32_t H1 = KB_NEG8, H2 = KB_NEG8, E1 = KB_NEG8, E2 = KB_NEG8, F1 = KB_NEG8, F2 = KB_NEG8;
    uint32_t acc = 0;
    int tp = ((dlo + 1) >> 1) + lane, jp = -tp;  // anti-diagonal 0
    uint32_t srow = 0, sel = 0;
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
        srow = kb_score_row(P, c, tsr((tp < 1 ? 0 : (tp > tlen ? tlen : tp) - 1)));
        sel = kb_score_sel(qs((jp < 1 ? 0 : (jp > qlen ? qlen : jp) - 1)));
        kb_band_step<true>(P, c, lane, stepB, tp, jp, srow, sel, H1, H2, E1, E2, F1, F2, acc);
        if ((rp & 3) == 3) kb_st_u32(tbw + (rp >> 2) * 32 + lane, acc);
    }
    // interior: every in-range cell has real neighbours; cells past the far edges compute garbage nobody reads
    const int cert_bound = kb_band_bound64(P, qlen, tlen, dlo, dhi);
    for (; rp + 1 <= r_end; rp += 2) {
        ++jp;
        sel = kb_score_sel(qs((jp > qlen ? qlen : jp) - 1));
        kb_band_step<false>(P, c, lane, false, tp, jp, srow, sel, H1, H2, E1, E2, F1, F2, acc);
        if ((rp & 3) == 3) kb_st_u32(tbw + (rp >> 2) * 32 + lane, acc);
        ++tp;
        srow = kb_score_row(P, c, tsr((tp > tlen ? tlen : tp) - 1));
        kb_band_step<false>(P, c, lane, true, tp, jp, srow, sel, H1, H2, E1, E2, F1, F2, acc);
        if (((rp + 1) & 3) == 3) kb_st_u32(tbw + ((rp + 1) >> 2) * 32 + lane, acc);
        if (((rp >> 1) & 15) == 15) {
            // give up early once the certificate is out of reach: a path ends on anti-diagonal rp or rp + 1 and gains at most
            // +a per remaining column.  The score handed back is then an extrapolation (same loss per column for the rest),
            // good enough to choose the wider window, whose own pass is certified on its real score.
            const int m = __reduce_max_sync(0xffffffffu, H1 > H2 ? H1 : H2) >> 3;
            const int rem = (r_end - rp + 1) >> 1;
            if (m + P.a * rem <= cert_bound) {
                const int done = (rp + 2) >> 1, loss = P.a * done - m;
                ez.score = P.a * (done + rem) - (int)((int64_t)loss * (done + rem) / (done > 0 ? done : 1));
                if (cell_counter && lane == 0) *cell_counter += (int64_t)32 * (rp + 2);
                return 0;
            }
        }
    }
    if (rp == r_end) {
        ++jp;
        sel = kb_score_sel(qs((jp > qlen ? qlen : jp) - 1));
        kb_band_step<false>(P, c, lane, false, tp, jp, srow, sel, H1, H2, E1, E2, F1, F2, acc);
        if ((rp & 3) == 3) kb_st_u32(tbw + (rp >> 2) * 32 + lane, acc);
    }
    if ((r_end & 3) != 3) kb_st_u32(tbw + (r_end >> 2) * 32 + lane, acc >> (8 * (3 - (r_end & 3))));
    if (cell_counter && lane == 0) *cell_counter += (int64_t)32 * (r_end + 1);
    const int score = __shfl_sync(0xffffffffu, H1, (tlen - ((r_end + dlo + 1) >> 1)) & 31) >> 3;
    if (!(score > cert_bound)) {  // certificate
        ez.score = score;  // a valid alignment's score all the same: a lower bound for any wider band
        return 0;
    }
    __syncwarp();
    int bad = 0;
    const int n_cigar = kb_backtrack_warp(lane, tlen - 1, qlen - 1, 0, flag, S.ezcig, [&](int i, int j) -> uint32_t {
        const int r2 = i + j + 2, l = i + 1 - ((r2 + dlo + 1) >> 1);  // a diagonal keeps its lane index
        if ((unsigned)l > 31u) {
            bad = 1;  // cannot happen once certified
            return 0xffu;
        }
        return (kb_ld_u32(tbw + (r2 >> 2) * 32 + l) >> ((r2 & 3) * 8)) & 0xffu;
    });
    if (__any_sync(0xffffffffu, bad)) return 0;
    ez.max = 0, ez.max_q = ez.max_t = -1, ez.zdropped = 0;
    ez.score = score, ez.n_cigar = n_cigar;
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// Wider certified bands: the same sliding window with K cells per lane (64 K diagonals, K = 2 or 4), for gap fills
// whose 64-diagonal pass scored too low to certify itself.  Lane l holds the K consecutive cells l K .. l K + K - 1
// of the window, so on an "A" step only slot 0 needs a shuffle (cell c - 1 lives in the lane's own slot m - 1) and on
// a "B" step only slot K - 1 does.  The target / query bases ride through the window like a systolic array: a B step
// moves every target base one cell down (the top cell of lane 31 loads the new one), an A step moves every query
// base one cell up.  kb_band_bound gives the certificate threshold for a window of K slots, so the caller can pick
// the smallest K the first pass's score (a lower bound of every wider pass's score) already certifies.
__device__ __forceinline__ bool kb_band_geometry(int qlen, int tlen, int K, int &dlo, int &dhi)
{
    const int d1 = tlen - qlen, lo_d = d1 < 0 ? d1 : 0, hi_d = d1 > 0 ? d1 : 0;
    const int margin = (64 * K - 1 - (hi_d - lo_d)) >> 1;
    dlo = lo_d - margin, dhi = dlo + 64 * K - 1;
    return margin >= KB_BAND_MIN_MARGIN;
}
__device__ __forceinline__ int kb_band_bound(const KbDpConst &P, int qlen, int tlen, int dlo, int dhi)
{
    return kb_band_bound64(P, qlen, tlen, dlo, dhi);
}

template <int K, class SQ, class ST>
static __device__ __noinline__ int kb_global_bandK(const KbDpConst P, int lane, int qlen, const SQ qs, int tlen, const ST ts, int flag, KbEz &ez,
                                                   const KbAlignScratch S, int64_t *cell_counter)
{
    int dlo, dhi;
    const int r_end = tlen + qlen;
    ez.score = KB_NEG_INF;
    if (!kb_band_geometry(qlen, tlen, K, dlo, dhi) || (int64_t)32 * K * (r_end + 8) > P.max_sw_cells) return 0;
    const KbC8 c = kb_c8(P, 0);
    uint8_t *tb = S.tb;  // [r'][lane][K]
    int32_t H1[K], H2[K], E1[K], E2[K], F1[K], F2[K];
    uint32_t srow[K], sel[K];
#pragma unroll
    for (int m = 0; m < K; ++m) H1[m] = H2[m] = E1[m] = E2[m] = F1[m] = F2[m] = KB_NEG8, srow[m] = sel[m] = 0;
    const int T0 = (dlo + 1) >> 1;  // T(0)
    int tp0 = T0 + lane * K, jp0 = -tp0;  // t', j' of the lane's slot 0 on anti-diagonal 0 (slot m: tp0 + m, jp0 - m)
    KbSeqReader<ST> tsr(ts);
    auto tbase = [&](int tp) { return kb_score_row(P, c, tsr((tp < 1 ? 0 : (tp > tlen ? tlen : tp) - 1))); };
    auto qbase = [&](int jp) { return kb_score_sel(qs((jp < 1 ? 0 : (jp > qlen ? qlen : jp) - 1))); };
#pragma unroll
    for (int m = 0; m < K; ++m) srow[m] = tbase(tp0 + m), sel[m] = qbase(jp0 - m);
    const int r_edge = (-dlo > dhi ? -dlo : dhi) + 1;  // boundary cells can occur up to here
    for (int rp = 0; rp <= r_end; ++rp) {
        const bool stepB = (rp + dlo) & 1;
        const bool edge = rp <= r_edge;
        uint32_t tbw = 0;
        if (rp > 0) {
            if (stepB) {  // every cell moves one target base on; slot K-1 takes the next lane's slot 0
                ++tp0;
                uint32_t nx = __shfl_down_sync(0xffffffffu, srow[0], 1);
                if (lane == 31) nx = tbase(tp0 + K - 1);
#pragma unroll
                for (int m = 0; m + 1 < K; ++m) srow[m] = srow[m + 1];
                srow[K - 1] = nx;
            } else {  // every cell moves one query base on; slot 0 takes the previous lane's slot K-1
                ++jp0;
                uint32_t nx = __shfl_up_sync(0xffffffffu, sel[K - 1], 1);
                if (lane == 0) nx = qbase(jp0);
#pragma unroll
                for (int m = K - 1; m > 0; --m) sel[m] = sel[m - 1];
                sel[0] = nx;
            }
        }
        if (!stepB) {  // A: (t-1, j) is cell c - 1; descending so that slot m reads slot m - 1 before it is overwritten
            int32_t uH = __shfl_up_sync(0xffffffffu, H1[K - 1], 1), uE1 = __shfl_up_sync(0xffffffffu, E1[K - 1], 1),
                    uE2 = __shfl_up_sync(0xffffffffu, E2[K - 1], 1);
            if (lane == 0) uH = uE1 = uE2 = KB_NEG8;
#pragma unroll
            for (int m = K - 1; m >= 0; --m) {
                const int32_t hu = m > 0 ? H1[m > 0 ? m - 1 : 0] : uH;
                int32_t e1 = m > 0 ? E1[m > 0 ? m - 1 : 0] : uE1, e2 = m > 0 ? E2[m > 0 ? m - 1 : 0] : uE2;
                uint32_t d;
                int32_t z = kb_cell8(c, hu, e1, e2, H1[m], F1[m], F2[m], H2[m], srow[m], sel[m], d);
                if (edge) {
                    const int tp = tp0 + m, jp = jp0 - m;
                    if (tp <= 0 || jp <= 0) {
                        const int mm = tp > jp ? tp : jp;
                        z = (tp < 0 || jp < 0) ? KB_NEG8 : (mm == 0 ? 0 : -8 * kb_gapcost2(P, mm));
                        e1 = e2 = F1[m] = F2[m] = KB_NEG8, d = 0;
                    }
                }
                H2[m] = H1[m], H1[m] = z, E1[m] = e1, E2[m] = e2;
                tbw |= d << (8 * m);
            }
        } else {  // B: (t, j-1) is cell c + 1; ascending
            int32_t lH = __shfl_down_sync(0xffffffffu, H1[0], 1), lF1 = __shfl_down_sync(0xffffffffu, F1[0], 1),
                    lF2 = __shfl_down_sync(0xffffffffu, F2[0], 1);
            if (lane == 31) lH = lF1 = lF2 = KB_NEG8;
#pragma unroll
            for (int m = 0; m < K; ++m) {
                const int32_t hl = m + 1 < K ? H1[m + 1 < K ? m + 1 : 0] : lH;
                int32_t f1 = m + 1 < K ? F1[m + 1 < K ? m + 1 : 0] : lF1, f2 = m + 1 < K ? F2[m + 1 < K ? m + 1 : 0] : lF2;
                uint32_t d;
                int32_t z = kb_cell8(c, H1[m], E1[m], E2[m], hl, f1, f2, H2[m], srow[m], sel[m], d);
                if (edge) {
                    const int tp = tp0 + m, jp = jp0 - m;
                    if (tp <= 0 || jp <= 0) {
                        const int mm = tp > jp ? tp : jp;
                        z = (tp < 0 || jp < 0) ? KB_NEG8 : (mm == 0 ? 0 : -8 * kb_gapcost2(P, mm));
                        E1[m] = E2[m] = f1 = f2 = KB_NEG8, d = 0;
                    }
                }
                H2[m] = H1[m], H1[m] = z, F1[m] = f1, F2[m] = f2;
                tbw |= d << (8 * m);
            }
        }
        uint8_t *dst = tb + ((size_t)rp * 32 + lane) * K;
        if (K == 4) kb_st_u32(reinterpret_cast<uint32_t *>(dst), tbw);
        else asm volatile("st.global.u16 [%0], %1;" ::"l"(__cvta_generic_to_global(dst)), "h"((unsigned short)tbw) : "memory");
    }
    if (cell_counter && lane == 0) *cell_counter += (int64_t)32 * K * (r_end + 1);
    const int cend = tlen - ((r_end + dlo + 1) >> 1);  // window index of the cell (tlen, qlen)
    int32_t hv = H1[0];
#pragma unroll
    for (int m = 1; m < K; ++m)
        if (m == cend % K) hv = H1[m];
    const int score = __shfl_sync(0xffffffffu, hv, (cend / K) & 31) >> 3;
    if (!(score > kb_band_bound(P, qlen, tlen, dlo, dhi))) {
        ez.score = score;  // still a valid alignment's score: a lower bound for any wider band
        return 0;
    }
    __syncwarp();
    int bad = 0;
    const int n_cigar = kb_backtrack_warp(lane, tlen - 1, qlen - 1, 0, flag, S.ezcig, [&](int i, int j) -> uint32_t {
        const int r2 = i + j + 2, cc = i + 1 - ((r2 + dlo + 1) >> 1);
        if ((unsigned)cc >= (unsigned)(32 * K)) {
            bad = 1;  // cannot happen once certified
            return 0xffu;
        }
        return (uint32_t)kb_ld_u8(tb + (size_t)r2 * 32 * K + cc);
    });
    if (__any_sync(0xffffffffu, bad)) return 0;
    ez.max = 0, ez.max_q = ez.max_t = -1, ez.zdropped = 0;
    ez.score = score, ez.n_cigar = n_cigar;
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// Row-stripe wavefront: the DP of a whole rectangle (any band half-width w of the spec, |d| = |t - j| <= w).
//
// The target is cut into tiles of 256 columns (the last one 32 * K' columns, K' = 1..8, just wide enough); inside a
// tile lane l owns the K consecutive columns t0 .. t0 + K - 1 (t0 = T0 + l * K) and at step s computes row
// j = s - l of them, left to right.  Along the row the (t-1, j) neighbour and its two E states are the values the
// lane has just produced (for its first column: what lane l-1 produced one step earlier, by shuffle), (t, j-1) and
// its F states are the column's own registers, (t-1, j-1) is the left column's previous H.  Per lane 3 * K state
// registers, per step 3 shuffles for 32 * K cells, and only 31 ramp steps per tile.  The last column of a full tile
// (H, E1, E2 per row) is spilled for lane 0 of the next tile.  Traceback bytes go to tb[tile][s][lane][8]: 256
// contiguous bytes per step.
//
// Extension / z-drop mode (TRACK): ksw2 evaluates, per anti-diagonal r, the maximum H and the lowest t attaining
// it.  Every cell folds key = (H + 2^19) << 12 | (4095 - t) into a per-warp shared-memory ring (red.shared.max,
// conflict free: lanes sit K - 1 anti-diagonals apart), the ring is drained into rmax[r] every 256 steps, and the
// sequential z-drop rule runs over rmax[] after the last tile.  Cells beyond a z-drop are computed in vain but
// never influence the result (a traceback only moves towards smaller r).
template <bool MASK, bool TRACK, int M>
__device__ __forceinline__ void kb_rows_cell(const KbC8 &c, int w, int d0, int nvm, unsigned ring, int rbase, int32_t ckey, uint32_t sel,
                                             int32_t &hu, int32_t &e1, int32_t &e2, int32_t &hd, int32_t (&Hc)[8], int32_t (&F1)[8],
                                             int32_t (&F2)[8], const uint32_t (&srow)[8], uint32_t (&tbw)[2])
{
    uint32_t d;
    const int32_t hl = Hc[M];
    int32_t z = kb_cell8(c, hu, e1, e2, hl, F1[M], F2[M], hd, srow[M], sel, d);
    bool ok = true;
    if (MASK) {
        ok = (unsigned)(d0 + M + w) <= (unsigned)(2 * w);
        if (!ok) z = e1 = e2 = F1[M] = F2[M] = KB_NEG8, d = 0;
    }
    hd = hl, Hc[M] = z, hu = z;
    tbw[M >> 2] += d << (8 * (M & 3));
    if (TRACK) {
        // `ring` is the byte address of the lane's slot-0 anti-diagonal in the ring (< 512 words in); slot M is M words on,
        // possibly in the 8 alias words past the end, which drain() folds back
        if (MASK) {
            if (M < nvm && ok) {
                const uint32_t key = (uint32_t)(z * 512 + (ckey - M));
                asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(ring + 4 * M), "r"(key) : "memory");
            }
        } else
            asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 k;\n\t"
                         "setp.lt.s32 p, %3, %4;\n\tmad.lo.s32 k, %0, 512, %1;\n\t"
                         "@p red.shared.max.u32 [%2+%5], k;\n\t}" ::"r"(z), "r"(ckey - M), "r"(ring), "n"(M), "r"(nvm), "n"(4 * M)
                         : "memory");
    }
}
// One row of a lane's stripe.  A stripe of kact < 8 columns lives in the LAST kact slots, so the row is one jump
// into the unrolled sequence (slot 8 - kact) and no per-cell test.
template <bool MASK, bool TRACK>
__device__ __forceinline__ void kb_rows_body(const KbC8 &c, int kact, int w, int d0, int nvm, unsigned ring, int rbase, int32_t ckey,
                                             uint32_t sel, int32_t &hu, int32_t &e1, int32_t &e2, int32_t &hd, int32_t (&Hc)[8],
                                             int32_t (&F1)[8], int32_t (&F2)[8], const uint32_t (&srow)[8], uint32_t (&tbw)[2])
{
    tbw[0] = tbw[1] = 0;
#define KB_RC(M) kb_rows_cell<MASK, TRACK, M>(c, w, d0, nvm, ring, rbase, ckey, sel, hu, e1, e2, hd, Hc, F1, F2, srow, tbw)
    switch (kact) {
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

#define KB_ROWS_KEY_BIAS (int32_t)(0x80000000u + 4095u)
template <bool TRACK, class SQ, class ST>
static __device__ __noinline__ void kb_rows(const KbDpConst P, int lane, int qlen, const SQ qs, int tlen, const ST ts, int w, int zdrop,
                                            int flag, KbEz &ez, const KbAlignScratch S, int64_t *cell_counter)
{
    const int rb = (flag & KB_EZ_RIGHT) ? 1 : 0;
    const KbC8 c = kb_c8(P, rb);
    const int ntile = (tlen + 255) >> 8, nstep = qlen + 31, n_diag = qlen + tlen - 1;
    const int klast = (tlen - ((ntile - 1) << 8) + 31) >> 5;  // columns per lane in the last tile
    const size_t tile_bytes = (size_t)nstep * 256;
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
        if (lane < 8) {  // alias words first: word 512 + x stands for word x
            const uint32_t a = S.wmax[512 + lane];
            if (a) {
                S.wmax[512 + lane] = 0;
                if (a > S.wmax[lane]) S.wmax[lane] = a;
            }
        }
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
        const bool spill = tile + 1 < ntile;  // then the tile is full width and its last column is lane 31's last
        const int kact = spill ? 8 : klast, koff = 8 - kact;
        const int T0 = tile << 8, t0 = T0 + lane * kact, t0s = t0 - koff;  // slot m holds column t0s + m (m >= koff)
        const int32_t *ein = edge + (size_t)((tile & 1) ^ 1) * 3 * KB_DP_MAXLEN;
        int32_t *eout = edge + (size_t)(tile & 1) * 3 * KB_DP_MAXLEN;
        int32_t Hc[8], F1[8], F2[8];
        uint32_t srow[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int t = t0s + m;
            Hc[m] = -8 * kb_gapcost2(P, t + 1), F1[m] = kb_neg_f1(c), F2[m] = kb_neg_f2(c);  // virtual row j = -1
            srow[m] = kb_score_row(P, c, (m >= koff && t < tlen) ? ts(t) : 4);
        }
        int nval = tlen - t0;
        nval = nval < 0 ? 0 : (nval > kact ? kact : nval);
        const int nvm = nval > 0 ? koff + nval : 0;  // slots below nvm hold real columns
        // what the lane offers to lane + 1: (H, E1, E2) of its last column in the row it has just finished
        int32_t oh = -8 * kb_gapcost2(P, t0 + kact), oe1 = kb_neg_e1(c), oe2 = kb_neg_e2(c);
        int32_t dg = T0 == 0 ? 0 : -8 * kb_gapcost2(P, T0);  // lane 0: H(T0 - 1, -1); other lanes: set by the first shuffle
        uint8_t *tbt = tb + (size_t)tile * tile_bytes + lane * 8;
        int cq_next = qs((lane == 0 ? 0 : qlen - 1));  // row 0 is lane 0's first; the others reload before use
        for (int s = 0; s < nstep; ++s) {
            const int cq = cq_next;
            {
                int jn = s + 1 - lane;
                jn = jn < 0 ? 0 : (jn >= qlen ? qlen - 1 : jn);
                cq_next = qs(jn);
            }
            int32_t uh = __shfl_up_sync(0xffffffffu, oh, 1), ue1 = __shfl_up_sync(0xffffffffu, oe1, 1), ue2 = __shfl_up_sync(0xffffffffu, oe2, 1);
            const int j = s - lane;
            const bool act = (unsigned)j < (unsigned)qlen;
            if (lane == 0 && act) {
                if (T0 == 0) uh = -8 * kb_gapcost2(P, j + 1), ue1 = kb_neg_e1(c), ue2 = kb_neg_e2(c);
                else uh = kb_ld_s32(ein + j), ue1 = kb_ld_s32(ein + KB_DP_MAXLEN + j), ue2 = kb_ld_s32(ein + 2 * KB_DP_MAXLEN + j);
            }
            const int d0 = t0s - j;  // diagonal of slot 0
            const bool edge_lane = banded && act && (d0 + koff < -w || d0 + 7 > w);
            const bool any_edge = banded && __any_sync(0xffffffffu, edge_lane);
            if (act) {
                const uint32_t sel = kb_score_sel(cq);
                int32_t hu = uh, e1 = ue1, e2 = ue2, hd = dg;
                uint32_t tbw[2];
                const unsigned rslot = ring + (((unsigned)(t0s + j) & 511u) << 2);
                if (any_edge) kb_rows_body<true, TRACK>(c, kact, w, d0, nvm, rslot, t0s + j, KB_ROWS_KEY_BIAS - t0s, sel, hu, e1, e2, hd, Hc, F1, F2, srow, tbw);
                else kb_rows_body<false, TRACK>(c, kact, w, d0, nvm, rslot, t0s + j, KB_ROWS_KEY_BIAS - t0s, sel, hu, e1, e2, hd, Hc, F1, F2, srow, tbw);
                oh = hu, oe1 = e1, oe2 = e2;
                uint32_t *dst = reinterpret_cast<uint32_t *>(tbt);  // a lane's slot is 8 bytes wide whatever kact is
                kb_st_u32(dst + 1, tbw[1]);
                if (kact > 4) kb_st_u32(dst, tbw[0]);
                if (spill && lane == 31) kb_st_s32(eout + j, oh), kb_st_s32(eout + KB_DP_MAXLEN + j, oe1), kb_st_s32(eout + 2 * KB_DP_MAXLEN + j, oe2);
            }
            dg = uh;
            tbt += 256;
            if (TRACK && (s & 255) == 255) drain(T0 + s - 287);
        }
        if (TRACK) drain(T0 + nstep - 288), drain(T0 + nstep + 224);
        if (!spill) {  // H(tlen - 1, qlen - 1): the last row of column tlen - 1
            const int cc = tlen - 1 - T0, ms = koff + cc % kact;
            int32_t hv = Hc[0];
#pragma unroll
            for (int m = 1; m < 8; ++m)
                if (m == ms) hv = Hc[m];
            score = __shfl_sync(0xffffffffu, hv, cc / kact) >> 3;
        }
        __syncwarp();  // spilled column visible to lane 0 of the next tile
    }
    KbEz z_;
    z_.max = 0, z_.max_q = z_.max_t = -1, z_.score = KB_NEG_INF, z_.zdropped = 0, z_.n_cigar = 0;
    if (TRACK) {
        // [mm2:ksw2.h:ksw_apply_zdrop] over the per-anti-diagonal maxima, in order.  The running state (max so far, first cell
        // attaining it) before anti-diagonal r is a prefix maximum, so every lane takes a contiguous chunk of r: pass 1 finds
        // the chunk's first maximum, a warp scan turns that into the state at the chunk's start, pass 2 replays the chunk with
        // the exact rule and reports the first anti-diagonal where the spec stops (z-drop, or an anti-diagonal the band leaves
        // empty); the earliest one over the warp wins.
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
    if (!z_.zdropped) z_.score = score;
    if (cell_counter && lane == 0) *cell_counter += (int64_t)qlen * tlen;
    int i = -1, j = -1;
    if (!z_.zdropped && !(flag & KB_EZ_EXTZ_ONLY)) i = tlen - 1, j = qlen - 1;
    else if (z_.max_t >= 0 && z_.max_q >= 0) i = z_.max_t, j = z_.max_q;
    z_.n_cigar = kb_backtrack_warp(lane, i, j, rb, flag, S.ezcig, [&](int ii, int jj) -> uint32_t {
        const int tile = ii >> 8, cc = ii & 255, kk = tile + 1 < ntile ? 8 : klast;
        const int l = cc / kk;
        return (uint32_t)kb_ld_u8(tb + (size_t)tile * tile_bytes + (size_t)(jj + l) * 256 + l * 8 + (8 - kk) + (cc - l * kk));
    });
    ez = z_;
}

#endif  // __CUDACC__
