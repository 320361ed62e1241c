// kb_band16.cuh -- the certified band pass of kb_align_reg.cuh with TWO gap fills per warp (device only).
//
// Same sliding window as kb_global_band / kb_global_bandK (64 K diagonals, lane l holds the K window cells l K .. l K + K - 1,
// one anti-diagonal per step, "A" steps move every cell one query base on and "B" steps one target base), same certificate
// (kb_band_bound64), same traceback bytes -- but every DP value is an x8-domain score in 16 bits and the low halfword of every
// register belongs to job A, the high halfword to job B (kb_cell16, the packed DPX cell of kb_align_reg16.cuh).  What makes two
// arbitrary fills fit one instruction stream:
//   * dlo is always odd (kb_band16_geometry gives one diagonal of the margin to the low side when needed): step 1 is an A step for
//     every job, so both jobs of a pair do the same kind of step at the same time whatever their shapes;
//   * no boundary code: anti-diagonal 0 holds H = 0 in the corner cell and a low constant everywhere else, and the ordinary
//     recurrences then produce the virtual row and column themselves (H(t, 0) = max(E1, E2) = -gapcost2(t): what the 32-bit pass
//     imposes); cells before the virtual row / column hold values no real cell can pick up;
//   * both jobs' target bases for window position x live in ONE staged selector (shared memory, built once per pair), both jobs'
//     query bases for a row in one byte that indexes a 25-entry table of score-word pairs: a step fetches one of them, and the
//     bases ride through the window by shuffle like a systolic array.
// Range: a job stays in the pass only while its best cell keeps the certificate within reach (checked every 32 steps; the 32-bit
// pass does the same for K = 1), and every in-band cell is within one gap of the band's width of that best cell, so live values
// stay far inside 16 bits (kb_band16_eligible); a job that has ended or given up keeps computing in its halfword, possibly
// wrapping, which nothing reads -- packed DPX arithmetic never carries between the halves.  An ambiguous TARGET base cannot be
// expressed in the selector: such pairs (found while staging) are handed to the 32-bit kernel.
#pragma once
#ifdef __CUDACC__

#define KB_B16_REND_MAX 3000                       // qlen + tlen of an eligible fill
#define KB_B16_X (KB_B16_REND_MAX / 2 + 128 + 16)  // staged window positions per pair (K <= 4)
#define KB_B16_NEG_INIT (-20000)                   // cells before the virtual row / column on anti-diagonal 0
#define KB_B16_SMEM_BYTES (KB_B16_X * 3 + 4)       // per warp: u16 selector + u8 query pair per position (multiple of 8)

// band of 64 K diagonals around the two corners, dlo odd
KB_HD bool kb_band16_geometry(int qlen, int tlen, int K, int &dlo, int &dhi)
{
    const int d1 = tlen - qlen, lo_d = d1 < 0 ? d1 : 0, hi_d = d1 > 0 ? d1 : 0;
    const int margin = (64 * K - 1 - (hi_d - lo_d)) >> 1;
    dlo = lo_d - margin;
    if (!(dlo & 1)) --dlo;
    dhi = dlo + 64 * K - 1;
    return lo_d - dlo >= KB_BAND_MIN_MARGIN && dhi - hi_d >= KB_BAND_MIN_MARGIN;
}
KB_HD bool kb_band16_eligible(const KbDpConst &P, int qlen, int tlen)
{
    if (qlen + tlen > KB_B16_REND_MAX) return false;
    int dlo, dhi;
    if (!kb_band16_geometry(qlen, tlen, 1, dlo, dhi)) return false;
    // a live job's best cell is within the certificate's slack of a * (columns done); any other in-band cell within a gap of the window's
    // width (and as many columns) of it
    const int slack = 256 * P.a + 2 * kb_gapcost2(P, 129) + 128 * P.a + kb_gapcost2(P, 512) + 64 * (P.a + P.b) + 256;
    const int hi = P.a * ((qlen < tlen ? qlen : tlen) + 136);
    return slack < 3500 && hi < 4000 && P.a <= 15 && P.b <= 15 && P.sc_ambi <= 15;
}

struct KbB16Job {
    int qlen, tlen, dlo, T0, r_end, bound;  // T0: t' of window cell 0 on anti-diagonal 0; bound: the certificate's threshold
};
struct KbB16Out {
    int score[2];  // H(tlen, qlen) of the band DP (state 1) or an extrapolation (state 0: gave up)
    int state[2];
    int steps[2];  // anti-diagonals the job took part in
};

__device__ __forceinline__ KbB16Job kb_b16_job(const KbDpConst &P, int K, int qlen, int tlen)
{
    KbB16Job J{qlen, tlen, 0, 0, qlen + tlen, 0};
    int dhi;
    kb_band16_geometry(qlen, tlen, K, J.dlo, dhi);
    J.T0 = (J.dlo + 1) >> 1;
    J.bound = kb_band_bound64(P, qlen, tlen, J.dlo, dhi);
    return J;
}

// Stage the pair's bases: sel[x] = selector of (A's target base at window position x, B's), qc[y + 32 K] = cqA + 5 cqB for row offset y.
// Returns true if a target base of either job is ambiguous (the pair cannot run packed).
template <class SQ, class ST>
__device__ __forceinline__ bool kb_band16_stage(int lane, int K, const KbB16Job &A, const SQ &qsA, const ST &tsA, const KbB16Job &B, const SQ &qsB,
                                                const ST &tsB, uint16_t *sel, uint8_t *qc)
{
    const int rmax = A.r_end > B.r_end ? A.r_end : B.r_end, nx = 32 * K + (rmax >> 1) + 6;
    bool amb = false;
    for (int x = lane; x < nx; x += 32) {
        int ta = A.T0 + x, tb = B.T0 + x;
        ta = (ta < 1 ? 1 : (ta > A.tlen ? A.tlen : ta)) - 1, tb = (tb < 1 ? 1 : (tb > B.tlen ? B.tlen : tb)) - 1;
        const int ca = tsA(ta), cb = tsB(tb);
        amb |= ca > 3 || cb > 3;
        sel[x] = (uint16_t)kb_sel16(ca, cb);
        const int y = x - 32 * K;  // row offset: j' = y - T0
        int ja = y - A.T0, jb = y - B.T0;
        ja = (ja < 1 ? 1 : (ja > A.qlen ? A.qlen : ja)) - 1, jb = (jb < 1 ? 1 : (jb > B.qlen ? B.qlen : jb)) - 1;
        int qa = qsA(ja), qb = qsB(jb);
        qa = qa > 4 ? 4 : qa, qb = qb > 4 ? 4 : qb;
        qc[x] = (uint8_t)(qa + 5 * qb);
    }
    __syncwarp();
    return __any_sync(0xffffffffu, amb);
}

// The DP pass of a pair.  sm_sel / sm_q / sm_lut: shared-memory addresses of the staged arrays and of the table of score-word pairs
// (25 x 8 bytes: {kb_qrow16(cqA), kb_qrow16(cqB)}).  tbA / tbB: traceback bytes, word ((r' - 1) / G) * 32 + lane holds the window
// cells of G = 4 / K consecutive anti-diagonals of one lane.
template <int K>
static __device__ __noinline__ void kb_band16_pass(const KbDpConst P, int lane, const KbB16Job A, const KbB16Job B, bool haveB, unsigned sm_sel,
                                                   unsigned sm_q, unsigned sm_lut, uint32_t *tbA, uint32_t *tbB, KbB16Out &out)
{
    constexpr int G = 4 / K, NST = K == 1 ? 4 : 2;
    const KbC16 c = kb_c16(P, 0, 0);
    const uint32_t NEGW = kb_pack2(KB_NEG16, KB_NEG16), NEGI = kb_pack2(KB_B16_NEG_INIT, KB_B16_NEG_INIT);
    auto lds8 = [](unsigned a) {
        unsigned v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
        return v;
    };
    auto lds16 = [](unsigned a) {
        unsigned v;
        asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(a));
        return v;
    };
    uint32_t H1[K], H2[K], E1[K], E2[K], F1[K], F2[K], sel[K], qlo[K], qhi[K];
#pragma unroll
    for (int m = 0; m < K; ++m) {
        const int c0 = lane * K + m;
        H1[m] = kb_pack2(A.T0 + c0 == 0 ? 0 : KB_B16_NEG_INIT, B.T0 + c0 == 0 ? 0 : KB_B16_NEG_INIT);
        H2[m] = E1[m] = E2[m] = F1[m] = F2[m] = NEGI;
        sel[m] = lds16(sm_sel + 2 * c0);
        const unsigned q = lds8(sm_q + (32 * K - c0));
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(qlo[m]), "=r"(qhi[m]) : "r"(sm_lut + 8 * q));
    }
    int doneA = 0, doneB = haveB ? 0 : 1;
    out.score[0] = out.score[1] = KB_NEG_INF, out.state[0] = out.state[1] = 0, out.steps[0] = out.steps[1] = 0;
    int r_stop = 0, cap = 0;  // last anti-diagonal any live job needs; next anti-diagonal on which a live job ends
    auto live = [&]() {
        const int ea = doneA ? 0x7fffffff : A.r_end, eb = doneB ? 0x7fffffff : B.r_end;
        cap = ea < eb ? ea : eb;
        const int sa = doneA ? 0 : A.r_end, sb = doneB ? 0 : B.r_end;
        r_stop = sa > sb ? sa : sb;
    };
    live();
    unsigned aq = sm_q + 32 * K + 1, at = sm_sel + 2 * (32 * K);  // what the next A step (row offset 1) / B step (position 32 K) takes in
    uint32_t *wa = tbA + lane, *wb = tbB + lane;
    uint32_t t01 = 0, t23 = 0;
    int rp = 1;
    for (;;) {
#pragma unroll
        for (int i = 0; i < NST; ++i) {
            if ((i % G) == 0) t01 = t23 = 0;
            if (!(i & 1)) {  // A: (t - 1, j) is window cell c - 1
                const unsigned q = lds8(aq);
                uint32_t nlo, nhi;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(nlo), "=r"(nhi) : "r"(sm_lut + 8 * q));
                ++aq;
                uint32_t uH = __shfl_up_sync(0xffffffffu, H1[K - 1], 1), uE1 = __shfl_up_sync(0xffffffffu, E1[K - 1], 1),
                         uE2 = __shfl_up_sync(0xffffffffu, E2[K - 1], 1);
                const uint32_t plo = __shfl_up_sync(0xffffffffu, qlo[K - 1], 1), phi = __shfl_up_sync(0xffffffffu, qhi[K - 1], 1);
#pragma unroll
                for (int m = K - 1; m > 0; --m) qlo[m] = qlo[m - 1], qhi[m] = qhi[m - 1];
                qlo[0] = lane == 0 ? nlo : plo, qhi[0] = lane == 0 ? nhi : phi;
                if (lane == 0) uH = uE1 = uE2 = NEGW;
#pragma unroll
                for (int m = K - 1; m >= 0; --m) {
                    const uint32_t hu = m > 0 ? H1[m > 0 ? m - 1 : 0] : uH;
                    uint32_t e1 = m > 0 ? E1[m > 0 ? m - 1 : 0] : uE1, e2 = m > 0 ? E2[m > 0 ? m - 1 : 0] : uE2;
                    const uint32_t s = (uint32_t)kb_prmt(qlo[m], qhi[m], sel[m]);
                    uint32_t d;
                    const uint32_t z = kb_cell16(c, hu, e1, e2, H1[m], F1[m], F2[m], H2[m], s, d);
                    H2[m] = H1[m], H1[m] = z, E1[m] = e1, E2[m] = e2;
                    const int q4 = (i % G) * K + m;  // byte of the traceback word
                    if (q4 == 0) t01 += d;
                    else if (q4 == 1) t01 += d << 8;
                    else if (q4 == 2) t23 += d;
                    else t23 += d << 8;
                }
            } else {  // B: (t, j - 1) is window cell c + 1
                const uint32_t nsel = lds16(at);
                at += 2;
                uint32_t lH = __shfl_down_sync(0xffffffffu, H1[0], 1), lF1 = __shfl_down_sync(0xffffffffu, F1[0], 1),
                         lF2 = __shfl_down_sync(0xffffffffu, F2[0], 1);
                const uint32_t psel = __shfl_down_sync(0xffffffffu, sel[0], 1);
#pragma unroll
                for (int m = 0; m + 1 < K; ++m) sel[m] = sel[m + 1];
                sel[K - 1] = lane == 31 ? nsel : psel;
                if (lane == 31) lH = lF1 = lF2 = NEGW;
#pragma unroll
                for (int m = 0; m < K; ++m) {
                    const uint32_t hl = m + 1 < K ? H1[m + 1 < K ? m + 1 : 0] : lH;
                    uint32_t f1 = m + 1 < K ? F1[m + 1 < K ? m + 1 : 0] : lF1, f2 = m + 1 < K ? F2[m + 1 < K ? m + 1 : 0] : lF2;
                    const uint32_t s = (uint32_t)kb_prmt(qlo[m], qhi[m], sel[m]);
                    uint32_t d;
                    const uint32_t z = kb_cell16(c, H1[m], E1[m], E2[m], hl, f1, f2, H2[m], s, d);
                    H2[m] = H1[m], H1[m] = z, F1[m] = f1, F2[m] = f2;
                    const int q4 = (i % G) * K + m;
                    if (q4 == 0) t01 += d;
                    else if (q4 == 1) t01 += d << 8;
                    else if (q4 == 2) t23 += d;
                    else t23 += d << 8;
                }
            }
            if ((i % G) == G - 1) {  // a traceback word of either job is complete
                kb_st_u32(wa, (uint32_t)kb_prmt(t01, t23, 0x5410)), kb_st_u32(wb, (uint32_t)kb_prmt(t01, t23, 0x7632));
                wa += 32, wb += 32;
            }
            if (rp == cap) {  // H(tlen, qlen) of a job is in the window now
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const KbB16Job &J = w ? B : A;
                    if ((w ? doneB : doneA) || rp != J.r_end) continue;
                    const int cend = J.tlen - (J.T0 + (rp >> 1));
                    uint32_t hv = __shfl_sync(0xffffffffu, H1[0], (cend / K) & 31);  // one shuffle per slot: indexing H1[] would put it in local memory
#pragma unroll
                    for (int m = 1; m < K; ++m) {
                        const uint32_t o = __shfl_sync(0xffffffffu, H1[m], (cend / K) & 31);
                        hv = m == cend % K ? o : hv;
                    }
                    out.score[w] = (w ? kb_hi16(hv) : kb_lo16(hv)) >> 3, out.state[w] = 1, out.steps[w] = rp;
                    if (w) doneB = 1;
                    else doneA = 1;
                }
                live();
            }
            ++rp;
        }
        if (doneA && doneB) break;
        if ((rp & 31) == 1) {
            // a path ends on one of the last two anti-diagonals and gains at most +a per remaining column: give a job up once the
            // certificate is out of reach.  The score handed back is then an extrapolation (the same loss per column for the rest), good
            // enough to choose the wider window, whose own pass is certified on its real score.
            uint32_t v = __vmaxs2(H1[0], H2[0]);
#pragma unroll
            for (int m = 1; m < K; ++m) v = __vmaxs2(v, __vmaxs2(H1[m], H2[m]));
            const int mA = __reduce_max_sync(0xffffffffu, kb_lo16(v)) >> 3, mB = __reduce_max_sync(0xffffffffu, kb_hi16(v)) >> 3;
            const int last = rp - 1, done = (last + 1) >> 1;
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                const KbB16Job &J = w ? B : A;
                if (w ? doneB : doneA) continue;
                const int mx = w ? mB : mA, rem = (J.r_end - last + 2) >> 1;
                if (mx + P.a * rem <= J.bound) {
                    const int loss = P.a * done - mx;
                    out.score[w] = P.a * (done + rem) - (int)((int64_t)loss * (done + rem) / (done > 0 ? done : 1)), out.state[w] = 0;
                    out.steps[w] = last;
                    if (w) doneB = 1;
                    else doneA = 1;
                }
            }
            if (doneA && doneB) break;
            live();
        }
        if (rp > r_stop) break;
    }
}

// traceback of one certified job of the pass into S.ezcig; returns the CIGAR length, or -1 (cannot happen once certified)
template <int K>
static __device__ __forceinline__ int kb_band16_backtrack(int lane, const KbB16Job &J, int flag, const uint32_t *tbw, uint32_t *ezcig)
{
    constexpr int G = 4 / K;
    const uint8_t *tb = reinterpret_cast<const uint8_t *>(tbw);
    int bad = 0;
    const int n_cigar = kb_backtrack_warp(lane, J.tlen - 1, J.qlen - 1, 0, flag, ezcig, [&](int i, int j) -> uint32_t {
        const int r2 = i + j + 2, cc = i + 1 - (J.T0 + (r2 >> 1));
        if ((unsigned)cc >= (unsigned)(32 * K)) {
            bad = 1;
            return 0xffu;
        }
        const int l = cc / K, m = cc - l * K;
        return (uint32_t)kb_ld_u8(tb + ((size_t)((r2 - 1) / G) * 32 + l) * 4 + ((r2 - 1) % G) * K + m);
    });
    return __any_sync(0xffffffffu, bad) ? -1 : n_cigar;
}

#endif  // __CUDACC__
