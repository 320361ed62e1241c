// kb_scan.cu -- the seeding scan: minimizer sketch of every contig + lookup of
// each minimizer in the gene hash + anchor emission, in one pass over the
// 2-bit packed assembly.  This is the HBM-streaming kernel the roofline is
// quoted on (DESIGN.md "Scan kernel").
//
// Work decomposition: one warp per chunk of KB_CHUNK_BASES (8192) bases of one
// contig, walked in tiles of 32 consecutive positions, one position per lane
// (see "position-parallel mm_sketch" below): the tile's two sequence words and
// one mask word are read once, one tile ahead.  An emitted
// minimizer asks the presence bitmap (load issued one tile ahead); survivors go
// to a per-warp shared-memory queue; every time the queue holds >= 32 the warp
// drains it: one hash-table probe per lane (the table lives in L2),
// warp-aggregated allocation of anchor slots (one global atomic per drain),
// cooperative expansion of multi-occurrence minimizers.
#include "kb_scan.cuh"
#include "kb_kernels.h"

#ifndef KB_SCAN_CTAS
#define KB_SCAN_CTAS 6  // CTAs of 128 threads per SM (measured: 8 CTAs at 64 registers are 5 % slower, the kernel is ALU-pipe bound)
#endif
#define KB_QCAP 512  // queue capacity per warp: 31 leftover + two tiles of 32 regular pushes in front, the rare identical-k-mer pushes from the back

// key/val of one anchor: assembly minimizer (rid-local pos `tpos`, strand tz) x gene entry e
KB_HD void kb_make_anchor(const KbEntry &en, int asm_id, int vpos, int tz, uint32_t eidx, uint64_t *key, uint32_t *val)
{
    int rev = (int)(en.qpos_z & 1u) != tz;
    *key = ((uint64_t)asm_id << KB_KEY_ASM_SHIFT) | ((uint64_t)en.gene << KB_KEY_GENE_SHIFT) |
           ((uint64_t)rev << KB_KEY_REV_SHIFT) | (uint64_t)vpos;
    *val = eidx;
}

KB_HD bool kb_ht_lookup(const uint64_t *ht, uint32_t ht_mask, uint32_t hash, uint32_t *start, uint32_t *count)
{
    uint32_t slot = hash & ht_mask;
    for (;;) {
        uint64_t e = ht[slot];
        if (e == 0) return false;
        if ((uint32_t)(e >> KB_HT_KEY_SHIFT) == hash) {
            *start = (uint32_t)(e >> KB_HT_START_SHIFT) & KB_HT_START_MASK;
            *count = (uint32_t)e & KB_HT_COUNT_MASK;
            return true;
        }
        slot = (slot + 1) & ht_mask;
    }
}

#ifdef __CUDACC__

struct ScanQueue {
    uint32_t x[KB_QCAP];
    uint32_t y[KB_QCAP];
};

// ---- position-parallel mm_sketch ------------------------------------------------------------------------------------
// mm_sketch (minimap2 sketch.c) is written as a state machine over the sequence, but what it emits at step i depends only
// on a bounded neighbourhood (scripts/proto_sketch_parallel.py checks this restatement against the oracle):
//   x[i], y[i]  hash and (position << 1 | strand) of the k-mer ending at i, MAX unless the last l[i] >= k bases are unambiguous
//   M[i]        the minimum of x over [i-w+1, i], the rightmost one among equals (mm_sketch's `min` after step i)
//   E1  x[i] <= M[i-1].x, l[i] >= w+k, M[i-1].x != MAX    -> M[i-1] is written
//   E2  x[i-w] < M[i].x, l[i] >= w+k-1                   -> M[i-1] (= entry i-w) is written, then every other entry of the
//                                                           window equal to M[i].x
//   E0  l[i] == w+k-1, M[i-1].x != MAX                   -> every entry of [i-w+1, i-1] equal to M[i-1].x except M[i-1]
//   End i == len-1, M[i].x != MAX                        -> M[i]
// so lane L of a warp takes position 32*t + L of tile t: the k-mer comes straight out of the packed words (funnel shift,
// bit reversal for the forward strand), l[i] out of the mask words (count leading zeros), and M[i] from a doubling ladder
// over shared-memory rings (pairs, fours, eights, then eight + two = w = 10).  Equal hashes inside a window (E0 and the
// second half of E2) are rare: every ladder comparison that sees equal minima raises a warp-wide flag, and only then the
// lanes concerned walk their window.
struct ScanRing {
    uint32_t x[64], y[64], a2x[64], a2y[64], a4x[64], a4y[64], mx[64], my[64];  // [tile parity][lane]
};

// Ring accesses by 32-bit shared address + immediate offset: the six lane addresses below are computed once per kernel
// (the compiler would otherwise rebuild them from the thread index in every tile, a sixth of the loop's instructions).
#define KB_LDS(dst, addr, off) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(dst) : "r"((addr) + (uint32_t)(off)) : "memory")
#define KB_STS(addr, off, v) asm volatile("st.shared.u32 [%0], %1;" ::"r"((addr) + (uint32_t)(off)), "r"(v) : "memory")
enum { RX = 0, RY = 256, RA2X = 512, RA2Y = 768, RA4X = 1024, RA4Y = 1280, RMX = 1536, RMY = 1792 };  // byte offsets in ScanRing

template <int W, int K>
__global__ void __launch_bounds__(128, KB_SCAN_CTAS) kb_scan_kernel(KbIndexView ix, KbBatchView bt, uint64_t *akey, uint32_t *aval,
                                                      unsigned long long *counters, int64_t anchor_cap,
                                                      uint32_t *mz_hash, int32_t *mz_ctg, uint32_t *mz_pos,
                                                      int64_t mz_cap, int32_t mz_asm)
{
    static_assert(W == 10 && K == 15, "the ladder below is written for w = 10 (8 + 2) and k-mers that fit one 32-bit word");
    extern __shared__ __align__(16) unsigned char kb_scan_dyn[];  // the four per-warp queues
    ScanQueue *queues = reinterpret_cast<ScanQueue *>(kb_scan_dyn);
    __shared__ int qtail[4];
    __shared__ ScanRing rings[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ScanRing &R = rings[warp];
    ScanQueue &Q = queues[warp];
    int &tail = qtail[warp];
    const int64_t n_warps = (int64_t)gridDim.x * 4;
    unsigned long long n_min_local = 0;
    unsigned n_emit = 0;                       // this lane's regular emissions (all of them, whatever the presence filter says)
    const bool nofilter = mz_hash != nullptr;  // minimizer dump (parity tests): every minimizer has to reach the queue
    const uint32_t mask = (1u << (2 * K)) - 1u;
    // shared addresses of the lane's ring slot in an even tile (a0) and of the slots d positions back from it (wrapping
    // into the odd half); in an odd tile the slot is a0 + 128 and the slots behind it are plain negative offsets
    uint32_t a0, w1, w2, w4, w8, w10;
    {
        const uint32_t rb = (uint32_t)__cvta_generic_to_shared(&R);
        a0 = rb + 4u * (uint32_t)lane;
        w1 = rb + 4u * (uint32_t)((lane - 1) & 63), w2 = rb + 4u * (uint32_t)((lane - 2) & 63), w4 = rb + 4u * (uint32_t)((lane - 4) & 63);
        w8 = rb + 4u * (uint32_t)((lane - 8) & 63), w10 = rb + 4u * (uint32_t)((lane - 10) & 63);
        asm volatile("" : "+r"(a0), "+r"(w1), "+r"(w2), "+r"(w4), "+r"(w8), "+r"(w10));
    }
#define KB_BACK(dst, D, WD, OFF)                                   \
    do {                                                           \
        if (PAR == 0) KB_LDS(dst, WD, OFF);                        \
        else KB_LDS(dst, a0, (OFF) + 128 - 4 * (D));               \
    } while (0)
    // loop invariants the compiler would otherwise re-read from the parameter bank (or rebuild from the thread index) per tile
    const uint32_t *bloom = ix.bloom;
    uint32_t bloom_mask = ix.bloom_mask;
    uint32_t nf_mask = nofilter ? 0xffffffffu : 0u;  // ORed into the bitmap word: the dump sees every minimizer
    uint32_t qb = (uint32_t)__cvta_generic_to_shared(&Q);
    asm volatile("" : "+l"(bloom), "+r"(bloom_mask), "+r"(nf_mask), "+r"(qb));
    // the lane's k-mer window [i-14, i] starts (lane + 2) bases into the word in front of the tile: the lane reads the two
    // words that hold it itself (through L1; the warp touches three consecutive words) and funnel-shifts them
    const int kword = ((lane + 2) >> 4) - 1;
    uint32_t ksh = 2 * ((lane + 2) & 15), msh = 31 - lane;
    asm volatile("" : "+r"(ksh), "+r"(msh));
    for (int64_t chunk = (int64_t)blockIdx.x * 4 + warp; chunk < bt.n_chunks; chunk += n_warps) {
        const int ctg = bt.chunk_ctg[chunk];
        const int cstart = bt.chunk_start[chunk];
        const int clen = bt.ctg_len[ctg];
        const int asm_id = bt.ctg_asm[ctg];
        const int vstart = bt.ctg_vstart[ctg];
        const int64_t soff = bt.ctg_soff[ctg];
        const int cend = cstart + KB_CHUNK_BASES < clen ? cstart + KB_CHUNK_BASES : clen;
        if (lane == 0) tail = 0;
#pragma unroll
        for (int h = 0; h < 64; h += 32) {
            R.x[h + lane] = R.y[h + lane] = R.a2x[h + lane] = R.a2y[h + lane] = KB_MAXU;
            R.a4x[h + lane] = R.a4y[h + lane] = R.mx[h + lane] = R.my[h + lane] = KB_MAXU;
        }
        __syncwarp();
        int front = 0;  // warp-uniform: entries pushed at the front of the queue
        auto slow = [&](uint32_t x, uint32_t y) {
            int k = atomicAdd(&tail, 1);  // `tail` counts the slow entries, stored from the back
            if (front + k < KB_QCAP - 96) Q.x[KB_QCAP - 1 - k] = x, Q.y[KB_QCAP - 1 - k] = y;
            else atomicOr(&counters[6], 1ull);
        };
        // two silent tiles in front of every chunk but the first rebuild the window state (w + k - 1 = 24 bases are needed;
        // 64 keep the pair loads below 16-byte aligned)
        const int n_sil = cstart > 0 ? 2 : 0;
        const int n_tiles = n_sil + ((cend - cstart + 31) >> 5);
        int pos = cstart - 32 * n_sil;  // contig position of lane 0 of the current tile
        const uint32_t *gw = bt.seq2 + ((soff + pos) >> 4);
        const uint32_t *gm = bt.nmask + ((soff + pos) >> 5);
        gw += kword;  // per lane from here on
        uint32_t mprev = __ldg(gm - 1);
        uint32_t k0 = __ldg(gw), k1 = __ldg(gw + 1);             // the lane's sequence words of the current tile
        uint2 mk = __ldg(reinterpret_cast<const uint2 *>(gm));  // mask words of the current pair of tiles
        bool pe = false;  // deferred emission: valid, hash, position, bitmap word
        uint32_t px = 0, py = 0, pw = 0;
        uint32_t lmx = KB_MAXU, lmy = KB_MAXU;  // M[i] of the last tile
        bool ptie = true;
        auto push = [&](bool pass, uint32_t x, uint32_t y) {
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            if (pass) {
                const uint32_t o = qb + 4u * (uint32_t)(front + __popc(bal & ((1u << lane) - 1u)));
                KB_STS(o, 0, x);
                KB_STS(o, 4 * KB_QCAP, y);
            }
            front += __popc(bal);
        };
        auto tile = [&](const int PAR, const bool live, const uint32_t mp, const uint32_t mcur) {
            const uint32_t ao = a0 + PAR * 128;  // folded into the address immediates
            const int i = pos + lane;
            const uint32_t win = __funnelshift_r(k0, k1, ksh) & mask;  // base i-14 in bits 0-1 ... base i in bits 28-29
            gw += 2;
            k0 = __ldg(gw), k1 = __ldg(gw + 1);  // next tile (the storage is padded by 128 bases)
            uint32_t t = __brev(win);
            t = ((t >> 1) & 0x55555555u) | ((t & 0x55555555u) << 1);
            const uint32_t fwd = t >> 2, rev = ~win & mask;
            const uint32_t z = fwd < rev ? 0u : 1u;
            const int l = __clz((int)__funnelshift_l(PAR == 0 && pos == 0 ? 0xffffffffu : mp, mcur, msh));  // unambiguous run ending at i (<= 32)
            const uint32_t h = kb_hash32(min(fwd, rev), mask);
            const uint32_t x = l >= K ? h : KB_MAXU, y = l >= K ? (((uint32_t)i << 1) | z) : KB_MAXU;
            KB_STS(ao, RX, x);
            KB_STS(ao, RY, y);
            pos += 32;
            __syncwarp();
            bool tie;
            uint32_t ax = x, ay = y, bx, x10;
            // ladder step: (ax, ay) the later entries, b the earlier ones -> the minimum, the later one among equals
            KB_BACK(bx, 1, w1, RX);
            KB_BACK(x10, 10, w10, RX);
            tie = ax == bx;
            if (ax > bx) KB_BACK(ay, 1, w1, RY);
            ax = min(ax, bx);
            KB_STS(ao, RA2X, ax);
            KB_STS(ao, RA2Y, ay);
            __syncwarp();
            uint32_t cx, cy;
            KB_BACK(bx, 2, w2, RA2X);
            KB_BACK(cx, 8, w8, RA2X);
            KB_BACK(cy, 8, w8, RA2Y);
            tie |= ax == bx;
            if (ax > bx) KB_BACK(ay, 2, w2, RA2Y);
            ax = min(ax, bx);
            KB_STS(ao, RA4X, ax);
            KB_STS(ao, RA4Y, ay);
            __syncwarp();
            KB_BACK(bx, 4, w4, RA4X);
            tie |= ax == bx;
            if (ax > bx) KB_BACK(ay, 4, w4, RA4Y);
            ax = min(ax, bx);
            tie |= ax == cx;
            ay = ax > cx ? cy : ay;
            ax = min(ax, cx);
            KB_STS(ao, RMX, ax);
            KB_STS(ao, RMY, ay);
            __syncwarp();
            uint32_t omx, omy;
            KB_BACK(omx, 1, w1, RMX);
            KB_BACK(omy, 1, w1, RMY);
            lmx = ax, lmy = ay;
            const bool act = live && i < cend;
            const bool e2 = x10 < ax && l >= W + K - 1;
            // E1 needs omx != MAX: as signed numbers MAX is -1 and a valid x (l >= k) is never below it
            const bool e = act && (e2 || ((int32_t)x <= (int32_t)omx && l >= W + K));
            n_emit += e ? 1u : 0u;
            const bool anyt = __any_sync(0xffffffffu, tie);
            if (anyt || ptie) {  // equal hashes somewhere near: the identical-k-mer rules of mm_sketch
                const int own = PAR * 32 + lane;
                if (act && l == W + K - 1 && omx != KB_MAXU)
                    for (int d = 1; d < W; ++d) {
                        const uint32_t jx = R.x[(own - d) & 63], jy = R.y[(own - d) & 63];
                        if (jx == omx && jy != omy) slow(jx, jy);
                    }
                if (act && e2 && ax != KB_MAXU)
                    for (int d = 0; d < W; ++d) {
                        const uint32_t jx = R.x[(own - d) & 63], jy = R.y[(own - d) & 63];
                        if (jx == ax && jy != ay) slow(jx, jy);
                    }
                __syncwarp();
            }
            ptie = anyt;
            // presence filter, one tile deferred so that the bitmap load (L2) is not waited for
            push(pe && (((pw | nf_mask) >> (px & 31u)) & 1u), px, py);
            pe = e, px = omx, py = omy;
            if (e) pw = __ldg(bloom + ((omx & bloom_mask) >> 5));  // consumed in the next tile: nothing may touch pw before
        };
        auto drain = [&]() {
            __syncwarp();
            const int n_slow = tail;
            const int n = front + n_slow;
            n_min_local += (lane == 0) ? (unsigned long long)n_slow : 0ull;  // fast emissions are counted per lane (n_emit)
            for (int base = 0; base < n; base += 32) {
                int qi = base + lane;
                bool have = qi < n;
                int qslot = qi < front ? qi : KB_QCAP - 1 - (qi - front);
                uint32_t hx = have ? Q.x[qslot] : 0, hy = have ? Q.y[qslot] : 0;
                uint32_t est = 0, ecnt = 0;
                bool hit = have && kb_ht_lookup(ix.ht, ix.ht_mask, hx, &est, &ecnt);
                if (mz_hash && have && (asm_id == mz_asm || mz_asm == -2)) {
                    // minimizer dump: one assembly with contig / position (parity tests), or every chunk given with the assembly id (census)
                    unsigned long long o = atomicAdd(&counters[7], 1ull);
                    if ((int64_t)o < mz_cap) {
                        mz_hash[o] = hx, mz_ctg[o] = mz_asm == -2 ? asm_id : ctg - bt.asm_ctg_start[asm_id];
                        if (mz_pos) mz_pos[o] = hy;
                    }
                }
                uint32_t cnt = hit ? ecnt : 0;
                // warp-aggregated slot allocation
                uint32_t incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                if (total == 0) continue;
                unsigned long long gbase = 0;
                if (lane == 0) gbase = atomicAdd(&counters[1], (unsigned long long)total);
                gbase = __shfl_sync(0xffffffffu, gbase, 0);
                uint32_t excl = incl - cnt;
                unsigned hm = __ballot_sync(0xffffffffu, cnt > 0);
                while (hm) {  // expand each hit cooperatively: coalesced entry reads and anchor writes
                    int src = __ffs(hm) - 1;
                    hm &= hm - 1;
                    uint32_t s_st = __shfl_sync(0xffffffffu, est, src);
                    uint32_t s_cnt = __shfl_sync(0xffffffffu, cnt, src);
                    uint32_t s_y = __shfl_sync(0xffffffffu, hy, src);
                    uint32_t s_off = __shfl_sync(0xffffffffu, excl, src);
                    for (uint32_t j = lane; j < s_cnt; j += 32) {
                        int64_t o = (int64_t)gbase + s_off + j;
                        if (o < anchor_cap) {
                            KbEntry en = ix.ent[s_st + j];
                            uint64_t key;
                            uint32_t val;
                            kb_make_anchor(en, asm_id, vstart + (int)(s_y >> 1), (int)(s_y & 1u), s_st + j, &key, &val);
                            akey[o] = key, aval[o] = val;
                        }
                    }
                }
            }
            __syncwarp();
            front = 0;
            if (lane == 0) tail = 0;
            __syncwarp();
        };
        for (int u = 0; u < n_tiles; u += 2) {
            gm += 2;
            const uint2 nmk = __ldg(reinterpret_cast<const uint2 *>(gm));  // next pair
            tile(0, u >= n_sil, mprev, mk.x);
            if (u + 1 < n_tiles) tile(1, u >= n_sil, mk.x, mk.y);
            mprev = mk.y, mk = nmk;
            __syncwarp();
            if (front + tail >= 32) drain();
        }
        push(pe && (((pw | nf_mask) >> (px & 31u)) & 1u), px, py);  // the deferred emission of the last tile
        {  // mm_sketch's final push, by the lane that holds the last base of the contig
            const bool e = cend == clen && clen > 0 && lane == ((clen - 1) & 31) && lmx != KB_MAXU;
            push(e, lmx, lmy);
            n_emit += e ? 1u : 0u;
        }
        __syncwarp();
        if (front + tail > 0) drain();
    }
    {
        unsigned long long t = n_emit;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        n_min_local += t;
    }
    if (lane == 0 && n_min_local) atomicAdd(&counters[0], n_min_local);
}

void kb_launch_scan(const KbIndexView &ix, const KbBatchView &bt, uint64_t *akey, uint32_t *aval,
                    unsigned long long *counters, int64_t anchor_cap, uint32_t *mz_hash, int32_t *mz_ctg,
                    uint32_t *mz_pos, int64_t mz_cap, int32_t mz_asm, int n_sm, cudaStream_t st)
{
    if (bt.n_chunks == 0) return;
    int64_t want = (bt.n_chunks + 3) / 4;
    int64_t grid = (int64_t)n_sm * KB_SCAN_CTAS;  // grid-stride over the chunks
    if (grid > want) grid = want;
    cudaFuncSetAttribute(kb_scan_kernel<10, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(ScanQueue)));  // per device
    kb_scan_kernel<10, 15><<<(unsigned)grid, 128, 4 * sizeof(ScanQueue), st>>>(ix, bt, akey, aval, counters, anchor_cap, mz_hash, mz_ctg,
                                                           mz_pos, mz_cap, mz_asm);
}

#endif  // __CUDACC__
