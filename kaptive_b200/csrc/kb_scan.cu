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

// (ax, ay): later entries, (bx, by): earlier ones -> the minimum, the later one among equals
__device__ __forceinline__ void kb_ladder(uint32_t &ax, uint32_t &ay, uint32_t bx, uint32_t by, bool &tie)
{
    tie |= ax == bx;
    ay = ax <= bx ? ay : by;
    ax = min(ax, bx);
}

template <int W, int K>
__global__ void __launch_bounds__(128, 6) kb_scan_kernel(KbIndexView ix, KbBatchView bt, uint64_t *akey, uint32_t *aval,
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
    // ring slots of the entry d positions back, for even and odd tiles
    int s0[2], s1[2], s2[2], s4[2], s8[2], s10[2];
#pragma unroll
    for (int par = 0; par < 2; ++par) {
        const int b = par * 32 + lane;
        s0[par] = b, s1[par] = (b - 1) & 63, s2[par] = (b - 2) & 63, s4[par] = (b - 4) & 63, s8[par] = (b - 8) & 63, s10[par] = (b - 10) & 63;
    }
    // the lane's k-mer window [i-14, i] inside (previous word : tile low word : tile high word)
    const bool sel_a = lane < 14, sel_c = lane >= 30;
    const int ksh = 2 * ((lane + 2) & 15);

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
        // one silent tile in front of every chunk but the first rebuilds the window state (w + k - 1 = 24 < 32 bases)
        const int n_sil = cstart > 0 ? 1 : 0;
        const int n_tiles = n_sil + ((cend - cstart + 31) >> 5);
        int pos = cstart - 32 * n_sil;  // contig position of lane 0 of the current tile
        const uint32_t *gw = bt.seq2 + ((soff + pos) >> 4);
        const uint32_t *gm = bt.nmask + ((soff + pos) >> 5);
        uint32_t wprev = __ldg(gw - 1), mprev = __ldg(gm - 1);
        uint32_t wlo = __ldg(gw), whi = __ldg(gw + 1), mcur = __ldg(gm);
        bool pe = false;  // deferred emission: valid, hash, position, bitmap word
        uint32_t px = 0, py = 0, pw = 0;
        uint32_t lmx = KB_MAXU, lmy = KB_MAXU;  // M[i] of the last tile
        bool ptie = true;
        auto push = [&](bool pass, uint32_t x, uint32_t y) {
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            if (pass) {
                const int o = front + __popc(bal & ((1u << lane) - 1u));
                Q.x[o] = x, Q.y[o] = y;
            }
            front += __popc(bal);
        };
        auto tile = [&](const int par, const bool live) {
            const int i = pos + lane;
            const uint32_t lo = sel_a ? wprev : (sel_c ? whi : wlo), hi = sel_a ? wlo : (sel_c ? 0u : whi);
            const uint32_t win = __funnelshift_r(lo, hi, ksh) & mask;  // base i-14 in bits 0-1 ... base i in bits 28-29
            uint32_t t = __brev(win);
            t = ((t >> 1) & 0x55555555u) | ((t & 0x55555555u) << 1);
            const uint32_t fwd = t >> 2, rev = ~win & mask;
            const uint32_t z = fwd < rev ? 0u : 1u;
            const int l = __clz((int)__funnelshift_l(pos == 0 ? 0xffffffffu : mprev, mcur, 31 - lane));  // unambiguous run ending at i (<= 32)
            uint32_t x = KB_MAXU, y = KB_MAXU;
            if (l >= K) x = kb_hash32(z ? rev : fwd, mask), y = ((uint32_t)i << 1) | z;
            R.x[s0[par]] = x, R.y[s0[par]] = y;
            wprev = whi, mprev = mcur;
            gw += 2, gm += 1, pos += 32;
            wlo = __ldg(gw), whi = __ldg(gw + 1), mcur = __ldg(gm);  // next tile (the storage is padded by 128 bases)
            __syncwarp();
            bool tie = false;
            uint32_t ax = x, ay = y;
            const uint32_t x10 = R.x[s10[par]];
            kb_ladder(ax, ay, R.x[s1[par]], R.y[s1[par]], tie);
            R.a2x[s0[par]] = ax, R.a2y[s0[par]] = ay;
            __syncwarp();
            const uint32_t cx = R.a2x[s8[par]], cy = R.a2y[s8[par]];
            kb_ladder(ax, ay, R.a2x[s2[par]], R.a2y[s2[par]], tie);
            R.a4x[s0[par]] = ax, R.a4y[s0[par]] = ay;
            __syncwarp();
            kb_ladder(ax, ay, R.a4x[s4[par]], R.a4y[s4[par]], tie);
            kb_ladder(ax, ay, cx, cy, tie);
            R.mx[s0[par]] = ax, R.my[s0[par]] = ay;
            __syncwarp();
            const uint32_t omx = R.mx[s1[par]], omy = R.my[s1[par]];
            lmx = ax, lmy = ay;
            const bool act = live && i < cend;
            const bool e2 = x10 < ax && l >= W + K - 1;
            const bool e = act && (e2 || (x <= omx && l >= W + K && omx != KB_MAXU));
            n_emit += e ? 1u : 0u;
            const bool anyt = __any_sync(0xffffffffu, tie);
            if (anyt || ptie) {  // equal hashes somewhere near: the identical-k-mer rules of mm_sketch
                if (act && l == W + K - 1 && omx != KB_MAXU)
                    for (int d = 1; d < W; ++d) {
                        const uint32_t jx = R.x[(s0[par] - d) & 63], jy = R.y[(s0[par] - d) & 63];
                        if (jx == omx && jy != omy) slow(jx, jy);
                    }
                if (act && e2 && ax != KB_MAXU)
                    for (int d = 0; d < W; ++d) {
                        const uint32_t jx = R.x[(s0[par] - d) & 63], jy = R.y[(s0[par] - d) & 63];
                        if (jx == ax && jy != ay) slow(jx, jy);
                    }
                __syncwarp();
            }
            ptie = anyt;
            // presence filter, one tile deferred so that the bitmap load (L2) is not waited for
            push(pe && (nofilter || ((pw >> (px & 31u)) & 1u)), px, py);
            pe = e, px = omx, py = omy;
            if (e && !nofilter) pw = __ldg(ix.bloom + ((omx & ix.bloom_mask) >> 5));
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
                if (mz_hash && have && asm_id == mz_asm) {  // debug / parity dump of one assembly's minimizers
                    unsigned long long o = atomicAdd(&counters[7], 1ull);
                    if ((int64_t)o < mz_cap) mz_hash[o] = hx, mz_ctg[o] = ctg - bt.asm_ctg_start[asm_id], mz_pos[o] = hy;
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
            tile(0, u >= n_sil);
            if (u + 1 < n_tiles) tile(1, true);
            __syncwarp();
            if (front + tail >= 32) drain();
        }
        push(pe && (nofilter || ((pw >> (px & 31u)) & 1u)), px, py);  // the deferred emission of the last tile
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
    int64_t grid = (int64_t)n_sm * 6;  // 6 CTAs of 128 threads (24 KB of shared memory each) per SM, grid-stride over the chunks
    if (grid > want) grid = want;
    cudaFuncSetAttribute(kb_scan_kernel<10, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(ScanQueue)));  // per device
    kb_scan_kernel<10, 15><<<(unsigned)grid, 128, 4 * sizeof(ScanQueue), st>>>(ix, bt, akey, aval, counters, anchor_cap, mz_hash, mz_ctg,
                                                           mz_pos, mz_cap, mz_asm);
}

#endif  // __CUDACC__
