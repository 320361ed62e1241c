// kb_scan.cu -- the seeding scan: minimizer sketch of every contig + lookup of
// each minimizer in the gene hash + anchor emission, in one pass over the
// 2-bit packed assembly.  This is the HBM-streaming kernel the roofline is
// quoted on (DESIGN.md "Scan kernel").
//
// Work decomposition: one warp per chunk of KB_CHUNK_BASES (8192) bases of one
// contig; lane L owns bases [L*256, (L+1)*256) of the chunk and reads them 16 at
// a time (one 32-bit word of 2-bit sequence per 16 steps, one mask word per 32:
// every byte of the batch is read exactly once, through L1), starts 24 bases
// early in silent mode to rebuild the sketch state, keeps the rolling k-mers and
// minima in registers and the w-entry window in shared memory.  An emitted
// minimizer asks the presence bitmap (load issued one step ahead); survivors go
// to a per-warp shared-memory queue; every time the queue holds >= 32 the warp
// drains it: one hash-table probe per lane (the table lives in L2),
// warp-aggregated allocation of anchor slots (one global atomic per drain),
// cooperative expansion of multi-occurrence minimizers.
#include "kb_scan.cuh"
#include "kb_kernels.h"

#define KB_QCAP 512  // queue capacity per warp: 31 leftover + 32 lanes * up to (W+1) pushes in one step

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

// The sketch step of kb_fast_step (kb_scan.cuh) with the window in shared memory instead of registers: the slot is a
// run-time value, so one copy of the step serves all W slots and the rare identical-k-mer paths are real loops.  The
// whole scan loop then fits the instruction cache (the W-fold unrolled register version is ~80 KB of code and spends a
// third of its time waiting for instructions).  Arrays are [slot][lane]: conflict free.
struct KbSketchRegs {
    uint32_t px, py, pdup;  // prefix minimum of the current block
    uint32_t mx, my;        // mm_sketch's `min` after the previous step
    uint32_t sdup;          // bit k: the hash of sx[k] occurs more than once in slots k..W-1
    uint32_t fwd, rev;
    int l;
};
template <int W, int K, class Slow>
__device__ __forceinline__ bool kb_fast_step_sm(KbSketchRegs &s, uint32_t *bx, uint32_t *by, uint32_t *sx, uint32_t *sy, int u, int i, int c,
                                                bool live, uint32_t *ex, uint32_t *ey, Slow &slow)
{
    const uint32_t mask = (1u << (2 * K)) - 1u;
    const int shift1 = 2 * (K - 1);
    uint32_t ix = KB_MAXU, iy = KB_MAXU;
    if (c < 4) {
        s.fwd = ((s.fwd << 2) | (uint32_t)c) & mask;
        s.rev = (s.rev >> 2) | ((3u ^ (uint32_t)c) << shift1);
        const int z = s.fwd < s.rev ? 0 : 1;
        ++s.l;
        if (s.l >= K) ix = kb_hash32(z ? s.rev : s.fwd, mask), iy = ((uint32_t)i << 1) | (uint32_t)z;
    } else s.l = 0;
    bx[u * 32] = ix, by[u * 32] = iy;
    const uint32_t omx = s.mx, omy = s.my;  // `min` before this step
    // first full window: mm_sketch emits the entries equal to the old minimum that are not the minimum itself
    if (s.l == W + K - 1 && omx != KB_MAXU && live) {
        for (int j = u + 1; j < W; ++j)
            if (omx == bx[j * 32] && by[j * 32] != omy) slow(bx[j * 32], by[j * 32]);
        for (int j = 0; j < u; ++j)
            if (omx == bx[j * 32] && by[j * 32] != omy) slow(bx[j * 32], by[j * 32]);
    }
    // prefix minimum of the current block (the latest entry wins ties)
    if (u == 0) s.px = ix, s.py = iy, s.pdup = 0;
    else if (ix <= s.px) s.pdup = (ix == s.px), s.px = ix, s.py = iy;
    // window minimum = suffix of the previous block (slots u+1..W-1) combined with the prefix (later, so it wins ties)
    uint32_t nx = s.px, ny = s.py, ndup = s.pdup;
    if (u + 1 < W) {
        const uint32_t qx = sx[(u + 1) * 32], qy = sy[(u + 1) * 32], qd = (s.sdup >> (u + 1)) & 1u;
        if (qx < nx) nx = qx, ny = qy, ndup = qd;
        else if (qx == nx) ndup = 1;
    }
    bool emit = false;
    if (ix <= omx) {  // new minimum: write the old one
        emit = s.l >= W + K && omx != KB_MAXU;
    } else if ((omy >> 1) == (uint32_t)(i - W)) {  // the old minimum left the window (it is valid here: ix > omx)
        emit = s.l >= W + K - 1;
        if (s.l >= W + K - 1 && nx != KB_MAXU && ndup && live) {  // identical k-mers of the new minimum
            for (int j = u + 1; j < W; ++j)
                if (nx == bx[j * 32] && ny != by[j * 32]) slow(bx[j * 32], by[j * 32]);
            for (int j = 0; j <= u; ++j)
                if (nx == bx[j * 32] && ny != by[j * 32]) slow(bx[j * 32], by[j * 32]);
        }
    }
    *ex = omx, *ey = omy;
    s.mx = nx, s.my = ny;
    if (u == W - 1) {  // block complete: rebuild the suffix minima (later entries win ties)
        uint32_t nxt = ix, nyt = iy, dup = 0, dk = 0;
        sx[(W - 1) * 32] = nxt, sy[(W - 1) * 32] = nyt;
#pragma unroll
        for (int k = W - 2; k >= 0; --k) {
            const uint32_t cx = bx[k * 32];
            if (cx < nxt) nxt = cx, nyt = by[k * 32], dk = 0;
            else dk = (cx == nxt) ? 1u : dk;
            dup |= dk << k;
            sx[k * 32] = nxt, sy[k * 32] = nyt;
        }
        s.sdup = dup;
    }
    return emit && live;
}

template <int W, int K>
__global__ void __launch_bounds__(128, 6) kb_scan_kernel(KbIndexView ix, KbBatchView bt, uint64_t *akey, uint32_t *aval,
                                                      unsigned long long *counters, int64_t anchor_cap,
                                                      uint32_t *mz_hash, int32_t *mz_ctg, uint32_t *mz_pos,
                                                      int64_t mz_cap, int32_t mz_asm)
{
    extern __shared__ __align__(16) unsigned char kb_scan_dyn[];  // the four per-warp queues (dynamic: the kernel needs > 48 KB in all)
    ScanQueue *queues = reinterpret_cast<ScanQueue *>(kb_scan_dyn);
    __shared__ int qtail[4];
    __shared__ uint32_t window[4][4 * W * 32];  // bx, by, sx, sy: [slot][lane]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *const bx = &window[warp][lane], *const by = bx + W * 32, *const sx = by + W * 32, *const sy = sx + W * 32;
    ScanQueue &Q = queues[warp];
    int &tail = qtail[warp];
    const int64_t n_warps = (int64_t)gridDim.x * 4;
    unsigned long long n_min_local = 0;
    unsigned n_emit = 0;                    // this lane's regular emissions (all of them, whatever the presence filter says)
    const bool nofilter = mz_hash != nullptr;  // minimizer dump (parity tests): every minimizer has to reach the queue

    for (int64_t chunk = (int64_t)blockIdx.x * 4 + warp; chunk < bt.n_chunks; chunk += n_warps) {
        const int ctg = bt.chunk_ctg[chunk];
        const int cstart = bt.chunk_start[chunk];
        const int clen = bt.ctg_len[ctg];
        const int asm_id = bt.ctg_asm[ctg];
        const int vstart = bt.ctg_vstart[ctg];
        const int64_t soff = bt.ctg_soff[ctg];
        const int lstart = cstart + lane * KB_LANE_BASES;
        int lend = lstart + KB_LANE_BASES;
        if (lend > clen) lend = clen;
        if (lane == 0) tail = 0;
        __syncwarp();

        // ---- the lane's sequence / mask words: 16 bases (32 with the mask) per 32-bit load straight from global memory, one
        // load per 16 (32) steps; gw[-2], gw[-1] and gm[-1] hold the 24-base look-back (the previous lane's last bases, or
        // padding before the first base of a contig, which is replaced by "ambiguous" below)
        const uint32_t *gw = bt.seq2 + ((soff + lstart) >> 4);
        const uint32_t *gm = bt.nmask + ((soff + lstart) >> 5);
        // ---- sketch: all lanes step together; at most one regular minimizer per lane and step goes to the front of
        // the queue through a ballot (no atomics); the rare identical-k-mer emissions of mm_sketch go to the back of
        // the queue through a shared-memory counter.  The queue is drained whenever it holds >= 32 entries.
        KbSketchRegs s;
        s.px = s.py = s.mx = s.my = KB_MAXU, s.pdup = s.sdup = 0, s.fwd = s.rev = 0, s.l = 0;
#pragma unroll
        for (int j = 0; j < W; ++j) bx[j * 32] = by[j * 32] = sx[j * 32] = sy[j * 32] = KB_MAXU;
        const bool active = lstart < lend;
        // every lane runs the same 24 silent look-back steps; before the first base of a contig they see ambiguous bases,
        // which leave the sketch in its reset state (l = 0, empty window)
        const int p0pos = lstart - KB_SCAN_LOOKBACK;
        const int i_end = active ? lend : 0;
        int front = 0;  // warp-uniform: entries pushed at the front of the queue
        auto slow = [&](uint32_t x, uint32_t y) {
            int k = atomicAdd(&tail, 1);  // `tail` counts the slow entries, stored from the back
            if (front + k < KB_QCAP - 64) Q.x[KB_QCAP - 1 - k] = x, Q.y[KB_QCAP - 1 - k] = y;
            else atomicOr(&counters[6], 1ull);
        };
        const int n_iter = (KB_LANE_BASES + KB_SCAN_LOOKBACK + W - 1) / W;
        uint32_t word = 0, mword = 0;  // the lane's current sequence / mask word, shifted down as bases are consumed
        bool pe = false;               // deferred emission: valid, hash, position, bitmap word
        uint32_t px = 0, py = 0, pw = 0;
        for (int it = 0; it <= n_iter; ++it) {
            if (it < n_iter) {
                for (int u = 0; u < W; ++u) {
                    const int step = it * W + u;
                    const int i = p0pos + step;
                    const int r = step + 32 - KB_SCAN_LOOKBACK;  // offset of base i from the start of gw[-2]: the same in every lane
                    if ((r & 15) == 0 || step == 0) word = i < i_end ? __ldg(gw + (r >> 4) - 2) >> (2 * (r & 15)) : 0u;
                    if ((r & 31) == 0 || step == 0) mword = i < i_end ? __ldg(gm + (r >> 5) - 1) >> (r & 31) : 0u;
                    const int c = ((mword & 1u) || i < 0) ? 4 : (int)(word & 3u);
                    word >>= 2, mword >>= 1;
                    uint32_t ex = 0, ey = 0;
                    bool e = false;
                    if (i < i_end) e = kb_fast_step_sm<W, K>(s, bx, by, sx, sy, u, i, c, i >= lstart, &ex, &ey, slow);
                    n_emit += e ? 1u : 0u;
                    // presence filter, one step deferred so that the bitmap load (L2) is not waited for: this step's emission
                    // only issues its load, the previous step's emission is tested and, if the gene index may hold it, queued
                    const bool pass = pe && (nofilter || ((pw >> (px & 31u)) & 1u));
                    const unsigned bal = __ballot_sync(0xffffffffu, pass);
                    if (pass) {
                        const int o = front + __popc(bal & ((1u << lane) - 1u));
                        Q.x[o] = px, Q.y[o] = py;
                    }
                    front += __popc(bal);
                    pe = e, px = ex, py = ey;
                    if (e && !nofilter) pw = __ldg(ix.bloom + ((ex & ix.bloom_mask) >> 5));
                }
                if (it == n_iter - 1) {  // flush the deferred emission of the last step
                    const bool pass = pe && (nofilter || ((pw >> (px & 31u)) & 1u));
                    const unsigned bal = __ballot_sync(0xffffffffu, pass);
                    if (pass) {
                        const int o = front + __popc(bal & ((1u << lane) - 1u));
                        Q.x[o] = px, Q.y[o] = py;
                    }
                    front += __popc(bal);
                    pe = false;
                }
            } else {  // mm_sketch's final push, by the lane that owns the end of the contig
                const bool e = active && lend == clen && s.mx != KB_MAXU;
                const unsigned bal = __ballot_sync(0xffffffffu, e);
                if (e) Q.x[front] = s.mx, Q.y[front] = s.my;  // at most one lane
                front += __popc(bal);
                n_emit += e ? 1u : 0u;
            }
            __syncwarp();
            const int n_slow = tail;
            if (front + n_slow >= 32 || (it == n_iter && front + n_slow > 0)) {
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
            }
        }
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
    int64_t grid = (int64_t)n_sm * 6;  // 6 CTAs of 128 threads (36 KB of shared memory each) per SM, grid-stride over the chunks
    if (grid > want) grid = want;
    cudaFuncSetAttribute(kb_scan_kernel<10, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(ScanQueue)));  // per device
    kb_scan_kernel<10, 15><<<(unsigned)grid, 128, 4 * sizeof(ScanQueue), st>>>(ix, bt, akey, aval, counters, anchor_cap, mz_hash, mz_ctg,
                                                           mz_pos, mz_cap, mz_asm);
}

#endif  // __CUDACC__
