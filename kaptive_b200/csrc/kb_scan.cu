// kb_scan.cu -- the seeding scan: minimizer sketch of every contig + lookup of
// each minimizer in the gene hash + anchor emission, in one pass over the
// 2-bit packed assembly.  This is the HBM-streaming kernel the roofline is
// quoted on (DESIGN.md "Scan kernel").
//
// Work decomposition: one warp per chunk of KB_CHUNK_BASES (8192) bases of one
// contig; lane L owns bases [L*256, (L+1)*256) of the chunk.  Each lane pulls
// its 64 B of 2-bit sequence and 32 B of ambiguity mask with 16-byte loads
// (a warp reads 2 KB + 1 KB contiguous), obtains the 24-base look-back from
// its left neighbour by shuffle, and runs the sketch state machine entirely in
// registers.  Minimizers go to a per-warp shared-memory queue; every time the
// queue holds >= 32 the warp drains it: one hash-table probe per lane (the
// table lives in L2), warp-aggregated allocation of anchor slots (one global
// atomic per drain), cooperative expansion of multi-occurrence minimizers.
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

// Per-lane view of the staged chunk in shared memory.  Lane L owns 18 sequence
// words (2 look-back + 16) at stride 19 and 9 mask words (1 look-back + 8) at
// stride 9: odd strides keep the 32 lanes on 32 different banks.
#define KB_SEQ_STRIDE 19
#define KB_MSK_STRIDE 9
struct LaneFetch {
    const uint32_t *w;  // this lane's 18 sequence words
    const uint32_t *m;  // this lane's 9 mask words
    int base0;          // contig position of bit 0 of w[2] (lane start, multiple of 256)
    __device__ __forceinline__ int operator()(int i) const
    {
        int r = i - base0 + 32;  // >= 8 because i >= base0 - 24
        uint32_t word = w[r >> 4];
        uint32_t mw = m[r >> 5];
        int c = (int)((word >> (2 * (r & 15))) & 3u);
        return ((mw >> (r & 31)) & 1u) ? 4 : c;
    }
};

// W sketch steps (compile-time window slots) with the ballot-based queue push between them.
template <int W, int K, int U, class Slow>
struct KbScanUnroll {
    __device__ __forceinline__ static void run(KbFastSketch<W, K> &s, int i0, int i_end, int live_from, const LaneFetch &F, Slow &slow,
                                               ScanQueue &Q, int &front, int lane)
    {
        const int i = i0 + U;
        uint32_t ex = 0, ey = 0;
        bool e = false;
        if (i < i_end) e = kb_fast_step<W, K, U>(s, i, F(i), i >= live_from, &ex, &ey, slow);
        const unsigned bal = __ballot_sync(0xffffffffu, e);
        if (e) {
            const int o = front + __popc(bal & ((1u << lane) - 1u));
            Q.x[o] = ex, Q.y[o] = ey;
        }
        front += __popc(bal);
        KbScanUnroll<W, K, U + 1, Slow>::run(s, i0, i_end, live_from, F, slow, Q, front, lane);
    }
};
template <int W, int K, class Slow>
struct KbScanUnroll<W, K, W, Slow> {
    __device__ __forceinline__ static void run(KbFastSketch<W, K> &, int, int, int, const LaneFetch &, Slow &, ScanQueue &, int &, int) {}
};

template <int W, int K>
__global__ void __launch_bounds__(128) kb_scan_kernel(KbIndexView ix, KbBatchView bt, uint64_t *akey, uint32_t *aval,
                                                      unsigned long long *counters, int64_t anchor_cap,
                                                      uint32_t *mz_hash, int32_t *mz_ctg, uint32_t *mz_pos,
                                                      int64_t mz_cap, int32_t mz_asm)
{
    __shared__ ScanQueue queues[4];
    __shared__ int qtail[4];
    __shared__ uint32_t stage_seq[4][32 * KB_SEQ_STRIDE];
    __shared__ uint32_t stage_msk[4][32 * KB_MSK_STRIDE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ScanQueue &Q = queues[warp];
    int &tail = qtail[warp];
    const int64_t n_warps = (int64_t)gridDim.x * 4;
    unsigned long long n_min_local = 0;

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

        // ---- loads: 4 x 16 B of sequence + 2 x 16 B of mask per lane, coalesced across the warp
        uint32_t *sw = &stage_seq[warp][lane * KB_SEQ_STRIDE];
        uint32_t *sm = &stage_msk[warp][lane * KB_MSK_STRIDE];
        LaneFetch F;
        F.w = sw, F.m = sm, F.base0 = lstart;
        {
            const uint4 *sp = reinterpret_cast<const uint4 *>(bt.seq2 + ((soff + lstart) >> 4));
            const uint4 *mp = reinterpret_cast<const uint4 *>(bt.nmask + ((soff + lstart) >> 5));
            const bool in = lstart < clen;  // contig storage is padded to 128 bases: a started 64-base (sequence) or 128-base (mask) group is in bounds
            const uint4 z = make_uint4(0, 0, 0, 0);
            uint4 v[4], mv[2];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) v[q4] = (in && lstart + q4 * 64 < clen) ? __ldg(sp + q4) : z;
#pragma unroll
            for (int q2 = 0; q2 < 2; ++q2) mv[q2] = (in && lstart + q2 * 128 < clen) ? __ldg(mp + q2) : z;
            // look-back from the left neighbour (its last two sequence words / last mask word)
            uint32_t p0 = __shfl_up_sync(0xffffffffu, v[3].z, 1);
            uint32_t p1 = __shfl_up_sync(0xffffffffu, v[3].w, 1);
            uint32_t pm = __shfl_up_sync(0xffffffffu, mv[1].w, 1);
            if (lane == 0) {
                if (cstart > 0) {
                    const uint32_t *s1 = bt.seq2 + ((soff + lstart) >> 4);
                    p0 = __ldg(s1 - 2), p1 = __ldg(s1 - 1);
                    pm = __ldg(bt.nmask + ((soff + lstart) >> 5) - 1);
                } else p0 = p1 = pm = 0;
            }
            sw[0] = p0, sw[1] = p1, sm[0] = pm;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
                sw[2 + q4 * 4 + 0] = v[q4].x, sw[2 + q4 * 4 + 1] = v[q4].y, sw[2 + q4 * 4 + 2] = v[q4].z, sw[2 + q4 * 4 + 3] = v[q4].w;
#pragma unroll
            for (int q2 = 0; q2 < 2; ++q2)
                sm[1 + q2 * 4 + 0] = mv[q2].x, sm[1 + q2 * 4 + 1] = mv[q2].y, sm[1 + q2 * 4 + 2] = mv[q2].z, sm[1 + q2 * 4 + 3] = mv[q2].w;
        }
        __syncwarp();

        // ---- sketch: all lanes step together; at most one regular minimizer per lane and step goes to the front of
        // the queue through a ballot (no atomics); the rare identical-k-mer emissions of mm_sketch go to the back of
        // the queue through a shared-memory counter.  The queue is drained whenever it holds >= 32 entries.
        KbFastSketch<W, K> s;
        s.reset();
        const bool active = lstart < lend;
        const int p0pos = lstart >= KB_SCAN_LOOKBACK ? lstart - KB_SCAN_LOOKBACK : 0;
        int front = 0;  // warp-uniform: entries pushed at the front of the queue
        auto slow = [&](uint32_t x, uint32_t y) {
            int k = atomicAdd(&tail, 1);  // `tail` counts the slow entries, stored from the back
            if (front + k < KB_QCAP - 64) Q.x[KB_QCAP - 1 - k] = x, Q.y[KB_QCAP - 1 - k] = y;
            else atomicOr(&counters[6], 1ull);
        };
        const int n_iter = (KB_LANE_BASES + KB_SCAN_LOOKBACK + W - 1) / W;
        for (int it = 0; it <= n_iter; ++it) {
            if (it < n_iter) {
                const int i0 = p0pos + it * W;
                KbScanUnroll<W, K, 0, decltype(slow)>::run(s, i0, active ? lend : 0, lstart, F, slow, Q, front, lane);
            } else {  // mm_sketch's final push, by the lane that owns the end of the contig
                const bool e = active && lend == clen && s.mx != KB_MAXU;
                const unsigned bal = __ballot_sync(0xffffffffu, e);
                if (e) Q.x[front] = s.mx, Q.y[front] = s.my;  // at most one lane
                front += __popc(bal);
            }
            __syncwarp();
            const int n_slow = tail;
            if (front + n_slow >= 32 || (it == n_iter && front + n_slow > 0)) {
                const int n = front + n_slow;
                n_min_local += (lane == 0) ? (unsigned long long)n : 0ull;
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
    if (lane == 0 && n_min_local) atomicAdd(&counters[0], n_min_local);
}

void kb_launch_scan(const KbIndexView &ix, const KbBatchView &bt, uint64_t *akey, uint32_t *aval,
                    unsigned long long *counters, int64_t anchor_cap, uint32_t *mz_hash, int32_t *mz_ctg,
                    uint32_t *mz_pos, int64_t mz_cap, int32_t mz_asm, int n_sm, cudaStream_t st)
{
    if (bt.n_chunks == 0) return;
    int64_t want = (bt.n_chunks + 3) / 4;
    int64_t grid = (int64_t)n_sm * 8;  // 8 CTAs of 128 threads per SM: 32 warps resident
    if (grid > want) grid = want;
    kb_scan_kernel<10, 15><<<(unsigned)grid, 128, 0, st>>>(ix, bt, akey, aval, counters, anchor_cap, mz_hash, mz_ctg,
                                                           mz_pos, mz_cap, mz_asm);
}

#endif  // __CUDACC__
