// kb_pipeline.cu -- kernels after the seeding scan: ASCII packing, group discovery, chaining
// (one thread per query group), base-level alignment (one warp per chain, persistent,
// dynamic work queue), per-query finalisation and the SoA scatter.  Sorting / selection /
// prefix sums between the stages are CUB device primitives (plumbing, like cuBLAS would be
// for a GEMM); every stage that carries the path's arithmetic is the hand-written logic in
// kb_scan.cuh / kb_chain.cuh / kb_align.cuh / kb_final.cuh.
#include <cub/cub.cuh>
// DP statistics (debug / profiling aid): calls and cells per (kind, path); kind 0 = gap fill, 1 = end extension,
// 2 = gap fill redone with z-drop; path 0 = band certified, 1 = band rejected, 2 = register single pass,
// 3 = register tiled, 4 = scratch-memory DP.  Slot = 2 * (5 * kind + path) (+1 for cells).
__device__ unsigned long long g_kb_dp_stats[32];
#define KB_DP_STAT(kind, path, cells)                                                       \
    do {                                                                                    \
        atomicAdd(&g_kb_dp_stats[2 * (5 * (kind) + (path))], 1ull);                         \
        atomicAdd(&g_kb_dp_stats[2 * (5 * (kind) + (path)) + 1], (unsigned long long)(cells)); \
    } while (0)
#define KB_DP_STAT_RAW(slot, v) atomicAdd(&g_kb_dp_stats[slot], (unsigned long long)(v))
#include "kb_final.cuh"
#include "kb_stage.cuh"
#include "kb_kernels.h"

// ------------------------------------------------------------------ pack: ASCII -> 2 bit + N mask
// One thread per 32 bases of storage (2 sequence words + 1 mask word).  Storage of a contig is
// padded to 128 bases; padding is marked ambiguous.
__global__ void kb_pack_kernel(const uint8_t *ascii, int64_t ascii_base, const int64_t *ctg_off, const int64_t *ctg_soff,
                               const int32_t *ctg_len, int32_t ctg_begin, int32_t ctg_end, int64_t grp_begin, int64_t grp_end,
                               uint32_t *seq2, uint32_t *nmask)
{
    for (int64_t g = grp_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < grp_end; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b0 = g << 5;
        // contig whose storage contains b0: last c in [ctg_begin, ctg_end) with soff[c] <= b0
        int lo = ctg_begin, hi = ctg_end - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (ctg_soff[mid] <= b0) lo = mid;
            else hi = mid - 1;
        }
        uint32_t w0 = 0, w1 = 0, m = 0xffffffffu;
        const int64_t rel = b0 - ctg_soff[lo];
        const int32_t L = ctg_len[lo];
        if (rel >= 0 && rel < L) {
            const uint8_t *s = ascii + (ctg_off[lo] - ascii_base) + rel;
            int n = L - rel < 32 ? (int)(L - rel) : 32;
            for (int i = 0; i < n; ++i) {
                uint32_t c = kb_nt4(__ldg(s + i));
                if (c < 4) {
                    m &= ~(1u << i);
                    if (i < 16) w0 |= c << (2 * i);
                    else w1 |= c << (2 * (i - 16));
                }
            }
        }
        seq2[2 * g] = w0, seq2[2 * g + 1] = w1, nmask[g] = m;
    }
}

void kb_launch_pack(const uint8_t *ascii, int64_t ascii_base, const int64_t *ctg_off, const int64_t *ctg_soff, const int32_t *ctg_len,
                    int32_t ctg_begin, int32_t ctg_end, int64_t grp_begin, int64_t grp_end, uint32_t *seq2, uint32_t *nmask,
                    cudaStream_t st)
{
    int64_t n = grp_end - grp_begin;
    if (n <= 0) return;
    int64_t grid = (n + 255) / 256;
    if (grid > 148 * 64) grid = 148 * 64;
    kb_pack_kernel<<<(unsigned)grid, 256, 0, st>>>(ascii, ascii_base, ctg_off, ctg_soff, ctg_len, ctg_begin, ctg_end, grp_begin, grp_end, seq2, nmask);
}

// ------------------------------------------------------------------ occurrence counts
// occ[asm][entry] += 1 for every anchor; flags assemblies in which some gene minimizer occurs more
// than min_mid_occ times (only those can be affected by the exact value of mid_occ).
__global__ void kb_occ_kernel(const uint64_t *akey, const uint32_t *aval, int64_t n, int64_t n_entries, uint32_t *occ32,
                              int32_t min_mid_occ, int32_t *need_census)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t a = (int64_t)(akey[i] >> KB_KEY_ASM_SHIFT);
        int64_t idx = a * n_entries + aval[i];
        int sh = (int)(idx & 1) * 16;
        uint32_t old = atomicAdd(&occ32[idx >> 1], 1u << sh);
        uint32_t c = ((old >> sh) & 0xffffu) + 1;
        if (c == (uint32_t)min_mid_occ + 1) need_census[a] = 1;
        if (c >= 0xfff0u) need_census[a] = 2;  // 16-bit counter about to wrap: reported as a limit error
    }
}

void kb_launch_occ(const uint64_t *akey, const uint32_t *aval, int64_t n, int64_t n_entries, uint32_t *occ32, int32_t min_mid_occ,
                   int32_t *need_census, cudaStream_t st)
{
    if (n <= 0) return;
    int64_t grid = (n + 255) / 256;
    if (grid > 148 * 32) grid = 148 * 32;
    kb_occ_kernel<<<(unsigned)grid, 256, 0, st>>>(akey, aval, n, n_entries, occ32, min_mid_occ, need_census);
}

// ------------------------------------------------------------------ group discovery
__global__ void kb_group_flag_kernel(const uint64_t *skey, int64_t n, uint8_t *flag)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        flag[i] = (i == 0 || (skey[i] >> KB_KEY_GENE_SHIFT) != (skey[i - 1] >> KB_KEY_GENE_SHIFT)) ? 1 : 0;
}

size_t kb_sort_pairs_temp_bytes(int64_t n)
{
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, n, 0, 64);
    return b;
}
cudaError_t kb_sort_pairs(void *tmp, size_t tmp_bytes, const uint64_t *kin, uint64_t *kout, const uint32_t *vin, uint32_t *vout,
                          int64_t n, int end_bit, cudaStream_t st)
{
    return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kin, kout, vin, vout, n, 0, end_bit, st);
}
size_t kb_sort_keys32_temp_bytes(int64_t n)
{
    size_t b = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, b, (const uint32_t *)nullptr, (uint32_t *)nullptr, n, 0, 32);
    return b;
}
cudaError_t kb_sort_keys32(void *tmp, size_t tmp_bytes, const uint32_t *kin, uint32_t *kout, int64_t n, int end_bit, cudaStream_t st)
{
    return cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, kin, kout, n, 0, end_bit, st);
}
size_t kb_select_temp_bytes(int64_t n)
{
    size_t b = 0;
    cub::CountingInputIterator<int64_t> it(0);
    cub::DeviceSelect::Flagged(nullptr, b, it, (const uint8_t *)nullptr, (int64_t *)nullptr, (int64_t *)nullptr, n);
    return b;
}
cudaError_t kb_select_flagged(void *tmp, size_t tmp_bytes, const uint8_t *flag, int64_t *out, int64_t *n_out, int64_t n, cudaStream_t st)
{
    cub::CountingInputIterator<int64_t> it(0);
    return cub::DeviceSelect::Flagged(tmp, tmp_bytes, it, flag, out, n_out, n, st);
}
size_t kb_scan_temp_bytes(int64_t n)
{
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int32_t *)nullptr, (int64_t *)nullptr, n);
    return b;
}
cudaError_t kb_exclusive_sum(void *tmp, size_t tmp_bytes, const int32_t *in, int64_t *out, int64_t n, cudaStream_t st)
{
    return cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, n, st);
}
size_t kb_rle_temp_bytes(int64_t n)
{
    size_t b = 0;
    cub::DeviceRunLengthEncode::Encode(nullptr, b, (const uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (int64_t *)nullptr, n);
    return b;
}
cudaError_t kb_rle(void *tmp, size_t tmp_bytes, const uint32_t *in, uint32_t *uniq, uint32_t *counts, int64_t *n_runs, int64_t n, cudaStream_t st)
{
    return cub::DeviceRunLengthEncode::Encode(tmp, tmp_bytes, in, uniq, counts, n_runs, n, st);
}

// ------------------------------------------------------------------ occurrence census of many assemblies at once
// [mm2:index.c:mm_idx_cal_max_occ] per assembly: the (1 - f) quantile of the occurrence counts of its distinct minimizers.  Keys
// asm << 32 | hash are sorted and run-length encoded, the runs re-keyed asm << 32 | count and sorted again; one thread per assembly
// then reads its quantile.
size_t kb_sort_keys64_temp_bytes(int64_t n)
{
    size_t b = 0, b2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, b, (const uint64_t *)nullptr, (uint64_t *)nullptr, n, 0, 64);
    cub::DeviceRunLengthEncode::Encode(nullptr, b2, (const uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr, (int64_t *)nullptr, n);
    return b > b2 ? b : b2;
}
cudaError_t kb_sort_keys64(void *tmp, size_t tmp_bytes, const uint64_t *kin, uint64_t *kout, int64_t n, int end_bit, cudaStream_t st)
{
    return cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, kin, kout, n, 0, end_bit, st);
}
cudaError_t kb_rle64(void *tmp, size_t tmp_bytes, const uint64_t *in, uint64_t *uniq, uint32_t *counts, int64_t *n_runs, int64_t n, cudaStream_t st)
{
    return cub::DeviceRunLengthEncode::Encode(tmp, tmp_bytes, in, uniq, counts, n_runs, n, st);
}
// key = (assembly - a_lo) << 30 | hash: as few radix passes as the group needs
__global__ void kb_census_key_kernel(const uint32_t *hash, const int32_t *asm_id, int64_t n, int32_t a_lo, uint64_t *key)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        key[i] = (uint64_t)(uint32_t)(asm_id[i] - a_lo) << 30 | hash[i];
}
// histogram of the occurrence counts >= 2 per assembly (counts of 1, nearly all of them, are what is left of the assembly's runs);
// hist: n_group x KB_CENSUS_BINS, the last bin collects every count that does not fit
#define KB_CENSUS_BINS 65536
__global__ void kb_census_hist_kernel(const uint64_t *uniq, const uint32_t *counts, const int64_t *n_runs, uint32_t *hist)
{
    const int64_t n = *n_runs;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t c = counts[i];
        if (c >= 2) atomicAdd(&hist[(uniq[i] >> 30) * KB_CENSUS_BINS + (c < KB_CENSUS_BINS - 1 ? c : KB_CENSUS_BINS - 1)], 1u);
    }
}
// one WARP per listed assembly: its runs are uniq[b, e) (binary search on the assembly field), the quantile is read off the histogram
__global__ void kb_census_quantile_kernel(const uint64_t *uniq, const int64_t *n_runs_p, const uint32_t *hist, const int32_t *asm_list, int n_list,
                                          int32_t a_lo, float mid_occ_frac, int32_t min_mid_occ, int32_t max_mid_occ, int32_t *mid_occ,
                                          unsigned long long *overflow)
{
    const int lane = threadIdx.x & 31, i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n_list) return;
    const int32_t a = asm_list[i];
    const int64_t n_runs = *n_runs_p;
    auto lower = [&](uint64_t v) {  // first run with key >= v
        int64_t lo = 0, hi = n_runs;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (uniq[mid] < v) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    };
    const int64_t b = lower((uint64_t)(uint32_t)(a - a_lo) << 30), e = lower((uint64_t)((uint32_t)(a - a_lo) + 1u) << 30), n = e - b;
    int32_t mid = INT32_MAX;  // f <= 0 or no minimizers: no cut-off
    if (mid_occ_frac > 0.f && n > 0) {
        int64_t kth = (int64_t)((1. - mid_occ_frac) * (double)n);
        kth = kth > n - 1 ? n - 1 : (kth < 0 ? 0 : kth);
        const uint32_t *h = hist + (int64_t)(a - a_lo) * KB_CENSUS_BINS;
        // runs above the quantile: n - 1 - kth of them.  Walk the histogram from the top until more than that many runs are passed.
        const int64_t above = n - 1 - kth;
        int64_t passed = 0;
        int32_t val = 1;  // every bin from 2 up holds too few runs: the quantile is a count of 1
        for (int top = KB_CENSUS_BINS - 1; top >= 2 && val == 1; top -= 32) {
            const int v = top - lane;
            const uint32_t c = v >= 2 ? h[v] : 0u;
            // inclusive prefix over lanes (lane 0 = highest count)
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const unsigned hit = __ballot_sync(0xffffffffu, passed + (int64_t)incl > above);
            if (hit) {
                const int l = __ffs(hit) - 1;
                val = top - l;
                if (val == KB_CENSUS_BINS - 1 && lane == 0) atomicAdd(overflow, 1ull);  // the quantile itself is a count beyond the histogram
            }
            passed += (int64_t)__shfl_sync(0xffffffffu, incl, 31);
        }
        mid = val + 1;
    }
    if (mid < min_mid_occ) mid = min_mid_occ;
    if (max_mid_occ > min_mid_occ && mid > max_mid_occ) mid = max_mid_occ;
    if (lane == 0) mid_occ[a] = mid;
}
void kb_launch_census_keys(const uint32_t *hash, const int32_t *asm_id, int64_t n, int32_t a_lo, uint64_t *key, cudaStream_t st)
{
    if (n > 0) kb_census_key_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(hash, asm_id, n, a_lo, key);
}
size_t kb_census_hist_bytes(int n_group) { return (size_t)n_group * KB_CENSUS_BINS * 4; }
// uniq / counts / n_runs: the run-length encoding of the sorted keys (n_runs on the device); overflow: a device counter
void kb_launch_census_quantile(const uint64_t *uniq, const uint32_t *counts, const int64_t *n_runs, int64_t max_runs, uint32_t *hist, int n_group,
                               const int32_t *asm_list, int n_list, int32_t a_lo, float mid_occ_frac, int32_t min_mid_occ, int32_t max_mid_occ,
                               int32_t *mid_occ, unsigned long long *overflow, cudaStream_t st)
{
    cudaMemsetAsync(hist, 0, kb_census_hist_bytes(n_group), st);
    if (max_runs > 0) kb_census_hist_kernel<<<(unsigned)std::min<int64_t>((max_runs + 255) / 256, 148 * 16), 256, 0, st>>>(uniq, counts, n_runs, hist);
    if (n_list > 0)
        kb_census_quantile_kernel<<<(unsigned)((n_list * 32 + 127) / 128), 128, 0, st>>>(uniq, n_runs, hist, asm_list, n_list, a_lo, mid_occ_frac,
                                                                                        min_mid_occ, max_mid_occ, mid_occ, overflow);
}

// ------------------------------------------------------------------ chaining: one thread per group
__global__ void __launch_bounds__(128) kb_chain_kernel(KbIndexView ix, KbBatchView bt, const uint64_t *skey, const uint32_t *sval,
                                                       const int64_t *gstart, int64_t n_groups, int64_t n_anchors,
                                                       const uint16_t *occ, const int32_t *mid_occ, KbChainWork W, uint64_t *cx,
                                                       uint64_t *cy, KbGroupInfo *ginfo, KbChainRec *chains,
                                                       unsigned long long *counters, int64_t chain_cap, const int32_t *occ_skip)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    int64_t gs = gstart[g], ge = g + 1 < n_groups ? gstart[g + 1] : n_anchors;
    kb_chain_group(ix, bt, skey, sval, gs, ge, occ, mid_occ, W, cx, cy, &ginfo[g], chains, &counters[3], chain_cap, (int32_t)g, occ_skip);
}

void kb_launch_chain(const KbIndexView &ix, const KbBatchView &bt, const uint64_t *skey, const uint32_t *sval, const int64_t *gstart,
                     int64_t n_groups, int64_t n_anchors, const uint16_t *occ, const int32_t *mid_occ, const KbChainWork &W,
                     uint64_t *cx, uint64_t *cy, KbGroupInfo *ginfo, KbChainRec *chains, unsigned long long *counters,
                     int64_t chain_cap, const int32_t *occ_skip, cudaStream_t st)
{
    if (n_groups <= 0) return;
    kb_chain_kernel<<<(unsigned)((n_groups + 127) / 128), 128, 0, st>>>(ix, bt, skey, sval, gstart, n_groups, n_anchors, occ, mid_occ,
                                                                        W, cx, cy, ginfo, chains, counters, chain_cap, occ_skip);
}

// ------------------------------------------------------------------ alignment: one warp per chain, persistent
#ifndef KB_ALIGN_MINB
#define KB_ALIGN_MINB 3
#endif
__global__ void __launch_bounds__(128, KB_ALIGN_MINB) kb_align_kernel(KbIndexView ix, KbBatchView bt, const KbChainRec *chains, int64_t n_chains,
                                                       const KbGroupInfo *ginfo, const uint64_t *cx, uint64_t *cy, uint8_t *scratch,
                                                       size_t scratch_bytes, KbRawHit *raw, int64_t raw_cap, uint32_t *pool,
                                                       int64_t pool_cap, unsigned long long *counters, unsigned long long *next_chain,
                                                       const int32_t *list, const unsigned long long *n_list)
{
    if (list) n_chains = (int64_t)*n_list;  // only the chains the staged path handed back
    const int lane = threadIdx.x & 31;
    const int64_t wg = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    __shared__ uint32_t wmax_ring[4][KB_RING_WORDS];
    KbAlignScratch S = kb_align_scratch_at(scratch + (size_t)wg * scratch_bytes, ix.p.max_sw_cells);
    S.wmax = wmax_ring[threadIdx.x >> 5];
    for (int x = lane; x < KB_RING_WORDS; x += 32) S.wmax[x] = 0;
    __syncwarp();
    int64_t cells = 0;
    for (;;) {
        unsigned long long ci = 0;
        if (lane == 0) ci = atomicAdd(next_chain, 1ull);
        ci = __shfl_sync(0xffffffffu, ci, 0);
        if ((int64_t)ci >= n_chains) break;
        if (list) ci = (unsigned long long)list[ci];
        const KbChainRec c = chains[ci];
        const KbGroupInfo gi = ginfo[c.group];
        KbReg r;
        const int reg_idx = (int)((int64_t)ci - gi.chain_base);
        r.as = c.as, r.cnt = c.cnt, r.score = c.score, r.score0 = c.score0, r.mlen = c.mlen, r.blen = c.blen, r.parent = c.parent, r.id = reg_idx;
        r.hash = c.hash, r.rev = c.rev, r.rid = c.rid, r.rs = c.rs, r.re = c.re, r.qs = c.qs, r.qe = c.qe;
        r.has_p = 0, r.dp_score = 0, r.dp_max = 0, r.n_ambi = 0, r.n_cigar = 0;
        for (int split = 0;; ++split) {
            KbReg r2;
            r2.cnt = 0;
            int e = kb_align1<32>(ix, bt, lane, gi.asm_id, gi.gene, r, r2, gi.n_a, cx + gi.a_base, cy + gi.a_base, S, &cells);
            const int ncg = e ? 0 : r.n_cigar;
            unsigned long long slot = 0, coff = 0;
            if (lane == 0) {
                slot = atomicAdd(&counters[4], 1ull);
                coff = atomicAdd(&counters[5], (unsigned long long)ncg);
            }
            slot = __shfl_sync(0xffffffffu, slot, 0);
            coff = __shfl_sync(0xffffffffu, coff, 0);
            if ((int64_t)slot < raw_cap && lane == 0) {
                KbRawHit h;
                h.group = c.group, h.reg_idx = reg_idx, h.split_idx = split;
                h.cnt = r.cnt, h.score = r.score, h.score0 = r.score0, h.hash = r.hash;
                h.rev = r.rev, h.rid = r.rid, h.rs = r.rs, h.re = r.re, h.qs = r.qs, h.qe = r.qe;
                h.has_p = r.has_p, h.dp_score = r.dp_score, h.dp_max = r.dp_max, h.dp_max2 = 0, h.n_ambi = r.n_ambi, h.mlen = r.mlen, h.blen = r.blen;
                h.parent = r.parent, h.subsc = c.subsc, h.n_sub = c.n_sub, h.mapq = 0, h.n_cigar = ncg, h.cigar_off = (int64_t)coff;
                h.err = e, h.pad = 0;
                raw[slot] = h;
            }
            if ((int64_t)(coff + ncg) <= pool_cap)
                for (int i = lane; i < ncg; i += 32) pool[coff + i] = S.cigar[i];
            __syncwarp();
            if (e == 0 && r2.cnt > 0) r = r2;
            else break;
        }
    }
    if (lane == 0 && cells) atomicAdd(&counters[8], (unsigned long long)cells);
}

void kb_launch_align(const KbIndexView &ix, const KbBatchView &bt, const KbChainRec *chains, int64_t n_chains, const KbGroupInfo *ginfo,
                     const uint64_t *cx, uint64_t *cy, uint8_t *scratch, size_t scratch_bytes, int n_warps, KbRawHit *raw,
                     int64_t raw_cap, uint32_t *pool, int64_t pool_cap, unsigned long long *counters, unsigned long long *next_chain,
                     const int32_t *list, const unsigned long long *n_list, cudaStream_t st)
{
    if (n_chains <= 0) return;
    kb_align_kernel<<<(unsigned)(n_warps / 4), 128, 0, st>>>(ix, bt, chains, n_chains, ginfo, cx, cy, scratch, scratch_bytes, raw, raw_cap,
                                                             pool, pool_cap, counters, next_chain, list, n_list);
}

// ------------------------------------------------------------------ staged alignment (kb_stage.cuh)
// counters used by the staged path (d_counters + KB_SC_*)
#define KB_SC_JOBS 9
#define KB_SC_BAND 10
#define KB_SC_ROWS 11
#define KB_SC_SLOW 12
#define KB_SC_JOBCIG 13
#define KB_SC_TMPCIG 14
#define KB_SC_QBAND 32
#define KB_SC_QROWS 33

// The rows kernel's queue has two ends: rectangles of at least big_thr cells are stored from the back of the list and taken
// first (longest jobs first keeps the tail of the persistent kernel short), the others from the front in planning order.
#define KB_SC_ROWS_BIG 34
// the packed 16-bit wavefront (kb_rows16) has its own queue and kernel: rectangles it is eligible for go there
#define KB_SC_R16 35
#define KB_SC_R16_BIG 36
#define KB_SC_QR16 37
// the packed 16-bit band pass (kb_band16): one queue per window width K = 1, 2, 4 (+0..2: jobs queued, +3..5: queue cursors), each sorted
// by length before its kernel runs; what a pass rejects goes to the next wider queue or to the rows queues
#define KB_SC_B16 40
struct KbBandQueues {
    int32_t *list[3];
    uint32_t *key[3];   // qlen + tlen of every queued job (> 0; unused slots hold 0)
    int32_t *band32;    // the 32-bit band kernel's list: fills outside kb_band16's range, pairs with an ambiguous target base
    int use16;
};
__device__ __forceinline__ void kb_band16_enqueue(const KbBandQueues &BQ, unsigned long long *counters, int stage, const KbJob &J, int32_t jid)
{
    const unsigned long long k = atomicAdd(&counters[KB_SC_B16 + stage], 1ull);
    BQ.list[stage][k] = jid, BQ.key[stage][k] = (uint32_t)(J.qlen + J.tlen);
}
struct KbRowsQueues {
    int32_t *rows_list, *r16_list;
    uint32_t *r16_key;  // size class of every job in r16_list: the list is sorted by it (descending) and neighbours are paired
    int64_t job_cap, big_thr;
    int use16;
};
__device__ __forceinline__ void kb_rows_enqueue(const KbRowsQueues &Q, const KbDpConst &P, unsigned long long *counters, const KbJob &J, int32_t jid)
{
    const bool track = !(J.flag & KB_EZ_GLOBAL_NO_ZDROP);
    if (Q.use16 && kb_rows16_eligible(P, J.qlen, J.tlen, J.w, track)) {
        const unsigned long long k = atomicAdd(&counters[KB_SC_R16], 1ull);
        Q.r16_list[k] = jid;
        Q.r16_key[k] = (track ? 1u << 28 : 0u) | (uint32_t)J.qlen << 14 | (uint32_t)J.tlen;  // > 0; unused slots hold 0
    } else if (Q.big_thr > 0 && (int64_t)J.qlen * J.tlen >= Q.big_thr) Q.rows_list[Q.job_cap - 1 - (int64_t)atomicAdd(&counters[KB_SC_ROWS_BIG], 1ull)] = jid;
    else Q.rows_list[atomicAdd(&counters[KB_SC_ROWS], 1ull)] = jid;
}
__global__ void __launch_bounds__(128) kb_plan_kernel(KbIndexView ix, KbBatchView bt, const KbChainRec *chains, int64_t n_chains,
                                                      const KbGroupInfo *ginfo, const uint64_t *cx, uint64_t *cy, int32_t *kscratch,
                                                      KbPlan *plans, KbJob *jobs, int64_t job_cap, KbBandQueues BQ, KbRowsQueues Q,
                                                      int32_t *slow_list, unsigned long long *counters)
{
    const int64_t ci = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= n_chains) return;
    const KbChainRec c = chains[ci];
    const KbGroupInfo gi = ginfo[c.group];
    const uint64_t *ax = cx + gi.a_base;
    uint64_t *ay = cy + gi.a_base;
    int32_t *K = kscratch + gi.a_base + c.as;  // the chain's own anchors index a private stretch of the scratch
    KbPlan pl;
    KbJobCount count;
    int nj = kb_stage_plan(ix, bt, gi.asm_id, gi.gene, c.as, c.cnt, c.mlen, gi.n_a, ax, ay, K, pl, count);
    long long base = -1;
    if (nj >= 0) {
        base = (long long)atomicAdd(&counters[KB_SC_JOBS], (unsigned long long)nj);
        if (base + nj > job_cap) nj = -1;
    }
    if (nj < 0) {
        pl.ok = 0;
        plans[ci] = pl;
        slow_list[atomicAdd(&counters[KB_SC_SLOW], 1ull)] = (int32_t)ci;
        return;
    }
    KbJobWrite write{jobs + base, (int32_t)ci};
    kb_stage_plan(ix, bt, gi.asm_id, gi.gene, c.as, c.cnt, c.mlen, gi.n_a, ax, ay, K, pl, write);
    pl.job_base = base;
    plans[ci] = pl;
    const KbDpConst P = kb_dp_const(ix.p);
    for (int k = 0; k < nj; ++k) {
        const KbJob &J = jobs[base + k];
        const bool band = J.kind == KB_JOB_FILL && J.qlen + J.tlen <= 8184 && kb_band_eligible(ix.p.max_sw_cells, J.qlen, J.tlen, J.w, J.flag);
        if (band && BQ.use16 && kb_band16_eligible(P, J.qlen, J.tlen)) kb_band16_enqueue(BQ, counters, 0, J, (int32_t)(base + k));
        else if (band) BQ.band32[atomicAdd(&counters[KB_SC_BAND], 1ull)] = (int32_t)(base + k);
        else kb_rows_enqueue(Q, P, counters, J, (int32_t)(base + k));
    }
}

// results of one DP job -> its record, CIGAR into the job pool (all lanes copy)
static __device__ __forceinline__ void kb_job_finish(int lane, KbJob *J, const KbEz &ez, const uint32_t *ezcig, uint32_t *jobcig,
                                                     int64_t jobcig_cap, unsigned long long *counters)
{
    int n = ez.n_cigar;
    unsigned long long off = 0;
    if (lane == 0 && n > 0) off = atomicAdd(&counters[KB_SC_JOBCIG], (unsigned long long)n);
    off = __shfl_sync(0xffffffffu, off, 0);
    if (n > 0) {
        if ((int64_t)(off + n) > jobcig_cap) n = -2;  // no room: the chain goes to kb_align1
        else
            for (int i = lane; i < n; i += 32) jobcig[off + i] = ezcig[i];
    }
    __syncwarp();
    if (lane == 0) {
        J->score = ez.score, J->max = ez.max, J->max_t = ez.max_t, J->max_q = ez.max_q, J->zdropped = ez.zdropped;
        J->n_cigar = n, J->cigar_off = (int64_t)off, J->state = 1;
    }
}

// certified band pass over the gap fills: one warp per job, persistent, dynamic queue
#ifndef KB_BAND_MINB
#define KB_BAND_MINB 6
#endif
__global__ void __launch_bounds__(128, KB_BAND_MINB) kb_band_kernel(KbIndexView ix, KbBatchView bt, KbJob *jobs, const int32_t *band_list, KbRowsQueues Q,
                                                        uint8_t *scratch, size_t scratch_bytes, uint32_t *jobcig, int64_t jobcig_cap,
                                                        unsigned long long *counters)
{
    const int lane = threadIdx.x & 31;
    const int64_t wg = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    KbAlignScratch S;
    memset(&S, 0, sizeof(S));
    S.ezcig = reinterpret_cast<uint32_t *>(scratch + (size_t)wg * scratch_bytes);
    S.tb = reinterpret_cast<uint8_t *>(S.ezcig + KB_CIG_MAX);
    const KbDpConst P = kb_dp_const(ix.p);
    const long long n = (long long)counters[KB_SC_BAND];
    int64_t cells = 0;
    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(&counters[KB_SC_QBAND], 1ull);
        k = __shfl_sync(0xffffffffu, k, 0);
        if ((long long)k >= n) break;
        const int jid = band_list[k];
        KbJob *J = jobs + jid;
        const KbDirBytes sq{(J->qrev ? ix.gseq_rev : ix.gseq_fwd) + J->qbase + J->qoff, 1};
        const KbDirPack st{bt.seq2, bt.nmask, J->tpos, 1};
        KbEz ez;
        int ok = kb_global_band(P, lane, J->qlen, sq, J->tlen, st, J->flag, ez, S, &cells);
        if (lane == 0) KB_DP_STAT(0, ok ? 0 : 1, (int64_t)32 * (J->qlen + J->tlen + 1));
        if (!ok && ez.score > KB_NEG_INF) {  // the smallest wider band this score already certifies, if any
            int dlo, dhi, k2 = 0;
            if (kb_band_geometry(J->qlen, J->tlen, 2, dlo, dhi) && ez.score > kb_band_bound(P, J->qlen, J->tlen, dlo, dhi)) k2 = 2;
            else if (kb_band_geometry(J->qlen, J->tlen, 4, dlo, dhi) && ez.score > kb_band_bound(P, J->qlen, J->tlen, dlo, dhi)) k2 = 4;
            // worth it only while the window stays well below the rectangle
            if (k2 && 64 * k2 * 3 <= 2 * (J->qlen + J->tlen)) {
                ok = k2 == 2 ? kb_global_bandK<2>(P, lane, J->qlen, sq, J->tlen, st, J->flag, ez, S, &cells)
                             : kb_global_bandK<4>(P, lane, J->qlen, sq, J->tlen, st, J->flag, ez, S, &cells);
                if (lane == 0) KB_DP_STAT(2, ok ? 0 : 1, (int64_t)32 * k2 * (J->qlen + J->tlen + 1));
                // the first pass may have handed back an extrapolated score: a rejected 128-diagonal pass gets one more try
                if (!ok && k2 == 2 && 64 * 4 * 3 <= 2 * (J->qlen + J->tlen) && kb_band_geometry(J->qlen, J->tlen, 4, dlo, dhi) &&
                    ez.score > kb_band_bound(P, J->qlen, J->tlen, dlo, dhi)) {
                    ok = kb_global_bandK<4>(P, lane, J->qlen, sq, J->tlen, st, J->flag, ez, S, &cells);
                    if (lane == 0) KB_DP_STAT(2, ok ? 0 : 1, (int64_t)32 * 4 * (J->qlen + J->tlen + 1));
                }
            }
        }
        if (ok) kb_job_finish(lane, J, ez, S.ezcig, jobcig, jobcig_cap, counters);
        else if (lane == 0) kb_rows_enqueue(Q, P, counters, *J, jid);
    }
    if (lane == 0 && cells) atomicAdd(&counters[8], (unsigned long long)cells);
}


// packed certified band pass over the gap fills: one warp per PAIR of jobs (neighbours of the length-sorted queue), persistent
#ifndef KB_BAND16_MINB
#define KB_BAND16_MINB 6  // 80 registers, 24 warps per SM: align 343.8 / 343.1 / 340.2 ms per 2000 assemblies at 4 / 5 / 6 CTAs per SM
#endif
// a wider window is taken while 64 K * KB_B16_WIN_RULE <= 2 (qlen + tlen); measured flat for 1 .. 4 (336.5 - 337.6 ms align per 2000
// assemblies), worse above (5: 343.5, 6: 352.2, 8: 368.5): the window pass is cheaper per cell than the rectangle even where it has more cells
#ifndef KB_B16_WIN_RULE
#define KB_B16_WIN_RULE 3
#endif
template <int K>
__global__ void __launch_bounds__(128, KB_BAND16_MINB) kb_band16_kernel(KbIndexView ix, KbBatchView bt, KbJob *jobs, const int32_t *sorted, KbBandQueues BQ,
                                                                         KbRowsQueues Q, uint8_t *scratch, size_t scratch_bytes, uint32_t *jobcig,
                                                                         int64_t jobcig_cap, unsigned long long *counters)
{
    constexpr int STAGE = K == 1 ? 0 : (K == 2 ? 1 : 2);
    __shared__ __align__(8) uint8_t sm_stage[4][KB_B16_SMEM_BYTES];
    __shared__ uint2 sm_lut[25];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int64_t wg = (int64_t)blockIdx.x * (blockDim.x >> 5) + wi;
    const KbDpConst P = kb_dp_const(ix.p);
    if (threadIdx.x < 25) sm_lut[threadIdx.x] = make_uint2(kb_qrow16(P, 7, threadIdx.x % 5), kb_qrow16(P, 7, threadIdx.x / 5));
    __syncthreads();
    uint32_t *ezcig = reinterpret_cast<uint32_t *>(scratch + (size_t)wg * scratch_bytes);
    uint32_t *tbA = ezcig + KB_CIG_MAX, *tbB = tbA + ((scratch_bytes - (size_t)KB_CIG_MAX * 4) / 8);
    uint16_t *st_sel = reinterpret_cast<uint16_t *>(sm_stage[wi]);
    uint8_t *st_q = sm_stage[wi] + 2 * KB_B16_X;
    const unsigned a_sel = (unsigned)__cvta_generic_to_shared(st_sel), a_q = (unsigned)__cvta_generic_to_shared(st_q),
                   a_lut = (unsigned)__cvta_generic_to_shared(sm_lut);
    const long long n = (long long)counters[KB_SC_B16 + STAGE], n_pairs = (n + 1) >> 1;
    int64_t cells = 0;
    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(&counters[KB_SC_B16 + 3 + STAGE], 1ull);
        k = __shfl_sync(0xffffffffu, k, 0);
        if ((long long)k >= n_pairs) break;
        const int jid[2] = {sorted[2 * k], (long long)(2 * k + 1) < n ? sorted[2 * k + 1] : -1};
        KbJob *J[2] = {jobs + jid[0], jid[1] >= 0 ? jobs + jid[1] : jobs + jid[0]};
        const bool haveB = jid[1] >= 0;
        KbB16Job V[2];
        for (int w = 0; w < 2; ++w) V[w] = kb_b16_job(P, K, J[w]->qlen, J[w]->tlen);
        const KbDirBytes qsA{(J[0]->qrev ? ix.gseq_rev : ix.gseq_fwd) + J[0]->qbase + J[0]->qoff, 1},
            qsB{(J[1]->qrev ? ix.gseq_rev : ix.gseq_fwd) + J[1]->qbase + J[1]->qoff, 1};
        const KbDirPack tsA{bt.seq2, bt.nmask, J[0]->tpos, 1}, tsB{bt.seq2, bt.nmask, J[1]->tpos, 1};
        __syncwarp();
        if (kb_band16_stage(lane, K, V[0], qsA, tsA, V[1], qsB, tsB, st_sel, st_q)) {  // an ambiguous target base: the 32-bit kernel takes both
            if (lane == 0)
                for (int w = 0; w < (haveB ? 2 : 1); ++w) BQ.band32[atomicAdd(&counters[KB_SC_BAND], 1ull)] = jid[w];
            continue;
        }
        KbB16Out out;
        kb_band16_pass<K>(P, lane, V[0], V[1], haveB, a_sel, a_q, a_lut, tbA, tbB, out);
        __syncwarp();
        for (int w = 0; w < (haveB ? 2 : 1); ++w) {
            const int score = out.score[w];
            int ok = out.state[w] == 1 && score > V[w].bound, n_cigar = 0;
            if (ok) {
                n_cigar = kb_band16_backtrack<K>(lane, V[w], J[w]->flag, w ? tbB : tbA, ezcig);
                if (n_cigar < 0) ok = 0;
            }
            cells += (int64_t)32 * K * out.steps[w];
            if (lane == 0) KB_DP_STAT(K == 1 ? 0 : 2, ok ? 0 : 1, (int64_t)32 * K * (V[w].r_end + 1));
            if (ok) {
                KbEz ez;
                ez.max = 0, ez.max_q = ez.max_t = -1, ez.zdropped = 0, ez.score = score, ez.n_cigar = n_cigar;
                kb_job_finish(lane, J[w], ez, ezcig, jobcig, jobcig_cap, counters);
                continue;
            }
            if (lane == 0) {  // the smallest wider window this score (a lower bound of every wider pass's score) already certifies, if any
                const int ql = V[w].qlen, tl = V[w].tlen;
                int k2 = 0, dlo, dhi;
                if (K == 1 && kb_band16_geometry(ql, tl, 2, dlo, dhi) && score > kb_band_bound64(P, ql, tl, dlo, dhi)) k2 = 2;
                else if (K <= 2 && kb_band16_geometry(ql, tl, 4, dlo, dhi) && score > kb_band_bound64(P, ql, tl, dlo, dhi)) k2 = 4;
                // worth it only while the window stays well below the rectangle
                if (k2 && 64 * k2 * KB_B16_WIN_RULE <= 2 * (ql + tl)) kb_band16_enqueue(BQ, counters, k2 == 2 ? 1 : 2, *J[w], jid[w]);
                else kb_rows_enqueue(Q, P, counters, *J[w], jid[w]);
            }
        }
    }
    if (lane == 0 && cells) atomicAdd(&counters[8], (unsigned long long)cells);
}

// row-stripe wavefront over everything else (end extensions, fills the band pass could not certify)
#ifndef KB_ROWS_MINB
#define KB_ROWS_MINB 6  // 80 registers, 24 warps per SM: measured 8 % faster than 128 registers / 16 warps (profiles/r1_summary.md)
#endif
__global__ void __launch_bounds__(128, KB_ROWS_MINB) kb_rows_kernel(KbIndexView ix, KbBatchView bt, KbJob *jobs, const int32_t *rows_list, uint8_t *scratch,
                                                        size_t scratch_bytes, uint32_t *jobcig, int64_t jobcig_cap,
                                                        unsigned long long *counters)
{
    __shared__ uint32_t wmax_ring[4][KB_RING_WORDS];
    const int lane = threadIdx.x & 31;
    const int64_t wg = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    KbAlignScratch S = kb_align_scratch_at(scratch + (size_t)wg * scratch_bytes, ix.p.max_sw_cells);
    S.wmax = wmax_ring[threadIdx.x >> 5];
    for (int x = lane; x < KB_RING_WORDS; x += 32) S.wmax[x] = 0;
    __syncwarp();
    const KbDpConst P = kb_dp_const(ix.p);
    const long long n_big = (long long)counters[KB_SC_ROWS_BIG], n = (long long)counters[KB_SC_ROWS] + n_big;
    const int64_t job_cap = jobcig_cap / 16;
    int64_t cells = 0;
    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(&counters[KB_SC_QROWS], 1ull);
        k = __shfl_sync(0xffffffffu, k, 0);
        if ((long long)k >= n) break;
        KbJob *J = jobs + ((long long)k < n_big ? rows_list[job_cap - 1 - (long long)k] : rows_list[(long long)k - n_big]);
        const int dir = J->kind == KB_JOB_LEFT ? -1 : 1;
        const KbDirBytes sq{(J->qrev ? ix.gseq_rev : ix.gseq_fwd) + J->qbase + J->qoff + (dir < 0 ? -1 : 0), dir};
        const KbDirPack st{bt.seq2, bt.nmask, J->tpos + (dir < 0 ? -1 : 0), dir};
        KbEz ez;
        const bool track = !(J->flag & KB_EZ_GLOBAL_NO_ZDROP);
        if (lane == 0) KB_DP_STAT(track ? 1 : 0, J->tlen > 256 ? 3 : 2, (int64_t)J->qlen * J->tlen);
        if (track) kb_rows<true>(P, lane, J->qlen, sq, J->tlen, st, J->w, J->zdrop, J->flag, ez, S, &cells);
        else kb_rows<false>(P, lane, J->qlen, sq, J->tlen, st, J->w, J->zdrop, J->flag, ez, S, &cells);
        kb_job_finish(lane, J, ez, S.ezcig, jobcig, jobcig_cap, counters);
    }
    if (lane == 0 && cells) atomicAdd(&counters[8], (unsigned long long)cells);
}

// the rectangles kb_rows16 takes: two jobs per warp (neighbours of the size-sorted list), two cells per DPX instruction
#ifndef KB_ROWS16_MINB
#define KB_ROWS16_MINB 4  // 128 registers, 16 warps per SM: at 96 the selector and state arrays spill into the hot loop (measured 260 vs 234 ms align)
#endif
__global__ void __launch_bounds__(128, KB_ROWS16_MINB) kb_rows16_kernel(KbIndexView ix, KbBatchView bt, KbJob *jobs, const int32_t *r16_sorted, uint8_t *scratch,
                                                          size_t scratch_bytes, uint32_t *jobcig, int64_t jobcig_cap,
                                                          unsigned long long *counters)
{
    __shared__ uint32_t wmax_ring[4][KB_R16_SMEM_WORDS];
    const int lane = threadIdx.x & 31;
    const int64_t wg = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    KbAlignScratch S = kb_align_scratch_at(scratch + (size_t)wg * scratch_bytes, ix.p.max_sw_cells);
    S.wmax = wmax_ring[threadIdx.x >> 5];
    for (int x = lane; x < KB_R16_SMEM_WORDS; x += 32) S.wmax[x] = 0;
    __syncwarp();
    const KbDpConst P = kb_dp_const(ix.p);
    const long long n = (long long)counters[KB_SC_R16], n_pairs = (n + 1) >> 1;
    int64_t cells = 0;
    typedef KbPairJob<KbDirBytes, KbDirPack> PJ;
    auto view = [&](const KbJob *J) {
        const int dir = J->kind == KB_JOB_LEFT ? -1 : 1;
        return PJ{J->qlen, J->tlen, J->w, J->zdrop, J->flag,
                  KbDirBytes{(J->qrev ? ix.gseq_rev : ix.gseq_fwd) + J->qbase + J->qoff + (dir < 0 ? -1 : 0), dir},
                  KbDirPack{bt.seq2, bt.nmask, J->tpos + (dir < 0 ? -1 : 0), dir}};
    };
    const PJ none{0, 0, 0, -1, KB_EZ_GLOBAL_NO_ZDROP, KbDirBytes{ix.gseq_fwd, 1}, KbDirPack{bt.seq2, bt.nmask, 0, 1}};
    auto run = [&](KbJob *Ja, KbJob *Jb) {  // Jb may be null
        const PJ A = view(Ja), B = Jb ? view(Jb) : none;
        const bool track = !(A.flag & KB_EZ_GLOBAL_NO_ZDROP) || (Jb && !(B.flag & KB_EZ_GLOBAL_NO_ZDROP));
        KbPairOut out;
        KbEz ez;
        if (track) kb_rows16_dp<true>(P, lane, A, B, out, S);
        else kb_rows16_dp<false>(P, lane, A, B, out, S);
        if (track) kb_rows16_finish<true>(P, lane, A, 0, B.tlen, out, ez, S, &cells);
        else kb_rows16_finish<false>(P, lane, A, 0, B.tlen, out, ez, S, &cells);
        kb_job_finish(lane, Ja, ez, S.ezcig, jobcig, jobcig_cap, counters);
        if (Jb) {
            if (track) kb_rows16_finish<true>(P, lane, B, 1, A.tlen, out, ez, S, &cells);
            else kb_rows16_finish<false>(P, lane, B, 1, A.tlen, out, ez, S, &cells);
            kb_job_finish(lane, Jb, ez, S.ezcig, jobcig, jobcig_cap, counters);
        }
    };
    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(&counters[KB_SC_QR16], 1ull);
        k = __shfl_sync(0xffffffffu, k, 0);
        if ((long long)k >= n_pairs) break;
        KbJob *Ja = jobs + r16_sorted[2 * k], *Jb = (long long)(2 * k + 1) < n ? jobs + r16_sorted[2 * k + 1] : nullptr;
        if (lane == 0) {  // path 4: the packed 16-bit wavefront
            KB_DP_STAT((Ja->flag & KB_EZ_GLOBAL_NO_ZDROP) ? 0 : 1, 4, (int64_t)Ja->qlen * Ja->tlen);
            if (Jb) KB_DP_STAT((Jb->flag & KB_EZ_GLOBAL_NO_ZDROP) ? 0 : 1, 4, (int64_t)Jb->qlen * Jb->tlen);
        }
        if (Jb && !kb_rows16_pair_fits(P, Ja->qlen, Ja->tlen, Jb->qlen, Jb->tlen)) run(Ja, nullptr), run(Jb, nullptr);
        else run(Ja, Jb);
    }
    if (lane == 0 && cells) atomicAdd(&counters[8], (unsigned long long)cells);
}

// one thread per planned chain: concatenate, test, mm_update_extra, emit the raw hit
__global__ void __launch_bounds__(128) kb_assemble_kernel(KbIndexView ix, KbBatchView bt, const KbChainRec *chains, int64_t n_chains,
                                                          const KbGroupInfo *ginfo, const KbPlan *plans, const KbJob *jobs,
                                                          const uint32_t *jobcig, uint32_t *tmpcig, int64_t tmpcig_cap, KbRawHit *raw,
                                                          int64_t raw_cap, uint32_t *pool, int64_t pool_cap, int32_t *slow_list,
                                                          unsigned long long *counters)
{
    const int64_t ci = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= n_chains) return;
    const KbPlan pl = plans[ci];
    if (!pl.ok) return;
    const KbChainRec c = chains[ci];
    const KbGroupInfo gi = ginfo[c.group];
    const KbJob *J = jobs + pl.job_base;
    long long total = 0;
    for (int k = 0; k < pl.n_jobs; ++k) total += J[k].n_cigar > 0 ? J[k].n_cigar : 0;
    const long long toff = (long long)atomicAdd(&counters[KB_SC_TMPCIG], (unsigned long long)(total + 1));
    KbReg r;
    const int reg_idx = (int)(ci - gi.chain_base);
    r.as = c.as, r.cnt = c.cnt, r.score = c.score, r.score0 = c.score0, r.mlen = c.mlen, r.blen = c.blen, r.parent = c.parent, r.id = reg_idx;
    r.hash = c.hash, r.rev = c.rev, r.rid = c.rid, r.rs = c.rs, r.re = c.re, r.qs = c.qs, r.qe = c.qe;
    r.has_p = 0, r.dp_score = 0, r.dp_max = 0, r.n_ambi = 0, r.n_cigar = 0;
    // every lane that got here calls it (it contains the warp's re-convergence point); one without scratch room is disabled
    const bool room = toff + total + 1 <= tmpcig_cap;
    const int rc = kb_stage_assemble(ix, bt, gi.gene, pl, J, jobcig, r, tmpcig + (room ? toff : 0), room);
    if (rc != 0) {
        slow_list[atomicAdd(&counters[KB_SC_SLOW], 1ull)] = (int32_t)ci;
        return;
    }
    const int ncg = r.n_cigar;
    const unsigned long long slot = atomicAdd(&counters[4], 1ull);
    const unsigned long long coff = atomicAdd(&counters[5], (unsigned long long)ncg);
    if ((int64_t)slot < raw_cap) {
        KbRawHit h;
        h.group = c.group, h.reg_idx = reg_idx, h.split_idx = 0;
        h.cnt = r.cnt, h.score = r.score, h.score0 = r.score0, h.hash = r.hash;
        h.rev = r.rev, h.rid = r.rid, h.rs = r.rs, h.re = r.re, h.qs = r.qs, h.qe = r.qe;
        h.has_p = r.has_p, h.dp_score = r.dp_score, h.dp_max = r.dp_max, h.dp_max2 = 0, h.n_ambi = r.n_ambi, h.mlen = r.mlen, h.blen = r.blen;
        h.parent = r.parent, h.subsc = c.subsc, h.n_sub = c.n_sub, h.mapq = 0, h.n_cigar = ncg, h.cigar_off = (int64_t)coff;
        h.err = 0, h.pad = 0;
        raw[slot] = h;
    }
    if ((int64_t)(coff + ncg) <= pool_cap)
        for (int i = 0; i < ncg; ++i) pool[coff + i] = tmpcig[toff + i];
}

size_t kb_band_scratch_bytes() { return (size_t)KB_CIG_MAX * 4 + 4 * 32 * 8200 + 256; }
size_t kb_sizeof_job() { return sizeof(KbJob); }
size_t kb_sizeof_plan() { return sizeof(KbPlan); }

static int64_t kb_rows_big_thr()
{
    // measured on the bench workload: 255.6 ms (off) -> 251.9 ms (150 k cells) -> 252.3 ms (300 k cells) for the align stage
    static const int64_t v = getenv("KAPTIVE_B200_ROWS_BIG") ? atoll(getenv("KAPTIVE_B200_ROWS_BIG")) : 150000;
    return v;
}
static int kb_use_rows16()
{
    static const int v = !(getenv("KAPTIVE_B200_ROWS16") && getenv("KAPTIVE_B200_ROWS16")[0] == '0');
    return v;
}
static int kb_use_band16()
{
    static const int v = !(getenv("KAPTIVE_B200_BAND16") && getenv("KAPTIVE_B200_BAND16")[0] == '0');
    return v;
}
static KbBandQueues kb_band_queues(const KbStageLists &L)
{
    KbBandQueues BQ;
    for (int i = 0; i < 3; ++i) BQ.list[i] = L.b16_list[i], BQ.key[i] = L.b16_key[i];
    BQ.band32 = L.band_list, BQ.use16 = kb_use_band16();
    return BQ;
}
void kb_launch_stage_plan(const KbIndexView &ix, const KbBatchView &bt, const KbChainRec *chains, int64_t n_chains, const KbGroupInfo *ginfo,
                          const uint64_t *cx, uint64_t *cy, int32_t *kscratch, void *plans, void *jobs, const KbStageLists &L, int32_t *slow_list,
                          unsigned long long *counters, cudaStream_t st)
{
    if (n_chains <= 0) return;
    // unused slots of the sorted queues must hold key 0
    cudaMemsetAsync(L.r16_key, 0, (size_t)L.job_cap * 4, st);
    for (int i = 0; i < 3; ++i) cudaMemsetAsync(L.b16_key[i], 0, (size_t)L.job_cap * 4, st);
    const KbRowsQueues Q{L.rows_list, L.r16_list, L.r16_key, L.job_cap, kb_rows_big_thr(), kb_use_rows16()};
    kb_plan_kernel<<<(unsigned)((n_chains + 127) / 128), 128, 0, st>>>(ix, bt, chains, n_chains, ginfo, cx, cy, kscratch, (KbPlan *)plans,
                                                                       (KbJob *)jobs, L.job_cap, kb_band_queues(L), Q, slow_list, counters);
}
size_t kb_r16_sort_temp_bytes(int64_t n)
{
    size_t b = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, b, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const int32_t *)nullptr,
                                              (int32_t *)nullptr, n, 0, 29);
    return b;
}
// every queue that is sorted: job_cap entries (unused ones hold key 0), sorted copies in r16_key2 / r16_list2
void kb_launch_stage_dp(const KbIndexView &ix, const KbBatchView &bt, void *jobs, const KbStageLists &L, uint8_t *band_scratch, int band_warps,
                        uint8_t *rows_scratch, size_t rows_scratch_bytes, int rows_warps, uint8_t *side_scratch, int side_warps, uint32_t *jobcig,
                        int64_t jobcig_cap, unsigned long long *counters, cudaStream_t st)
{
    const KbRowsQueues Q{L.rows_list, L.r16_list, L.r16_key, L.job_cap, kb_rows_big_thr(), kb_use_rows16()};
    const KbBandQueues BQ = kb_band_queues(L);
    size_t tmp = L.sort_tmp_bytes;
    if (BQ.use16) {  // K = 1, then what it handed on at K = 2, then K = 4: each queue sorted by length, neighbours share a warp
        const unsigned g = (unsigned)(band_warps / 4);
        cub::DeviceRadixSort::SortPairsDescending(L.sort_tmp, tmp, L.b16_key[0], L.r16_key2, L.b16_list[0], L.r16_list2, L.job_cap, 0, 13, st);
        kb_band16_kernel<1><<<g, 128, 0, st>>>(ix, bt, (KbJob *)jobs, L.r16_list2, BQ, Q, band_scratch, kb_band_scratch_bytes(), jobcig, jobcig_cap, counters);
        cub::DeviceRadixSort::SortPairsDescending(L.sort_tmp, tmp, L.b16_key[1], L.r16_key2, L.b16_list[1], L.r16_list2, L.job_cap, 0, 13, st);
        kb_band16_kernel<2><<<g, 128, 0, st>>>(ix, bt, (KbJob *)jobs, L.r16_list2, BQ, Q, band_scratch, kb_band_scratch_bytes(), jobcig, jobcig_cap, counters);
        cub::DeviceRadixSort::SortPairsDescending(L.sort_tmp, tmp, L.b16_key[2], L.r16_key2, L.b16_list[2], L.r16_list2, L.job_cap, 0, 13, st);
        kb_band16_kernel<4><<<g, 128, 0, st>>>(ix, bt, (KbJob *)jobs, L.r16_list2, BQ, Q, band_scratch, kb_band_scratch_bytes(), jobcig, jobcig_cap, counters);
    }
    kb_band_kernel<<<(unsigned)(band_warps / 4), 128, 0, st>>>(ix, bt, (KbJob *)jobs, L.band_list, Q, band_scratch, kb_band_scratch_bytes(), jobcig,
                                                               jobcig_cap, counters);
    // The 32-bit rows kernel is left with the few rectangles outside the packed kernel's range: long single-warp jobs (14 ms for ~20 jobs
    // per 1000 assemblies).  It runs on a side stream next to the packed kernel, in its own scratch (side_scratch, one slot per warp of
    // side_warps), instead of after it.
    int r16_warps = rows_warps / 4 * 4;
    cudaStream_t s2 = st;
    cudaEvent_t e_fork = nullptr, e_join = nullptr;
    if (kb_use_rows16() && side_scratch) {
        static thread_local struct Side {
            int dev = -1;
            cudaStream_t s = nullptr;
            cudaEvent_t a = nullptr, b = nullptr;
        } side;
        int dev = 0;
        cudaGetDevice(&dev);
        if (side.dev != dev) {
            if (side.s) cudaStreamDestroy(side.s), cudaEventDestroy(side.a), cudaEventDestroy(side.b);
            cudaStreamCreateWithFlags(&side.s, cudaStreamNonBlocking);
            cudaEventCreateWithFlags(&side.a, cudaEventDisableTiming), cudaEventCreateWithFlags(&side.b, cudaEventDisableTiming);
            side.dev = dev;
        }
        s2 = side.s, e_fork = side.a, e_join = side.b;
        cudaEventRecord(e_fork, st);
        cudaStreamWaitEvent(s2, e_fork, 0);
        kb_rows_kernel<<<(unsigned)(side_warps / 4), 128, 0, s2>>>(ix, bt, (KbJob *)jobs, L.rows_list, side_scratch, rows_scratch_bytes, jobcig, jobcig_cap,
                                                                  counters);
        cudaEventRecord(e_join, s2);
    }
    if (kb_use_rows16()) {
        cub::DeviceRadixSort::SortPairsDescending(L.sort_tmp, tmp, L.r16_key, L.r16_key2, L.r16_list, L.r16_list2, L.job_cap, 0, 29, st);
        kb_rows16_kernel<<<(unsigned)(r16_warps / 4), 128, 0, st>>>(ix, bt, (KbJob *)jobs, L.r16_list2, rows_scratch, rows_scratch_bytes, jobcig,
                                                                    jobcig_cap, counters);
    }
    if (s2 != st) cudaStreamWaitEvent(st, e_join, 0);
    else
        kb_rows_kernel<<<(unsigned)(rows_warps / 4), 128, 0, st>>>(ix, bt, (KbJob *)jobs, L.rows_list, rows_scratch, rows_scratch_bytes, jobcig, jobcig_cap,
                                                                   counters);
}
void kb_launch_stage_assemble(const KbIndexView &ix, const KbBatchView &bt, const KbChainRec *chains, int64_t n_chains, const KbGroupInfo *ginfo,
                              const void *plans, const void *jobs, const uint32_t *jobcig, uint32_t *tmpcig, int64_t tmpcig_cap, KbRawHit *raw,
                              int64_t raw_cap, uint32_t *pool, int64_t pool_cap, int32_t *slow_list, unsigned long long *counters,
                              cudaStream_t st)
{
    if (n_chains <= 0) return;
    kb_assemble_kernel<<<(unsigned)((n_chains + 127) / 128), 128, 0, st>>>(ix, bt, chains, n_chains, ginfo, (const KbPlan *)plans,
                                                                           (const KbJob *)jobs, jobcig, tmpcig, tmpcig_cap, raw, raw_cap, pool,
                                                                           pool_cap, slow_list, counters);
}

// ------------------------------------------------------------------ finalisation
__global__ void kb_rawkey_kernel(const KbRawHit *raw, int64_t n, uint64_t *key, uint32_t *idx, unsigned long long *n_err)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const KbRawHit &h = raw[i];
        if (h.err) atomicAdd(n_err, 1ull);  // an alignment that ran into an internal limit: dropped by the finaliser, counted here
        key[i] = (uint64_t)(uint32_t)h.group << 32 | (uint64_t)(h.reg_idx & 0xfffff) << 12 | (uint64_t)(h.split_idx & 0xfff);
        idx[i] = (uint32_t)i;
    }
}
__global__ void kb_gather_raw_kernel(const KbRawHit *raw, const uint32_t *idx, int64_t n, KbRawHit *out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = raw[idx[i]];
}
// thread i works only if sorted hit i is the first of its query group
__global__ void __launch_bounds__(128) kb_finalize_kernel(kb_params_t P, KbRawHit *hits, int64_t n, const KbGroupInfo *ginfo, int32_t *w,
                                                          uint64_t *cov, int32_t *keep_flag)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i > 0 && hits[i - 1].group == hits[i].group) return;
    int64_t e = i + 1;
    while (e < n && hits[e].group == hits[i].group) ++e;
    const int32_t grp = hits[i].group;
    int kept = kb_finalize_group(P, static_cast<KbHitView *>(hits + i), (int)(e - i), ginfo[grp].rep_len, w + i, cov + i);
    for (int64_t j = i; j < e; ++j) {
        keep_flag[j] = (j - i) < kept ? 1 : 0;
        hits[j].group = grp;                         // dropped slots keep the group id so segment detection stays valid
        if ((j - i) < kept) hits[j].pad = (int32_t)(j - i);  // rank within the query
    }
}
__global__ void kb_scatter_kernel(const KbRawHit *hits, const int32_t *keep_flag, const int64_t *out_idx, int64_t n,
                                  const KbGroupInfo *ginfo, KbBatchView bt, kb_hits_t o)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (!keep_flag[i]) continue;
        const KbRawHit &h = hits[i];
        const KbGroupInfo &gi = ginfo[h.group];
        const int64_t k = out_idx[i];
        o.asm_id[k] = gi.asm_id, o.gene[k] = gi.gene;
        o.q_start[k] = h.qs, o.q_end[k] = h.qe;
        o.t_ctg[k] = h.rid, o.t_len[k] = bt.ctg_len[bt.asm_ctg_start[gi.asm_id] + h.rid], o.t_start[k] = h.rs, o.t_end[k] = h.re;
        o.strand[k] = h.rev ? -1 : 1;
        o.score[k] = h.dp_score, o.matches[k] = h.mlen, o.block_len[k] = h.blen, o.edit_distance[k] = h.blen - h.mlen + h.n_ambi;
        o.mapq[k] = (uint8_t)h.mapq, o.is_primary[k] = (uint8_t)(h.parent == h.pad);
        o.cigar_off[k] = h.cigar_off, o.n_cigar[k] = h.n_cigar;
    }
}

void kb_launch_rawkey(const KbRawHit *raw, int64_t n, uint64_t *key, uint32_t *idx, unsigned long long *n_err, cudaStream_t st)
{
    if (n > 0) kb_rawkey_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(raw, n, key, idx, n_err);
}
void kb_launch_gather_raw(const KbRawHit *raw, const uint32_t *idx, int64_t n, KbRawHit *out, cudaStream_t st)
{
    if (n > 0) kb_gather_raw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(raw, idx, n, out);
}
void kb_launch_finalize(const kb_params_t &P, KbRawHit *hits, int64_t n, const KbGroupInfo *ginfo, int32_t *w, uint64_t *cov,
                        int32_t *keep_flag, cudaStream_t st)
{
    if (n > 0) kb_finalize_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(P, hits, n, ginfo, w, cov, keep_flag);
}
void kb_launch_scatter(const KbRawHit *hits, const int32_t *keep_flag, const int64_t *out_idx, int64_t n, const KbGroupInfo *ginfo,
                       const KbBatchView &bt, const kb_hits_t &o, cudaStream_t st)
{
    if (n > 0) kb_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hits, keep_flag, out_idx, n, ginfo, bt, o);
}
void kb_launch_group_flag(const uint64_t *skey, int64_t n, uint8_t *flag, cudaStream_t st)
{
    if (n > 0) {
        int64_t grid = (n + 255) / 256;
        if (grid > 148 * 32) grid = 148 * 32;
        kb_group_flag_kernel<<<(unsigned)grid, 256, 0, st>>>(skey, n, flag);
    }
}

extern "C" int kb_debug_dp_stats(int64_t *out32, int reset)
{
    if (out32 && cudaMemcpyFromSymbol(out32, g_kb_dp_stats, sizeof(g_kb_dp_stats)) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long z[32] = {0};
        if (cudaMemcpyToSymbol(g_kb_dp_stats, z, sizeof(z)) != cudaSuccess) return -1;
    }
    return 0;
}

// ------------------------------------------------------------------ DP kernels one by one (diagnostic / parity tests)
// One warp per job; mode 0 = scratch-memory DP (kb_extd2, the statement closest to the oracle), 1 = kb_rows, 2 = kb_rows16,
// 3 = certified band pass (kb_global_band), 4 = kb_rows16 paired, 5 / 6 / 7 = packed band pass (kb_band16) with K = 1 / 2 / 4, paired.
// out: n x 8 = score, max, max_t, max_q, zdropped, n_cigar, ran (0: not eligible, 2: ran, not certified), pass state (kb_band16).
__global__ void __launch_bounds__(128) kb_debug_dp_kernel(kb_params_t pp, const uint8_t *q, const int64_t *qoff, const int32_t *qlen,
                                                          const uint8_t *t, const int64_t *toff, const int32_t *tlen, const int32_t *flag,
                                                          const int32_t *w, const int32_t *zdrop, int n, int mode, uint8_t *scratch,
                                                          size_t scratch_bytes, int32_t *out, uint32_t *cig, int cig_stride)
{
    __shared__ uint32_t wmax_ring[4][KB_R16_SMEM_WORDS];
    __shared__ uint2 dbg_lut[25];
    const int lane = threadIdx.x & 31;
    const int64_t wg = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    {
        const KbDpConst P0 = kb_dp_const(pp);
        if (threadIdx.x < 25) dbg_lut[threadIdx.x] = make_uint2(kb_qrow16(P0, 7, threadIdx.x % 5), kb_qrow16(P0, 7, threadIdx.x / 5));
        __syncthreads();
    }
    KbAlignScratch S = kb_align_scratch_at(scratch + (size_t)wg * scratch_bytes, pp.max_sw_cells);
    S.wmax = wmax_ring[threadIdx.x >> 5];
    for (int x = lane; x < KB_R16_SMEM_WORDS; x += 32) S.wmax[x] = 0;
    __syncwarp();
    const KbDpConst P = kb_dp_const(pp);
    for (int64_t k = wg; k < n; k += nw) {
        const KbPtrSeq sq{q + qoff[k]}, st{t + toff[k]};
        const int ql = qlen[k], tl = tlen[k], fl = flag[k], ww = w[k], zd = zdrop[k];
        const bool track = !(fl & KB_EZ_GLOBAL_NO_ZDROP);
        KbEz ez;
        ez.max = 0, ez.max_q = ez.max_t = -1, ez.score = KB_NEG_INF, ez.zdropped = 0, ez.n_cigar = 0;
        int ran = 1, o7 = 0;
        if (mode == 0) kb_extd2<32>(P, lane, ql, sq.p, tl, st.p, ww, zd, fl, ez, S, nullptr);
        else if (mode == 1) {
            if (!kb_rows_eligible(P.max_sw_cells, ql, tl, ww, track)) ran = 0;
            else if (track) kb_rows<true>(P, lane, ql, sq, tl, st, ww, zd, fl, ez, S, nullptr);
            else kb_rows<false>(P, lane, ql, sq, tl, st, ww, zd, fl, ez, S, nullptr);
        } else if (mode == 2 || mode == 4) {
            // mode 2: the job alone in the low half; mode 4: paired with job k ^ 1 (this job in the low half when k is even)
            typedef KbPairJob<KbPtrSeq, KbPtrSeq> PJ;
            const int64_t ko = (mode == 4 && (k ^ 1) < n) ? (k ^ 1) : -1;
            const PJ me{ql, tl, ww, zd, fl, sq, st};
            PJ other{0, 0, 0, -1, KB_EZ_GLOBAL_NO_ZDROP, sq, st};
            bool pair = false;
            if (ko >= 0) {
                const PJ o{qlen[ko], tlen[ko], w[ko], zdrop[ko], flag[ko], KbPtrSeq{q + qoff[ko]}, KbPtrSeq{t + toff[ko]}};
                pair = kb_rows16_eligible(P, o.qlen, o.tlen, o.w, !(o.flag & KB_EZ_GLOBAL_NO_ZDROP)) && kb_rows16_pair_fits(P, ql, tl, o.qlen, o.tlen);
                if (pair) other = o;
            }
            if (!kb_rows16_eligible(P, ql, tl, ww, track) || (mode == 4 && !pair)) ran = 0;
            else {
                const bool me_hi = pair && (k & 1);
                const PJ &A = me_hi ? other : me, &B = me_hi ? me : other;
                const bool tr = !(A.flag & KB_EZ_GLOBAL_NO_ZDROP) || (pair && !(B.flag & KB_EZ_GLOBAL_NO_ZDROP));
                KbPairOut po;
                if (tr) kb_rows16_dp<true>(P, lane, A, B, po, S), kb_rows16_finish<true>(P, lane, me, me_hi ? 1 : 0, other.tlen, po, ez, S, nullptr);
                else kb_rows16_dp<false>(P, lane, A, B, po, S), kb_rows16_finish<false>(P, lane, me, me_hi ? 1 : 0, other.tlen, po, ez, S, nullptr);
            }
        } else if (mode >= 5) {
            // modes 5, 6, 7: the packed band pass with K = 1, 2, 4, paired with job k ^ 1 (this job in the low half when k is even)
            const int K = mode == 5 ? 1 : (mode == 6 ? 2 : 4);
            const int64_t ko = (k ^ 1) < n ? (k ^ 1) : -1;
            int dlo, dhi;
            auto fits = [&](int64_t x) {
                return kb_band_eligible(P.max_sw_cells, qlen[x], tlen[x], w[x], flag[x]) && kb_band16_eligible(P, qlen[x], tlen[x]) &&
                       kb_band16_geometry(qlen[x], tlen[x], K, dlo, dhi);
            };
            if (!fits(k)) ran = 0;
            else {
                const bool pair = ko >= 0 && fits(ko), me_hi = pair && (k & 1);
                const int64_t ka = me_hi ? ko : k, kb = pair ? (me_hi ? k : ko) : k;
                const KbB16Job A = kb_b16_job(P, K, qlen[ka], tlen[ka]), B = kb_b16_job(P, K, qlen[kb], tlen[kb]);
                uint16_t *st_sel = reinterpret_cast<uint16_t *>(S.wmax);
                uint8_t *st_q = reinterpret_cast<uint8_t *>(S.wmax) + 2 * KB_B16_X;
                uint32_t *tbA = reinterpret_cast<uint32_t *>(S.tb), *tbB = tbA + (P.max_sw_cells >> 3);
                const bool amb = kb_band16_stage(lane, K, A, KbPtrSeq{q + qoff[ka]}, KbPtrSeq{t + toff[ka]}, B, KbPtrSeq{q + qoff[kb]},
                                                 KbPtrSeq{t + toff[kb]}, st_sel, st_q);
                if (amb) ran = 0;
                else {
                    KbB16Out po;
                    const unsigned a_sel = (unsigned)__cvta_generic_to_shared(st_sel), a_q = (unsigned)__cvta_generic_to_shared(st_q),
                                   a_lut = (unsigned)__cvta_generic_to_shared(dbg_lut);
                    if (K == 1) kb_band16_pass<1>(P, lane, A, B, pair, a_sel, a_q, a_lut, tbA, tbB, po);
                    else if (K == 2) kb_band16_pass<2>(P, lane, A, B, pair, a_sel, a_q, a_lut, tbA, tbB, po);
                    else kb_band16_pass<4>(P, lane, A, B, pair, a_sel, a_q, a_lut, tbA, tbB, po);
                    __syncwarp();
                    const int wme = me_hi ? 1 : 0;
                    const KbB16Job &M = me_hi ? B : A;
                    ez.score = po.score[wme], ran = 2;
                    if (po.state[wme] == 1 && po.score[wme] > M.bound) {
                        const uint32_t *tbm = me_hi ? tbB : tbA;
                        const int nc = K == 1 ? kb_band16_backtrack<1>(lane, M, fl, tbm, S.ezcig)
                                              : (K == 2 ? kb_band16_backtrack<2>(lane, M, fl, tbm, S.ezcig) : kb_band16_backtrack<4>(lane, M, fl, tbm, S.ezcig));
                        if (nc >= 0) ez.n_cigar = nc, ran = 1;
                    }
                    o7 = po.state[wme];
                }
                // the staging area doubles as kb_rows16's rings: leave it zeroed
                __syncwarp();
                for (int x = lane; x < KB_R16_SMEM_WORDS; x += 32) S.wmax[x] = 0;
                __syncwarp();
            }
        } else {
            if (!kb_band_eligible(P.max_sw_cells, ql, tl, ww, fl)) ran = 0;
            else ran = kb_global_band(P, lane, ql, sq, tl, st, fl, ez, S, nullptr) ? 1 : 2;  // 2: ran, not certified
        }
        __syncwarp();
        if (lane == 0) {
            int32_t *o = out + k * 8;
            o[0] = ez.score, o[1] = ez.max, o[2] = ez.max_t, o[3] = ez.max_q, o[4] = ez.zdropped, o[5] = ez.n_cigar, o[6] = ran, o[7] = o7;
        }
        if (ran == 1)
            for (int i = lane; i < ez.n_cigar && i < cig_stride; i += 32) cig[k * (int64_t)cig_stride + i] = S.ezcig[i];
        __syncwarp();
    }
}
extern "C" int kb_debug_dp(const kb_params_t *pp, int device, const uint8_t *q, const int64_t *qoff, const int32_t *qlen, const uint8_t *t,
                           const int64_t *toff, const int32_t *tlen, const int32_t *flag, const int32_t *w, const int32_t *zdrop, int32_t n,
                           int32_t mode, int32_t *out, uint32_t *cig, int32_t cig_stride)
{
    if (!pp || n <= 0 || !out || !cig) return KB_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return KB_ERR_CUDA;
    int64_t qb = 0, tbts = 0;
    for (int i = 0; i < n; ++i) qb = std::max<int64_t>(qb, qoff[i] + qlen[i]), tbts = std::max<int64_t>(tbts, toff[i] + tlen[i]);
    const int n_warps = 148 * 4;
    kb_params_t fastp = *pp;  // the register-resident kernels work within the fast limit of the mapping path (kb_api.cu)
    if (fastp.max_sw_cells > 4000000) fastp.max_sw_cells = 4000000;
    pp = &fastp;
    const size_t sbytes = kb_align_scratch_bytes(pp->max_sw_cells);
    uint8_t *dq = nullptr, *dt = nullptr, *scr = nullptr;
    int64_t *dqo = nullptr, *dto = nullptr;
    int32_t *dql = nullptr, *dtl = nullptr, *dfl = nullptr, *dw = nullptr, *dz = nullptr, *dout = nullptr;
    uint32_t *dcig = nullptr;
    int rc = KB_OK;
    auto up = [&](void **d, const void *h, size_t bytes) {
        if (cudaMalloc(d, bytes + 16) != cudaSuccess || (h && cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice) != cudaSuccess)) rc = KB_ERR_CUDA;
    };
    up((void **)&dq, q, (size_t)qb), up((void **)&dt, t, (size_t)tbts), up((void **)&dqo, qoff, (size_t)n * 8), up((void **)&dto, toff, (size_t)n * 8);
    up((void **)&dql, qlen, (size_t)n * 4), up((void **)&dtl, tlen, (size_t)n * 4), up((void **)&dfl, flag, (size_t)n * 4);
    up((void **)&dw, w, (size_t)n * 4), up((void **)&dz, zdrop, (size_t)n * 4);
    up((void **)&dout, nullptr, (size_t)n * 32), up((void **)&dcig, nullptr, (size_t)n * cig_stride * 4), up((void **)&scr, nullptr, (size_t)n_warps * sbytes);
    if (rc == KB_OK) {
        kb_debug_dp_kernel<<<n_warps / 4, 128>>>(*pp, dq, dqo, dql, dt, dto, dtl, dfl, dw, dz, n, mode, scr, sbytes, dout, dcig, cig_stride);
        if (cudaDeviceSynchronize() != cudaSuccess) rc = KB_ERR_CUDA;
        else if (cudaMemcpy(out, dout, (size_t)n * 32, cudaMemcpyDeviceToHost) != cudaSuccess ||
                 cudaMemcpy(cig, dcig, (size_t)n * cig_stride * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = KB_ERR_CUDA;
    }
    (void)cudaGetLastError();
    for (void *p : {(void *)dq, (void *)dt, (void *)dqo, (void *)dto, (void *)dql, (void *)dtl, (void *)dfl, (void *)dw, (void *)dz, (void *)dout,
                    (void *)dcig, (void *)scr})
        if (p) cudaFree(p);
    return rc;
}

// ------------------------------------------------------------------ stage dumps for parity tests
__global__ void kb_dump_anchors_kernel(KbBatchView bt, const KbGroupInfo *ginfo, int64_t n_groups, const int64_t *gstart,
                                       const uint32_t *wx, const int32_t *wy, const int64_t *out_off, int32_t *out)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const KbGroupInfo &gi = ginfo[g];
    for (int32_t i = 0; i < gi.n_seed; ++i) {
        uint32_t xv = wx[gstart[g] + i];
        int32_t yv = wy[gstart[g] + i];
        int32_t vpos = (int32_t)(xv & KB_VPOS_MASK);
        int32_t c = kb_vpos_to_ctg(bt, gi.asm_id, vpos);
        int32_t *o = out + (out_off[g] + i) * 7;
        o[0] = gi.asm_id, o[1] = gi.gene, o[2] = (int32_t)(xv >> KB_KEY_REV_SHIFT), o[3] = c - bt.asm_ctg_start[gi.asm_id];
        o[4] = vpos - bt.ctg_vstart[c], o[5] = yv & 0x3fffffff, o[6] = (yv >> 30) & 1;
    }
}
__global__ void kb_group_nseed_kernel(const KbGroupInfo *ginfo, int64_t n_groups, int32_t *n_seed)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_groups) n_seed[g] = ginfo[g].n_seed;
}
void kb_launch_group_nseed(const KbGroupInfo *ginfo, int64_t n_groups, int32_t *n_seed, cudaStream_t st)
{
    if (n_groups > 0) kb_group_nseed_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, st>>>(ginfo, n_groups, n_seed);
}
void kb_launch_dump_anchors(const KbBatchView &bt, const KbGroupInfo *ginfo, int64_t n_groups, const int64_t *gstart, const uint32_t *wx,
                            const int32_t *wy, const int64_t *out_off, int32_t *out, cudaStream_t st)
{
    if (n_groups > 0) kb_dump_anchors_kernel<<<(unsigned)((n_groups + 127) / 128), 128, 0, st>>>(bt, ginfo, n_groups, gstart, wx, wy, out_off, out);
}
