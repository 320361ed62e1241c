// kb_post.cu -- post-mapping numerics of the typing path on the GPU (first "next" row).
//
//   kb_post_extract        replaces _extract_ragged_kernel   (src/kaptive/core/seq.py:612-668; called serotyping/core.py:333,352)
//   kb_post_translate      replaces _translate_ragged_kernel (src/kaptive/core/seq.py:671-741; called serotyping/core.py:360)
//   kb_post_protein_align  replaces _batched_banded_gotoh    (src/kaptive/core/pairwise.py:395-584; called serotyping/core.py:378)
//   kb_post_cull_overlaps  replaces _cull_overlaps_kernel    (src/kaptive/core/interval.py:698-751; via Alignments.cull_overlaps)
//   kb_post_cluster        replaces _cluster_kernel          (src/kaptive/core/interval.py:595-639; via Intervals.cluster_spatial)
//
// The reference launches each of these once per assembly on ~20-60 items, which costs it ~6 ms of numba parallel-region
// wake-up per call; here one call handles the items of a whole batch of assemblies.  Bit-exact: integer/byte work only.
// Byte gathers are one thread per output byte (coalesced writes); the protein DP is one thread per pair with the two
// live rows and a 1-byte/cell packed traceback in global scratch -- thousands of independent pairs per call are the
// parallelism, exactly as in the reference's prange.
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "kb_common.cuh"
#include "kb_post_tables.h"

namespace {
thread_local std::string g_post_err;

struct PostTables {
    uint8_t comp[256], chr[256], codon[128];
    int8_t aa_idx[256];  // index into the 25-letter alphabet or -1
    int8_t blosum[25 * 25];
};
__constant__ PostTables c_tab;
bool g_tables_uploaded[64] = {false};

void build_tables(PostTables &T)
{
    const char *from = "ACGTUacgtu", *to = "TGCAAtgcaa", *ord = "TCAG", *alpha = KB_AA_ALPHABET;
    for (int i = 0; i < 256; ++i) T.comp[i] = (uint8_t)i, T.chr[i] = 4, T.aa_idx[i] = -1;
    for (int i = 0; from[i]; ++i) T.comp[(uint8_t)from[i]] = (uint8_t)to[i];
    for (int i = 0; i < 4; ++i) T.chr[(uint8_t)"ACGT"[i]] = (uint8_t)i, T.chr[(uint8_t)"acgt"[i]] = (uint8_t)i;
    T.chr['U'] = T.chr['u'] = 3;
    for (int i = 0; i < 128; ++i) T.codon[i] = 'X';
    int idx[256] = {0}, k = 0;
    idx['A'] = 0, idx['C'] = 1, idx['G'] = 2, idx['T'] = 3;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
            for (int c = 0; c < 4; ++c) T.codon[idx[(int)ord[a]] * 25 + idx[(int)ord[b]] * 5 + idx[(int)ord[c]]] = (uint8_t)KB_CODE_TCAG[k++];
    for (int a = 0; a < 25; ++a) T.aa_idx[(uint8_t)alpha[a]] = (int8_t)a;
    for (int i = 0; i < 625; ++i) T.blosum[i] = KB_BLOSUM62[i];
}

#define PCU(x)                                                                              \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            g_post_err = std::string(#x) + " failed: " + cudaGetErrorString(e_);            \
            throw g_post_err;                                                               \
        }                                                                                   \
    } while (0)

struct Dev {  // tiny RAII helper: device buffers freed on scope exit
    std::vector<void *> p;
    template <class T>
    T *alloc(size_t n)
    {
        void *q = nullptr;
        PCU(cudaMalloc(&q, (n ? n : 1) * sizeof(T)));
        p.push_back(q);
        return (T *)q;
    }
    template <class T>
    T *upload(const T *h, size_t n)
    {
        T *d = alloc<T>(n);
        if (n) PCU(cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyDefault));
        return d;
    }
    ~Dev()
    {
        for (void *q : p) cudaFree(q);
    }
};

void ensure_device()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        g_post_err = "no CUDA device: libkaptive_b200 has no CPU fallback";
        throw g_post_err;
    }
    int dev = 0;
    PCU(cudaGetDevice(&dev));
    if (dev < 64 && !g_tables_uploaded[dev]) {
        PostTables T;
        build_tables(T);
        PCU(cudaMemcpyToSymbol(c_tab, &T, sizeof(T)));
        g_tables_uploaded[dev] = true;
    }
}

// ---- extract: item i = bytes [starts[i], ends[i]) of parent indices[i], reverse-complemented when strands[i] < 0
__global__ void extract_kernel(const uint8_t *seqs, const int64_t *poff, const int32_t *indices, const int32_t *starts, const int32_t *ends,
                               const int8_t *strands, const int64_t *out_off, int32_t n, int64_t total, uint8_t *out)
{
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        int lo = 0, hi = n - 1;  // last item whose output starts at or before o and is not empty before o
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (out_off[mid] <= o) lo = mid;
            else hi = mid - 1;
        }
        const int i = lo;
        const int64_t c = o - out_off[i];
        const int64_t base = poff[indices[i]];
        out[o] = strands[i] >= 0 ? seqs[base + starts[i] + c] : c_tab.comp[seqs[base + ends[i] - 1 - c]];
    }
}

// ---- translate, pass 1: number of codons before the first stop (to_stop) per item; one thread per item
__global__ void translate_count_kernel(const uint8_t *seqs, const int64_t *off, const int32_t *len, const int8_t *frames, int32_t n,
                                       int to_stop, int32_t *out_len)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t L = len[i], f = frames[i], nc = 0;
    if (L > f) {
        int32_t adj = L - f, mx = adj >= 3 ? adj / 3 : 0;
        if (!to_stop) nc = mx;
        else {
            const uint8_t *p = seqs + off[i] + f;
            for (; nc < mx; ++nc, p += 3)
                if (c_tab.codon[c_tab.chr[p[0]] * 25 + c_tab.chr[p[1]] * 5 + c_tab.chr[p[2]]] == 42) break;
        }
    }
    out_len[i] = nc;
}
// pass 2: one thread per output residue
__global__ void translate_fill_kernel(const uint8_t *seqs, const int64_t *off, const int8_t *frames, const int64_t *out_off, int32_t n,
                                      int64_t total, uint8_t *out)
{
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        int lo = 0, hi = n - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (out_off[mid] <= o) lo = mid;
            else hi = mid - 1;
        }
        const uint8_t *p = seqs + off[lo] + frames[lo] + 3 * (o - out_off[lo]);
        out[o] = c_tab.codon[c_tab.chr[p[0]] * 25 + c_tab.chr[p[1]] * 5 + c_tab.chr[p[2]]];
    }
}

// ---- banded local Gotoh with traceback, one thread per pair (the reference's loop body, rolling rows)
// traceback byte: bits 0-1 tb_M (0 diag, 1 D, 2 I, 3 stop), bit 2 tb_D extended, bit 3 tb_I extended
__global__ void gotoh_kernel(const uint8_t *q, const int64_t *q_off, const int32_t *q_len, const uint8_t *t, const int64_t *t_off,
                             const int32_t *t_len, int32_t n, int32_t k, int32_t go, int32_t ge, const int64_t *tb_off, uint8_t *tb,
                             const int64_t *row_off, int32_t *rowbuf, int32_t *res)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int32_t INF = -1000000000;
    const uint8_t *s1 = q + q_off[idx], *s2 = t + t_off[idx];
    const int32_t len1 = q_len[idx], len2 = t_len[idx], rows = len1 + 1, cols = len2 + 1;
    const int32_t dl = len1 > len2 ? len1 - len2 : len2 - len1;
    const int32_t kl = k > dl + 1 ? k : dl + 1, bw = 2 * kl + 3;
    uint8_t *T = tb + tb_off[idx];
    int32_t *Mp = rowbuf + row_off[idx], *Dp = Mp + bw, *Mc = Dp + bw, *Dc = Mc + bw;  // previous / current rows of M and D
    int32_t max_score = 0, max_i = 0, max_j = 0;
    // row 0: every in-range cell is the initial state (M = 0, D = -inf, stop)
    for (int32_t jm = 0; jm < bw; ++jm) Mp[jm] = 0, Dp[jm] = INF;
    for (int32_t i = 1; i < rows; ++i) {
        const int32_t sj = i - kl > 1 ? i - kl : 1, ej = i + kl + 1 < cols ? i + kl + 1 : cols;
        const int32_t sp = i - 1 - kl - 1 > 0 ? i - 1 - kl - 1 : 0, sc = i - kl - 1 > 0 ? i - kl - 1 : 0;
        // cells of the previous row that were initialised but not computed read as the initial state; the same for this row
        for (int32_t jm = 0; jm < bw; ++jm) Mc[jm] = 0, Dc[jm] = INF;
        if (sj < cols && ej > 1) {
            const int32_t ai = c_tab.aa_idx[s1[i - 1]];
            int32_t m_left = 0, i_left = INF;  // (i, sj - 1) is an initialised, never computed cell
            // the previous row was computed for columns [psj, pej); outside of it the initial state applies
            const int32_t psj = i - 1 >= 1 ? (i - 1 - kl > 1 ? i - 1 - kl : 1) : cols, pej = i - 1 >= 1 ? (i + kl < cols ? i + kl : cols) : 0;
            for (int32_t j = sj; j < ej; ++j) {
                const bool top_ok = j >= psj && j < pej, tl_ok = j - 1 >= psj && j - 1 < pej;
                const int32_t m_top = top_ok ? Mp[j - sp] : 0, d_top = top_ok ? Dp[j - sp] : INF, m_tl = tl_ok ? Mp[j - 1 - sp] : 0;
                const int32_t d_open = m_top - go - ge, d_ext = d_top - ge;
                int32_t dcur, icur, tbv = 0;
                if (d_open >= d_ext) dcur = d_open;
                else dcur = d_ext, tbv |= 4;
                const int32_t i_open = m_left - go - ge, i_ext = i_left - ge;
                if (i_open >= i_ext) icur = i_open;
                else icur = i_ext, tbv |= 8;
                const int32_t bi = c_tab.aa_idx[s2[j - 1]];
                const int32_t sub = (ai >= 0 && bi >= 0) ? (int32_t)c_tab.blosum[ai * 25 + bi] : -128;
                int32_t best = m_tl + sub, tbm = 0;
                if (dcur > best) best = dcur, tbm = 1;
                if (icur > best) best = icur, tbm = 2;
                int32_t mcur;
                if (best <= 0) mcur = 0, tbm = 3;
                else {
                    mcur = best;
                    if (best > max_score) max_score = best, max_i = i, max_j = j;
                }
                Mc[j - sc] = mcur, Dc[j - sc] = dcur;
                T[(int64_t)i * bw + (j - sc)] = (uint8_t)(tbv | tbm);
                m_left = mcur, i_left = icur;
            }
        }
        int32_t *x = Mp;
        Mp = Mc, Mc = x, x = Dp, Dp = Dc, Dc = x;
    }
    int32_t i = max_i, j = max_j, matches = 0, mismatches = 0, gaps = 0, state = 0;
    while (i > 0 && j > 0) {
        const int32_t sj = i - kl > 1 ? i - kl : 1, ej = i + kl + 1 < cols ? i + kl + 1 : cols;
        const int32_t sc = i - kl - 1 > 0 ? i - kl - 1 : 0;
        // a cell that was never computed keeps the initial traceback state: stop
        const int32_t v = (j >= sj && j < ej) ? T[(int64_t)i * bw + (j - sc)] : 3;
        if (state == 0) {
            const int32_t tbm = v & 3;
            if (tbm == 3) break;
            else if (tbm == 0) {
                if (s1[i - 1] == s2[j - 1]) ++matches;
                else ++mismatches;
                --i, --j;
            } else state = tbm;
        } else if (state == 1) {
            ++gaps, --i;
            if (!(v & 4)) state = 0;
        } else {
            ++gaps, --j;
            if (!(v & 8)) state = 0;
        }
    }
    int32_t *r = res + (int64_t)idx * 8;
    r[0] = max_score, r[1] = matches, r[2] = mismatches, r[3] = gaps, r[4] = i, r[5] = max_i, r[6] = j, r[7] = max_j;
}

// ---- typing numerics on the device-resident batch (kb_type.cpp): extract + translate fused, straight from the 2-bit packed
// contigs.  Item i = bases [ts, te) of contig `ctg` (batch-global index), reverse-complemented when strand < 0, read in frame
// `frame`, translated with table 11 up to the first stop codon (seq.py:612-741 fused; a base that is not A/C/G/T/U is code 4 in
// the reference's char_map too, so the packed form loses nothing).  One thread per item; out at out_off[i], length to prot_len[i].
__global__ void type_translate_kernel(KbBatchView bv, const int32_t *ctg, const int32_t *ts, const int32_t *te, const int8_t *strand,
                                      const int8_t *frame, const int64_t *out_off, int64_t n, uint8_t *out, int32_t *prot_len)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t L = te[i] - ts[i], f = frame[i];
    int32_t nc = 0;
    if (L > f) {
        const int32_t adj = L - f, mx = adj >= 3 ? adj / 3 : 0;
        const int64_t base = bv.ctg_soff[ctg[i]];
        const bool rev = strand[i] < 0;
        uint8_t *o = out + out_off[i];
        auto code = [&](int32_t x) {  // base x of the extracted (oriented) sequence
            const int c = kb_fetch_base(bv.seq2, bv.nmask, rev ? base + te[i] - 1 - x : base + ts[i] + x);
            return rev && c < 4 ? 3 - c : c;
        };
        for (int32_t x = f; nc < mx; ++nc, x += 3) {
            const uint8_t aa = c_tab.codon[code(x) * 25 + code(x + 1) * 5 + code(x + 2)];
            if (aa == 42) break;
            o[nc] = aa;
        }
    }
    prot_len[i] = nc;
}

// ---- banded local Gotoh for the typing pass: the same recurrences and tie rules as gotoh_kernel, but the three numbers the
// typing logic consumes (matches, mismatches, gaps of the traceback path) are carried FORWARD with every state instead of being
// recovered from a stored traceback: a state's counts are those of the predecessor the reference's traceback would step to
// (M <- diag / D / I by `tbm`, D <- M when d_open >= d_ext else D, I <- M when i_open >= i_ext else I), so the counts at the first
// maximum cell are exactly what the traceback from that cell yields -- with two rows of state per pair and no traceback memory
// (the stored form needs (len + 1) * band bytes per pair: 15 KB for a typical gene, several GB per batch).
// counts packed as matches | mismatches << 21 | gaps << 42.  res: n x 4 = score, matches, mismatches, gaps.
// Thread `tid` works on pair perm[tid] (pairs sorted by band width, so the lanes of a warp run rows of similar length); the row
// state of a warp's 32 pairs is interleaved (element jm of lane l at tile + jm * 32 + l), which makes every access of the inner
// loop one coalesced line per warp.
__global__ void gotoh_counts_kernel(const uint8_t *q, const int64_t *q_off, const int32_t *q_len, const uint8_t *t, const int64_t *t_off,
                                    const int32_t *t_len, int32_t n, int32_t k, int32_t go, int32_t ge, const int32_t *perm, const int64_t *tile_off,
                                    const int32_t *tile_bw, int32_t *rowbuf, unsigned long long *cntbuf, int32_t *res)
{
    // the substitution tables in shared memory: lanes index them with different residues, which a __constant__ bank serialises
    __shared__ int8_t s_aa[256], s_bl[625];
    for (int x = threadIdx.x; x < 256; x += blockDim.x) s_aa[x] = c_tab.aa_idx[x];
    for (int x = threadIdx.x; x < 625; x += blockDim.x) s_bl[x] = c_tab.blosum[x];
    __syncthreads();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    const int idx = perm[tid];
    const int64_t tile = tile_off[tid >> 5];
    const int32_t tbw = tile_bw[tid >> 5], lane = tid & 31;
    typedef unsigned long long u64;
    const int32_t INF = -1000000000;
    const u64 ONE_M = 1ull, ONE_X = 1ull << 21, ONE_G = 1ull << 42;
    const uint8_t *s1 = q + q_off[idx], *s2 = t + t_off[idx];
    const int32_t len1 = q_len[idx], len2 = t_len[idx], rows = len1 + 1, cols = len2 + 1;
    const int32_t dl = len1 > len2 ? len1 - len2 : len2 - len1;
    const int32_t kl = k > dl + 1 ? k : dl + 1, bw = 2 * kl + 3;
    // four arrays of tbw x 32 elements each per warp tile; this lane's element jm sits at [jm * 32]
    int32_t *Mp = rowbuf + tile * 4 + lane, *Dp = Mp + (int64_t)tbw * 32, *Mc = Dp + (int64_t)tbw * 32, *Dc = Mc + (int64_t)tbw * 32;
    u64 *CMp = cntbuf + tile * 4 + lane, *CDp = CMp + (int64_t)tbw * 32, *CMc = CDp + (int64_t)tbw * 32, *CDc = CMc + (int64_t)tbw * 32;
    int32_t max_score = 0;
    u64 max_cnt = 0;
    // no initialisation of the rows: every read below is guarded by the range the previous row actually computed ([psj, pej))
    for (int32_t i = 1; i < rows; ++i) {
        const int32_t sj = i - kl > 1 ? i - kl : 1, ej = i + kl + 1 < cols ? i + kl + 1 : cols;
        const int32_t sp = i - 1 - kl - 1 > 0 ? i - 1 - kl - 1 : 0, sc = i - kl - 1 > 0 ? i - kl - 1 : 0;
        if (sj < cols && ej > 1) {
            const uint8_t a = s1[i - 1];
            const int32_t ai = s_aa[a];
            int32_t m_left = 0, i_left = INF;
            u64 cm_left = 0, ci_left = 0;
            const int32_t psj = i - 1 >= 1 ? (i - 1 - kl > 1 ? i - 1 - kl : 1) : cols, pej = i - 1 >= 1 ? (i + kl < cols ? i + kl : cols) : 0;
            for (int32_t j = sj; j < ej; ++j) {
                const bool top_ok = j >= psj && j < pej, tl_ok = j - 1 >= psj && j - 1 < pej;
                const int32_t m_top = top_ok ? Mp[(j - sp) * 32] : 0, d_top = top_ok ? Dp[(j - sp) * 32] : INF, m_tl = tl_ok ? Mp[(j - 1 - sp) * 32] : 0;
                const u64 cm_top = top_ok ? CMp[(j - sp) * 32] : 0, cd_top = top_ok ? CDp[(j - sp) * 32] : 0, cm_tl = tl_ok ? CMp[(j - 1 - sp) * 32] : 0;
                const int32_t d_open = m_top - go - ge, d_ext = d_top - ge;
                int32_t dcur, icur;
                u64 cd, ci;
                if (d_open >= d_ext) dcur = d_open, cd = cm_top + ONE_G;
                else dcur = d_ext, cd = cd_top + ONE_G;
                const int32_t i_open = m_left - go - ge, i_ext = i_left - ge;
                if (i_open >= i_ext) icur = i_open, ci = cm_left + ONE_G;
                else icur = i_ext, ci = ci_left + ONE_G;
                const uint8_t b = s2[j - 1];
                const int32_t bi = s_aa[b];
                const int32_t sub = (ai >= 0 && bi >= 0) ? (int32_t)s_bl[ai * 25 + bi] : -128;
                int32_t best = m_tl + sub;
                u64 cb = cm_tl + (a == b ? ONE_M : ONE_X);
                if (dcur > best) best = dcur, cb = cd;
                if (icur > best) best = icur, cb = ci;
                int32_t mcur;
                u64 cm;
                if (best <= 0) mcur = 0, cm = 0;
                else {
                    mcur = best, cm = cb;
                    if (best > max_score) max_score = best, max_cnt = cb;
                }
                Mc[(j - sc) * 32] = mcur, Dc[(j - sc) * 32] = dcur, CMc[(j - sc) * 32] = cm, CDc[(j - sc) * 32] = cd;
                m_left = mcur, i_left = icur, cm_left = cm, ci_left = ci;
            }
        }
        int32_t *x = Mp;
        Mp = Mc, Mc = x, x = Dp, Dp = Dc, Dc = x;
        u64 *y = CMp;
        CMp = CMc, CMc = y, y = CDp, CDp = CDc, CDc = y;
    }
    int32_t *r = res + (int64_t)idx * 4;
    r[0] = max_score, r[1] = (int32_t)(max_cnt & 0x1fffff), r[2] = (int32_t)((max_cnt >> 21) & 0x1fffff), r[3] = (int32_t)(max_cnt >> 42);
}

// The same DP, one WARP per pair, row state in registers.  Lane l holds the K band slots b = l K .. l K + K - 1 of the current row
// (slot b of row i is column j = i - kl + b, so (i - 1, j - 1) is the same slot of the previous row and (i - 1, j) the next one: one
// shuffle per quantity and row).  The horizontal gap state, the only dependency along a row, is a prefix maximum: opening from a cell
// that was itself reached by a horizontal gap never beats extending that gap (gap_open > 0), so
//     I(b) = max over b' < b of [ max(M'(b'), 0) - go - ge (b - b') ],   M' = max(diagonal, vertical) as the reference computes it,
// where ties go to the nearest b' (the reference prefers "open" on ties) and the cell left of the band counts as M = 0; the counts
// ride along as (count - gaps * b').  A warp scan (5 shuffle steps) gives every lane the prefix of the lanes before it.  The best
// cell is the first one in row-major order with the highest score, as in the sequential loop.  Pairs with more than 32 K band slots
// or a target longer than GW_TMAX stay with the thread-per-pair kernel.
#define GW_TMAX 2048
template <int K>
__global__ void __launch_bounds__(128) gotoh_counts_warp_kernel(const uint8_t *q, const int64_t *q_off, const int32_t *q_len, const uint8_t *t,
                                                                const int64_t *t_off, const int32_t *t_len, int32_t first, int32_t n_end, int32_t k,
                                                                int32_t go, int32_t ge, const int32_t *perm, int32_t *res)
{
    __shared__ int8_t s_aa[256], s_bl[625];
    __shared__ uint8_t s_t[4][GW_TMAX + 8];
    for (int x = threadIdx.x; x < 256; x += blockDim.x) s_aa[x] = c_tab.aa_idx[x];
    for (int x = threadIdx.x; x < 625; x += blockDim.x) s_bl[x] = c_tab.blosum[x];
    __syncthreads();
    typedef unsigned long long u64;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int64_t nw = (int64_t)gridDim.x * 4;
    const int32_t INF = -1000000000;
    const u64 ONE_M = 1ull, ONE_X = 1ull << 21, ONE_G = 1ull << 42;
    const unsigned FULL = 0xffffffffu;
    for (int64_t pi = first + (int64_t)blockIdx.x * 4 + wi; pi < n_end; pi += nw) {
        const int idx = perm[pi];
        const uint8_t *s1 = q + q_off[idx], *s2 = t + t_off[idx];
        const int32_t len1 = q_len[idx], len2 = t_len[idx], rows = len1 + 1, cols = len2 + 1;
        const int32_t dl = len1 > len2 ? len1 - len2 : len2 - len1;
        const int32_t kl = k > dl + 1 ? k : dl + 1, nslots = 2 * kl + 1;
        __syncwarp();
        for (int x = lane; x < len2; x += 32) s_t[wi][x] = s2[x];
        __syncwarp();
        int32_t Mp[K], Dp[K];
        u64 CMp[K], CDp[K];
#pragma unroll
        for (int m = 0; m < K; ++m) Mp[m] = 0, Dp[m] = INF, CMp[m] = 0, CDp[m] = 0;  // row 0: M = 0, no vertical gap
        int32_t best_s = 0, best_pos = 0x7fffffff;
        u64 best_c = 0;
        uint8_t a_next = rows > 1 ? s1[0] : 0;
        for (int32_t i = 1; i < rows; ++i) {
            const uint8_t a = a_next;
            if (i + 1 < rows) a_next = s1[i];
            const int32_t ai = s_aa[a], o = i - kl;
            int32_t nM = __shfl_down_sync(FULL, Mp[0], 1), nD = __shfl_down_sync(FULL, Dp[0], 1);
            u64 nCM = __shfl_down_sync(FULL, CMp[0], 1), nCD = __shfl_down_sync(FULL, CDp[0], 1);
            if (lane == 31) nM = 0, nD = INF, nCM = 0, nCD = 0;
            int32_t Mq[K], Dn[K], key[K];
            u64 CMq[K], CDn[K], cadj[K];
#pragma unroll
            for (int m = 0; m < K; ++m) {
                const int32_t b = lane * K + m, j = o + b;
                const bool valid = j >= 1 && j < cols && b < nslots;
                const int32_t m_top = m + 1 < K ? Mp[m + 1 < K ? m + 1 : 0] : nM, d_top = m + 1 < K ? Dp[m + 1 < K ? m + 1 : 0] : nD;
                const u64 cm_top = m + 1 < K ? CMp[m + 1 < K ? m + 1 : 0] : nCM, cd_top = m + 1 < K ? CDp[m + 1 < K ? m + 1 : 0] : nCD;
                const int32_t d_open = m_top - go - ge, d_ext = d_top - ge;
                int32_t dcur;
                u64 cd;
                if (d_open >= d_ext) dcur = d_open, cd = cm_top + ONE_G;
                else dcur = d_ext, cd = cd_top + ONE_G;
                const uint8_t br = valid ? s_t[wi][j - 1] : (uint8_t)0;
                const int32_t bi = s_aa[br];
                const int32_t sub = (ai >= 0 && bi >= 0) ? (int32_t)s_bl[ai * 25 + bi] : -128;
                int32_t best = Mp[m] + sub;
                u64 cb = CMp[m] + (a == br ? ONE_M : ONE_X);
                if (dcur > best) best = dcur, cb = cd;
                Mq[m] = best, CMq[m] = cb, Dn[m] = dcur, CDn[m] = cd;
                const bool pos = valid && best > 0;  // what a horizontal gap can open from: the cell's M unless the gap state itself wins there
                key[m] = (pos ? best : 0) + ge * b, cadj[m] = (pos ? cb : 0ull) - ONE_G * (u64)(int64_t)b;
            }
            int32_t ak = key[0];
            u64 ac = cadj[0];
#pragma unroll
            for (int m = 1; m < K; ++m)
                if (key[m] >= ak) ak = key[m], ac = cadj[m];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int32_t ok = __shfl_up_sync(FULL, ak, d);
                const u64 oc = __shfl_up_sync(FULL, ac, d);
                if (lane >= d && ok > ak) ak = ok, ac = oc;
            }
            int32_t run_k = __shfl_up_sync(FULL, ak, 1);
            u64 run_c = __shfl_up_sync(FULL, ac, 1);
            if (lane == 0 || -ge > run_k) run_k = -ge, run_c = ONE_G;  // the cell left of the band (slot -1): M = 0, no counts
#pragma unroll
            for (int m = 0; m < K; ++m) {
                const int32_t b = lane * K + m, j = o + b;
                const bool valid = j >= 1 && j < cols && b < nslots;
                const int32_t icur = run_k - go - ge * b;
                const u64 ci = run_c + ONE_G * (u64)(int64_t)b;
                int32_t best = Mq[m];
                u64 cb = CMq[m];
                if (icur > best) best = icur, cb = ci;
                if (valid && best > best_s) best_s = best, best_c = cb, best_pos = i * cols + j;
                const bool keep = valid && best > 0;
                Mp[m] = keep ? best : 0, CMp[m] = keep ? cb : 0ull;
                Dp[m] = valid ? Dn[m] : INF, CDp[m] = valid ? CDn[m] : 0ull;
                if (key[m] >= run_k) run_k = key[m], run_c = cadj[m];
            }
        }
        // the first cell in row-major order among those with the highest score
        const int32_t top = __reduce_max_sync(FULL, best_s);
        const int32_t wpos = __reduce_min_sync(FULL, best_s == top ? best_pos : 0x7fffffff);
        if (best_s == top && best_pos == wpos) {
            int32_t *r = res + (int64_t)idx * 4;
            r[0] = best_s, r[1] = (int32_t)(best_c & 0x1fffff), r[2] = (int32_t)((best_c >> 21) & 0x1fffff), r[3] = (int32_t)(best_c >> 42);
        }
    }
}

// The same recurrences for WIDE bands (a translated hit much shorter or longer than the database protein: band = twice the length
// difference): one warp per pair, the row state of all band slots in shared memory, a row processed in chunks of 64 slots from left
// to right (two slots per lane; the horizontal-gap prefix of a chunk continues from the last lane of the chunk before).  The state
// is updated in place: slot b of row i needs slots b and b + 1 of row i - 1, which a chunk reads before it writes and which later
// chunks have not touched yet.  Slots outside the matrix are never written and keep the "no cell" state they start with.
// Dynamic shared memory per warp: (slot_cap + 2) x 24 bytes + GW_TMAX + 8.
__global__ void __launch_bounds__(128) gotoh_counts_wide_kernel(const uint8_t *q, const int64_t *q_off, const int32_t *q_len, const uint8_t *t,
                                                                const int64_t *t_off, const int32_t *t_len, int32_t first, int32_t n_end, int32_t k,
                                                                int32_t go, int32_t ge, const int32_t *perm, int32_t slot_cap, int32_t *res)
{
    __shared__ int8_t s_aa[256], s_bl[625];
    extern __shared__ __align__(16) uint8_t gw_dyn[];
    for (int x = threadIdx.x; x < 256; x += blockDim.x) s_aa[x] = c_tab.aa_idx[x];
    for (int x = threadIdx.x; x < 625; x += blockDim.x) s_bl[x] = c_tab.blosum[x];
    __syncthreads();
    typedef unsigned long long u64;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int64_t nw = (int64_t)gridDim.x * wpc;
    const size_t per_warp = (size_t)(slot_cap + 2) * 24 + GW_TMAX + 8;
    uint8_t *base = gw_dyn + (size_t)wi * per_warp;
    u64 *CM = reinterpret_cast<u64 *>(base), *CD = CM + (slot_cap + 2);
    int32_t *M = reinterpret_cast<int32_t *>(CD + (slot_cap + 2)), *D = M + (slot_cap + 2);
    uint8_t *s_t = reinterpret_cast<uint8_t *>(D + (slot_cap + 2));
    const int32_t INF = -1000000000;
    const u64 ONE_M = 1ull, ONE_X = 1ull << 21, ONE_G = 1ull << 42;
    const unsigned FULL = 0xffffffffu;
    for (int64_t pi = first + (int64_t)blockIdx.x * wpc + wi; pi < n_end; pi += nw) {
        const int idx = perm[pi];
        const uint8_t *s1 = q + q_off[idx], *s2 = t + t_off[idx];
        const int32_t len1 = q_len[idx], len2 = t_len[idx], rows = len1 + 1, cols = len2 + 1;
        const int32_t dl = len1 > len2 ? len1 - len2 : len2 - len1;
        const int32_t kl = k > dl + 1 ? k : dl + 1, nslots = 2 * kl + 1;
        __syncwarp();
        for (int x = lane; x < len2; x += 32) s_t[x] = s2[x];
        for (int x = lane; x < nslots + 2; x += 32) M[x] = 0, D[x] = INF, CM[x] = 0, CD[x] = 0;
        __syncwarp();
        int32_t best_s = 0, best_pos = 0x7fffffff;
        u64 best_c = 0;
        uint8_t a_next = rows > 1 ? s1[0] : 0;
        for (int32_t i = 1; i < rows; ++i) {
            const uint8_t a = a_next;
            if (i + 1 < rows) a_next = s1[i];
            const int32_t ai = s_aa[a], o = i - kl;
            int32_t b_lo = 1 - o > 0 ? 1 - o : 0, b_hi = cols - 1 - o < nslots - 1 ? cols - 1 - o : nslots - 1;  // valid slots of this row
            if (b_hi < b_lo) continue;
            // the cell left of the first valid one counts as M = 0 without counts
            int32_t run_k = ge * (b_lo - 1);
            u64 run_c = 0ull - ONE_G * (u64)(int64_t)(b_lo - 1);
            for (int32_t cb0 = b_lo; cb0 <= b_hi; cb0 += 64) {
                const int32_t b0 = cb0 + 2 * lane;
                int32_t pm[3], pd[3];
                u64 pcm[3], pcd[3];
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    const int32_t b = b0 + x < nslots + 1 ? b0 + x : nslots + 1;  // slots nslots, nslots + 1: never written ("no cell")
                    pm[x] = M[b], pd[x] = D[b], pcm[x] = CM[b], pcd[x] = CD[b];
                }
                __syncwarp();
                int32_t Mq[2], Dn[2], key[2];
                u64 CMq[2], CDn[2], cadj[2];
                bool valid[2];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const int32_t b = b0 + m, j = o + b;
                    valid[m] = b <= b_hi;
                    const int32_t d_open = pm[m + 1] - go - ge, d_ext = pd[m + 1] - ge;
                    int32_t dcur;
                    u64 cd;
                    if (d_open >= d_ext) dcur = d_open, cd = pcm[m + 1] + ONE_G;
                    else dcur = d_ext, cd = pcd[m + 1] + ONE_G;
                    const uint8_t br = valid[m] ? s_t[j - 1] : (uint8_t)0;
                    const int32_t bi = s_aa[br];
                    const int32_t sub = (ai >= 0 && bi >= 0) ? (int32_t)s_bl[ai * 25 + bi] : -128;
                    int32_t best = pm[m] + sub;
                    u64 cb = pcm[m] + (a == br ? ONE_M : ONE_X);
                    if (dcur > best) best = dcur, cb = cd;
                    Mq[m] = best, CMq[m] = cb, Dn[m] = dcur, CDn[m] = cd;
                    const bool pos = valid[m] && best > 0;
                    key[m] = valid[m] ? (pos ? best : 0) + ge * b : INF;  // slots past the row's end: never a source
                    cadj[m] = (pos ? cb : 0ull) - ONE_G * (u64)(int64_t)b;
                }
                int32_t ak = key[0];
                u64 ac = cadj[0];
                if (key[1] >= ak) ak = key[1], ac = cadj[1];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int32_t ok = __shfl_up_sync(FULL, ak, d);
                    const u64 oc = __shfl_up_sync(FULL, ac, d);
                    if (lane >= d && ok > ak) ak = ok, ac = oc;
                }
                int32_t ek = __shfl_up_sync(FULL, ak, 1);
                u64 ec = __shfl_up_sync(FULL, ac, 1);
                if (lane == 0 || run_k > ek) ek = run_k, ec = run_c;  // what came before this chunk wins only when strictly better
                // the prefix after this chunk, for the next one
                {
                    const int32_t lk = __shfl_sync(FULL, ak, 31);
                    const u64 lc = __shfl_sync(FULL, ac, 31);
                    if (lk >= run_k) run_k = lk, run_c = lc;
                }
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const int32_t b = b0 + m, j = o + b;
                    const int32_t icur = ek - go - ge * b;
                    const u64 ci = ec + ONE_G * (u64)(int64_t)b;
                    int32_t best = Mq[m];
                    u64 cb = CMq[m];
                    if (icur > best) best = icur, cb = ci;
                    if (valid[m]) {
                        if (best > best_s) best_s = best, best_c = cb, best_pos = i * cols + j;
                        const bool keep = best > 0;
                        M[b] = keep ? best : 0, CM[b] = keep ? cb : 0ull, D[b] = Dn[m], CD[b] = CDn[m];
                    }
                    if (key[m] >= ek) ek = key[m], ec = cadj[m];
                }
                __syncwarp();
            }
        }
        const int32_t top = __reduce_max_sync(FULL, best_s);
        const int32_t wpos = __reduce_min_sync(FULL, best_s == top ? best_pos : 0x7fffffff);
        if (best_s == top && best_pos == wpos) {
            int32_t *r = res + (int64_t)idx * 4;
            r[0] = best_s, r[1] = (int32_t)(best_c & 0x1fffff), r[2] = (int32_t)((best_c >> 21) & 0x1fffff), r[3] = (int32_t)(best_c >> 42);
        }
    }
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// Row a8: greedy overlap cull (one warp per segment: the hit under test is compared with the kept list 32 entries at a
// time) and 1-D single-linkage clustering (one thread per segment, a sequential sweep in the given order).
__global__ void kb_cull_kernel(const int32_t *order, const int32_t *g1, const int32_t *g2, const int32_t *starts, const int32_t *ends,
                               double frac, const int64_t *seg_off, int32_t n_seg, int32_t *kept_list, uint8_t *kept)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = warp; s < n_seg; s += n_warps) {
        const int64_t b = seg_off[s];
        const int32_t n = (int32_t)(seg_off[s + 1] - b);
        int32_t *kl = kept_list + b;  // global indices of the kept hits, in evaluation order
        int32_t nk = 0;
        for (int32_t i = lane; i < n; i += 32) kept[b + i] = 0;
        __syncwarp();
        for (int32_t i = 0; i < n; ++i) {
            const int64_t idx = b + order[b + i];
            const int32_t a1 = g1[idx], a2 = g2[idx], st = starts[idx], en = ends[idx], length = en - st;
            if (length <= 0) continue;  // warp-uniform
            bool found = false;
            for (int32_t j = lane; j < nk; j += 32) {
                const int64_t p = kl[j];
                if (g1[p] != a1 || g2[p] != a2) continue;
                const int32_t ks = starts[p], ke = ends[p];
                const int32_t ov = min(en, ke) - max(st, ks), mn = min(length, ke - ks);
                if (ov > 0 && ((double)ov / (double)mn) > frac) found = true;
            }
            if (!__any_sync(0xffffffffu, found)) {
                if (lane == 0) kept[idx] = 1, kl[nk] = (int32_t)idx;
                ++nk;
                __syncwarp();
            }
        }
    }
}

__global__ void kb_cluster_kernel(const int32_t *starts, const int32_t *ends, const int32_t *groups, int32_t tolerance, const int32_t *order,
                                  const int64_t *seg_off, int32_t n_seg, int32_t *cluster_ids)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const int64_t b = seg_off[s];
    const int32_t n = (int32_t)(seg_off[s + 1] - b);
    if (n == 0) return;
    int32_t cur = 0, cur_e = ends[b + order[b]], cur_g = groups[b + order[b]];
    cluster_ids[b + order[b]] = 0;
    for (int32_t i = 1; i < n; ++i) {
        const int64_t idx = b + order[b + i];
        const int32_t st = starts[idx], en = ends[idx], g = groups[idx];
        if (g == cur_g && st <= cur_e + tolerance) cur_e = max(cur_e, en);
        else ++cur, cur_e = en, cur_g = g;
        cluster_ids[idx] = cur;
    }
}

extern "C" {

const char *kb_post_last_error(void) { return g_post_err.c_str(); }

int kb_post_extract(const uint8_t *seqs, int64_t n_seq_bytes, const int64_t *parent_off, int32_t n_parents, const int32_t *indices,
                    const int32_t *starts, const int32_t *ends, const int8_t *strands, int32_t n, uint8_t *out, int64_t out_cap,
                    int64_t *out_off, int32_t *out_len)
{
    if (n < 0 || (n > 0 && (!seqs || !parent_off || !indices || !starts || !ends || !strands || !out_off || !out_len))) return KB_ERR_ARG;
    try {
        ensure_device();
        int64_t total = 0;
        for (int32_t i = 0; i < n; ++i) {
            int32_t L = ends[i] - starts[i];
            if (L < 0 || indices[i] < 0 || indices[i] >= n_parents) {
                g_post_err = "extract: bad interval";
                return KB_ERR_ARG;
            }
            out_off[i] = total, out_len[i] = L, total += L;
        }
        if (total > out_cap) {
            g_post_err = "extract: output buffer too small";
            return KB_ERR_CAPACITY;
        }
        if (total == 0) return KB_OK;
        Dev D;
        const uint8_t *d_seq = D.upload(seqs, (size_t)n_seq_bytes);
        const int64_t *d_poff = D.upload(parent_off, (size_t)n_parents);
        const int32_t *d_idx = D.upload(indices, (size_t)n), *d_st = D.upload(starts, (size_t)n), *d_en = D.upload(ends, (size_t)n);
        const int8_t *d_sd = D.upload(strands, (size_t)n);
        const int64_t *d_oo = D.upload(out_off, (size_t)n);
        uint8_t *d_out = D.alloc<uint8_t>((size_t)total);
        int64_t grid = (total + 255) / 256;
        if (grid > 148 * 16) grid = 148 * 16;
        extract_kernel<<<(unsigned)grid, 256>>>(d_seq, d_poff, d_idx, d_st, d_en, d_sd, d_oo, n, total, d_out);
        PCU(cudaGetLastError());
        PCU(cudaMemcpy(out, d_out, (size_t)total, cudaMemcpyDeviceToHost));
    } catch (const std::string &) {
        return KB_ERR_CUDA;
    }
    return KB_OK;
}

int kb_post_translate(const uint8_t *seqs, int64_t n_seq_bytes, const int64_t *offsets, const int32_t *lengths, const int8_t *frames,
                      int32_t n, int32_t to_stop, uint8_t *out, int64_t out_cap, int64_t *out_off, int32_t *out_len, int64_t *n_out)
{
    if (n < 0 || (n > 0 && (!seqs || !offsets || !lengths || !frames || !out_off || !out_len)) || !n_out) return KB_ERR_ARG;
    *n_out = 0;
    if (n == 0) return KB_OK;
    try {
        ensure_device();
        Dev D;
        const uint8_t *d_seq = D.upload(seqs, (size_t)n_seq_bytes);
        const int64_t *d_off = D.upload(offsets, (size_t)n);
        const int32_t *d_len = D.upload(lengths, (size_t)n);
        const int8_t *d_fr = D.upload(frames, (size_t)n);
        int32_t *d_olen = D.alloc<int32_t>((size_t)n);
        translate_count_kernel<<<(n + 127) / 128, 128>>>(d_seq, d_off, d_len, d_fr, n, to_stop, d_olen);
        PCU(cudaGetLastError());
        PCU(cudaMemcpy(out_len, d_olen, (size_t)n * 4, cudaMemcpyDeviceToHost));
        int64_t total = 0;
        for (int32_t i = 0; i < n; ++i) out_off[i] = total, total += out_len[i];
        *n_out = total;
        if (total > out_cap) {
            g_post_err = "translate: output buffer too small";
            return KB_ERR_CAPACITY;
        }
        if (total == 0) return KB_OK;
        const int64_t *d_oo = D.upload(out_off, (size_t)n);
        uint8_t *d_out = D.alloc<uint8_t>((size_t)total);
        int64_t grid = (total + 255) / 256;
        if (grid > 148 * 16) grid = 148 * 16;
        translate_fill_kernel<<<(unsigned)grid, 256>>>(d_seq, d_off, d_fr, d_oo, n, total, d_out);
        PCU(cudaGetLastError());
        PCU(cudaMemcpy(out, d_out, (size_t)total, cudaMemcpyDeviceToHost));
    } catch (const std::string &) {
        return KB_ERR_CUDA;
    }
    return KB_OK;
}

int kb_post_protein_align(const uint8_t *q, const int64_t *q_off, const int32_t *q_len, const uint8_t *t, const int64_t *t_off,
                          const int32_t *t_len, int32_t n, int32_t k, int32_t gap_open, int32_t gap_extend, int32_t *res)
{
    if (n < 0 || (n > 0 && (!q_off || !q_len || !t_off || !t_len || !res))) return KB_ERR_ARG;
    if (n == 0) return KB_OK;
    try {
        ensure_device();
        std::vector<int64_t> tb_off((size_t)n), row_off((size_t)n);
        int64_t tb_total = 0, row_total = 0, qb = 0, tbytes = 0;
        for (int32_t i = 0; i < n; ++i) {
            if (q_len[i] < 0 || t_len[i] < 0) return KB_ERR_ARG;
            int64_t dl = q_len[i] > t_len[i] ? q_len[i] - t_len[i] : t_len[i] - q_len[i];
            int64_t kl = k > dl + 1 ? k : dl + 1, bw = 2 * kl + 3;
            tb_off[(size_t)i] = tb_total, tb_total += ((int64_t)q_len[i] + 1) * bw;
            row_off[(size_t)i] = row_total, row_total += 4 * bw;
            qb = qb > q_off[i] + q_len[i] ? qb : q_off[i] + q_len[i];
            tbytes = tbytes > t_off[i] + t_len[i] ? tbytes : t_off[i] + t_len[i];
        }
        Dev D;
        const uint8_t *d_q = D.upload(q, (size_t)qb), *d_t = D.upload(t, (size_t)tbytes);
        const int64_t *d_qo = D.upload(q_off, (size_t)n), *d_to = D.upload(t_off, (size_t)n);
        const int32_t *d_ql = D.upload(q_len, (size_t)n), *d_tl = D.upload(t_len, (size_t)n);
        const int64_t *d_tbo = D.upload(tb_off.data(), (size_t)n), *d_ro = D.upload(row_off.data(), (size_t)n);
        uint8_t *d_tb = D.alloc<uint8_t>((size_t)tb_total);
        int32_t *d_rows = D.alloc<int32_t>((size_t)row_total), *d_res = D.alloc<int32_t>((size_t)n * 8);
        gotoh_kernel<<<(n + 63) / 64, 64>>>(d_q, d_qo, d_ql, d_t, d_to, d_tl, n, k, gap_open, gap_extend, d_tbo, d_tb, d_ro, d_rows, d_res);
        PCU(cudaGetLastError());
        PCU(cudaMemcpy(res, d_res, (size_t)n * 32, cudaMemcpyDeviceToHost));
    } catch (const std::string &) {
        return KB_ERR_CUDA;
    }
    return KB_OK;
}

/* _cull_overlaps_kernel (src/kaptive/core/interval.py:698-751), one call for the hits of many assemblies. */
int kb_post_cull_overlaps(const int32_t *order, const int32_t *group1, const int32_t *group2, const int32_t *starts, const int32_t *ends,
                          double max_overlap_fraction, const int64_t *seg_off, int32_t n_seg, uint8_t *kept)
{
    if (n_seg < 0 || !seg_off || (n_seg > 0 && (!order || !group1 || !group2 || !starts || !ends || !kept))) {
        g_post_err = "null argument";
        return KB_ERR_ARG;
    }
    try {
        ensure_device();
        const int64_t n = n_seg ? seg_off[n_seg] : 0;
        if (n == 0) return KB_OK;
        Dev D;
        const int32_t *d_order = D.upload(order, (size_t)n), *d_g1 = D.upload(group1, (size_t)n), *d_g2 = D.upload(group2, (size_t)n);
        const int32_t *d_st = D.upload(starts, (size_t)n), *d_en = D.upload(ends, (size_t)n);
        const int64_t *d_off = D.upload(seg_off, (size_t)n_seg + 1);
        int32_t *d_kl = D.alloc<int32_t>((size_t)n);
        uint8_t *d_kept = D.alloc<uint8_t>((size_t)n);
        const int64_t blocks = std::min<int64_t>(((int64_t)n_seg + 3) / 4, 148 * 16);
        kb_cull_kernel<<<(unsigned)blocks, 128>>>(d_order, d_g1, d_g2, d_st, d_en, max_overlap_fraction, d_off, n_seg, d_kl, d_kept);
        PCU(cudaGetLastError());
        PCU(cudaMemcpy(kept, d_kept, (size_t)n, cudaMemcpyDeviceToHost));
    } catch (const std::string &) {
        return KB_ERR_CUDA;
    }
    return KB_OK;
}

/* _cluster_kernel (src/kaptive/core/interval.py:595-639), one call for the hits of many assemblies; ids restart at 0 per segment. */
int kb_post_cluster(const int32_t *starts, const int32_t *ends, const int32_t *groups, int32_t tolerance, const int32_t *order,
                    const int64_t *seg_off, int32_t n_seg, int32_t *cluster_ids)
{
    if (n_seg < 0 || !seg_off || (n_seg > 0 && (!order || !groups || !starts || !ends || !cluster_ids))) {
        g_post_err = "null argument";
        return KB_ERR_ARG;
    }
    try {
        ensure_device();
        const int64_t n = n_seg ? seg_off[n_seg] : 0;
        if (n == 0) return KB_OK;
        Dev D;
        const int32_t *d_st = D.upload(starts, (size_t)n), *d_en = D.upload(ends, (size_t)n), *d_g = D.upload(groups, (size_t)n);
        const int32_t *d_order = D.upload(order, (size_t)n);
        const int64_t *d_off = D.upload(seg_off, (size_t)n_seg + 1);
        int32_t *d_ids = D.alloc<int32_t>((size_t)n);
        kb_cluster_kernel<<<(unsigned)((n_seg + 127) / 128), 128>>>(d_st, d_en, d_g, tolerance, d_order, d_off, n_seg, d_ids);
        PCU(cudaGetLastError());
        PCU(cudaMemcpy(cluster_ids, d_ids, (size_t)n * 4, cudaMemcpyDeviceToHost));
    } catch (const std::string &) {
        return KB_ERR_CUDA;
    }
    return KB_OK;
}

}  // extern "C"


// ---------------------------------------------------------------------------------------------------------------
// Typing numerics for a batch (internal, called by kb_type.cpp): translate every item from the packed batch, then align each
// protein with the database translation of its gene (banded local Gotoh, the kernel above).  Device-resident inputs (batch,
// translations); items are processed in chunks so that the traceback scratch stays bounded.  prot_len: n; res: n x 8.
int kb_post_type_numerics(const KbBatchView &bv, int device, const int32_t *ctg, const int32_t *ts, const int32_t *te, const int8_t *strand,
                          const int8_t *frame, const int32_t *gene, int64_t n, const uint8_t *d_trans, const int64_t *d_trans_off,
                          const int32_t *h_trans_len, int32_t k, int32_t go, int32_t ge, int32_t *prot_len, int32_t *res)
{
    if (n <= 0) return KB_OK;
    if (n > 0x7fffffff) return KB_ERR_LIMIT;
    cudaStream_t st = nullptr;
    std::vector<void *> owned;
    auto dalloc = [&](size_t bytes) {
        void *p = nullptr;
        PCU(cudaMallocAsync(&p, bytes ? bytes : 16, st));
        owned.push_back(p);
        return p;
    };
    auto up = [&](const void *h, size_t bytes) {
        void *p = dalloc(bytes);
        PCU(cudaMemcpyAsync(p, h, bytes, cudaMemcpyHostToDevice, st));
        return p;
    };
    int rc = KB_OK;
    const bool dbg = getenv("KAPTIVE_B200_DEBUG_TYPE") != nullptr;
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_0 = dbg ? now_ms() : 0;
    try {
        PCU(cudaSetDevice(device));
        ensure_device();
        PCU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        std::vector<int64_t> aa_off((size_t)n), t_off((size_t)n), row_off((size_t)n);
        std::vector<int32_t> t_len((size_t)n);
        int64_t aa_total = 0;
        int32_t gmax = 0;
        for (int64_t i = 0; i < n; ++i) {
            const int32_t L = te[i] - ts[i];
            aa_off[(size_t)i] = aa_total, aa_total += (L > 0 ? L / 3 : 0) + 1;
            gmax = gene[i] > gmax ? gene[i] : gmax;
        }
        std::vector<int64_t> table((size_t)gmax + 2);
        PCU(cudaMemcpyAsync(table.data(), d_trans_off, ((size_t)gmax + 2) * 8, cudaMemcpyDeviceToHost, st));
        const int32_t *d_ctg = (const int32_t *)up(ctg, (size_t)n * 4), *d_ts = (const int32_t *)up(ts, (size_t)n * 4), *d_te = (const int32_t *)up(te, (size_t)n * 4);
        const int8_t *d_st = (const int8_t *)up(strand, (size_t)n), *d_fr = (const int8_t *)up(frame, (size_t)n);
        const int64_t *d_ao = (const int64_t *)up(aa_off.data(), (size_t)n * 8);
        uint8_t *d_aa = (uint8_t *)dalloc((size_t)aa_total);
        int32_t *d_pl = (int32_t *)dalloc((size_t)n * 4);
        type_translate_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(bv, d_ctg, d_ts, d_te, d_st, d_fr, d_ao, n, d_aa, d_pl);
        PCU(cudaGetLastError());
        PCU(cudaMemcpyAsync(prot_len, d_pl, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        PCU(cudaStreamSynchronize(st));  // the band of every pair depends on the length of its translated hit
        const double t_tr = dbg ? now_ms() : 0;
        // pairs by kernel, then by band width (widest first): 0 = thread per pair (more than 2048 band slots or a target over GW_TMAX residues:
        // a warp's tile is as wide as its widest pair), 1 / 2 / 3 = warp per pair with the row state in shared memory (up to 2048 / 1024 / 512
        // slots: 1 / 2 / 3 CTAs per SM), 4 / 5 = warp per pair with the row state in registers (up to 128 / 64 slots)
        std::vector<int32_t> bwv((size_t)n), perm((size_t)n);
        std::vector<uint8_t> cls((size_t)n);
        const bool use_warp = go > 0 && ge > 0 && !(getenv("KAPTIVE_B200_GOTOH_WARP") && getenv("KAPTIVE_B200_GOTOH_WARP")[0] == '0');
        int64_t n_cls[6] = {0, 0, 0, 0, 0, 0};
        for (int64_t i = 0; i < n; ++i) {
            const int32_t l1 = prot_len[i], l2 = h_trans_len[gene[i]];
            const int64_t dl = l1 > l2 ? l1 - l2 : l2 - l1, kl = k > dl + 1 ? k : dl + 1, slots = 2 * kl + 1;
            bwv[(size_t)i] = (int32_t)(2 * kl + 3);
            const int c = (!use_warp || slots > 2048 || l2 > GW_TMAX) ? 0 : (slots > 1024 ? 1 : (slots > 512 ? 2 : (slots > 128 ? 3 : (slots > 64 ? 4 : 5))));
            cls[(size_t)i] = (uint8_t)c, ++n_cls[c];
            t_len[(size_t)i] = l2, t_off[(size_t)i] = table[(size_t)gene[i]];
            perm[(size_t)i] = (int32_t)i;
        }
        {
            // order: kernel class, then band width and hit length descending (the balance of the kernels needs no finer order than 4096
            // steps of either); a stable three-pass radix sort of 27-bit keys instead of a comparison sort of 300,000 pairs
            std::vector<uint32_t> key((size_t)n), key2((size_t)n);
            std::vector<int32_t> perm2((size_t)n);
            for (int64_t i = 0; i < n; ++i) {
                const uint32_t b = (uint32_t)std::min<int32_t>(bwv[(size_t)i], 4095), l = (uint32_t)std::min<int32_t>(std::max<int32_t>(prot_len[i], 0), 4095);
                key[(size_t)i] = (uint32_t)cls[(size_t)i] << 24 | (4095u - b) << 12 | (4095u - l);
            }
            for (int pass = 0; pass < 3; ++pass) {
                const int sh = 9 * pass;
                int64_t cnt[513] = {0};
                for (int64_t i = 0; i < n; ++i) ++cnt[((key[(size_t)i] >> sh) & 511u) + 1];
                for (int x = 0; x < 512; ++x) cnt[x + 1] += cnt[x];
                for (int64_t i = 0; i < n; ++i) {
                    const int64_t o = cnt[(key[(size_t)i] >> sh) & 511u]++;
                    key2[(size_t)o] = key[(size_t)i], perm2[(size_t)o] = perm[(size_t)i];
                }
                key.swap(key2), perm.swap(perm2);
            }
        }
        const int64_t n_thr = n_cls[0];
        const int64_t n_warp = (n_thr + 31) / 32;
        std::vector<int64_t> tile_off((size_t)n_warp + 1);
        std::vector<int32_t> tile_bw((size_t)n_warp + 1);
        int64_t row_total = 0;  // in elements of one array
        for (int64_t w = 0; w < n_warp; ++w) {
            int32_t widest = 0;
            for (int64_t x = w * 32; x < std::min<int64_t>(n_thr, w * 32 + 32); ++x) widest = std::max(widest, bwv[(size_t)perm[(size_t)x]]);
            tile_bw[(size_t)w] = widest;
            tile_off[(size_t)w] = row_total, row_total += (int64_t)widest * 32;
        }
        (void)row_off;
        const int64_t *d_to = (const int64_t *)up(t_off.data(), (size_t)n * 8), *d_tile = (const int64_t *)up(tile_off.data(), ((size_t)n_warp + 1) * 8);
        const int32_t *d_tl = (const int32_t *)up(t_len.data(), (size_t)n * 4), *d_perm = (const int32_t *)up(perm.data(), (size_t)n * 4);
        const int32_t *d_tbw = (const int32_t *)up(tile_bw.data(), ((size_t)n_warp + 1) * 4);
        int32_t *d_rows = (int32_t *)dalloc((size_t)row_total * 4 * 4), *d_res = (int32_t *)dalloc((size_t)n * 16);
        unsigned long long *d_cnt = (unsigned long long *)dalloc((size_t)row_total * 4 * 8);
        const double t_prep = dbg ? now_ms() : 0;
        if (n_thr > 0)
            gotoh_counts_kernel<<<(unsigned)((n_thr + 63) / 64), 64, 0, st>>>(d_aa, d_ao, d_pl, d_trans, d_to, d_tl, (int32_t)n_thr, k, go, ge, d_perm, d_tile,
                                                                             d_tbw, d_rows, d_cnt, d_res);
        {
            auto grid = [](int64_t pairs, int per_sm) { return (unsigned)std::min<int64_t>((pairs + 3) / 4, 148 * (int64_t)per_sm); };
            int64_t lo = n_thr;
            for (int c = 1; c <= 3; ++c) {  // wide bands: shared-memory rows
                const int64_t hi = lo + n_cls[c];
                if (hi > lo) {
                    const int slot_cap = c == 1 ? 2048 : (c == 2 ? 1024 : 512);
                    const size_t smem = 4 * ((size_t)(slot_cap + 2) * 24 + GW_TMAX + 8);
                    static bool attr_set[4] = {false, false, false, false};
                    if (!attr_set[c]) {
                        PCU(cudaFuncSetAttribute(gotoh_counts_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (2050 * 24 + GW_TMAX + 8)));
                        attr_set[c] = true;
                    }
                    gotoh_counts_wide_kernel<<<grid(hi - lo, c), 128, smem, st>>>(d_aa, d_ao, d_pl, d_trans, d_to, d_tl, (int32_t)lo, (int32_t)hi, k, go,
                                                                                              ge, d_perm, slot_cap, d_res);
                }
                lo = hi;
            }
            if (n_cls[4] > 0)
                gotoh_counts_warp_kernel<4><<<grid(n_cls[4], 16), 128, 0, st>>>(d_aa, d_ao, d_pl, d_trans, d_to, d_tl, (int32_t)lo, (int32_t)(lo + n_cls[4]), k, go, ge,
                                                                               d_perm, d_res);
            lo += n_cls[4];
            if (n_cls[5] > 0)
                gotoh_counts_warp_kernel<2><<<grid(n_cls[5], 16), 128, 0, st>>>(d_aa, d_ao, d_pl, d_trans, d_to, d_tl, (int32_t)lo, (int32_t)(lo + n_cls[5]), k, go, ge,
                                                                               d_perm, d_res);
        }
        PCU(cudaGetLastError());
        std::vector<int32_t> r4((size_t)n * 4);
        PCU(cudaMemcpyAsync(r4.data(), d_res, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
        PCU(cudaStreamSynchronize(st));
        if (dbg)
            fprintf(stderr, "[type numerics] %lld pairs (thread %lld, wide %lld + %lld + %lld, registers %lld + %lld): translate %.1f ms, host prep %.1f ms, gotoh %.1f ms\n",
                    (long long)n, (long long)n_cls[0], (long long)n_cls[1], (long long)n_cls[2], (long long)n_cls[3], (long long)n_cls[4], (long long)n_cls[5], t_tr - t_0,
                    t_prep - t_tr, now_ms() - t_prep);
        for (int64_t i = 0; i < n; ++i) {  // the n x 8 layout of kb_post_protein_align; coordinates are not produced here
            int32_t *o = res + i * 8;
            o[0] = r4[(size_t)i * 4], o[1] = r4[(size_t)i * 4 + 1], o[2] = r4[(size_t)i * 4 + 2], o[3] = r4[(size_t)i * 4 + 3];
            o[4] = o[5] = o[6] = o[7] = -1;
        }
    } catch (const std::string &) {
        rc = KB_ERR_CUDA;
    }
    for (void *p : owned) cudaFreeAsync(p, st);
    if (st) cudaStreamSynchronize(st), cudaStreamDestroy(st);
    return rc;
}

// small device-memory helpers for kb_type.cpp (host-only translation unit)
void *kb_type_dev_upload(int device, const void *h, size_t bytes)
{
    void *d = nullptr;
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&d, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    if (h && bytes && cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(d);
        return nullptr;
    }
    return d;
}
void kb_type_dev_free(int device, void *p)
{
    if (p && cudaSetDevice(device) == cudaSuccess) cudaFree(p);
}
int kb_type_dev_download(int device, void *h, const void *dptr, size_t bytes)
{
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    return cudaMemcpy(h, dptr, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}
