// kb_api.cu -- the C-ABI of libkaptive_b200.so (include/kaptive_b200.h) and the host-side
// orchestration of one mapping call: scan -> sort -> groups -> occurrence filter (+ census
// where needed) -> chain -> align -> finalise -> SoA.  Every call runs on its own CUDA
// stream; stage times come from CUDA events recorded on that stream.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstring>
#include <string>
#include <vector>
#include "kb_final.cuh"
#include "kb_host.h"
#include "kb_kernels.h"

extern "C" int kb_packed_layout(const int32_t *contig_len, int64_t n_contigs, int64_t *contig_soff, int64_t *storage_bases);

// ---- prototypes of the launchers in kb_pipeline.cu
void kb_launch_pack(const uint8_t *, int64_t, const int64_t *, const int64_t *, const int32_t *, int32_t, int32_t, int64_t, int64_t,
                    uint32_t *, uint32_t *, cudaStream_t);
void kb_launch_occ(const uint64_t *, const uint32_t *, int64_t, int64_t, uint32_t *, int32_t, int32_t *, cudaStream_t);
void kb_launch_group_flag(const uint64_t *, int64_t, uint8_t *, cudaStream_t);
size_t kb_sort_pairs_temp_bytes(int64_t);
cudaError_t kb_sort_pairs(void *, size_t, const uint64_t *, uint64_t *, const uint32_t *, uint32_t *, int64_t, int, cudaStream_t);
size_t kb_sort_keys32_temp_bytes(int64_t);
cudaError_t kb_sort_keys32(void *, size_t, const uint32_t *, uint32_t *, int64_t, int, cudaStream_t);
size_t kb_select_temp_bytes(int64_t);
cudaError_t kb_select_flagged(void *, size_t, const uint8_t *, int64_t *, int64_t *, int64_t, cudaStream_t);
size_t kb_scan_temp_bytes(int64_t);
cudaError_t kb_exclusive_sum(void *, size_t, const int32_t *, int64_t *, int64_t, cudaStream_t);
size_t kb_rle_temp_bytes(int64_t);
cudaError_t kb_rle(void *, size_t, const uint32_t *, uint32_t *, uint32_t *, int64_t *, int64_t, cudaStream_t);
void kb_launch_chain(const KbIndexView &, const KbBatchView &, const uint64_t *, const uint32_t *, const int64_t *, int64_t, int64_t,
                     const uint16_t *, const int32_t *, const KbChainWork &, uint64_t *, uint64_t *, KbGroupInfo *, KbChainRec *,
                     unsigned long long *, int64_t, const int32_t *, cudaStream_t);
void kb_launch_align(const KbIndexView &, const KbBatchView &, const KbChainRec *, int64_t, const KbGroupInfo *, const uint64_t *,
                     uint64_t *, uint8_t *, size_t, int, KbRawHit *, int64_t, uint32_t *, int64_t, unsigned long long *,
                     unsigned long long *, const int32_t *, const unsigned long long *, cudaStream_t);
size_t kb_sort_keys64_temp_bytes(int64_t);
cudaError_t kb_sort_keys64(void *, size_t, const uint64_t *, uint64_t *, int64_t, int, cudaStream_t);
cudaError_t kb_rle64(void *, size_t, const uint64_t *, uint64_t *, uint32_t *, int64_t *, int64_t, cudaStream_t);
void kb_launch_census_keys(const uint32_t *, const int32_t *, int64_t, int32_t, uint64_t *, cudaStream_t);
size_t kb_census_hist_bytes(int);
void kb_launch_census_quantile(const uint64_t *, const uint32_t *, const int64_t *, int64_t, uint32_t *, int, const int32_t *, int, int32_t, float,
                               int32_t, int32_t, int32_t *, unsigned long long *, cudaStream_t);
size_t kb_band_scratch_bytes();
size_t kb_sizeof_job();
size_t kb_sizeof_plan();
void kb_launch_stage_plan(const KbIndexView &, const KbBatchView &, const KbChainRec *, int64_t, const KbGroupInfo *, const uint64_t *,
                          uint64_t *, int32_t *, void *, void *, const KbStageLists &, int32_t *, unsigned long long *, cudaStream_t);
size_t kb_r16_sort_temp_bytes(int64_t);
void kb_launch_stage_dp(const KbIndexView &, const KbBatchView &, void *, const KbStageLists &, uint8_t *, int, uint8_t *, size_t, int, uint8_t *,
                        int, uint32_t *, int64_t, unsigned long long *, cudaStream_t);
void kb_launch_stage_assemble(const KbIndexView &, const KbBatchView &, const KbChainRec *, int64_t, const KbGroupInfo *, const void *,
                              const void *, const uint32_t *, uint32_t *, int64_t, KbRawHit *, int64_t, uint32_t *, int64_t, int32_t *,
                              unsigned long long *, cudaStream_t);
void kb_launch_rawkey(const KbRawHit *, int64_t, uint64_t *, uint32_t *, unsigned long long *, cudaStream_t);
void kb_launch_gather_raw(const KbRawHit *, const uint32_t *, int64_t, KbRawHit *, cudaStream_t);
void kb_launch_finalize(const kb_params_t &, KbRawHit *, int64_t, const KbGroupInfo *, int32_t *, uint64_t *, int32_t *, cudaStream_t);
void kb_launch_scatter(const KbRawHit *, const int32_t *, const int64_t *, int64_t, const KbGroupInfo *, const KbBatchView &,
                       const kb_hits_t &, cudaStream_t);
void kb_launch_group_nseed(const KbGroupInfo *, int64_t, int32_t *, cudaStream_t);
void kb_launch_dump_anchors(const KbBatchView &, const KbGroupInfo *, int64_t, const int64_t *, const uint32_t *, const int32_t *,
                            const int64_t *, int32_t *, cudaStream_t);

static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
#define CU(x)                                                                                         \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) {                                                                      \
            char b_[512];                                                                             \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            throw std::string(b_);                                                                    \
        }                                                                                             \
    } while (0)

// Workspace arenas: one big device allocation per concurrent mapping call, kept for the life of the process and handed
// out by bumping a pointer, so a steady-state call never reaches the CUDA allocator (whose latency is anything between
// microseconds and a fraction of a second once pools grow or trim).  A call that needs more than its arena holds falls
// back to cudaMallocAsync for the excess and the arena is regrown when the call returns it.
struct Arena {
    uint8_t *base = nullptr;
    size_t cap = 0, off = 0, demand = 0;
    int device = 0;
};
static std::mutex g_arena_mu;
static std::vector<Arena *> g_arenas;  // idle arenas
static Arena *arena_acquire(int device)
{
    std::lock_guard<std::mutex> lk(g_arena_mu);
    int best = -1;
    for (int i = 0; i < (int)g_arenas.size(); ++i)
        if (g_arenas[(size_t)i]->device == device && (best < 0 || g_arenas[(size_t)i]->cap > g_arenas[(size_t)best]->cap)) best = i;
    Arena *a;
    if (best >= 0) a = g_arenas[(size_t)best], g_arenas.erase(g_arenas.begin() + best);
    else a = new Arena(), a->device = device;
    a->off = a->demand = 0;
    return a;
}
static void arena_release(Arena *a)  // the caller has synchronised the stream that used it
{
    if (!a) return;
    if (a->demand > a->cap) {
        if (a->base) cudaFree(a->base);
        a->base = nullptr, a->cap = 0;
        size_t want = a->demand + a->demand / 8 + (64u << 20);
        if (cudaMalloc((void **)&a->base, want) == cudaSuccess) a->cap = want;
        else (void)cudaGetLastError();  // stay on the stream-ordered allocator
    }
    std::lock_guard<std::mutex> lk(g_arena_mu);
    g_arenas.push_back(a);
}

struct DevPool {  // per-call scratch: bump allocation from an arena when it has one, stream-ordered allocations otherwise
    cudaStream_t st = nullptr;
    Arena *arena = nullptr;
    std::vector<void *> ptrs;  // stream-ordered allocations to free
    template <class T>
    T *get(size_t n)
    {
        void *p = nullptr;
        size_t bytes = n * sizeof(T);
        if (bytes == 0) bytes = 16;
        if (arena) {
            const size_t al = (bytes + 255) & ~(size_t)255;
            arena->demand += al;
            if (arena->off + al <= arena->cap) {
                p = arena->base + arena->off;
                arena->off += al;
                return (T *)p;
            }
        }
        CU(cudaMallocAsync(&p, bytes, st));
        ptrs.push_back(p);
        return (T *)p;
    }
    void release(void *p)  // arena memory simply stays put until the call ends
    {
        for (size_t i = 0; i < ptrs.size(); ++i)
            if (ptrs[i] == p) {
                cudaFreeAsync(p, st);
                ptrs.erase(ptrs.begin() + (long)i);
                return;
            }
    }
    void clear()
    {
        for (size_t i = ptrs.size(); i-- > 0;) cudaFreeAsync(ptrs[i], st);
        ptrs.clear();
    }
    ~DevPool() { clear(); }
};

struct kb_index {
    KbHostIndex host;
    int device = 0;
    uint8_t *d_blob = nullptr;
    KbIndexView view;
};

struct kb_batch {
    KbHostBatchLayout L;
    int device = 0;
    int n_sm = 148;
    uint8_t *d_blob = nullptr;  // layout arrays
    uint32_t *seq2 = nullptr, *nmask = nullptr;
    KbBatchView view;
    std::atomic<int64_t> last_anchor_count{0};  // sizing hint for repeated calls on the same batch (the only mutable field; relaxed)
};

struct kb_result {
    int device = 0;
    cudaStream_t st = nullptr;
    int64_t n_hits = 0, n_cigar = 0;
    kb_hits_t d;  // device SoA
    uint32_t *pool = nullptr;
    std::vector<void *> owned;
    float stage_ms[KB_N_STAGES] = {0};
    int64_t counters[16] = {0};
    std::vector<int32_t> mid_occ;
    // stage dumps
    int32_t *d_anchor_dump = nullptr;
    int64_t n_anchor_dump = 0;
    std::vector<int32_t> chain_dump;
};

// internal: the device view of a batch for the typing numerics (kb_type.cpp)
const KbBatchView *kb_batch_view_internal(const kb_batch *b, int *device)
{
    if (!b) return nullptr;
    if (device) *device = b->device;
    return &b->view;
}

static void setup_mempool(int device)
{
    cudaMemPool_t mp;
    if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr);
    }
}

template <class T>
static size_t put_blob(std::vector<uint8_t> &blob, const std::vector<T> &v)
{
    size_t off = (blob.size() + 255) & ~(size_t)255;
    blob.resize(off + v.size() * sizeof(T) + 16);
    if (!v.empty()) memcpy(blob.data() + off, v.data(), v.size() * sizeof(T));
    return off;
}

static int index_upload(kb_index *ix, int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(KB_ERR_CUDA, "no CUDA device: libkaptive_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(KB_ERR_ARG, "bad device ordinal");
    try {
        CU(cudaSetDevice(device));
        setup_mempool(device);
        const KbHostIndex &h = ix->host;
        std::vector<uint8_t> blob;
        size_t o_ht = put_blob(blob, h.ht), o_ent = put_blob(blob, h.ent), o_gl = put_blob(blob, h.gene_len);
        size_t o_gn = put_blob(blob, h.gene_nmin), o_gmo = put_blob(blob, h.gene_min_off), o_gq = put_blob(blob, h.gm_qpos_z);
        size_t o_go = put_blob(blob, h.gm_qocc), o_gh = put_blob(blob, h.gene_hash), o_gso = put_blob(blob, h.gene_seq_off);
        size_t o_f = put_blob(blob, h.gseq_fwd), o_r = put_blob(blob, h.gseq_rev);
        // presence bitmap over the low hash bits (>= 32 bits per distinct minimizer, at least 2 MB): the scan asks it first
        uint32_t bbits = 1u << 24;
        while ((uint64_t)bbits < (uint64_t)h.ht.size() * 16 && bbits < (1u << 30)) bbits <<= 1;
        std::vector<uint32_t> bloom(bbits / 32, 0);
        for (uint64_t e : h.ht)
            if (e) {
                uint32_t hv = (uint32_t)(e >> KB_HT_KEY_SHIFT) & (bbits - 1);
                bloom[hv >> 5] |= 1u << (hv & 31);
            }
        size_t o_bl = put_blob(blob, bloom);
        CU(cudaMalloc((void **)&ix->d_blob, blob.size()));
        CU(cudaMemcpy(ix->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
        KbIndexView &v = ix->view;
        v.p = h.p, v.n_genes = h.n_genes, v.n_entries = (int64_t)h.ent.size(), v.ht_mask = (uint32_t)h.ht.size() - 1;
        uint8_t *b = ix->d_blob;
        v.ht = (const uint64_t *)(b + o_ht), v.ent = (const KbEntry *)(b + o_ent), v.gene_len = (const int32_t *)(b + o_gl);
        v.gene_nmin = (const int32_t *)(b + o_gn), v.gene_min_off = (const int64_t *)(b + o_gmo), v.gm_qpos_z = (const uint32_t *)(b + o_gq);
        v.gm_qocc = (const int32_t *)(b + o_go), v.gene_hash = (const uint32_t *)(b + o_gh), v.gene_seq_off = (const int64_t *)(b + o_gso);
        v.gseq_fwd = b + o_f, v.gseq_rev = b + o_r;
        v.bloom = (const uint32_t *)(b + o_bl), v.bloom_mask = bbits - 1;
        ix->device = device;
    } catch (const std::string &e) {
        return fail(KB_ERR_CUDA, e);
    }
    return KB_OK;
}

extern "C" {

const char *kb_last_error(void) { return g_err.c_str(); }
int kb_version(void) { return 100; }
int kb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int kb_index_create(const uint8_t *gene_seqs, const int64_t *offsets, const int32_t *lengths, int32_t n_genes, const kb_params_t *params,
                    int device, kb_index_t **out)
{
    if (!out || !params || (n_genes > 0 && (!gene_seqs || !offsets || !lengths))) return fail(KB_ERR_ARG, "null argument");
    if (params->max_gap >= KB_CTG_VGAP - 64 || params->bw >= KB_CTG_VGAP - 64) return fail(KB_ERR_LIMIT, "max_gap/bw must stay below 8128");
    if (params->max_sw_cells <= 0 || params->max_sw_cells > 100000000) return fail(KB_ERR_LIMIT, "max_sw_cells out of range (1 .. 100,000,000 = minimap2 max_sw_mat)");
    kb_index *ix = new kb_index();
    std::string err = ix->host.build(gene_seqs, offsets, lengths, n_genes, *params);
    if (!err.empty()) {
        delete ix;
        return fail(KB_ERR_LIMIT, err);
    }
    int rc = index_upload(ix, device);
    if (rc) {
        delete ix;
        return rc;
    }
    *out = ix;
    return KB_OK;
}

void kb_index_destroy(kb_index_t *ix)
{
    if (!ix) return;
    if (ix->d_blob) {
        cudaSetDevice(ix->device);
        cudaFree(ix->d_blob);
    }
    delete ix;
}
int32_t kb_index_n_genes(const kb_index_t *ix) { return ix ? ix->host.n_genes : 0; }
int64_t kb_index_n_minimizers(const kb_index_t *ix) { return ix ? (int64_t)ix->host.ent.size() : 0; }
int64_t kb_index_serialized_size(const kb_index_t *ix) { return ix ? ix->host.serialized_size() : 0; }
int kb_index_serialize(const kb_index_t *ix, uint8_t *buf, int64_t cap)
{
    if (!ix || !buf) return fail(KB_ERR_ARG, "null argument");
    if (cap < ix->host.serialized_size()) return fail(KB_ERR_CAPACITY, "buffer smaller than kb_index_serialized_size()");
    ix->host.serialize(buf);
    return KB_OK;
}
int kb_index_deserialize(const uint8_t *buf, int64_t n, int device, kb_index_t **out)
{
    if (!buf || !out) return fail(KB_ERR_ARG, "null argument");
    kb_index *ix = new kb_index();
    std::string err = ix->host.deserialize(buf, n);
    if (!err.empty()) {
        delete ix;
        return fail(KB_ERR_ARG, err);
    }
    int rc = index_upload(ix, device);
    if (rc) {
        delete ix;
        return rc;
    }
    *out = ix;
    return KB_OK;
}

}  // extern "C"

// Common part of the two batch constructors: layout, device arrays, view.  `fill(b, st, d_off)` puts the sequence in place.
template <class Fill>
static int batch_create_common(const int64_t *contig_off, const int32_t *contig_len, const int32_t *asm_contig_start, int32_t n_asm, int device,
                               kb_batch_t **out, Fill fill)
{
    if (!out || !asm_contig_start || n_asm < 0) return fail(KB_ERR_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(KB_ERR_CUDA, "no CUDA device: libkaptive_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(KB_ERR_ARG, "bad device ordinal");
    kb_batch *b = new kb_batch();
    std::string err = b->L.build(contig_off, contig_len, asm_contig_start, n_asm);
    if (!err.empty()) {
        delete b;
        return fail(KB_ERR_LIMIT, err);
    }
    cudaStream_t st = nullptr;
    try {
        CU(cudaSetDevice(device));
        setup_mempool(device);
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        b->n_sm = prop.multiProcessorCount;
        b->device = device;
        const KbHostBatchLayout &L = b->L;
        std::vector<int64_t> off_v;
        if (contig_off) off_v.assign(contig_off, contig_off + L.n_ctg);
        std::vector<uint8_t> blob;
        size_t o_so = put_blob(blob, L.ctg_soff), o_len = put_blob(blob, L.ctg_len), o_asm = put_blob(blob, L.ctg_asm);
        size_t o_vs = put_blob(blob, L.ctg_vstart), o_acs = put_blob(blob, L.asm_ctg_start), o_cc = put_blob(blob, L.chunk_ctg);
        size_t o_cs = put_blob(blob, L.chunk_start), o_off = put_blob(blob, off_v);
        CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        // stream-ordered allocations: neither creating nor destroying a batch synchronises the device, so batches of
        // concurrent calls (kb_map_assemblies' slabs) overlap with each other's kernels
        CU(cudaMallocAsync((void **)&b->d_blob, blob.size(), st));
        CU(cudaMemcpyAsync(b->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));
        CU(cudaMallocAsync((void **)&b->seq2, (size_t)(L.storage_bases >> 4) * 4 + 64, st));
        CU(cudaMallocAsync((void **)&b->nmask, (size_t)(L.storage_bases >> 5) * 4 + 64, st));
        uint8_t *db = b->d_blob;
        KbBatchView &v = b->view;
        v.n_asm = L.n_asm, v.n_ctg = L.n_ctg, v.n_chunks = (int64_t)L.chunk_ctg.size(), v.total_bases = L.total_bases;
        v.seq2 = b->seq2, v.nmask = b->nmask;
        v.ctg_soff = (const int64_t *)(db + o_so), v.ctg_len = (const int32_t *)(db + o_len), v.ctg_asm = (const int32_t *)(db + o_asm);
        v.ctg_vstart = (const int32_t *)(db + o_vs), v.asm_ctg_start = (const int32_t *)(db + o_acs);
        v.chunk_ctg = (const int32_t *)(db + o_cc), v.chunk_start = (const int32_t *)(db + o_cs);
        fill(b, st, (const int64_t *)(db + o_off));
        CU(cudaStreamSynchronize(st));
        CU(cudaStreamDestroy(st));
    } catch (const std::string &e) {
        if (st) cudaStreamSynchronize(st), cudaStreamDestroy(st);
        if (b->d_blob) cudaFreeAsync(b->d_blob, 0);
        if (b->seq2) cudaFreeAsync(b->seq2, 0);
        if (b->nmask) cudaFreeAsync(b->nmask, 0);
        delete b;
        return fail(KB_ERR_CUDA, e);
    }
    *out = b;
    return KB_OK;
}

extern "C" {

int kb_batch_create(const uint8_t *contig_seqs, const int64_t *contig_off, const int32_t *contig_len, const int32_t *asm_contig_start,
                    int32_t n_asm, int device, kb_batch_t **out)
{
    if (n_asm > 0 && asm_contig_start && asm_contig_start[n_asm] > 0 && (!contig_seqs || !contig_off || !contig_len)) return fail(KB_ERR_ARG, "null argument");
    return batch_create_common(contig_off, contig_len, asm_contig_start, n_asm, device, out, [&](kb_batch *b, cudaStream_t st, const int64_t *d_off) {
        const KbHostBatchLayout &L = b->L;
        const KbBatchView &v = b->view;
        // lead-in / tail padding groups
        CU(cudaMemsetAsync(b->seq2, 0, (size_t)(L.storage_bases >> 4) * 4 + 64, st));
        CU(cudaMemsetAsync(b->nmask, 0xff, (size_t)(L.storage_bases >> 5) * 4 + 64, st));
        // pack in slabs of <= 1 GiB of ASCII so the transient staging buffer stays bounded
        const int64_t slab = (int64_t)1 << 30;
        uint8_t *d_ascii = nullptr;
        int c0 = 0;
        int64_t staged_cap = 0;
        try {
            while (c0 < L.n_ctg) {
                int64_t lo = contig_off[c0], hi = lo;
                int c1 = c0;
                while (c1 < L.n_ctg) {
                    int64_t a = contig_off[c1], e = a + L.ctg_len[c1];
                    int64_t nlo = a < lo ? a : lo, nhi = e > hi ? e : hi;
                    if (c1 > c0 && nhi - nlo > slab) break;
                    lo = nlo, hi = nhi, ++c1;
                }
                int64_t bytes = hi - lo;
                if (bytes > staged_cap) {
                    if (d_ascii) CU(cudaFreeAsync(d_ascii, st));
                    d_ascii = nullptr;
                    staged_cap = bytes + 64;
                    CU(cudaMallocAsync((void **)&d_ascii, (size_t)staged_cap, st));
                }
                if (bytes > 0) CU(cudaMemcpyAsync(d_ascii, contig_seqs + lo, (size_t)bytes, cudaMemcpyDefault, st));
                int64_t g0 = L.ctg_soff[c0] >> 5;
                int64_t g1 = (c1 < L.n_ctg ? L.ctg_soff[c1] : L.storage_bases - 128) >> 5;
                kb_launch_pack(d_ascii, lo, d_off, v.ctg_soff, v.ctg_len, c0, c1, g0, g1, b->seq2, b->nmask, st);
                CU(cudaGetLastError());
                c0 = c1;
            }
        } catch (...) {
            if (d_ascii) cudaFreeAsync(d_ascii, st);
            throw;
        }
        if (d_ascii) CU(cudaFreeAsync(d_ascii, st));
    });
}

// Contigs that are already 2-bit packed on the host (kb_fasta_ingest_pack): seq2 / nmask are the arrays of a whole ingest call in
// the layout of kb_packed_layout, `first_soff` the storage offset (in bases) of this batch's first contig in them.  The batch's own
// layout is the same rule started afresh, so its storage is the host range [first_soff, first_soff + storage - 256) moved to 128.
int kb_batch_create_packed(const uint32_t *seq2, const uint32_t *nmask, int64_t first_soff, const int32_t *contig_len,
                           const int32_t *asm_contig_start, int32_t n_asm, int device, kb_batch_t **out)
{
    if (!seq2 || !nmask || first_soff < 128 || (first_soff & 127)) return fail(KB_ERR_ARG, "bad packed buffers / first_soff");
    return batch_create_common(nullptr, contig_len, asm_contig_start, n_asm, device, out, [&](kb_batch *b, cudaStream_t st, const int64_t *) {
        const KbHostBatchLayout &L = b->L;
        const size_t ws = (size_t)(L.storage_bases >> 4), wm = (size_t)(L.storage_bases >> 5);
        // lead-in and tail groups (128 bases = 8 sequence words, 4 mask words each) are padding; the rest comes from the host
        CU(cudaMemsetAsync(b->seq2, 0, 8 * 4, st));
        CU(cudaMemsetAsync(b->nmask, 0xff, 4 * 4, st));
        CU(cudaMemsetAsync(b->seq2 + ws - 8, 0, 8 * 4 + 64, st));
        CU(cudaMemsetAsync(b->nmask + wm - 4, 0xff, 4 * 4 + 64, st));
        const int64_t body = L.storage_bases - 256;  // bases
        // in pieces of 64 MiB so that a scan launched behind it on another stream is never starved of copy-engine slots for long
        const int64_t piece = (int64_t)1 << 28;  // bases
        for (int64_t o = 0; o < body; o += piece) {
            const int64_t nb = body - o < piece ? body - o : piece;
            CU(cudaMemcpyAsync(b->seq2 + 8 + (o >> 4), seq2 + ((first_soff + o) >> 4), (size_t)(nb >> 4) * 4, cudaMemcpyDefault, st));
            CU(cudaMemcpyAsync(b->nmask + 4 + (o >> 5), nmask + ((first_soff + o) >> 5), (size_t)(nb >> 5) * 4, cudaMemcpyDefault, st));
        }
    });
}

// the batch's packed words back on the host (layout of kb_packed_layout over the batch's own contigs): storage / 16 and
// storage / 32 words.  Parity tests compare the device pack kernel with the host packer through it; bench.py builds its pinned
// end-to-end input from the resident batch with it.
int kb_batch_download_packed(const kb_batch_t *b, uint32_t *seq2, uint32_t *nmask, int64_t *storage_bases)
{
    if (!b) return fail(KB_ERR_ARG, "null batch");
    if (storage_bases) *storage_bases = b->L.storage_bases;
    if (!seq2 && !nmask) return KB_OK;
    if (cudaSetDevice(b->device) != cudaSuccess) return fail(KB_ERR_CUDA, "cudaSetDevice failed");
    if (seq2 && cudaMemcpy(seq2, b->seq2, (size_t)(b->L.storage_bases >> 4) * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return fail(KB_ERR_CUDA, "D2H of seq2 failed");
    if (nmask && cudaMemcpy(nmask, b->nmask, (size_t)(b->L.storage_bases >> 5) * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        return fail(KB_ERR_CUDA, "D2H of nmask failed");
    return KB_OK;
}

void kb_batch_destroy(kb_batch_t *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->d_blob) cudaFreeAsync(b->d_blob, 0);
    if (b->seq2) cudaFreeAsync(b->seq2, 0);
    if (b->nmask) cudaFreeAsync(b->nmask, 0);
    delete b;
}
int32_t kb_batch_n_assemblies(const kb_batch_t *b) { return b ? b->L.n_asm : 0; }
int64_t kb_batch_total_bases(const kb_batch_t *b) { return b ? b->L.total_bases : 0; }
int64_t kb_batch_packed_bytes(const kb_batch_t *b) { return b ? (b->L.storage_bases >> 2) + (b->L.storage_bases >> 3) : 0; }

void kb_result_destroy(kb_result_t *r)
{
    if (!r) return;
    cudaSetDevice(r->device);
    for (void *p : r->owned) cudaFreeAsync(p, 0);  // stream-ordered pool memory: back to the pool without a device-wide sync
    delete r;
}

}  // extern "C"

// minimap2 mm_idx_cal_max_occ for one assembly, on the device: dump its minimizer hashes with the
// scan kernel, radix sort, run-length encode, sort the run lengths, read the quantile.
static int32_t census_one(const kb_index *ix, const kb_batch *bt, int asm_id, DevPool &P, unsigned long long *d_counters, cudaStream_t st)
{
    const kb_params_t &p = ix->host.p;
    const KbHostBatchLayout &L = bt->L;
    int64_t bases = 0;
    for (int c = L.asm_ctg_start[asm_id]; c < L.asm_ctg_start[asm_id + 1]; ++c) bases += L.ctg_len[c];
    int64_t cap = bases / 2 + 4 * (int64_t)(L.asm_ctg_start[asm_id + 1] - L.asm_ctg_start[asm_id]) + 1024;
    uint32_t *h = P.get<uint32_t>((size_t)cap), *h2 = P.get<uint32_t>((size_t)cap), *pos = P.get<uint32_t>((size_t)cap);
    int32_t *ctg = P.get<int32_t>((size_t)cap);
    uint32_t *cnt = P.get<uint32_t>((size_t)cap), *cnt2 = P.get<uint32_t>((size_t)cap);
    int64_t *d_n = P.get<int64_t>(1);
    CU(cudaMemsetAsync(d_counters + 16, 0, 16 * 8, st));
    // restrict the scan to this assembly's chunks
    KbBatchView v = bt->view;
    int64_t k0 = 0, k1 = 0;
    {
        const std::vector<int32_t> &cc = L.chunk_ctg;
        int lo_c = L.asm_ctg_start[asm_id], hi_c = L.asm_ctg_start[asm_id + 1];
        k0 = std::lower_bound(cc.begin(), cc.end(), lo_c) - cc.begin();
        k1 = std::lower_bound(cc.begin(), cc.end(), hi_c) - cc.begin();
    }
    v.chunk_ctg += k0, v.chunk_start += k0, v.n_chunks = k1 - k0;
    kb_launch_scan(ix->view, v, nullptr, nullptr, d_counters + 16, 0, h, ctg, pos, cap, asm_id, bt->n_sm, st);
    CU(cudaGetLastError());
    unsigned long long n_mz = 0;
    CU(cudaMemcpyAsync(&n_mz, d_counters + 23, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if ((int64_t)n_mz > cap) throw std::string("census buffer overflow");
    int32_t mid;
    if (n_mz == 0) mid = INT32_MAX;
    else {
        size_t tb = kb_sort_keys32_temp_bytes((int64_t)n_mz), tb2 = kb_rle_temp_bytes((int64_t)n_mz);
        if (tb2 > tb) tb = tb2;
        uint8_t *tmp = P.get<uint8_t>(tb);
        CU(kb_sort_keys32(tmp, tb, h, h2, (int64_t)n_mz, 30, st));
        CU(kb_rle(tmp, tb, h2, h, cnt, d_n, (int64_t)n_mz, st));
        int64_t n_runs = 0;
        CU(cudaMemcpyAsync(&n_runs, d_n, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        CU(kb_sort_keys32(tmp, tb, cnt, cnt2, n_runs, 32, st));
        if (p.mid_occ_frac <= 0.f || n_runs <= 0) mid = INT32_MAX;  // [mm2:index.c:mm_idx_cal_max_occ] f <= 0: no cut-off
        else {
            int64_t kth = (int64_t)((1. - p.mid_occ_frac) * (double)n_runs);
            if (kth > n_runs - 1) kth = n_runs - 1;
            if (kth < 0) kth = 0;
            uint32_t val = 0;
            CU(cudaMemcpyAsync(&val, cnt2 + kth, 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            mid = (int32_t)(val + 1);
        }
        P.release(tmp);
    }
    if (mid < p.min_mid_occ) mid = p.min_mid_occ;
    if (p.max_mid_occ > p.min_mid_occ && mid > p.max_mid_occ) mid = p.max_mid_occ;
    P.release(h), P.release(h2), P.release(pos), P.release(ctg), P.release(cnt), P.release(cnt2), P.release(d_n);
    return mid;
}

// The census of many assemblies (ascending ids in `list`), about a gigabase at a time: one scan over the chunks of a run of assemblies
// dumps (hash, assembly) of every minimizer; a sort and a run-length encode give the occurrence count of every distinct minimizer, a
// histogram of the counts per assembly its quantile (kb_census_quantile_kernel).  One stream synchronisation per group of assemblies
// instead of three per assembly.  d_mid: n_asm entries on the device, those of the listed assemblies are overwritten.
// d_counters[16..31] are used; [24] counts quantiles beyond the histogram (checked by the caller).  Returns the number of launches.
static int census_many(const kb_index *ix, const kb_batch *bt, const std::vector<int32_t> &list, int32_t *d_mid, DevPool &P,
                       unsigned long long *d_counters, cudaStream_t st)
{
    const kb_params_t &p = ix->host.p;
    const KbHostBatchLayout &L = bt->L;
    const int64_t group_bases = getenv("KAPTIVE_B200_CENSUS_GROUP") ? atoll(getenv("KAPTIVE_B200_CENSUS_GROUP")) : ((int64_t)1200 << 20);
    auto asm_bases = [&](int a) {
        int64_t b = 0;
        for (int c = L.asm_ctg_start[a]; c < L.asm_ctg_start[a + 1]; ++c) b += L.ctg_len[c];
        return b;
    };
    int launches = 0;
    unsigned long long *d_over = P.get<unsigned long long>(1);
    CU(cudaMemsetAsync(d_over, 0, 8, st));
    size_t i = 0;
    while (i < list.size()) {
        // a group: listed assemblies from list[i] on, with gaps of at most 4 unlisted ones, until the bases in the range reach the limit
        size_t j = i;
        int a_lo = list[i], a_hi = list[i];
        int64_t bases = asm_bases(a_lo);
        while (j + 1 < list.size() && list[j + 1] - a_hi <= 5 && list[j + 1] - a_lo < 4096) {
            int64_t add = 0;
            for (int a = a_hi + 1; a <= list[j + 1]; ++a) add += asm_bases(a);
            if (bases + add > group_bases) break;
            bases += add, a_hi = list[j + 1], ++j;
        }
        const int n_list = (int)(j - i + 1), n_group = a_hi - a_lo + 1;
        int group_bits = 1;
        while ((1 << group_bits) < n_group) ++group_bits;
        const int64_t n_ctg = L.asm_ctg_start[a_hi + 1] - L.asm_ctg_start[a_lo];
        const int64_t cap = bases / 2 + 4 * n_ctg + 1024;
        uint32_t *h = P.get<uint32_t>((size_t)cap), *cnt = P.get<uint32_t>((size_t)cap), *hist = P.get<uint32_t>(kb_census_hist_bytes(n_group) / 4);
        int32_t *as = P.get<int32_t>((size_t)cap), *d_list = P.get<int32_t>((size_t)n_list);
        uint64_t *k1 = P.get<uint64_t>((size_t)cap), *k2 = P.get<uint64_t>((size_t)cap);
        int64_t *d_n = P.get<int64_t>(1);
        const size_t tb = kb_sort_keys64_temp_bytes(cap);
        uint8_t *tmp = P.get<uint8_t>(tb);
        CU(cudaMemcpyAsync(d_list, list.data() + i, (size_t)n_list * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(d_counters + 16, 0, 16 * 8, st));
        CU(cudaMemsetAsync(d_n, 0, 8, st));
        KbBatchView v = bt->view;
        const std::vector<int32_t> &cc = L.chunk_ctg;
        const int64_t c0 = std::lower_bound(cc.begin(), cc.end(), L.asm_ctg_start[a_lo]) - cc.begin();
        const int64_t c1 = std::lower_bound(cc.begin(), cc.end(), L.asm_ctg_start[a_hi + 1]) - cc.begin();
        v.chunk_ctg += c0, v.chunk_start += c0, v.n_chunks = c1 - c0;
        kb_launch_scan(ix->view, v, nullptr, nullptr, d_counters + 16, 0, h, as, nullptr, cap, -2, bt->n_sm, st);
        CU(cudaGetLastError());
        unsigned long long n_mz = 0;
        CU(cudaMemcpyAsync(&n_mz, d_counters + 23, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if ((int64_t)n_mz > cap) throw std::string("census buffer overflow");
        if (n_mz > 0) {
            kb_launch_census_keys(h, as, (int64_t)n_mz, a_lo, k1, st);
            CU(kb_sort_keys64(tmp, tb, k1, k2, (int64_t)n_mz, 30 + group_bits, st));
            CU(kb_rle64(tmp, tb, k2, k1, cnt, d_n, (int64_t)n_mz, st));
        }
        kb_launch_census_quantile(k1, cnt, d_n, (int64_t)n_mz, hist, n_group, d_list, n_list, a_lo, p.mid_occ_frac, p.min_mid_occ, p.max_mid_occ, d_mid,
                                  d_over, st);
        CU(cudaGetLastError());
        launches += 8;
        P.release(h), P.release(cnt), P.release(hist), P.release(as), P.release(d_list), P.release(k1), P.release(k2), P.release(d_n), P.release(tmp);
        i = j + 1;
    }
    unsigned long long over = 0;
    CU(cudaMemcpyAsync(&over, d_over, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    P.release(d_over);
    if (over) throw std::string("occurrence census: the quantile of an assembly lies beyond 65534 occurrences (internal limit)");
    return launches;
}

static int map_batch_impl(const kb_index *ix, kb_batch *bt, kb_result **out, bool keep_stages)
{
    if (ix->device != bt->device) return fail(KB_ERR_ARG, "index and batch live on different devices");
    const kb_params_t &p = ix->host.p;
    kb_result *R = new kb_result();
    R->device = ix->device;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[KB_N_STAGES + 1];
    for (auto &e : ev) e = nullptr;
    DevPool P;
    int launches = 0;
    try {
        CU(cudaSetDevice(ix->device));
        CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        P.st = st;
        P.arena = arena_acquire(ix->device);
        for (auto &e : ev) CU(cudaEventCreate(&e));
        const KbBatchView &bv = bt->view;
        const KbIndexView &iv = ix->view;
        const int n_asm = bv.n_asm;
        unsigned long long *d_counters = P.get<unsigned long long>(KB_N_COUNTERS);
        unsigned long long hc[KB_N_COUNTERS];

        // ---------------- scan (+ rerun once if the anchor buffer was too small)
        const int64_t hint = bt->last_anchor_count.load(std::memory_order_relaxed);
        int64_t anchor_cap = hint > 0 ? hint + hint / 16 + 1024 : bv.total_bases / 48 + (1 << 20);
        uint64_t *akey = nullptr;
        uint32_t *aval = nullptr;
        int64_t n_anchors = 0;
        CU(cudaEventRecord(ev[0], st));
        for (int attempt = 0;; ++attempt) {
            akey = P.get<uint64_t>((size_t)anchor_cap), aval = P.get<uint32_t>((size_t)anchor_cap);
            CU(cudaMemsetAsync(d_counters, 0, KB_N_COUNTERS * 8, st));
            kb_launch_scan(iv, bv, akey, aval, d_counters, anchor_cap, nullptr, nullptr, nullptr, 0, -1, bt->n_sm, st);
            ++launches;
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(hc, d_counters, KB_N_COUNTERS * 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            n_anchors = (int64_t)hc[1];
            if (hc[6] & 1) throw std::string("scan queue overflow (internal limit)");
            if (n_anchors <= anchor_cap) break;
            if (attempt >= 2) throw std::string("anchor buffer overflow");
            P.release(akey), P.release(aval);
            anchor_cap = n_anchors + n_anchors / 16 + 1024;
        }
        bt->last_anchor_count.store(n_anchors, std::memory_order_relaxed);
        if (n_anchors >= ((int64_t)1 << 31)) throw std::string("more than 2^31 anchors in one batch: split the batch");
        CU(cudaEventRecord(ev[1], st));
        R->counters[0] = (int64_t)hc[0], R->counters[1] = n_anchors;

        // ---------------- occurrence counts, census where needed
        std::vector<int32_t> mid_occ((size_t)n_asm, p.mid_occ > 0 ? p.mid_occ : p.min_mid_occ);
        std::vector<int32_t> occ_skip((size_t)n_asm + 1, 0), census_list;
        size_t occ_words = ((size_t)n_asm * (size_t)iv.n_entries + 1) / 2 + 4;
        uint32_t *occ32 = P.get<uint32_t>(occ_words);
        int32_t *d_need = P.get<int32_t>((size_t)n_asm + 1);
        CU(cudaMemsetAsync(occ32, 0, occ_words * 4, st));
        CU(cudaMemsetAsync(d_need, 0, ((size_t)n_asm + 1) * 4, st));
        kb_launch_occ(akey, aval, n_anchors, iv.n_entries, occ32, p.min_mid_occ, d_need, st);
        ++launches;
        CU(cudaGetLastError());
        {
            std::vector<int32_t> need((size_t)n_asm + 1, 0);
            CU(cudaMemcpyAsync(need.data(), d_need, (size_t)n_asm * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            const char *fc = getenv("KAPTIVE_B200_FORCE_CENSUS");
            bool force = (fc && fc[0] == '1') || ix->host.max_qocc > p.min_mid_occ;
            for (int a = 0; a < n_asm; ++a) occ_skip[(size_t)a] = need[(size_t)a] == 0 && mid_occ[(size_t)a] >= p.min_mid_occ;
            for (int a = 0; a < n_asm; ++a) {
                if (need[(size_t)a] == 2) throw std::string("a gene minimizer occurs more than 65519 times in one assembly (limit)");
                if (p.mid_occ <= 0 && (need[(size_t)a] || force)) census_list.push_back(a);
            }
        }
        int32_t *d_mid = P.get<int32_t>((size_t)n_asm + 1);
        CU(cudaMemcpyAsync(d_mid, mid_occ.data(), (size_t)n_asm * 4, cudaMemcpyHostToDevice, st));
        if (!census_list.empty()) {
            const char *one = getenv("KAPTIVE_B200_CENSUS_SERIAL");  // test hook: the per-assembly form
            if (one && one[0] == '1') {
                for (int a : census_list) mid_occ[(size_t)a] = census_one(ix, bt, a, P, d_counters, st), launches += 4;
                CU(cudaMemcpyAsync(d_mid, mid_occ.data(), (size_t)n_asm * 4, cudaMemcpyHostToDevice, st));
            } else {
                launches += census_many(ix, bt, census_list, d_mid, P, d_counters, st);
                CU(cudaMemcpyAsync(mid_occ.data(), d_mid, (size_t)n_asm * 4, cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
            }
        }
        // assemblies in which no gene minimizer occurs more than min_mid_occ times (and whose mid_occ is not below that floor):
        // the occurrence filter cannot fire there, the chain kernel skips its table look-ups
        int32_t *d_occ_skip = P.get<int32_t>((size_t)n_asm + 1);
        CU(cudaMemcpyAsync(d_occ_skip, occ_skip.data(), (size_t)n_asm * 4, cudaMemcpyHostToDevice, st));
        R->mid_occ = mid_occ;

        // ---------------- sort anchors by (asm, gene, strand, position)
        uint64_t *skey = P.get<uint64_t>((size_t)n_anchors + 1);
        uint32_t *sval = P.get<uint32_t>((size_t)n_anchors + 1);
        {
            size_t tb = kb_sort_pairs_temp_bytes(n_anchors);
            uint8_t *tmp = P.get<uint8_t>(tb);
            if (n_anchors > 0) CU(kb_sort_pairs(tmp, tb, akey, skey, aval, sval, n_anchors, 60, st));
            launches += 8;
            P.release(tmp);
        }
        P.release(akey), P.release(aval);
        // groups
        uint8_t *gflag = P.get<uint8_t>((size_t)n_anchors + 1);
        int64_t *gstart = P.get<int64_t>((size_t)n_anchors + 1);
        int64_t *d_ng = P.get<int64_t>(1);
        int64_t n_groups = 0;
        if (n_anchors > 0) {
            kb_launch_group_flag(skey, n_anchors, gflag, st);
            size_t tb = kb_select_temp_bytes(n_anchors);
            uint8_t *tmp = P.get<uint8_t>(tb);
            CU(kb_select_flagged(tmp, tb, gflag, gstart, d_ng, n_anchors, st));
            launches += 3;
            CU(cudaMemcpyAsync(&n_groups, d_ng, 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            P.release(tmp);
        }
        P.release(gflag);
        CU(cudaEventRecord(ev[2], st));
        R->counters[2] = n_groups;

        // ---------------- chaining
        KbChainWork W;
        W.x = P.get<uint32_t>((size_t)n_anchors + 1), W.y = P.get<int32_t>((size_t)n_anchors + 1);
        W.f = P.get<int32_t>((size_t)n_anchors + 1), W.p = P.get<int32_t>((size_t)n_anchors + 1);
        W.v = P.get<int32_t>((size_t)n_anchors + 1), W.t = P.get<int32_t>((size_t)n_anchors + 1);
        W.z = P.get<uint64_t>((size_t)n_anchors + 1), W.u = P.get<uint64_t>((size_t)n_anchors + 1);
        uint64_t *cx = P.get<uint64_t>((size_t)n_anchors + 1), *cy = P.get<uint64_t>((size_t)n_anchors + 1);
        KbGroupInfo *ginfo = P.get<KbGroupInfo>((size_t)n_groups + 1);
        int64_t chain_cap = n_anchors / 3 + 16;
        KbChainRec *chains = P.get<KbChainRec>((size_t)chain_cap);
        kb_launch_chain(iv, bv, skey, sval, gstart, n_groups, n_anchors, (const uint16_t *)occ32, d_mid, W, cx, cy, ginfo, chains,
                        d_counters, chain_cap, d_occ_skip, st);
        ++launches;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(hc, d_counters, KB_N_COUNTERS * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        const int64_t n_chains = (int64_t)hc[3];
        if (n_chains > chain_cap) throw std::string("chain buffer overflow");
        CU(cudaEventRecord(ev[3], st));
        R->counters[3] = n_chains;
        if (keep_stages && n_groups > 0) {  // anchors after the occurrence filters, chains before alignment
            int32_t *nseed = P.get<int32_t>((size_t)n_groups + 1);
            int64_t *aoff = P.get<int64_t>((size_t)n_groups + 1);
            kb_launch_group_nseed(ginfo, n_groups, nseed, st);
            size_t tb = kb_scan_temp_bytes(n_groups);
            uint8_t *tmp = P.get<uint8_t>(tb);
            CU(kb_exclusive_sum(tmp, tb, nseed, aoff, n_groups, st));
            int64_t last_off = 0;
            int32_t last_n = 0;
            CU(cudaMemcpyAsync(&last_off, aoff + n_groups - 1, 8, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(&last_n, nseed + n_groups - 1, 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            R->n_anchor_dump = last_off + last_n;
            CU(cudaMallocAsync((void **)&R->d_anchor_dump, (size_t)(R->n_anchor_dump + 1) * 7 * 4, st));
            R->owned.push_back(R->d_anchor_dump);
            kb_launch_dump_anchors(bv, ginfo, n_groups, gstart, W.x, W.y, aoff, R->d_anchor_dump, st);
            std::vector<KbChainRec> hch((size_t)n_chains);
            std::vector<KbGroupInfo> hg((size_t)n_groups);
            CU(cudaMemcpyAsync(hch.data(), chains, (size_t)n_chains * sizeof(KbChainRec), cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(hg.data(), ginfo, (size_t)n_groups * sizeof(KbGroupInfo), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (int64_t g = 0; g < n_groups; ++g)
                for (int i = 0; i < hg[(size_t)g].n_chains; ++i) {
                    const KbChainRec &c = hch[(size_t)(hg[(size_t)g].chain_base + i)];
                    int32_t rec[10] = {hg[(size_t)g].asm_id, hg[(size_t)g].gene, c.score, c.cnt, c.rev, c.rid, c.rs, c.re, c.qs, c.qe};
                    R->chain_dump.insert(R->chain_dump.end(), rec, rec + 10);
                }
            P.release(tmp), P.release(nseed), P.release(aoff);
        }
        P.release(skey), P.release(sval), P.release(gstart), P.release(occ32);
        P.release(W.x), P.release(W.y), P.release(W.f), P.release(W.p), P.release(W.v), P.release(W.t), P.release(W.z), P.release(W.u);

        // ---------------- alignment
        int64_t raw_cap = n_chains + n_anchors / 3 + 16;
        KbRawHit *raw = P.get<KbRawHit>((size_t)raw_cap);
        int64_t pool_cap = raw_cap * 24 + (1 << 16);
        int64_t n_raw = 0, n_pool = 0;
        uint32_t *pool = nullptr;
        // Two scratch sizes.  The staged kernels (and the thousands of warps they keep resident) work with DP problems of up to
        // KB_FAST_SW_CELLS cells -- all but a handful; a chain with a larger problem (minimap2 allows max_sw_mat = 100 M cells: an end
        // extension over >= 1.4 kb of unaligned gene) is handed to the one-warp-per-chain kernel, which gets a few warps with scratch
        // for the full p.max_sw_cells, allocated only when such a chain exists.
        const int64_t KB_FAST_SW_CELLS = 4000000;
        const int32_t fast_cells = (int32_t)std::min<int64_t>(p.max_sw_cells, KB_FAST_SW_CELLS);
        KbIndexView iv_fast = iv;
        iv_fast.p.max_sw_cells = fast_cells;
        const size_t sbytes = kb_align_scratch_bytes(fast_cells), sbytes_full = kb_align_scratch_bytes(p.max_sw_cells);
        int n_warps = bt->n_sm * 24;  // resident warps of the DP kernels (6 CTAs of 4 warps per SM)
        {
            int64_t need_warps = ((n_chains + 3) / 4) * 4;
            if (need_warps < 4) need_warps = 4;
            if (n_warps > need_warps) n_warps = (int)need_warps;
        }
        auto slow_warps_for = [&](int64_t n_slow) {  // warps of the full-size kernel: at most 8 GB of scratch, at least one CTA
            int64_t w = std::max<int64_t>(4, ((int64_t)8 << 30) / (int64_t)sbytes_full / 4 * 4);
            w = std::min<int64_t>(w, std::max<int64_t>(4, (n_slow + 3) / 4 * 4));
            return (int)std::min<int64_t>(w, n_warps);
        };
        unsigned long long *d_next = P.get<unsigned long long>(1);
        // staged path (kb_stage.cuh): plan -> band / rows DP kernels -> assemble; what it hands back goes to kb_align_kernel
        const char *sg = getenv("KAPTIVE_B200_STAGED");
        const bool staged = !(sg && sg[0] == '0') && n_chains > 0;
        if (!staged) n_warps = slow_warps_for(n_chains);  // every chain goes through the full-size kernel
        uint8_t *scratch = P.get<uint8_t>((size_t)n_warps * (staged ? sbytes : sbytes_full));
        uint8_t *slow_scratch = nullptr;
        const int64_t job_cap = n_chains * 12 + 4096, jobcig_cap = job_cap * 16;
        int band_warps = bt->n_sm * 24, rows_warps = n_warps;
        if (const char *e = getenv("KAPTIVE_B200_BAND_WARPS")) band_warps = bt->n_sm * atoi(e);
        if ((int64_t)band_warps > 4 * n_chains + 4) band_warps = (int)(((4 * n_chains + 4) + 3) / 4 * 4);  // small calls: small scratch
        if (const char *e = getenv("KAPTIVE_B200_ROWS_WARPS")) rows_warps = std::min(n_warps, bt->n_sm * atoi(e));
        void *plans = nullptr, *jobs = nullptr;
        int32_t *slow_list = nullptr, *kscratch = nullptr;
        KbStageLists L;
        memset(&L, 0, sizeof(L));
        L.job_cap = job_cap;
        uint32_t *jobcig = nullptr, *tmpcig = nullptr;
        uint8_t *band_scratch = nullptr, *side_scratch = nullptr;
        int side_warps = 0;
        if (staged) {
            plans = P.get<uint8_t>((size_t)n_chains * kb_sizeof_plan());
            jobs = P.get<uint8_t>((size_t)job_cap * kb_sizeof_job());
            L.band_list = P.get<int32_t>((size_t)job_cap), L.rows_list = P.get<int32_t>((size_t)job_cap), L.r16_list = P.get<int32_t>((size_t)job_cap);
            L.r16_list2 = P.get<int32_t>((size_t)job_cap), L.r16_key = P.get<uint32_t>((size_t)job_cap), L.r16_key2 = P.get<uint32_t>((size_t)job_cap);
            for (int i = 0; i < 3; ++i) L.b16_list[i] = P.get<int32_t>((size_t)job_cap), L.b16_key[i] = P.get<uint32_t>((size_t)job_cap);
            L.sort_tmp_bytes = kb_r16_sort_temp_bytes(job_cap), L.sort_tmp = P.get<uint8_t>(L.sort_tmp_bytes);
            slow_list = P.get<int32_t>((size_t)n_chains + 1), kscratch = P.get<int32_t>((size_t)n_anchors + 8);
            jobcig = P.get<uint32_t>((size_t)jobcig_cap), tmpcig = P.get<uint32_t>((size_t)jobcig_cap + (size_t)n_chains + 8);
            band_scratch = P.get<uint8_t>((size_t)band_warps * kb_band_scratch_bytes());
            // the 32-bit rows kernel next to the packed one (kb_launch_stage_dp): its own scratch, one CTA per SM is plenty for what is left to it
            side_warps = std::min(rows_warps, bt->n_sm * 4) / 4 * 4;
            if (side_warps >= 4 && rows_warps >= bt->n_sm * 4) side_scratch = P.get<uint8_t>((size_t)side_warps * sbytes);
        }
        for (int attempt = 0;; ++attempt) {
            CU(cudaMallocAsync((void **)&pool, (size_t)pool_cap * 4 + 16, st));
            CU(cudaMemsetAsync(d_next, 0, 8, st));
            CU(cudaMemsetAsync(d_counters + 4, 0, 16, st));
            CU(cudaMemsetAsync(d_counters + 8, 0, 8 * 8, st));
            CU(cudaMemsetAsync(d_counters + 32, 0, 16 * 8, st));
            if (staged) {
                kb_launch_stage_plan(iv_fast, bv, chains, n_chains, ginfo, cx, cy, kscratch, plans, jobs, L, slow_list, d_counters, st);
                kb_launch_stage_dp(iv_fast, bv, jobs, L, band_scratch, band_warps, scratch, sbytes, rows_warps, side_scratch, side_warps, jobcig,
                                   jobcig_cap, d_counters, st);
                kb_launch_stage_assemble(iv_fast, bv, chains, n_chains, ginfo, plans, jobs, jobcig, tmpcig, jobcig_cap + n_chains, raw, raw_cap, pool,
                                         pool_cap, slow_list, d_counters, st);
                launches += 14;
                // chains handed back (a DP problem over the fast limit, a z-drop inside a gap fill, a capacity limit): full-size kernel
                unsigned long long n_slow = 0;
                CU(cudaMemcpyAsync(&n_slow, d_counters + 12, 8, cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                if (n_slow > 0) {
                    const int sw = slow_warps_for((int64_t)n_slow);
                    if (!slow_scratch) CU(cudaMallocAsync((void **)&slow_scratch, (size_t)sw * sbytes_full, st));
                    kb_launch_align(iv, bv, chains, n_chains, ginfo, cx, cy, slow_scratch, sbytes_full, sw, raw, raw_cap, pool, pool_cap, d_counters,
                                    d_next, slow_list, d_counters + 12, st);
                    ++launches;
                }
            } else {
                kb_launch_align(iv, bv, chains, n_chains, ginfo, cx, cy, scratch, sbytes_full, n_warps, raw, raw_cap, pool, pool_cap, d_counters, d_next,
                                nullptr, nullptr, st);
                ++launches;
            }
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(hc, d_counters, KB_N_COUNTERS * 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            n_raw = (int64_t)hc[4], n_pool = (int64_t)hc[5];
            if (n_raw > raw_cap) throw std::string("raw hit buffer overflow");
            if (n_pool <= pool_cap) break;
            if (attempt >= 1) throw std::string("cigar pool overflow");
            // NB: the seed-filter flags written into cy[] are idempotent, so the stage can simply be re-run
            CU(cudaFreeAsync(pool, st));
            pool_cap = n_pool + 1024;
        }
        R->pool = pool, R->owned.push_back(pool), R->n_cigar = n_pool;
        R->counters[7] = (int64_t)hc[12];  // chains the staged path handed back to kb_align_kernel
        P.release(scratch);
        if (slow_scratch) CU(cudaFreeAsync(slow_scratch, st));
        if (staged) {
            P.release(plans), P.release(jobs), P.release(L.band_list), P.release(L.rows_list), P.release(L.r16_list), P.release(slow_list), P.release(kscratch);
            P.release(L.r16_list2), P.release(L.r16_key), P.release(L.r16_key2), P.release(L.sort_tmp);
            for (int i = 0; i < 3; ++i) P.release(L.b16_list[i]), P.release(L.b16_key[i]);
            P.release(jobcig), P.release(tmpcig), P.release(band_scratch);
            if (side_scratch) P.release(side_scratch);
        }
        CU(cudaEventRecord(ev[4], st));
        R->counters[4] = n_raw, R->counters[6] = (int64_t)hc[8];

        // ---------------- finalise: order raw hits by (group, reg, split), per-query filter/sort/parent/mapq, compact to SoA
        KbRawHit *sorted = P.get<KbRawHit>((size_t)n_raw + 1);
        int32_t *keep = P.get<int32_t>((size_t)n_raw + 1);
        int64_t *oidx = P.get<int64_t>((size_t)n_raw + 1);
        int64_t n_hits = 0;
        if (n_raw > 0) {
            uint64_t *rk = P.get<uint64_t>((size_t)n_raw), *rk2 = P.get<uint64_t>((size_t)n_raw);
            uint32_t *ri = P.get<uint32_t>((size_t)n_raw), *ri2 = P.get<uint32_t>((size_t)n_raw);
            kb_launch_rawkey(raw, n_raw, rk, ri, d_counters + 15, st);  // [15]: raw hits that carry an internal-limit error
            size_t tb = kb_sort_pairs_temp_bytes(n_raw), tb2 = kb_scan_temp_bytes(n_raw);
            if (tb2 > tb) tb = tb2;
            uint8_t *tmp = P.get<uint8_t>(tb);
            CU(kb_sort_pairs(tmp, tb, rk, rk2, ri, ri2, n_raw, 64, st));
            kb_launch_gather_raw(raw, ri2, n_raw, sorted, st);
            int32_t *fw = P.get<int32_t>((size_t)n_raw + 1);
            uint64_t *fcov = P.get<uint64_t>((size_t)n_raw + 1);
            kb_launch_finalize(p, sorted, n_raw, ginfo, fw, fcov, keep, st);
            CU(kb_exclusive_sum(tmp, tb, keep, oidx, n_raw, st));
            launches += 14;
            int64_t last_o = 0;
            int32_t last_k = 0;
            CU(cudaMemcpyAsync(&last_o, oidx + n_raw - 1, 8, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(&last_k, keep + n_raw - 1, 4, cudaMemcpyDeviceToHost, st));
            unsigned long long n_err = 0;
            CU(cudaMemcpyAsync(&n_err, d_counters + 15, 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            n_hits = last_o + last_k;
            R->counters[8] = (int64_t)n_err;
        }
        R->n_hits = n_hits;
        {
            size_t n = (size_t)n_hits + 1;
            // one stream-ordered allocation for the 17 result arrays (256-byte aligned slices)
            const size_t n4 = (n * 4 + 255) & ~(size_t)255, n1 = (n + 255) & ~(size_t)255, n8 = (n * 8 + 255) & ~(size_t)255;
            uint8_t *arena = nullptr;
            CU(cudaMallocAsync((void **)&arena, 13 * n4 + 3 * n1 + n8, st));
            R->owned.push_back(arena);
            auto own = [&](size_t bytes) {
                void *q = arena;
                arena += bytes;
                return q;
            };
            kb_hits_t &d = R->d;
            d.capacity = n_hits;
            d.asm_id = (int32_t *)own(n4), d.gene = (int32_t *)own(n4), d.q_start = (int32_t *)own(n4), d.q_end = (int32_t *)own(n4);
            d.t_ctg = (int32_t *)own(n4), d.t_len = (int32_t *)own(n4), d.t_start = (int32_t *)own(n4), d.t_end = (int32_t *)own(n4);
            d.strand = (int8_t *)own(n1), d.score = (int32_t *)own(n4), d.matches = (int32_t *)own(n4), d.block_len = (int32_t *)own(n4);
            d.edit_distance = (int32_t *)own(n4), d.mapq = (uint8_t *)own(n1), d.is_primary = (uint8_t *)own(n1);
            d.cigar_off = (int64_t *)own(n8), d.n_cigar = (int32_t *)own(n4);
            kb_launch_scatter(sorted, keep, oidx, n_raw, ginfo, bv, d, st);
            ++launches;
            CU(cudaGetLastError());
        }
        CU(cudaEventRecord(ev[5], st));
        CU(cudaStreamSynchronize(st));
        R->counters[5] = launches;
        for (int s = 0; s < KB_STAGE_TOTAL; ++s) CU(cudaEventElapsedTime(&R->stage_ms[s], ev[s], ev[s + 1]));
        CU(cudaEventElapsedTime(&R->stage_ms[KB_STAGE_TOTAL], ev[0], ev[5]));
        P.clear();
        CU(cudaStreamSynchronize(st));
        arena_release(P.arena), P.arena = nullptr;
        for (auto &e : ev) cudaEventDestroy(e);
        cudaStreamDestroy(st);
    } catch (const std::string &e) {
        P.clear();
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
        arena_release(P.arena), P.arena = nullptr;
        for (auto &e2 : ev)
            if (e2) cudaEventDestroy(e2);
        kb_result_destroy(R);
        return fail(KB_ERR_CUDA, e);
    }
    *out = R;
    return KB_OK;
}

extern "C" {

int kb_map_batch(const kb_index_t *ix, const kb_batch_t *bt, kb_result_t **out)
{
    if (!ix || !bt || !out) return fail(KB_ERR_ARG, "null argument");
    const char *ks = getenv("KAPTIVE_B200_KEEP_STAGES");
    return map_batch_impl(ix, const_cast<kb_batch *>(bt), out, ks && ks[0] == '1');
}

int kb_result_size(const kb_result_t *r, int64_t *n_hits, int64_t *n_cigar)
{
    if (!r) return fail(KB_ERR_ARG, "null result");
    if (n_hits) *n_hits = r->n_hits;
    if (n_cigar) *n_cigar = r->n_cigar;
    return KB_OK;
}

int kb_result_fetch(const kb_result_t *r, kb_hits_t *dst, uint32_t *cigar, int64_t cigar_cap)
{
    if (!r || !dst) return fail(KB_ERR_ARG, "null argument");
    if (dst->capacity < r->n_hits) return fail(KB_ERR_CAPACITY, "hit arrays smaller than kb_result_size()");
    if (cigar && cigar_cap < r->n_cigar) return fail(KB_ERR_CAPACITY, "cigar buffer smaller than kb_result_size()");
    try {
        CU(cudaSetDevice(r->device));
        const size_t n = (size_t)r->n_hits;
        const kb_hits_t &d = r->d;
#define CP(f, sz) \
    if (dst->f && n) CU(cudaMemcpy(dst->f, d.f, n * (sz), cudaMemcpyDeviceToHost))
        CP(asm_id, 4); CP(gene, 4); CP(q_start, 4); CP(q_end, 4); CP(t_ctg, 4); CP(t_len, 4); CP(t_start, 4); CP(t_end, 4);
        CP(strand, 1); CP(score, 4); CP(matches, 4); CP(block_len, 4); CP(edit_distance, 4); CP(mapq, 1); CP(is_primary, 1);
        CP(cigar_off, 8); CP(n_cigar, 4);
#undef CP
        if (cigar && r->n_cigar) CU(cudaMemcpy(cigar, r->pool, (size_t)r->n_cigar * 4, cudaMemcpyDeviceToHost));
    } catch (const std::string &e) {
        return fail(KB_ERR_CUDA, e);
    }
    return KB_OK;
}

int kb_result_stage_ms(const kb_result_t *r, float *ms)
{
    if (!r || !ms) return fail(KB_ERR_ARG, "null argument");
    memcpy(ms, r->stage_ms, sizeof(r->stage_ms));
    return KB_OK;
}
int kb_result_counters(const kb_result_t *r, int64_t *c)
{
    if (!r || !c) return fail(KB_ERR_ARG, "null argument");
    memcpy(c, r->counters, sizeof(r->counters));
    return KB_OK;
}
int kb_result_mid_occ(const kb_result_t *r, int32_t *out)
{
    if (!r || !out) return fail(KB_ERR_ARG, "null argument");
    if (!r->mid_occ.empty()) memcpy(out, r->mid_occ.data(), r->mid_occ.size() * 4);
    return KB_OK;
}
int kb_result_fetch_anchors(const kb_result_t *r, int32_t *out, int64_t cap, int64_t *n)
{
    if (!r || !n) return fail(KB_ERR_ARG, "null argument");
    *n = r->n_anchor_dump;
    if (!out) return KB_OK;
    if (cap < r->n_anchor_dump) return fail(KB_ERR_CAPACITY, "anchor dump buffer too small");
    if (r->n_anchor_dump) {
        cudaSetDevice(r->device);
        if (cudaMemcpy(out, r->d_anchor_dump, (size_t)r->n_anchor_dump * 28, cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail(KB_ERR_CUDA, "copy of the anchor dump failed");
    }
    return KB_OK;
}
int kb_result_fetch_chains(const kb_result_t *r, int32_t *out, int64_t cap, int64_t *n)
{
    if (!r || !n) return fail(KB_ERR_ARG, "null argument");
    *n = (int64_t)r->chain_dump.size() / 10;
    if (!out) return KB_OK;
    if (cap < *n) return fail(KB_ERR_CAPACITY, "chain dump buffer too small");
    if (*n) memcpy(out, r->chain_dump.data(), r->chain_dump.size() * 4);
    return KB_OK;
}

}  // extern "C"

// Host buffers in, host arrays out.  Calls of 512 assemblies or more are cut into slabs (ASCII input: equal slabs of at most
// KAPTIVE_B200_SLAB = 768 assemblies; packed input: a short first slab, then large ones), which bounds the device memory of a
// call and lets the H2D copy (and packing) of one slab run under the kernels of the previous one; hits come back in assembly
// order.  Slabs must stay large: small slabs lose more to half-empty persistent kernels than the overlap wins (measured,
// scripts/e2e_probe.py).  Results do not depend on the slab plan (assemblies are independent units;
// tests/test_gpu_parity.py::test_host_buffer_entry_point_slabs_equal_batch_path).
// make_batch(a0, a1, c0, acs) builds the device batch of assemblies [a0, a1) (contigs from c0, acs = their contig ranges from 0).
template <class MakeBatch>
static int map_slabs(const kb_index_t *ix, const int32_t *asm_contig_start, int32_t n_asm, kb_hits_t *dst, int64_t *n_hits, uint32_t *cigar,
                     int64_t cigar_cap, int64_t *n_cigar, bool packed, MakeBatch make_batch)
{
    if (!ix) return fail(KB_ERR_ARG, "null index");
    if (!asm_contig_start || n_asm < 0) return fail(KB_ERR_ARG, "null argument");
    // slab boundaries (assembly indices): bnd[k] .. bnd[k+1]
    std::vector<int> bnd{0};
    if (const char *e = getenv("KAPTIVE_B200_SLAB_PLAN")) {  // experiments: explicit slab sizes, e.g. "150,425,425" (the last one repeats)
        int last = 0;
        for (const char *q = e; bnd.back() < n_asm;) {
            int v = atoi(q);
            if (v > 0) last = v;
            if (last <= 0) break;
            bnd.push_back(std::min(n_asm, bnd.back() + last));
            const char *c = strchr(q, ',');
            q = c ? c + 1 : "";
        }
    } else if (packed && !getenv("KAPTIVE_B200_SLAB") && n_asm >= 1024) {
        // packed input: the copy of a slab (0.375 B per base) is several times shorter than its kernels, so only the FIRST slab's copy
        // is ever exposed: a short first slab, then slabs as large as possible (the persistent DP kernels lose ~10 % at 700
        // assemblies per launch against 2500; device memory is not a constraint at 1.9 MB per assembly)
        // sizes 512, 1536, then equal slabs of at most 2560: the copy of slab k + 1 (0.075 ms per assembly at 25 GB/s) always ends
        // before the kernels of slab k (0.27 ms per assembly) do
        bnd.push_back(std::min(512, n_asm / 4));
        if (n_asm - bnd.back() > 2 * 1536) bnd.push_back(bnd.back() + 1536);
        const int base = bnd.back(), rest = n_asm - base, n = (rest + 2559) / 2560, each = (rest + n - 1) / n;
        for (int k = 1; k <= n; ++k) bnd.push_back(std::min(n_asm, base + k * each));
    } else {
        int slab = 768, n;
        if (const char *e = getenv("KAPTIVE_B200_SLAB")) {  // explicit slab size: cut whenever the call is larger
            slab = atoi(e) > 0 ? atoi(e) : slab;
            n = n_asm <= slab ? 1 : (n_asm + slab - 1) / slab;
        } else n = n_asm < 512 ? 1 : std::max(2, (n_asm + slab - 1) / slab);
        const int each = n > 1 ? (n_asm + n - 1) / n : n_asm;  // equal slabs
        for (int k = 1; k <= n; ++k) bnd.push_back(std::min(n_asm, k * each));
    }
    if (bnd.back() < n_asm || bnd.size() < 2) bnd.assign({0, n_asm});
    const int n_slabs = (int)bnd.size() - 1;
    // A producer thread builds the device batches in order (host -> device copy, and the pack kernel for ASCII input), at most two
    // ahead of the consumer; this thread maps them one after the other and fetches each slab's hits as soon as they exist.  One
    // mapping call at a time keeps the persistent kernels of two slabs from competing for the SMs, while the copy engine works
    // under them (measured against two threads that each did copy + map: 3270 -> 3480 assemblies/s end to end at 10,000 per call).
    std::vector<kb_batch_t *> ready((size_t)n_slabs, nullptr);
    std::vector<int> made((size_t)n_slabs, 0), mk_rc((size_t)n_slabs, KB_OK);
    std::vector<std::string> mk_err((size_t)n_slabs);
    std::mutex mu;
    std::condition_variable cv;
    int consumed = 0;
    bool stop = false;
    std::thread producer([&]() {
        for (int k = 0; k < n_slabs; ++k) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || k < consumed + 2; });
                if (stop) return;
            }
            const int a0 = bnd[(size_t)k], a1 = bnd[(size_t)k + 1];
            const int c0 = asm_contig_start[a0];
            std::vector<int32_t> acs((size_t)(a1 - a0) + 1);
            for (int a = a0; a <= a1; ++a) acs[(size_t)(a - a0)] = asm_contig_start[a] - c0;
            kb_batch_t *b = nullptr;
            const int rc = make_batch(a0, a1, c0, acs.data(), &b);
            {
                std::lock_guard<std::mutex> lk(mu);
                ready[(size_t)k] = b, mk_rc[(size_t)k] = rc, made[(size_t)k] = 1;
                if (rc) mk_err[(size_t)k] = g_err;  // g_err is thread local
            }
            cv.notify_all();
            if (rc) return;
        }
    });
    int rc = KB_OK;
    int64_t oh = 0, oc = 0;
    bool cap_fail = false;
    for (int k = 0; k < n_slabs && !rc; ++k) {
        kb_batch_t *b = nullptr;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return made[(size_t)k] != 0; });
            b = ready[(size_t)k];
            if (mk_rc[(size_t)k]) rc = fail(mk_rc[(size_t)k], mk_err[(size_t)k]);
        }
        kb_result_t *r = nullptr;
        if (!rc) rc = kb_map_batch(ix, b, &r);
        if (b) kb_batch_destroy(b);
        {
            std::lock_guard<std::mutex> lk(mu);
            consumed = k + 1;
        }
        cv.notify_all();
        // too small: the remaining slabs are still mapped (not fetched) so that the sizes the caller needs for a retry are reported
        if (!rc && dst && (dst->capacity < oh + r->n_hits || (cigar && cigar_cap < oc + r->n_cigar))) cap_fail = true;
        if (!rc && dst && !cap_fail) {
            kb_hits_t v = *dst;  // a view of the caller's arrays starting at hit `oh`
            v.capacity = dst->capacity - oh;
#define OFF(f) \
    if (v.f) v.f += oh
            OFF(asm_id); OFF(gene); OFF(q_start); OFF(q_end); OFF(t_ctg); OFF(t_len); OFF(t_start); OFF(t_end); OFF(strand); OFF(score);
            OFF(matches); OFF(block_len); OFF(edit_distance); OFF(mapq); OFF(is_primary); OFF(cigar_off); OFF(n_cigar);
#undef OFF
            rc = kb_result_fetch(r, &v, cigar ? cigar + oc : nullptr, cigar_cap - oc);
            const int a0 = bnd[(size_t)k];
            if (!rc && (a0 || oc))
                for (int64_t i = 0; i < r->n_hits; ++i) {
                    if (v.asm_id) v.asm_id[i] += a0;
                    if (v.cigar_off) v.cigar_off[i] += oc;
                }
        }
        if (r) oh += r->n_hits, oc += r->n_cigar;
        kb_result_destroy(r);
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        stop = true;
    }
    cv.notify_all();
    producer.join();
    for (int k = 0; k < n_slabs; ++k)  // batches the producer made after a failure stopped the consumer
        if (made[(size_t)k] && k >= consumed && ready[(size_t)k]) kb_batch_destroy(ready[(size_t)k]);
    if (!rc) {
        if (n_hits) *n_hits = oh;
        if (n_cigar) *n_cigar = oc;
        if (cap_fail) rc = fail(KB_ERR_CAPACITY, "hit arrays / cigar buffer smaller than the call's hits (sizes reported in n_hits / n_cigar)");
    }
    return rc;
}

extern "C" {

int kb_map_assemblies(const kb_index_t *ix, const uint8_t *contig_seqs, const int64_t *contig_off, const int32_t *contig_len,
                      const int32_t *asm_contig_start, int32_t n_asm, kb_hits_t *dst, int64_t *n_hits, uint32_t *cigar,
                      int64_t cigar_cap, int64_t *n_cigar)
{
    return map_slabs(ix, asm_contig_start, n_asm, dst, n_hits, cigar, cigar_cap, n_cigar, false,
                     [&](int a0, int a1, int c0, const int32_t *acs, kb_batch_t **b) {
                         return kb_batch_create(contig_seqs, contig_off + c0, contig_len + c0, acs, a1 - a0, ix->device, b);
                     });
}

// The same for contigs the host has already packed (kb_fasta_ingest_pack): 0.375 B per base cross PCIe, no pack kernel.
int kb_map_assemblies_packed(const kb_index_t *ix, const uint32_t *seq2, const uint32_t *nmask, const int32_t *contig_len,
                             const int32_t *asm_contig_start, int32_t n_asm, kb_hits_t *dst, int64_t *n_hits, uint32_t *cigar,
                             int64_t cigar_cap, int64_t *n_cigar)
{
    if (!asm_contig_start || n_asm < 0) return fail(KB_ERR_ARG, "null argument");
    const int64_t n_ctg = n_asm ? asm_contig_start[n_asm] : 0;
    std::vector<int64_t> soff((size_t)n_ctg + 1);
    int64_t storage = 0;
    if (kb_packed_layout(contig_len, n_ctg, soff.data(), &storage) != KB_OK) return fail(KB_ERR_ARG, "bad contig lengths");
    soff[(size_t)n_ctg] = storage - 128;
    return map_slabs(ix, asm_contig_start, n_asm, dst, n_hits, cigar, cigar_cap, n_cigar, true,
                     [&](int a0, int a1, int c0, const int32_t *acs, kb_batch_t **b) {
                         return kb_batch_create_packed(seq2, nmask, soff[(size_t)c0], contig_len + c0, acs, a1 - a0, ix->device, b);
                     });
}

int kb_release_workspace(int device)
{
    std::vector<Arena *> drop;
    {
        std::lock_guard<std::mutex> lk(g_arena_mu);
        for (size_t i = g_arenas.size(); i-- > 0;)
            if (device < 0 || g_arenas[i]->device == device) drop.push_back(g_arenas[i]), g_arenas.erase(g_arenas.begin() + (long)i);
    }
    for (Arena *a : drop) {
        if (a->base) {
            cudaSetDevice(a->device);
            cudaFree(a->base);
        }
        delete a;
    }
    return KB_OK;
}

int kb_scan_minimizers(const kb_index_t *ix, const kb_batch_t *bt, int32_t asm_id, uint32_t *hash, int32_t *ctg, uint32_t *pos_strand,
                       int64_t cap, int64_t *n)
{
    if (!ix || !bt || !n) return fail(KB_ERR_ARG, "null argument");
    if (asm_id < 0 || asm_id >= bt->L.n_asm) return fail(KB_ERR_ARG, "assembly index out of range");
    cudaStream_t st = nullptr;
    DevPool P;
    try {
        CU(cudaSetDevice(ix->device));
        CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        P.st = st;
        unsigned long long *dc = P.get<unsigned long long>(KB_N_COUNTERS);
        CU(cudaMemsetAsync(dc, 0, KB_N_COUNTERS * 8, st));
        uint32_t *dh = P.get<uint32_t>((size_t)cap + 1), *dp = P.get<uint32_t>((size_t)cap + 1);
        int32_t *dcg = P.get<int32_t>((size_t)cap + 1);
        kb_launch_scan(ix->view, bt->view, nullptr, nullptr, dc + 16, 0, dh, dcg, dp, cap, asm_id, bt->n_sm, st);
        CU(cudaGetLastError());
        unsigned long long cnt = 0;
        CU(cudaMemcpyAsync(&cnt, dc + 23, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        *n = (int64_t)cnt;
        int64_t m = (int64_t)cnt < cap ? (int64_t)cnt : cap;
        if (m > 0 && hash) CU(cudaMemcpy(hash, dh, (size_t)m * 4, cudaMemcpyDeviceToHost));
        if (m > 0 && ctg) CU(cudaMemcpy(ctg, dcg, (size_t)m * 4, cudaMemcpyDeviceToHost));
        if (m > 0 && pos_strand) CU(cudaMemcpy(pos_strand, dp, (size_t)m * 4, cudaMemcpyDeviceToHost));
        P.clear();
        CU(cudaStreamSynchronize(st));
        cudaStreamDestroy(st);
    } catch (const std::string &e) {
        P.clear();
        if (st) cudaStreamDestroy(st);
        return fail(KB_ERR_CUDA, e);
    }
    return KB_OK;
}

int kb_bench_scan(const kb_index_t *ix, const kb_batch_t *bt, int iters, float *mean_ms, int64_t *n_anchors_out)
{
    if (!ix || !bt || !mean_ms || iters <= 0) return fail(KB_ERR_ARG, "bad argument");
    cudaStream_t st = nullptr;
    DevPool P;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    try {
        CU(cudaSetDevice(ix->device));
        CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        P.st = st;
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        const int64_t hint = bt->last_anchor_count.load(std::memory_order_relaxed);
        int64_t cap = hint > 0 ? hint + 1024 : bt->view.total_bases / 48 + (1 << 20);
        unsigned long long *dc = P.get<unsigned long long>(KB_N_COUNTERS);
        uint64_t *akey = P.get<uint64_t>((size_t)cap);
        uint32_t *aval = P.get<uint32_t>((size_t)cap);
        CU(cudaMemsetAsync(dc, 0, KB_N_COUNTERS * 8, st));
        kb_launch_scan(ix->view, bt->view, akey, aval, dc, cap, nullptr, nullptr, nullptr, 0, -1, bt->n_sm, st);  // warm-up
        CU(cudaStreamSynchronize(st));
        CU(cudaEventRecord(e0, st));
        for (int i = 0; i < iters; ++i) {
            CU(cudaMemsetAsync(dc, 0, KB_N_COUNTERS * 8, st));
            kb_launch_scan(ix->view, bt->view, akey, aval, dc, cap, nullptr, nullptr, nullptr, 0, -1, bt->n_sm, st);
        }
        CU(cudaEventRecord(e1, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        *mean_ms = ms / (float)iters;
        unsigned long long hc[KB_N_COUNTERS];
        CU(cudaMemcpy(hc, dc, sizeof hc, cudaMemcpyDeviceToHost));
        if (n_anchors_out) *n_anchors_out = (int64_t)hc[1];
        P.clear();
        CU(cudaStreamSynchronize(st));
        cudaEventDestroy(e0), cudaEventDestroy(e1);
        cudaStreamDestroy(st);
    } catch (const std::string &e) {
        P.clear();
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (st) cudaStreamDestroy(st);
        return fail(KB_ERR_CUDA, e);
    }
    return KB_OK;
}

}  // extern "C"
