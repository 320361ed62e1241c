// kb_common.cuh -- shared types and deterministic helpers for libkaptive_b200.
//
// Everything marked KB_HD compiles for both the device and the host: the host
// instantiation is what tests/host_emul runs on a machine without a GPU to
// check the logic against the oracle before any GPU time is spent.
#pragma once
#include <stdint.h>
#include <string.h>
#include "../../include/kaptive_b200.h"

#if defined(__CUDACC__)
#define KB_HD __host__ __device__ __forceinline__
#define KB_D __device__ __forceinline__
#else
#define KB_HD inline
#define KB_D inline
#endif

// Re-converge the lanes of a warp (device only).  One-thread-per-item kernels run long data-dependent loops; lanes that
// left an earlier loop at different times are not guaranteed to meet again before the next one unless told to.  Every
// lane of the warp that has not exited must reach the same call.
#ifdef __CUDA_ARCH__
#define KB_WARP_RECONVERGE() __syncwarp()
#else
#define KB_WARP_RECONVERGE() ((void)0)
#endif

#define KB_MAXU 0xffffffffu
#define KB_SEED_LONG_JOIN (1ULL << 40)
#define KB_SEED_IGNORE (1ULL << 41)
#define KB_SEED_TANDEM (1ULL << 42)
#define KB_NEG_INF (-0x20000000)

// anchor sort key: asm(17) | gene(15) | rev(1) | vpos(27)
#define KB_KEY_ASM_SHIFT 43
#define KB_KEY_GENE_SHIFT 28
#define KB_KEY_REV_SHIFT 27
#define KB_VPOS_MASK ((1u << 27) - 1)
#define KB_MAX_ASM (1 << 17)
#define KB_MAX_GENES (1 << 15)
#define KB_CTG_VGAP 8192  // virtual gap between contigs in vpos space: > max_gap + bw, so chains never cross

// hash table slot: key(30) << 34 | start(23) << 11 | count(11); 0 = empty
#define KB_HT_KEY_SHIFT 34
#define KB_HT_START_SHIFT 11
#define KB_HT_START_MASK ((1u << 23) - 1)
#define KB_HT_COUNT_MASK ((1u << 11) - 1)

// ------------------------------------------------------------ deterministic float
KB_HD float kb_u2f(uint32_t i)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(i);
#else
    float f;
    memcpy(&f, &i, 4);
    return f;
#endif
}
KB_HD uint32_t kb_f2u(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t i;
    memcpy(&i, &f, 4);
    return i;
#endif
}
// Round-to-nearest, never contracted into FMA (nvcc would fuse a*b+c otherwise).
KB_HD float kb_fmul(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}
KB_HD float kb_fadd(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
KB_HD float kb_fsub(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    volatile float r = a - b;
    return r;
#endif
}
KB_HD float kb_fdiv(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b;
    return r;
#endif
}
KB_HD double kb_dmul(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
KB_HD double kb_dadd(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}

// minimap2's fast log2 (mmpriv.h mg_log2); valid for x >= 2
KB_HD float kb_log2_fast(float x)
{
    uint32_t zi = kb_f2u(x);
    float log_2 = (float)((int)((zi >> 23) & 255) - 128);
    zi &= ~(255u << 23);
    zi += 127u << 23;
    float zf = kb_u2f(zi);
    float t = kb_fmul(-0.34484843f, zf);
    t = kb_fadd(t, 2.02466578f);
    t = kb_fmul(t, zf);
    t = kb_fsub(t, 0.67487759f);
    return kb_fadd(log_2, t);
}

// fdlibm logf polynomial, op by op; x > 0 finite. Same sequence as the oracle's.
KB_HD float kb_logf(float x)
{
    const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f;
    const float Lg1 = 0.66666662693f, Lg2 = 0.40000972152f, Lg3 = 0.28498786688f, Lg4 = 0.24279078841f;
    uint32_t ix = kb_f2u(x);
    if (ix == 0x3f800000u) return 0.0f;
    ix += 0x3f800000u - 0x3f3504f3u;
    int k = (int)(ix >> 23) - 0x7f;
    ix = (ix & 0x007fffffu) + 0x3f3504f3u;
    x = kb_u2f(ix);
    float f = kb_fsub(x, 1.0f);
    float s = kb_fdiv(f, kb_fadd(2.0f, f));
    float z = kb_fmul(s, s);
    float w = kb_fmul(z, z);
    float t1 = kb_fmul(w, kb_fadd(Lg2, kb_fmul(w, Lg4)));
    float t2 = kb_fmul(z, kb_fadd(Lg1, kb_fmul(w, Lg3)));
    float R = kb_fadd(t2, t1);
    float hfsq = kb_fmul(kb_fmul(0.5f, f), f);
    float dk = (float)k;
    float r = kb_fmul(s, kb_fadd(hfsq, R));
    r = kb_fadd(r, kb_fmul(dk, ln2_lo));
    r = kb_fsub(r, hfsq);
    r = kb_fadd(r, f);
    r = kb_fadd(r, kb_fmul(dk, ln2_hi));
    return r;
}

// minimap2 sketch.c hash64 restricted to <= 30 bits: pure 32-bit arithmetic
KB_HD uint32_t kb_hash32(uint32_t key, uint32_t mask)
{
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    // (key + (key << 31)) & mask is the identity for mask < 2^31
    return key;
}

KB_HD uint64_t kb_hash64_full(uint64_t key)
{
    key = ~key + (key << 21);
    key = key ^ key >> 24;
    key = (key + (key << 3)) + (key << 8);
    key = key ^ key >> 14;
    key = (key + (key << 2)) + (key << 4);
    key = key ^ key >> 28;
    key = key + (key << 31);
    return key;
}

KB_HD uint32_t kb_wang_hash(uint32_t key)
{
    key += ~(key << 15);
    key ^= (key >> 10);
    key += (key << 3);
    key ^= (key >> 6);
    key += ~(key << 11);
    key ^= (key >> 16);
    return key;
}

KB_HD uint8_t kb_nt4(uint8_t c)
{
    // A/a C/c G/g T/t U/u -> 0..3, anything else 4
    uint8_t u = c & 0xdf;  // upper-case
    return u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : (u == 'T' || u == 'U') ? 3 : 4;
}

// ------------------------------------------------------------ device views

// one gene minimizer occurrence ("entry"), entries sorted by (hash, gene, pos)
struct KbEntry {
    int32_t gene;
    uint32_t qpos_z;   // last base of the k-mer in the gene << 1 | strand
    uint32_t mi_flags; // ordinal of this minimizer within its gene | tandem << 31
    int32_t qocc;      // occurrences of this hash within the gene
};

struct KbIndexView {
    kb_params_t p;
    int32_t n_genes;
    int64_t n_entries;
    uint32_t ht_mask;            // slots - 1
    const uint64_t *ht;          // hash -> (start, count) into entries
    const KbEntry *ent;
    const int32_t *gene_len;
    const int32_t *gene_nmin;    // minimizers per gene (unfiltered)
    const int64_t *gene_min_off; // offset of the gene's minimizer list in gm_*
    const uint32_t *gm_qpos_z;   // per gene, in query order
    const int32_t *gm_qocc;
    const uint32_t *gene_hash;   // per-query tie-break hash
    const int64_t *gene_seq_off; // into gseq_fwd / gseq_rev
    const uint8_t *gseq_fwd;     // nt4 codes, 1 byte per base
    const uint8_t *gseq_rev;     // reverse complement
    const uint32_t *bloom;       // device only: 1 bit per (hash & bloom_mask), set for every indexed minimizer
    uint32_t bloom_mask;
};

struct KbBatchView {
    int32_t n_asm, n_ctg;
    int64_t n_chunks;
    int64_t total_bases;
    const uint32_t *seq2;        // 16 bases per word, base b at bits 2*(b&15)
    const uint32_t *nmask;       // 32 bases per word, bit set = ambiguous
    const int64_t *ctg_soff;     // storage offset in bases (multiple of 128)
    const int32_t *ctg_len;
    const int32_t *ctg_asm;
    const int32_t *ctg_vstart;   // virtual position of base 0 within its assembly
    const int32_t *asm_ctg_start; // n_asm + 1
    const int32_t *chunk_ctg;    // scan work list: one warp per chunk
    const int32_t *chunk_start;
};

#define KB_LANE_BASES 256
#define KB_CHUNK_BASES (32 * KB_LANE_BASES)
#define KB_SCAN_LOOKBACK 24  // w + k - 1: enough to rebuild the sketch state exactly

KB_HD int kb_fetch_base(const uint32_t *seq2, const uint32_t *nmask, int64_t b)
{
    if ((nmask[b >> 5] >> (b & 31)) & 1u) return 4;
    return (int)((seq2[b >> 4] >> (2 * (b & 15))) & 3u);
}
