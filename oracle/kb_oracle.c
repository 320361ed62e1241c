/* kb_oracle.c -- CPU oracle (see kb_oracle.h for scope, status and deviations).
 *
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (no rammappy here, no golden
 * vectors in the reference).  Each function cites what it restates:
 *   [mm2:<file>:<function>]  the published minimap2 routine (third party,
 *                            pinned via rammappy 0.1.3 "minimap2-based",
 *                            /root/reference/docs/serotyping/method.md:23-25)
 *   [ref:<file>:<line>]      the call site / consumer in /root/reference
 *
 * Direction follows the reference exactly: the ASSEMBLY is indexed
 * (ref:src/kaptive/core/genome.py:188-189) and every DB gene is a query
 * (ref:src/kaptive/serotyping/core.py:111-121,154).
 */
#include "kb_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <assert.h>

typedef struct { uint64_t x, y; } mm128_t;

#define SEED_LONG_JOIN (1ULL << 40)
#define SEED_IGNORE    (1ULL << 41)
#define SEED_TANDEM    (1ULL << 42)
#define NEG_INF        (-0x20000000)
#define PARENT_UNSET   (-1)
#define PARENT_TMP_PRI (-2)

/* ------------------------------------------------------------------ params */

void kbo_params_default(kbo_params_t *p) /* [mm2:options.c:mm_idxopt_init/mm_mapopt_init] + ref overrides */
{
    memset(p, 0, sizeof(*p));
    p->k = 15; p->w = 10;
    p->min_cnt = 3; p->min_chain_score = 40; p->bw = 500; p->max_gap = 5000;
    p->max_chain_skip = 25; p->max_chain_iter = 5000;
    p->chain_gap_scale = 0.8f;
    p->a = 2; p->b = 4; p->q = 4; p->e = 2; p->q2 = 24; p->e2 = 1; p->sc_ambi = 1;
    p->zdrop = 400; p->min_dp_max = 80; p->min_ksw_len = 200;
    p->mid_occ = 0; p->min_mid_occ = 10; p->max_mid_occ = 1000000;
    p->mid_occ_frac = 2e-4f; p->q_occ_frac = 0.01f; p->mask_level = 0.5f;
    p->mask_len = INT_MAX;
    p->seed = 11;
    p->ext_bw = (int)(500 * 1.5 + 1.);
    p->max_sw_cells = 100000000; /* [mm2:options.c:mm_mapopt_init] max_sw_mat */
}

/* ------------------------------------------------------- deterministic math */

static inline float u2f(uint32_t i) { float f; memcpy(&f, &i, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t i; memcpy(&i, &f, 4); return i; }

/* [mm2:mmpriv.h:mg_log2]; every operation is a separate IEEE binary32 op */
float kbo_log2_fast(float x)
{
    uint32_t zi = f2u(x);
    float log_2 = (float)((int)((zi >> 23) & 255) - 128);
    float zf, t;
    zi &= ~(255u << 23);
    zi += 127u << 23;
    zf = u2f(zi);
    t = -0.34484843f * zf;
    t = t + 2.02466578f;
    t = t * zf;
    t = t - 0.67487759f;
    return log_2 + t;
}

/* fdlibm/musl logf polynomial, evaluated op by op (no contraction), x > 0 finite */
float kbo_logf(float x)
{
    static const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f;
    static const float Lg1 = 0.66666662693f, Lg2 = 0.40000972152f, Lg3 = 0.28498786688f, Lg4 = 0.24279078841f;
    uint32_t ix = f2u(x);
    int k;
    float f, s, z, w, t1, t2, R, hfsq, dk, r;
    if (ix == 0x3f800000u) return 0.0f;
    ix += 0x3f800000u - 0x3f3504f3u;
    k = (int)(ix >> 23) - 0x7f;
    ix = (ix & 0x007fffffu) + 0x3f3504f3u;
    x = u2f(ix);
    f = x - 1.0f;
    s = f / (2.0f + f);
    z = s * s;
    w = z * z;
    t1 = w * Lg4; t1 = Lg2 + t1; t1 = w * t1;
    t2 = w * Lg3; t2 = Lg1 + t2; t2 = z * t2;
    R = t2 + t1;
    hfsq = 0.5f * f; hfsq = hfsq * f;
    dk = (float)k;
    r = hfsq + R; r = s * r;
    r = r + dk * ln2_lo;
    r = r - hfsq;
    r = r + f;
    r = r + dk * ln2_hi;
    return r;
}

/* [mm2:sketch.c:hash64] restricted to 2k<=30 bits so it is pure 32-bit arithmetic */
uint32_t kbo_hash32(uint32_t key, uint32_t mask)
{
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

static inline uint64_t hash64_mask(uint64_t key, uint64_t mask) /* [mm2:sketch.c:hash64] */
{
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

static inline uint64_t hash64_full(uint64_t key) /* [mm2:hit.c:hash64] */
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

static inline uint32_t wang_hash(uint32_t key) /* [mm2:khash.h:__ac_Wang_hash] */
{
    key += ~(key << 15);
    key ^= (key >> 10);
    key += (key << 3);
    key ^= (key >> 6);
    key += ~(key << 11);
    key ^= (key >> 16);
    return key;
}

static inline uint32_t x31_hash_string(const char *s) /* [mm2:khash.h:__ac_X31_hash_string] */
{
    uint32_t h = (uint32_t)*s;
    if (h) for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)*s;
    return h;
}

static inline uint8_t nt4(uint8_t c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return 4;
    }
}

/* ------------------------------------------------------------------ sketch */

typedef struct { size_t n, m; mm128_t *a; } mm128_v;

static inline void v_push(mm128_v *v, mm128_t e)
{
    if (v->n == v->m) { v->m = v->m ? v->m << 1 : 256; v->a = (mm128_t *)realloc(v->a, v->m * sizeof(mm128_t)); }
    v->a[v->n++] = e;
}

/* [mm2:sketch.c:mm_sketch], non-HPC; input already nt4-encoded */
static void sketch_nt4(const uint8_t *s, int len, int w, int k, uint32_t rid, mm128_v *p)
{
    uint64_t shift1 = 2 * (k - 1), mask = (1ULL << 2 * k) - 1, kmer[2] = {0, 0};
    int i, j, l, buf_pos, min_pos, kmer_span = 0;
    mm128_t buf[256], min = {UINT64_MAX, UINT64_MAX};
    if (len <= 0) return;
    memset(buf, 0xff, (size_t)w * 16);
    for (i = l = buf_pos = min_pos = 0; i < len; ++i) {
        int c = s[i];
        mm128_t info = {UINT64_MAX, UINT64_MAX};
        if (c < 4) {
            int z;
            kmer_span = l + 1 < k ? l + 1 : k;
            kmer[0] = (kmer[0] << 2 | (uint64_t)c) & mask;
            kmer[1] = (kmer[1] >> 2) | (3ULL ^ (uint64_t)c) << shift1;
            if (kmer[0] == kmer[1]) continue;
            z = kmer[0] < kmer[1] ? 0 : 1;
            ++l;
            if (l >= k && kmer_span < 256) {
                info.x = hash64_mask(kmer[z], mask) << 8 | (uint64_t)kmer_span;
                info.y = (uint64_t)rid << 32 | (uint32_t)i << 1 | (uint32_t)z;
            }
        } else l = 0, kmer_span = 0;
        buf[buf_pos] = info;
        if (l == w + k - 1 && min.x != UINT64_MAX) {
            for (j = buf_pos + 1; j < w; ++j)
                if (min.x == buf[j].x && buf[j].y != min.y) v_push(p, buf[j]);
            for (j = 0; j < buf_pos; ++j)
                if (min.x == buf[j].x && buf[j].y != min.y) v_push(p, buf[j]);
        }
        if (info.x <= min.x) {
            if (l >= w + k && min.x != UINT64_MAX) v_push(p, min);
            min = info, min_pos = buf_pos;
        } else if (buf_pos == min_pos) {
            if (l >= w + k - 1 && min.x != UINT64_MAX) v_push(p, min);
            for (j = buf_pos + 1, min.x = UINT64_MAX; j < w; ++j)
                if (min.x >= buf[j].x) min = buf[j], min_pos = j;
            for (j = 0; j <= buf_pos; ++j)
                if (min.x >= buf[j].x) min = buf[j], min_pos = j;
            if (l >= w + k - 1 && min.x != UINT64_MAX) {
                for (j = buf_pos + 1; j < w; ++j)
                    if (min.x == buf[j].x && min.y != buf[j].y) v_push(p, buf[j]);
                for (j = 0; j <= buf_pos; ++j)
                    if (min.x == buf[j].x && min.y != buf[j].y) v_push(p, buf[j]);
            }
        }
        if (++buf_pos == w) buf_pos = 0;
    }
    if (min.x != UINT64_MAX) v_push(p, min);
}

int64_t kbo_sketch(const uint8_t *seq, int32_t len, int32_t w, int32_t k, uint64_t *out_x, uint32_t *out_y, int64_t cap)
{
    mm128_v v = {0, 0, 0};
    uint8_t *s = (uint8_t *)malloc(len > 0 ? (size_t)len : 1);
    int64_t i, n;
    for (i = 0; i < len; ++i) s[i] = nt4(seq[i]);
    sketch_nt4(s, len, w, k, 0, &v);
    n = (int64_t)v.n;
    for (i = 0; i < n && i < cap; ++i) out_x[i] = v.a[i].x, out_y[i] = (uint32_t)v.a[i].y;
    free(v.a); free(s);
    return n;
}

/* --------------------------------------------------------------------- db */

typedef struct {
    kbo_params_t p;
    int32_t n_genes;
    int32_t *len;
    uint8_t **fwd, **rev;   /* nt4 codes, query and its reverse complement */
    mm128_t **mv;           /* unfiltered query minimizers */
    int32_t *n_mv;
    uint8_t **tandem;       /* per query minimizer: equal hash to a neighbour */
    int32_t **qocc;         /* per query minimizer: occurrences of its hash in this query */
    uint32_t *qhash;        /* per-query tie-break hash [mm2:map.c:mm_map_frag] */
} kbo_db_t;

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

void *kbo_db_create(const uint8_t *seqs, const int64_t *offsets, const int32_t *lengths, int32_t n_genes, const kbo_params_t *p)
{
    kbo_db_t *db = (kbo_db_t *)calloc(1, sizeof(kbo_db_t));
    int32_t g;
    db->p = *p; db->n_genes = n_genes;
    db->len = (int32_t *)calloc(n_genes, 4);
    db->fwd = (uint8_t **)calloc(n_genes, sizeof(void *));
    db->rev = (uint8_t **)calloc(n_genes, sizeof(void *));
    db->mv = (mm128_t **)calloc(n_genes, sizeof(void *));
    db->n_mv = (int32_t *)calloc(n_genes, 4);
    db->tandem = (uint8_t **)calloc(n_genes, sizeof(void *));
    db->qocc = (int32_t **)calloc(n_genes, sizeof(void *));
    db->qhash = (uint32_t *)calloc(n_genes, 4);
    for (g = 0; g < n_genes; ++g) {
        int32_t L = lengths[g], i;
        const uint8_t *s = seqs + offsets[g];
        mm128_v v = {0, 0, 0};
        char name[16];
        uint32_t h;
        db->len[g] = L;
        db->fwd[g] = (uint8_t *)malloc(L > 0 ? (size_t)L : 1);
        db->rev[g] = (uint8_t *)malloc(L > 0 ? (size_t)L : 1);
        for (i = 0; i < L; ++i) {
            db->fwd[g][i] = nt4(s[i]);
            db->rev[g][L - 1 - i] = db->fwd[g][i] < 4 ? 3 - db->fwd[g][i] : 4; /* [mm2:align.c:mm_align_skeleton] */
        }
        sketch_nt4(db->fwd[g], L, p->w, p->k, 0, &v);
        db->mv[g] = v.a; db->n_mv[g] = (int32_t)v.n;
        db->tandem[g] = (uint8_t *)calloc(v.n ? v.n : 1, 1);
        db->qocc[g] = (int32_t *)calloc(v.n ? v.n : 1, 4);
        /* [mm2:seed.c:mm_seed_collect_all] is_tandem; spec v1 evaluates it on the unfiltered list */
        for (i = 0; i < (int32_t)v.n; ++i) {
            if (i > 0 && v.a[i].x >> 8 == v.a[i - 1].x >> 8) db->tandem[g][i] = 1;
            if (i < (int32_t)v.n - 1 && v.a[i].x >> 8 == v.a[i + 1].x >> 8) db->tandem[g][i] = 1;
        }
        { /* occurrences of each hash inside the query, for [mm2:seed.c:mm_seed_mz_flt] */
            uint64_t *t = (uint64_t *)malloc((v.n ? v.n : 1) * 8);
            int32_t st;
            for (i = 0; i < (int32_t)v.n; ++i) t[i] = (v.a[i].x >> 8) << 24 | (uint64_t)i; /* hash is 30 bit, i < 2^24 */
            qsort(t, v.n, 8, cmp_u64);
            for (st = 0, i = 1; i <= (int32_t)v.n; ++i)
                if (i == (int32_t)v.n || t[i] >> 24 != t[st] >> 24) {
                    int32_t j;
                    for (j = st; j < i; ++j) db->qocc[g][t[j] & 0xffffff] = i - st;
                    st = i;
                }
            free(t);
        }
        /* the reference names query i str(i) (ref:serotyping/core.py:113) */
        {
            int n = 0, x = g; char tmp[16];
            if (x == 0) tmp[n++] = '0';
            while (x > 0) tmp[n++] = (char)('0' + x % 10), x /= 10;
            for (i = 0; i < n; ++i) name[i] = tmp[n - 1 - i];
            name[n] = 0;
        }
        h = x31_hash_string(name);
        h ^= wang_hash((uint32_t)L) + wang_hash((uint32_t)p->seed);
        db->qhash[g] = wang_hash(h);
    }
    return db;
}

void kbo_db_destroy(void *db_)
{
    kbo_db_t *db = (kbo_db_t *)db_;
    int32_t g;
    if (!db) return;
    for (g = 0; g < db->n_genes; ++g) { free(db->fwd[g]); free(db->rev[g]); free(db->mv[g]); free(db->tandem[g]); free(db->qocc[g]); }
    free(db->len); free(db->fwd); free(db->rev); free(db->mv); free(db->n_mv); free(db->tandem); free(db->qocc); free(db->qhash);
    free(db);
}

/* --------------------------------------------------------- assembly index */

typedef struct {
    int32_t n_ctg;
    const int32_t *ctg_len;
    uint8_t **ctg;        /* nt4 */
    mm128_t *mz; int64_t n_mz; /* sorted by (x, y) */
    int32_t mid_occ;
} asm_idx_t;

static int cmp_128xy(const void *a, const void *b)
{
    const mm128_t *p = (const mm128_t *)a, *q = (const mm128_t *)b;
    if (p->x != q->x) return p->x < q->x ? -1 : 1;
    return p->y < q->y ? -1 : p->y > q->y;
}

static int cmp_u32(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : x > y;
}

/* [mm2:index.c:mm_idx_cal_max_occ] + [mm2:options.c:mm_mapopt_update] */
static int32_t cal_mid_occ(const asm_idx_t *ai, const kbo_params_t *p)
{
    int64_t i, st, n = 0;
    uint32_t *a, thres;
    int32_t mid;
    if (p->mid_occ > 0) return p->mid_occ;
    if (ai->n_mz == 0) mid = INT32_MAX;
    else {
        a = (uint32_t *)malloc((size_t)ai->n_mz * 4);
        for (st = 0, i = 1; i <= ai->n_mz; ++i)
            if (i == ai->n_mz || ai->mz[i].x >> 8 != ai->mz[st].x >> 8) a[n++] = (uint32_t)(i - st), st = i;
        qsort(a, (size_t)n, 4, cmp_u32);
        thres = a[(uint32_t)((1. - p->mid_occ_frac) * n)] + 1;
        free(a);
        mid = (int32_t)thres;
    }
    if (mid < p->min_mid_occ) mid = p->min_mid_occ;
    if (p->max_mid_occ > p->min_mid_occ && mid > p->max_mid_occ) mid = p->max_mid_occ;
    return mid;
}

static void idx_get(const asm_idx_t *ai, uint64_t minier, int64_t *lo_, int64_t *hi_)
{ /* [mm2:index.c:mm_idx_get]: all entries whose hash equals minier */
    int64_t lo = 0, hi = ai->n_mz, a, b;
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (ai->mz[m].x >> 8 < minier) lo = m + 1; else hi = m; }
    a = lo; hi = ai->n_mz;
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (ai->mz[m].x >> 8 <= minier) lo = m + 1; else hi = m; }
    b = lo;
    *lo_ = a, *hi_ = b;
}

/* ---------------------------------------------------------------- chaining */

/* [mm2:lchain.c:comput_sc], n_seg==1, not cDNA */
static inline int32_t comput_sc(const mm128_t *ai, const mm128_t *aj, int32_t max_dist_x, int32_t max_dist_y, int32_t bw, float chn_pen_gap, float chn_pen_skip)
{
    int32_t dq = (int32_t)ai->y - (int32_t)aj->y, dr, dd, dg, q_span, sc;
    if (dq <= 0 || dq > max_dist_x) return INT32_MIN;
    dr = (int32_t)(ai->x - aj->x);
    if (dr == 0 || dq > max_dist_y) return INT32_MIN;
    dd = dr > dq ? dr - dq : dq - dr;
    if (dd > bw) return INT32_MIN;
    dg = dr < dq ? dr : dq;
    q_span = (int32_t)(aj->y >> 32 & 0xff);
    sc = q_span < dg ? q_span : dg;
    if (dd || dg > q_span) {
        float lin_pen, log_pen, t0, t1;
        t0 = chn_pen_gap * (float)dd;
        t1 = chn_pen_skip * (float)dg;
        lin_pen = t0 + t1;
        log_pen = dd >= 1 ? kbo_log2_fast((float)(dd + 1)) : 0.0f;
        t0 = .5f * log_pen;
        t0 = lin_pen + t0;
        sc -= (int)t0;
    }
    return sc;
}

typedef struct { int64_t key; int64_t idx; } zrec_t;
static int cmp_zrec(const void *a, const void *b)
{
    const zrec_t *p = (const zrec_t *)a, *q = (const zrec_t *)b;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    return p->idx < q->idx ? -1 : p->idx > q->idx;
}

/* [mm2:lchain.c:mg_chain_bk_end] */
static int64_t chain_bk_end(int32_t max_drop, const zrec_t *z, const int32_t *f, const int64_t *p, int32_t *t, int64_t k)
{
    int64_t i = z[k].idx, end_i = -1, max_i = i;
    int32_t max_s = 0;
    if (i < 0 || t[i] != 0) return i;
    do {
        int32_t s;
        t[i] = 2;
        end_i = i = p[i];
        s = i < 0 ? (int32_t)z[k].key : (int32_t)z[k].key - f[i];
        if (s > max_s) max_s = s, max_i = i;
        else if (max_s - s > max_drop) break;
    } while (i >= 0 && t[i] == 0);
    for (i = z[k].idx; i >= 0 && i != end_i; i = p[i]) t[i] = 0;
    return max_i;
}

/* [mm2:lchain.c:mg_chain_backtrack]; z ties ordered by anchor index */
static uint64_t *chain_backtrack(int64_t n, const int32_t *f, const int64_t *p, int32_t *v, int32_t *t, int32_t min_cnt, int32_t min_sc, int32_t max_drop, int32_t *n_u_, int32_t *n_v_)
{
    zrec_t *z;
    uint64_t *u;
    int64_t i, k, n_z, n_v;
    int32_t n_u;
    *n_u_ = *n_v_ = 0;
    for (i = 0, n_z = 0; i < n; ++i) if (f[i] >= min_sc) ++n_z;
    if (n_z == 0) return 0;
    z = (zrec_t *)malloc((size_t)n_z * sizeof(zrec_t));
    for (i = 0, k = 0; i < n; ++i) if (f[i] >= min_sc) z[k].key = f[i], z[k++].idx = i;
    qsort(z, (size_t)n_z, sizeof(zrec_t), cmp_zrec);
    memset(t, 0, (size_t)n * 4);
    for (k = n_z - 1, n_v = n_u = 0; k >= 0; --k) {
        if (t[z[k].idx] == 0) {
            int64_t n_v0 = n_v, end_i;
            int32_t sc;
            end_i = chain_bk_end(max_drop, z, f, p, t, k);
            for (i = z[k].idx; i != end_i; i = p[i]) ++n_v, t[i] = 1;
            sc = i < 0 ? (int32_t)z[k].key : (int32_t)z[k].key - f[i];
            if (sc >= min_sc && n_v > n_v0 && n_v - n_v0 >= min_cnt) ++n_u;
            else n_v = n_v0;
        }
    }
    u = (uint64_t *)malloc((size_t)(n_u ? n_u : 1) * 8);
    memset(t, 0, (size_t)n * 4);
    for (k = n_z - 1, n_v = n_u = 0; k >= 0; --k) {
        if (t[z[k].idx] == 0) {
            int64_t n_v0 = n_v, end_i;
            int32_t sc;
            end_i = chain_bk_end(max_drop, z, f, p, t, k);
            for (i = z[k].idx; i != end_i; i = p[i]) v[n_v++] = (int32_t)i, t[i] = 1;
            sc = i < 0 ? (int32_t)z[k].key : (int32_t)z[k].key - f[i];
            if (sc >= min_sc && n_v > n_v0 && n_v - n_v0 >= min_cnt) u[n_u++] = (uint64_t)sc << 32 | (uint64_t)(n_v - n_v0);
            else n_v = n_v0;
        }
    }
    free(z);
    *n_u_ = n_u, *n_v_ = (int32_t)n_v;
    return u;
}

/* [mm2:lchain.c:mg_lchain_dp] + [mm2:lchain.c:compact_a]; returns compacted anchors (malloc) */
static mm128_t *lchain_dp(const kbo_params_t *P, float chn_pen_gap, float chn_pen_skip, int64_t n, mm128_t *a, int *n_u_, uint64_t **_u)
{
    int32_t max_dist_x = P->max_gap, max_dist_y = P->max_gap, bw = P->bw, max_skip = P->max_chain_skip, max_iter = P->max_chain_iter;
    int32_t *f, *t, *v, n_u, n_v, max_drop = bw;
    int64_t *p, i, j, max_ii, st = 0;
    uint64_t *u;
    mm128_t *b;
    *_u = 0, *n_u_ = 0;
    if (n == 0) return 0;
    if (max_dist_x < bw) max_dist_x = bw;
    if (max_dist_y < bw) max_dist_y = bw;
    p = (int64_t *)malloc((size_t)n * 8);
    f = (int32_t *)malloc((size_t)n * 4);
    v = (int32_t *)malloc((size_t)n * 4);
    t = (int32_t *)calloc((size_t)n, 4);
    for (i = 0, max_ii = -1; i < n; ++i) {
        int64_t max_j = -1, end_j;
        int32_t max_f = (int32_t)(a[i].y >> 32 & 0xff), n_skip = 0;
        while (st < i && (a[i].x >> 32 != a[st].x >> 32 || a[i].x > a[st].x + (uint64_t)max_dist_x)) ++st;
        if (i - st > max_iter) st = i - max_iter;
        for (j = i - 1; j >= st; --j) {
            int32_t sc = comput_sc(&a[i], &a[j], max_dist_x, max_dist_y, bw, chn_pen_gap, chn_pen_skip);
            if (sc == INT32_MIN) continue;
            sc += f[j];
            if (sc > max_f) {
                max_f = sc, max_j = j;
                if (n_skip > 0) --n_skip;
            } else if (t[j] == (int32_t)i) {
                if (++n_skip > max_skip) break;
            }
            if (p[j] >= 0) t[p[j]] = (int32_t)i;
        }
        end_j = j;
        if (max_ii < 0 || a[i].x - a[max_ii].x > (uint64_t)max_dist_x) {
            int32_t max = INT32_MIN;
            max_ii = -1;
            for (j = i - 1; j >= st; --j)
                if (max < f[j]) max = f[j], max_ii = j;
        }
        if (max_ii >= 0 && max_ii < end_j) {
            int32_t tmp = comput_sc(&a[i], &a[max_ii], max_dist_x, max_dist_y, bw, chn_pen_gap, chn_pen_skip);
            if (tmp != INT32_MIN && max_f < tmp + f[max_ii]) max_f = tmp + f[max_ii], max_j = max_ii;
        }
        f[i] = max_f, p[i] = max_j;
        v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f;
        if (max_ii < 0 || (a[i].x - a[max_ii].x <= (uint64_t)max_dist_x && f[max_ii] < f[i])) max_ii = i;
    }
    u = chain_backtrack(n, f, p, v, t, P->min_cnt, P->min_chain_score, max_drop, &n_u, &n_v);
    free(p); free(f); free(t);
    *n_u_ = n_u, *_u = u;
    if (n_u == 0) { free(v); free(u); *_u = 0; return 0; }
    /* compact_a */
    b = (mm128_t *)malloc((size_t)n_v * sizeof(mm128_t));
    {
        int64_t k;
        uint64_t *u2;
        mm128_t *c;
        for (i = 0, k = 0; i < n_u; ++i) {
            int32_t k0 = (int32_t)k, ni = (int32_t)u[i];
            for (j = 0; j < ni; ++j) b[k++] = a[v[k0 + (ni - j - 1)]];
        }
        free(v);
        /* sort chains by the target position of their first anchor; ties by original order */
        c = (mm128_t *)malloc((size_t)n_v * sizeof(mm128_t));
        u2 = (uint64_t *)malloc((size_t)n_u * 8);
        {
            /* key must be compared unsigned: strand lives in bit 63 */
            typedef struct { uint64_t x; int64_t k; int32_t i; } wrec_t;
            wrec_t *ww = (wrec_t *)malloc((size_t)n_u * sizeof(wrec_t));
            int32_t ii, jj;
            for (i = k = 0; i < n_u; ++i) { ww[i].x = b[k].x; ww[i].k = k; ww[i].i = (int32_t)i; k += (int32_t)u[i]; }
            for (ii = 1; ii < n_u; ++ii) { /* insertion sort: n_u is small */
                wrec_t tmp = ww[ii];
                for (jj = ii - 1; jj >= 0 && (ww[jj].x > tmp.x || (ww[jj].x == tmp.x && ww[jj].k > tmp.k)); --jj) ww[jj + 1] = ww[jj];
                ww[jj + 1] = tmp;
            }
            for (i = k = 0; i < n_u; ++i) {
                int32_t nn = (int32_t)u[ww[i].i];
                u2[i] = u[ww[i].i];
                memcpy(&c[k], &b[ww[i].k], (size_t)nn * sizeof(mm128_t));
                k += nn;
            }
            free(ww);
        }
        memcpy(u, u2, (size_t)n_u * 8);
        free(u2); free(b);
        return c;
    }
}

/* -------------------------------------------------------------------- regs */

typedef struct {
    int32_t id, cnt, rid, score, qs, qe, rs, re, parent, subsc, as, mlen, blen, n_sub, score0;
    uint32_t hash;
    int32_t rev, mapq;
    /* mm_extra_t */
    int32_t has_p, dp_score, dp_max, dp_max2, n_ambi, n_cigar, m_cigar;
    uint32_t *cigar;
} reg_t;

/* [mm2:hit.c:mm_reg_set_coor] */
static void reg_set_coor(reg_t *r, int32_t qlen, const mm128_t *a)
{
    int32_t k = r->as, q_span = (int32_t)(a[k].y >> 32 & 0xff);
    r->rev = (int32_t)(a[k].x >> 63);
    r->rid = (int32_t)(a[k].x << 1 >> 33);
    r->rs = (int32_t)a[k].x + 1 > q_span ? (int32_t)a[k].x + 1 - q_span : 0;
    r->re = (int32_t)a[k + r->cnt - 1].x + 1;
    if (!r->rev) {
        r->qs = (int32_t)a[k].y + 1 - q_span;
        r->qe = (int32_t)a[k + r->cnt - 1].y + 1;
    } else {
        r->qs = qlen - ((int32_t)a[k + r->cnt - 1].y + 1);
        r->qe = qlen - ((int32_t)a[k].y + 1 - q_span);
    }
}

/* [mm2:hit.c:mm_cal_fuzzy_len] */
static void cal_fuzzy_len(reg_t *r, const mm128_t *a)
{
    int i;
    r->mlen = r->blen = 0;
    if (r->cnt <= 0) return;
    r->mlen = r->blen = (int32_t)(a[r->as].y >> 32 & 0xff);
    for (i = r->as + 1; i < r->as + r->cnt; ++i) {
        int span = (int)(a[i].y >> 32 & 0xff);
        int tl = (int32_t)a[i].x - (int32_t)a[i - 1].x;
        int ql = (int32_t)a[i].y - (int32_t)a[i - 1].y;
        r->blen += tl > ql ? tl : ql;
        r->mlen += tl > span && ql > span ? span : tl < ql ? tl : ql;
    }
}

typedef struct { uint64_t key; int32_t idx; } krec_t;
/* descending by key; equal keys: larger original index first (= stable ascending sort, reversed) */
static int cmp_krec_desc(const void *a, const void *b)
{
    const krec_t *p = (const krec_t *)a, *q = (const krec_t *)b;
    if (p->key != q->key) return p->key > q->key ? -1 : 1;
    return p->idx > q->idx ? -1 : p->idx < q->idx;
}

/* [mm2:hit.c:mm_gen_regs] */
static reg_t *gen_regs(uint32_t hash, int qlen, int n_u, const uint64_t *u, const mm128_t *a)
{
    krec_t *z;
    int32_t *ks;
    reg_t *r;
    int i, k;
    if (n_u == 0) return 0;
    z = (krec_t *)malloc((size_t)n_u * sizeof(krec_t));
    ks = (int32_t *)malloc((size_t)n_u * 4);
    for (i = k = 0; i < n_u; ++i) {
        uint32_t h = (uint32_t)hash64_full((hash64_full(a[k].x) + hash64_full(a[k].y)) ^ hash);
        z[i].key = u[i] ^ h;
        z[i].idx = i;
        ks[i] = k;
        k += (int32_t)u[i];
    }
    qsort(z, (size_t)n_u, sizeof(krec_t), cmp_krec_desc);
    r = (reg_t *)calloc((size_t)n_u, sizeof(reg_t));
    for (i = 0; i < n_u; ++i) {
        reg_t *ri = &r[i];
        ri->id = i;
        ri->parent = PARENT_UNSET;
        ri->score = ri->score0 = (int32_t)(z[i].key >> 32);
        ri->hash = (uint32_t)z[i].key;
        ri->cnt = (int32_t)u[z[i].idx];
        ri->as = ks[z[i].idx];
        reg_set_coor(ri, qlen, a);
        cal_fuzzy_len(ri, a);
    }
    free(z); free(ks);
    return r;
}

/* [mm2:hit.c:mm_set_parent], no alt contigs, hard_mask_level off */
static void set_parent(float mask_level, int mask_len, int n, reg_t *r, int sub_diff)
{
    int i, j, k, *w;
    uint64_t *cov;
    if (n <= 0) return;
    for (i = 0; i < n; ++i) r[i].id = i;
    cov = (uint64_t *)malloc((size_t)n * 8);
    w = (int *)malloc((size_t)n * sizeof(int));
    w[0] = 0, r[0].parent = 0;
    for (i = 1, k = 1; i < n; ++i) {
        reg_t *ri = &r[i];
        int si = ri->qs, ei = ri->qe, n_cov = 0, uncov_len = 0;
        for (j = 0; j < k; ++j) {
            reg_t *rp = &r[w[j]];
            int sj = rp->qs, ej = rp->qe;
            if (ej <= si || sj >= ei) continue;
            if (sj < si) sj = si;
            if (ej > ei) ej = ei;
            cov[n_cov++] = (uint64_t)sj << 32 | (uint32_t)ej;
        }
        if (n_cov > 0) {
            int jj, x = si;
            qsort(cov, (size_t)n_cov, 8, cmp_u64);
            for (jj = 0; jj < n_cov; ++jj) {
                if ((int)(cov[jj] >> 32) > x) uncov_len += (int)(cov[jj] >> 32) - x;
                x = (int32_t)cov[jj] > x ? (int32_t)cov[jj] : x;
            }
            if (ei > x) uncov_len += ei - x;
            for (j = 0; j < k; ++j) {
                reg_t *rp = &r[w[j]];
                int sj = rp->qs, ej = rp->qe, min, max, ol;
                if (ej <= si || sj >= ei) continue;
                min = ej - sj < ei - si ? ej - sj : ei - si;
                max = ej - sj > ei - si ? ej - sj : ei - si;
                ol = si < sj ? (ei < sj ? 0 : ei < ej ? ei - sj : ej - sj) : (ej < si ? 0 : ej < ei ? ej - si : ei - si);
                if ((float)ol / min - (float)uncov_len / max > mask_level && uncov_len <= mask_len) {
                    int cnt_sub = 0, sci = ri->score;
                    ri->parent = rp->parent;
                    rp->subsc = rp->subsc > sci ? rp->subsc : sci;
                    if (ri->cnt >= rp->cnt) cnt_sub = 1;
                    if (rp->has_p && ri->has_p && (rp->rid != ri->rid || rp->rs != ri->rs || rp->re != ri->re || ol != min)) {
                        sci = ri->dp_max;
                        rp->dp_max2 = rp->dp_max2 > sci ? rp->dp_max2 : sci;
                        if (rp->dp_max - ri->dp_max <= sub_diff) cnt_sub = 1;
                    }
                    if (cnt_sub) ++rp->n_sub;
                    break;
                }
            }
        } else j = k;
        if (j == k) w[k++] = i, ri->parent = i, ri->n_sub = 0;
    }
    free(cov); free(w);
}

/* [mm2:hit.c:mm_set_mapq2], not short-read, not spliced */
static void set_mapq(int n_regs, reg_t *regs, int min_chain_sc, int match_sc, int rep_len)
{
    static const float q_coef = 40.0f;
    int64_t sum_sc = 0;
    float uniq_ratio;
    int i;
    if (n_regs == 0) return;
    for (i = 0; i < n_regs; ++i)
        if (regs[i].parent == regs[i].id) sum_sc += regs[i].score;
    uniq_ratio = (float)sum_sc / (float)(sum_sc + rep_len);
    for (i = 0; i < n_regs; ++i) {
        reg_t *r = &regs[i];
        if (r->parent == r->id) {
            int mapq, subsc;
            float pen_s1 = (r->score > 100 ? 1.0f : 0.01f * (float)r->score) * uniq_ratio;
            float pen_cm = r->cnt > 10 ? 1.0f : 0.1f * (float)r->cnt;
            pen_cm = pen_s1 < pen_cm ? pen_s1 : pen_cm;
            subsc = r->subsc > min_chain_sc ? r->subsc : min_chain_sc;
            if (r->has_p && r->dp_max2 > 0 && r->dp_max > 0) {
                float identity = (float)r->mlen / (float)r->blen;
                float x = (float)r->dp_max2 * (float)subsc;
                float t, lg;
                int mapq_alt;
                x = x / (float)r->dp_max; x = x / (float)r->score0;
                lg = kbo_logf((float)r->dp_max / (float)match_sc);
                t = identity * pen_cm; t = t * q_coef; t = t * (1.0f - x * x); t = t * lg;
                mapq = (int)t;
                t = 6.02f * identity; t = t * identity; t = t * (float)(r->dp_max - r->dp_max2); t = t / (float)match_sc; t = t + .499f;
                mapq_alt = (int)t;
                mapq = mapq < mapq_alt ? mapq : mapq_alt;
            } else {
                float x = (float)subsc / (float)r->score0, t;
                if (r->has_p) {
                    float identity = (float)r->mlen / (float)r->blen;
                    t = identity * pen_cm; t = t * q_coef; t = t * (1.0f - x); t = t * kbo_logf((float)r->dp_max / (float)match_sc);
                    mapq = (int)t;
                } else {
                    t = pen_cm * q_coef; t = t * (1.0f - x); t = t * kbo_logf((float)r->score);
                    mapq = (int)t;
                }
            }
            {
                float t = 4.343f * kbo_logf((float)(r->n_sub + 1));
                t = t + .499f;
                mapq -= (int)t;
            }
            mapq = mapq > 0 ? mapq : 0;
            r->mapq = mapq < 60 ? mapq : 60;
            if (r->has_p && r->dp_max > r->dp_max2 && r->mapq == 0) r->mapq = 1;
        } else r->mapq = 0;
    }
}

/* ---------------------------------------------------------------- ksw-like */

#define EZ_EXTZ_ONLY 0x1
#define EZ_RIGHT     0x2
#define EZ_REV_CIGAR 0x4
#define EZ_GLOBAL_NO_ZDROP 0x8   /* first gap-fill pass: plain global alignment */

typedef struct {
    int32_t max, max_q, max_t, score, zdropped;
    int32_t n_cigar, m_cigar;
    uint32_t *cigar;
} ez_t;

static inline void push_cigar(ez_t *ez, uint32_t op, int len) /* [mm2:ksw2.h:ksw_push_cigar] */
{
    if (ez->n_cigar == 0 || op != (ez->cigar[ez->n_cigar - 1] & 0xf)) {
        if (ez->n_cigar == ez->m_cigar) {
            ez->m_cigar = ez->m_cigar ? ez->m_cigar << 1 : 16;
            ez->cigar = (uint32_t *)realloc(ez->cigar, (size_t)ez->m_cigar * 4);
        }
        ez->cigar[ez->n_cigar++] = (uint32_t)len << 4 | op;
    } else ez->cigar[ez->n_cigar - 1] += (uint32_t)len << 4;
}

static inline int32_t gapcost2(const kbo_params_t *P, int l)
{
    int32_t c1 = P->q + P->e * l, c2 = P->q2 + P->e2 * l;
    return c1 < c2 ? c1 : c2;
}

/* Dual-affine banded DP in anti-diagonal order.
 * Restates [mm2:ksw2_extd2_sse.c:ksw_extd2_sse] in absolute-score form:
 *   i = target index, j = query index, r = i + j
 *   E*(i,j) = max(H(i-1,j) - q*, E*(i-1,j)) - e*      (deletion: consumes target)
 *   F*(i,j) = max(H(i,j-1) - q*, F*(i,j-1)) - e*      (insertion: consumes query)
 *   H(i,j)  = max(H(i-1,j-1)+s, E1, F1, E2, F2); tie order diag,E1,F1,E2,F2
 *             (reversed, gap-preferring, with EZ_RIGHT)
 * Band: (r-w+1)>>1 <= i <= (r+w)>>1; outside = -inf (spec v1).
 * Max/z-drop are evaluated once per anti-diagonal on its maximum
 * (lowest i among equals) with [mm2:ksw2.h:ksw_apply_zdrop] using e2.
 * Traceback: [mm2:ksw2.h:ksw_backtrack].
 */
static void extd2(const kbo_params_t *P, int qlen, const uint8_t *qs, int tlen, const uint8_t *ts, int w, int zdrop, int flag, ez_t *ez)
{
    int r, t, n_diag, right = !!(flag & EZ_RIGHT);
    int32_t *H[3], *E1[2], *E2[2], *F1[2], *F2[2];
    int32_t *mem;
    int *off, *off_end;
    int64_t *ppos, tb_n = 0, tb_m;
    uint8_t *p;
    int last_st = 0, last_en = -1, last2_st = 0, last2_en = -1;
    int8_t mat[25];
    const int q = P->q, e = P->e, q2 = P->q2, e2 = P->e2;
    int end_r = -1;

    ez->max = 0, ez->max_q = ez->max_t = -1, ez->score = NEG_INF, ez->zdropped = 0, ez->n_cigar = 0;
    if (qlen <= 0 || tlen <= 0) return;
    if ((int64_t)qlen * tlen > P->max_sw_cells) { ez->zdropped = 1; return; } /* [mm2:align.c:mm_align_pair] max_sw_mat */
    {
        int i, j;
        for (i = 0; i < 4; ++i) { for (j = 0; j < 4; ++j) mat[i * 5 + j] = (int8_t)(i == j ? P->a : -P->b); mat[i * 5 + 4] = (int8_t)-P->sc_ambi; }
        for (j = 0; j < 5; ++j) mat[20 + j] = (int8_t)-P->sc_ambi;
    }
    n_diag = qlen + tlen - 1;
    mem = (int32_t *)malloc((size_t)tlen * 11 * 4);
    H[0] = mem; H[1] = mem + tlen; H[2] = mem + 2 * tlen;
    E1[0] = mem + 3 * tlen; E1[1] = mem + 4 * tlen; E2[0] = mem + 5 * tlen; E2[1] = mem + 6 * tlen;
    F1[0] = mem + 7 * tlen; F1[1] = mem + 8 * tlen; F2[0] = mem + 9 * tlen; F2[1] = mem + 10 * tlen;
    off = (int *)malloc((size_t)n_diag * 2 * sizeof(int)); off_end = off + n_diag;
    ppos = (int64_t *)malloc((size_t)n_diag * 8);
    tb_m = (int64_t)(qlen < tlen ? qlen : tlen) * 64 + 1024;
    p = (uint8_t *)malloc((size_t)tb_m);

    for (r = 0; r < n_diag; ++r) {
        int st = 0, en = tlen - 1, max_t = -1;
        int32_t max_H = INT32_MIN;
        int32_t *Hc = H[r % 3], *H1 = H[(r + 2) % 3], *Hd = H[(r + 1) % 3];
        int32_t *e1c = E1[r & 1], *e1p = E1[(r & 1) ^ 1], *e2c = E2[r & 1], *e2p = E2[(r & 1) ^ 1];
        int32_t *f1c = F1[r & 1], *f1p = F1[(r & 1) ^ 1], *f2c = F2[r & 1], *f2p = F2[(r & 1) ^ 1];
        uint8_t *pr;
        if (st < r - qlen + 1) st = r - qlen + 1;
        if (en > r) en = r;
        if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
        if (en > (r + w) >> 1) en = (r + w) >> 1;
        if (st > en) { ez->zdropped = 1; break; }
        off[r] = st, off_end[r] = en, ppos[r] = tb_n;
        if (tb_n + (en - st + 1) > tb_m) { tb_m = (tb_n + (en - st + 1)) * 2; p = (uint8_t *)realloc(p, (size_t)tb_m); }
        pr = p + tb_n - st;
        tb_n += en - st + 1;
        for (t = st; t <= en; ++t) {
            int j = r - t; /* query index; i = t */
            int32_t h_up, h_left, h_diag, a1, a2, b1, b2, z, hq, hq2;
            uint8_t d;
            /* (i-1, j): previous diagonal, index t-1 */
            if (t == 0) { h_up = -gapcost2(P, j + 1); a1 = a2 = NEG_INF; }
            else if (t - 1 >= last_st && t - 1 <= last_en) { h_up = H1[t - 1]; a1 = e1p[t - 1]; a2 = e2p[t - 1]; }
            else { h_up = NEG_INF; a1 = a2 = NEG_INF; }
            /* (i, j-1): previous diagonal, index t */
            if (j == 0) { h_left = -gapcost2(P, t + 1); b1 = b2 = NEG_INF; }
            else if (t >= last_st && t <= last_en) { h_left = H1[t]; b1 = f1p[t]; b2 = f2p[t]; }
            else { h_left = NEG_INF; b1 = b2 = NEG_INF; }
            /* (i-1, j-1): diagonal r-2, index t-1 */
            if (t == 0 && j == 0) h_diag = 0;
            else if (t == 0) h_diag = -gapcost2(P, j);
            else if (j == 0) h_diag = -gapcost2(P, t);
            else if (t - 1 >= last2_st && t - 1 <= last2_en) h_diag = Hd[t - 1];
            else h_diag = NEG_INF;
            a1 = (h_up - q > a1 ? h_up - q : a1) - e;
            a2 = (h_up - q2 > a2 ? h_up - q2 : a2) - e2;
            b1 = (h_left - q > b1 ? h_left - q : b1) - e;
            b2 = (h_left - q2 > b2 ? h_left - q2 : b2) - e2;
            z = h_diag + mat[ts[t] * 5 + qs[j]];
            d = 0;
            if (!right) {
                if (a1 > z) d = 1, z = a1;
                if (b1 > z) d = 2, z = b1;
                if (a2 > z) d = 3, z = a2;
                if (b2 > z) d = 4, z = b2;
                hq = z - q, hq2 = z - q2;
                if (a1 > hq) d |= 0x08;
                if (b1 > hq) d |= 0x10;
                if (a2 > hq2) d |= 0x20;
                if (b2 > hq2) d |= 0x40;
            } else {
                if (a1 >= z) d = 1, z = a1;
                if (b1 >= z) d = 2, z = b1;
                if (a2 >= z) d = 3, z = a2;
                if (b2 >= z) d = 4, z = b2;
                hq = z - q, hq2 = z - q2;
                if (a1 >= hq) d |= 0x08;
                if (b1 >= hq) d |= 0x10;
                if (a2 >= hq2) d |= 0x20;
                if (b2 >= hq2) d |= 0x40;
            }
            Hc[t] = z; e1c[t] = a1; e2c[t] = a2; f1c[t] = b1; f2c[t] = b2;
            pr[t] = d;
            if (z > max_H) max_H = z, max_t = t;
        }
        if (!(flag & EZ_GLOBAL_NO_ZDROP)) { /* [mm2:ksw2.h:ksw_apply_zdrop], is_rot */
            if (max_H > ez->max) {
                ez->max = max_H, ez->max_t = max_t, ez->max_q = r - max_t;
            } else if (max_t >= ez->max_t && r - max_t >= ez->max_q) {
                int tl = max_t - ez->max_t, ql = (r - max_t) - ez->max_q, l;
                l = tl > ql ? tl - ql : ql - tl;
                if (zdrop >= 0 && ez->max - max_H > zdrop + l * e2) { ez->zdropped = 1; end_r = r; break; }
            }
        }
        if (r == n_diag - 1 && en == tlen - 1) ez->score = Hc[tlen - 1];
        last2_st = last_st, last2_en = last_en;
        last_st = st, last_en = en;
        end_r = r;
    }
    (void)end_r;
    { /* [mm2:ksw2.h:ksw_backtrack], is_rot=1 */
        int i0 = -1, j0 = -1;
        if (!ez->zdropped && !(flag & EZ_EXTZ_ONLY)) i0 = tlen - 1, j0 = qlen - 1;
        else if (ez->max_t >= 0 && ez->max_q >= 0) i0 = ez->max_t, j0 = ez->max_q;
        if (i0 >= 0 && j0 >= 0) {
            int i = i0, j = j0, state = 0;
            while (i >= 0 && j >= 0) {
                int force_state = -1;
                uint32_t tmp;
                r = i + j;
                if (i < off[r]) force_state = 2;
                if (i > off_end[r]) force_state = 1;
                tmp = force_state < 0 ? p[ppos[r] + i - off[r]] : 0;
                if (state == 0) state = tmp & 7;
                else if (!(tmp >> (state + 2) & 1)) state = 0;
                if (state == 0) state = tmp & 7;
                if (force_state >= 0) state = force_state;
                if (state == 0) push_cigar(ez, 0, 1), --i, --j;
                else if (state == 1 || state == 3) push_cigar(ez, 2, 1), --i;
                else push_cigar(ez, 1, 1), --j;
            }
            if (i >= 0) push_cigar(ez, 2, i + 1);
            if (j >= 0) push_cigar(ez, 1, j + 1);
            if (!(flag & EZ_REV_CIGAR)) {
                int a;
                for (a = 0; a < ez->n_cigar >> 1; ++a) { uint32_t tmp = ez->cigar[a]; ez->cigar[a] = ez->cigar[ez->n_cigar - 1 - a]; ez->cigar[ez->n_cigar - 1 - a] = tmp; }
            }
        }
    }
    free(mem); free(off); free(ppos); free(p);
}

/* ------------------------------------------------------------------- align */

typedef struct {
    const kbo_params_t *P;
    const kbo_db_t *db;
    const asm_idx_t *ai;
} ctx_t;

static void append_cigar(reg_t *r, int n_cigar, const uint32_t *cigar) /* [mm2:align.c:mm_append_cigar] */
{
    if (n_cigar == 0) return;
    if (r->n_cigar + n_cigar > r->m_cigar) {
        r->m_cigar = (r->n_cigar + n_cigar) * 2;
        r->cigar = (uint32_t *)realloc(r->cigar, (size_t)r->m_cigar * 4);
    }
    r->has_p = 1;
    if (r->n_cigar > 0 && (r->cigar[r->n_cigar - 1] & 0xf) == (cigar[0] & 0xf)) {
        r->cigar[r->n_cigar - 1] += (cigar[0] >> 4) << 4;
        if (n_cigar > 1) memcpy(r->cigar + r->n_cigar, cigar + 1, (size_t)(n_cigar - 1) * 4);
        r->n_cigar += n_cigar - 1;
    } else {
        memcpy(r->cigar + r->n_cigar, cigar, (size_t)n_cigar * 4);
        r->n_cigar += n_cigar;
    }
}

/* [mm2:align.c:mm_fix_bad_ends] */
static void fix_bad_ends(const reg_t *r, const mm128_t *a, int bw, int min_match, int32_t *as, int32_t *cnt)
{
    int32_t i, l, m;
    *as = r->as, *cnt = r->cnt;
    if (r->cnt < 3) return;
    m = l = (int32_t)(a[r->as].y >> 32 & 0xff);
    for (i = r->as + 1; i < r->as + r->cnt - 1; ++i) {
        int32_t lq, lr, min, max;
        int32_t q_span = (int32_t)(a[i].y >> 32 & 0xff);
        if (a[i].y & SEED_LONG_JOIN) break;
        lr = (int32_t)a[i].x - (int32_t)a[i - 1].x;
        lq = (int32_t)a[i].y - (int32_t)a[i - 1].y;
        min = lr < lq ? lr : lq;
        max = lr > lq ? lr : lq;
        if (max - min > l >> 1) *as = i;
        l += min;
        m += min < q_span ? min : q_span;
        if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r->mlen >> 1) break;
    }
    *cnt = r->as + r->cnt - *as;
    m = l = (int32_t)(a[r->as + r->cnt - 1].y >> 32 & 0xff);
    for (i = r->as + r->cnt - 2; i > *as; --i) {
        int32_t lq, lr, min, max;
        int32_t q_span = (int32_t)(a[i + 1].y >> 32 & 0xff);
        if (a[i + 1].y & SEED_LONG_JOIN) break;
        lr = (int32_t)a[i + 1].x - (int32_t)a[i].x;
        lq = (int32_t)a[i + 1].y - (int32_t)a[i].y;
        min = lr < lq ? lr : lq;
        max = lr > lq ? lr : lq;
        if (max - min > l >> 1) *cnt = i + 1 - *as;
        l += min;
        m += min < q_span ? min : q_span;
        if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r->mlen >> 1) break;
    }
}

static inline int anchor_gap(const mm128_t *a, int i) /* query advance minus target advance between anchors i-1 and i */
{
    return ((int32_t)a[i].y - (int32_t)a[i - 1].y) - ((int32_t)a[i].x - (int32_t)a[i - 1].x);
}

/* [mm2:align.c:collect_long_gaps] */
static int *collect_long_gaps(int as1, int cnt1, const mm128_t *a, int min_gap, int *n_)
{
    int i, n, *K;
    *n_ = 0;
    for (i = 1, n = 0; i < cnt1; ++i) { int gap = anchor_gap(a, as1 + i); if (gap < -min_gap || gap > min_gap) ++n; }
    if (n <= 1) return 0;
    K = (int *)malloc((size_t)n * sizeof(int));
    for (i = 1, n = 0; i < cnt1; ++i) { int gap = anchor_gap(a, as1 + i); if (gap < -min_gap || gap > min_gap) K[n++] = i; }
    *n_ = n;
    return K;
}

/* [mm2:align.c:mm_filter_bad_seeds] */
static void filter_bad_seeds(int as1, int cnt1, mm128_t *a, int min_gap, int diff_thres, int max_ext_len, int max_ext_cnt)
{
    int max_st, max_en, n, i, k, max, *K;
    K = collect_long_gaps(as1, cnt1, a, min_gap, &n);
    if (K == 0) return;
    max = 0, max_st = max_en = -1;
    for (k = 0;; ++k) {
        int gap, l, n_ins = 0, n_del = 0, qs, rs, max_diff = 0, max_diff_l = -1;
        if (k == n || k >= max_en) {
            if (max_en > 0)
                for (i = K[max_st]; i < K[max_en]; ++i) a[as1 + i].y |= SEED_IGNORE;
            max = 0, max_st = max_en = -1;
            if (k == n) break;
        }
        i = K[k];
        gap = anchor_gap(a, as1 + i);
        if (gap > 0) n_ins += gap; else n_del += -gap;
        qs = (int32_t)a[as1 + i - 1].y;
        rs = (int32_t)a[as1 + i - 1].x;
        for (l = k + 1; l < n && l <= k + max_ext_cnt; ++l) {
            int j = K[l], diff;
            if ((int32_t)a[as1 + j].y - qs > max_ext_len || (int32_t)a[as1 + j].x - rs > max_ext_len) break;
            gap = anchor_gap(a, as1 + j);
            if (gap > 0) n_ins += gap; else n_del += -gap;
            diff = n_ins + n_del - abs(n_ins - n_del);
            if (max_diff < diff) max_diff = diff, max_diff_l = l;
        }
        if (max_diff > diff_thres && max_diff > max) max = max_diff, max_st = k, max_en = max_diff_l;
    }
    free(K);
}

/* [mm2:align.c:mm_filter_bad_seeds_alt] */
static void filter_bad_seeds_alt(int as1, int cnt1, mm128_t *a, int min_gap, int max_ext)
{
    int n, k, *K;
    K = collect_long_gaps(as1, cnt1, a, min_gap, &n);
    if (K == 0) return;
    for (k = 0; k < n;) {
        int i = K[k], l;
        int gap1 = anchor_gap(a, as1 + i);
        int re1 = (int32_t)a[as1 + i].x;
        int qe1 = (int32_t)a[as1 + i].y;
        gap1 = gap1 > 0 ? gap1 : -gap1;
        for (l = k + 1; l < n; ++l) {
            int j = K[l], gap2, q_span_pre, rs2, qs2, m;
            if ((int32_t)a[as1 + j].y - qe1 > max_ext || (int32_t)a[as1 + j].x - re1 > max_ext) break;
            gap2 = anchor_gap(a, as1 + j);
            q_span_pre = (int)(a[as1 + j - 1].y >> 32 & 0xff);
            rs2 = (int32_t)a[as1 + j - 1].x + q_span_pre;
            qs2 = (int32_t)a[as1 + j - 1].y + q_span_pre;
            m = rs2 - re1 < qs2 - qe1 ? rs2 - re1 : qs2 - qe1;
            gap2 = gap2 > 0 ? gap2 : -gap2;
            if (m > gap1 + gap2) break;
            re1 = (int32_t)a[as1 + j].x;
            qe1 = (int32_t)a[as1 + j].y;
            gap1 = gap2;
        }
        if (l > k + 1) {
            int j, end = K[l - 1];
            for (j = K[k]; j < end; ++j) a[as1 + j].y |= SEED_IGNORE;
            a[as1 + end].y |= SEED_LONG_JOIN;
        }
        k = l;
    }
    free(K);
}

/* [mm2:align.c:update_max_zdrop] */
static inline void update_max_zdrop(int32_t score, int i, int j, int32_t *max, int *max_i, int *max_j, int e, int *max_zdrop)
{
    if (score < *max) {
        int li = i - *max_i, lj = j - *max_j;
        int diff = li > lj ? li - lj : lj - li;
        int z = *max - score - diff * e;
        if (z > *max_zdrop) *max_zdrop = z;
    } else *max = score, *max_i = i, *max_j = j;
}

/* [mm2:align.c:mm_test_zdrop] without the inversion test */
static int test_zdrop(const kbo_params_t *P, const uint8_t *qseq, const uint8_t *tseq, int n_cigar, const uint32_t *cigar)
{
    int k;
    int32_t score = 0, max = INT32_MIN, max_i = -1, max_j = -1, i = 0, j = 0, max_zdrop = 0;
    for (k = 0; k < n_cigar; ++k) {
        uint32_t l, op = cigar[k] & 0xf, len = cigar[k] >> 4;
        if (op == 0) {
            for (l = 0; l < len; ++l) {
                int ct = tseq[i + l], cq = qseq[j + l];
                score += (ct > 3 || cq > 3) ? -P->sc_ambi : (ct == cq ? P->a : -P->b);
                update_max_zdrop(score, i + (int)l, j + (int)l, &max, &max_i, &max_j, P->e, &max_zdrop);
            }
            i += len, j += len;
        } else {
            score -= P->q + P->e * (int)len;
            if (op == 1) j += len; else i += len;
            update_max_zdrop(score, i, j, &max, &max_i, &max_j, P->e, &max_zdrop);
        }
    }
    return max_zdrop > P->zdrop ? 1 : 0;
}

/* [mm2:align.c:mm_fix_cigar] */
static void fix_cigar(reg_t *r, const uint8_t *qseq, const uint8_t *tseq, int *qshift, int *tshift)
{
    int32_t toff = 0, qoff = 0, to_shrink = 0;
    int k;
    *qshift = *tshift = 0;
    if (r->n_cigar <= 1) return;
    for (k = 0; k < r->n_cigar; ++k) {
        uint32_t op = r->cigar[k] & 0xf, len = r->cigar[k] >> 4;
        if (len == 0) to_shrink = 1;
        if (op == 0) {
            toff += len, qoff += len;
        } else if (op == 1 || op == 2) {
            if (k > 0 && k < r->n_cigar - 1 && (r->cigar[k - 1] & 0xf) == 0 && (r->cigar[k + 1] & 0xf) == 0) {
                int l, prev_len = (int)(r->cigar[k - 1] >> 4);
                if (op == 1) {
                    for (l = 0; l < prev_len; ++l)
                        if (qseq[qoff - 1 - l] != qseq[qoff + (int)len - 1 - l]) break;
                } else {
                    for (l = 0; l < prev_len; ++l)
                        if (tseq[toff - 1 - l] != tseq[toff + (int)len - 1 - l]) break;
                }
                if (l > 0) r->cigar[k - 1] -= (uint32_t)l << 4, r->cigar[k + 1] += (uint32_t)l << 4, qoff -= l, toff -= l;
                if (l == prev_len) to_shrink = 1;
            }
            if (op == 2) toff += len; else qoff += len;
        }
    }
    for (k = 0; k < r->n_cigar - 2; ++k) {
        if ((r->cigar[k] & 0xf) > 0 && (r->cigar[k] & 0xf) + (r->cigar[k + 1] & 0xf) == 3) {
            int l;
            uint32_t s[3] = {0, 0, 0};
            for (l = k; l < r->n_cigar; ++l) {
                uint32_t op = r->cigar[l] & 0xf;
                if (op == 1 || op == 2) s[op] += r->cigar[l] >> 4;
                else break;
            }
            if (s[1] > 0 && s[2] > 0 && l - k > 2) {
                r->cigar[k] = s[1] << 4 | 1;
                r->cigar[k + 1] = s[2] << 4 | 2;
                for (k += 2; k < l; ++k) r->cigar[k] &= 0xf;
                to_shrink = 1;
            }
            k = l;
        }
    }
    if (to_shrink) {
        int l = 0;
        for (k = 0; k < r->n_cigar; ++k)
            if (r->cigar[k] >> 4 != 0) r->cigar[l++] = r->cigar[k];
        r->n_cigar = l;
        for (k = l = 0; k < r->n_cigar; ++k)
            if (k == r->n_cigar - 1 || (r->cigar[k] & 0xf) != (r->cigar[k + 1] & 0xf)) r->cigar[l++] = r->cigar[k];
            else r->cigar[k + 1] += r->cigar[k] >> 4 << 4;
        r->n_cigar = l;
    }
    if (r->n_cigar > 0 && ((r->cigar[0] & 0xf) == 1 || (r->cigar[0] & 0xf) == 2)) {
        int32_t l = (int32_t)(r->cigar[0] >> 4);
        if ((r->cigar[0] & 0xf) == 1) {
            if (r->rev) r->qe -= l; else r->qs += l;
            *qshift = l;
        } else r->rs += l, *tshift = l;
        --r->n_cigar;
        memmove(r->cigar, r->cigar + 1, (size_t)r->n_cigar * 4);
    }
}

/* [mm2:align.c:mm_update_extra], log_gap = 1 */
static void update_extra(const kbo_params_t *P, reg_t *r, const uint8_t *qseq, const uint8_t *tseq)
{
    int k;
    uint32_t l;
    int32_t qshift, tshift, toff = 0, qoff = 0;
    double s = 0.0, max = 0.0;
    if (!r->has_p) return;
    fix_cigar(r, qseq, tseq, &qshift, &tshift);
    qseq += qshift, tseq += tshift;
    r->blen = r->mlen = 0; r->n_ambi = 0;
    for (k = 0; k < r->n_cigar; ++k) {
        uint32_t op = r->cigar[k] & 0xf, len = r->cigar[k] >> 4;
        if (op == 0) {
            int n_ambi = 0, n_diff = 0;
            for (l = 0; l < len; ++l) {
                int cq = qseq[qoff + l], ct = tseq[toff + l];
                if (ct > 3 || cq > 3) ++n_ambi, s += -P->sc_ambi;
                else if (ct != cq) ++n_diff, s += -P->b;
                else s += P->a;
                if (s < 0) s = 0;
                else max = max > s ? max : s;
            }
            r->blen += len - n_ambi, r->mlen += len - (n_ambi + n_diff), r->n_ambi += n_ambi;
            toff += len, qoff += len;
        } else if (op == 1) {
            int n_ambi = 0;
            double pen;
            for (l = 0; l < len; ++l) if (qseq[qoff + l] > 3) ++n_ambi;
            r->blen += len - n_ambi, r->n_ambi += n_ambi;
            pen = (double)P->e * (double)kbo_log2_fast((float)(1.0 + len));
            pen = (double)P->q + pen;
            s -= pen;
            if (s < 0) s = 0;
            qoff += len;
        } else if (op == 2) {
            int n_ambi = 0;
            double pen;
            for (l = 0; l < len; ++l) if (tseq[toff + l] > 3) ++n_ambi;
            r->blen += len - n_ambi, r->n_ambi += n_ambi;
            pen = (double)P->e * (double)kbo_log2_fast((float)(1.0 + len));
            pen = (double)P->q + pen;
            s -= pen;
            if (s < 0) s = 0;
            toff += len;
        }
    }
    r->dp_max = (int32_t)(max + .499);
}

/* [mm2:hit.c:mm_split_reg] */
static void split_reg(reg_t *r, reg_t *r2, int n, int qlen, const mm128_t *a)
{
    if (n <= 0 || n >= r->cnt) return;
    *r2 = *r;
    r2->id = -1;
    r2->has_p = 0; r2->cigar = 0; r2->n_cigar = r2->m_cigar = 0; r2->dp_score = r2->dp_max = r2->dp_max2 = r2->n_ambi = 0;
    r2->cnt = r->cnt - n;
    r2->score = (int32_t)((float)r->score * ((float)r2->cnt / (float)r->cnt) + .499f);
    r2->as = r->as + n;
    if (r->parent == r->id) r2->parent = PARENT_TMP_PRI;
    reg_set_coor(r2, qlen, a);
    r->cnt -= r2->cnt;
    r->score -= r2->score;
    reg_set_coor(r, qlen, a);
}

/* [mm2:align.c:mm_align1], long-read path (not sr, not splice) */
static void align1(const ctx_t *C, int g, reg_t *r, reg_t *r2, int n_a, mm128_t *a, ez_t *ez)
{
    const kbo_params_t *P = C->P;
    const int qlen = C->db->len[g];
    int32_t rid = (int32_t)(a[r->as].x << 1 >> 33), rev = (int32_t)(a[r->as].x >> 63), as1, cnt1;
    const uint8_t *tfull = C->ai->ctg[rid];
    const int32_t tlen_full = C->ai->ctg_len[rid];
    const uint8_t *qseq0 = rev ? C->db->rev[g] : C->db->fwd[g];
    uint8_t *qbuf, *tbuf;
    int32_t i, l, bw, bw_long, dropped = 0, rs0, re0, qs0, qe0;
    int32_t rs, re, qs, qe;
    int32_t rs1, qs1, re1, qe1;
    const int hk = P->k >> 1;

    r2->cnt = 0;
    if (r->cnt == 0) return;
    bw = P->ext_bw;
    bw_long = (int)(20000 * 1.5 + 1.);
    if (bw_long < bw) bw_long = bw;

    fix_bad_ends(r, a, P->bw, P->min_chain_score * 2, &as1, &cnt1);
    filter_bad_seeds(as1, cnt1, a, 10, 40, P->max_gap >> 1, 10);
    filter_bad_seeds_alt(as1, cnt1, a, 30, P->max_gap >> 1);

    /* [mm2:align.c:mm_adjust_minier], non-HPC */
    rs = (int32_t)a[as1].x - hk; qs = (int32_t)a[as1].y - hk;
    re = (int32_t)a[as1 + cnt1 - 1].x - hk; qe = (int32_t)a[as1 + cnt1 - 1].y - hk;

    /* compute rs0 and qs0 */
    rs0 = (int32_t)a[r->as].x + 1 - (int32_t)(a[r->as].y >> 32 & 0xff);
    qs0 = (int32_t)a[r->as].y + 1 - (int32_t)(a[r->as].y >> 32 & 0xff);
    if (rs0 < 0) rs0 = 0;
    rs1 = qs1 = 0;
    for (i = r->as - 1, l = 0; i >= 0 && a[i].x >> 32 == a[r->as].x >> 32; --i) {
        int32_t x = (int32_t)a[i].x + 1 - (int32_t)(a[i].y >> 32 & 0xff);
        int32_t y = (int32_t)a[i].y + 1 - (int32_t)(a[i].y >> 32 & 0xff);
        if (x < rs0 && y < qs0) {
            if (++l > P->min_cnt) {
                l = rs0 - x > qs0 - y ? rs0 - x : qs0 - y;
                rs1 = rs0 - l, qs1 = qs0 - l;
                if (rs1 < 0) rs1 = 0;
                break;
            }
        }
    }
    if (qs > 0 && rs > 0) {
        l = qs < P->max_gap ? qs : P->max_gap;
        qs1 = qs1 > qs - l ? qs1 : qs - l;
        qs0 = qs0 < qs1 ? qs0 : qs1;
        l += l * P->a > P->q ? (l * P->a - P->q) / P->e : 0;
        l = l < P->max_gap ? l : P->max_gap;
        l = l < rs ? l : rs;
        rs1 = rs1 > rs - l ? rs1 : rs - l;
        rs0 = rs0 < rs1 ? rs0 : rs1;
        rs0 = rs0 < rs ? rs0 : rs;
    } else rs0 = rs, qs0 = qs;
    /* compute re0 and qe0 */
    re0 = (int32_t)a[r->as + r->cnt - 1].x + 1;
    qe0 = (int32_t)a[r->as + r->cnt - 1].y + 1;
    re1 = tlen_full, qe1 = qlen;
    for (i = r->as + r->cnt, l = 0; i < n_a && a[i].x >> 32 == a[r->as].x >> 32; ++i) {
        int32_t x = (int32_t)a[i].x + 1;
        int32_t y = (int32_t)a[i].y + 1;
        if (x > re0 && y > qe0) {
            if (++l > P->min_cnt) {
                l = x - re0 > y - qe0 ? x - re0 : y - qe0;
                re1 = re0 + l, qe1 = qe0 + l;
                break;
            }
        }
    }
    if (qe < qlen && re < tlen_full) {
        l = qlen - qe < P->max_gap ? qlen - qe : P->max_gap;
        qe1 = qe1 < qe + l ? qe1 : qe + l;
        qe0 = qe0 > qe1 ? qe0 : qe1;
        l += l * P->a > P->q ? (l * P->a - P->q) / P->e : 0;
        l = l < P->max_gap ? l : P->max_gap;
        l = l < tlen_full - re ? l : tlen_full - re;
        re1 = re1 < re + l ? re1 : re + l;
        re0 = re0 > re1 ? re0 : re1;
    } else re0 = re, qe0 = qe;

    qbuf = (uint8_t *)malloc((size_t)qlen + 1);
    tbuf = (uint8_t *)malloc((size_t)(re0 - rs0 > 0 ? re0 - rs0 : 1) + 1);

    if (qs > 0 && rs > 0) { /* left extension */
        int ql = qs - qs0, tl = rs - rs0, x;
        for (x = 0; x < ql; ++x) qbuf[x] = qseq0[qs - 1 - x];
        for (x = 0; x < tl; ++x) tbuf[x] = tfull[rs - 1 - x];
        extd2(P, ql, qbuf, tl, tbuf, bw, P->zdrop, EZ_EXTZ_ONLY | EZ_RIGHT | EZ_REV_CIGAR, ez);
        if (ez->n_cigar > 0) {
            append_cigar(r, ez->n_cigar, ez->cigar);
            r->dp_score += ez->max;
        }
        rs1 = rs - (ez->max_t + 1);
        qs1 = qs - (ez->max_q + 1);
    } else rs1 = rs, qs1 = qs;
    re1 = rs, qe1 = qs;

    for (i = 1; i < cnt1; ++i) { /* gap filling */
        if ((a[as1 + i].y & (SEED_IGNORE | SEED_TANDEM)) && i != cnt1 - 1) continue;
        re = (int32_t)a[as1 + i].x - hk; qe = (int32_t)a[as1 + i].y - hk;
        re1 = re, qe1 = qe;
        if (i == cnt1 - 1 || (a[as1 + i].y & SEED_LONG_JOIN) || (qe - qs >= P->min_ksw_len && re - rs >= P->min_ksw_len)) {
            int j, bw1 = bw_long, zdrop_code;
            const uint8_t *tseq = tfull + rs, *qseq = qseq0 + qs;
            if (a[as1 + i].y & SEED_LONG_JOIN) bw1 = qe - qs > re - rs ? qe - qs : re - rs;
            extd2(P, qe - qs, qseq, re - rs, tseq, bw1, -1, EZ_GLOBAL_NO_ZDROP, ez);
            if ((zdrop_code = (ez->zdropped ? 1 : test_zdrop(P, qseq, tseq, ez->n_cigar, ez->cigar))) != 0)
                extd2(P, qe - qs, qseq, re - rs, tseq, bw1, P->zdrop, 0, ez);
            if (ez->n_cigar > 0) append_cigar(r, ez->n_cigar, ez->cigar);
            if (ez->zdropped) {
                r->has_p = 1;
                for (j = i - 1; j >= 0; --j)
                    if ((int32_t)a[as1 + j].x <= rs + ez->max_t) break;
                dropped = 1;
                if (j < 0) j = 0;
                r->dp_score += ez->max;
                re1 = rs + (ez->max_t + 1);
                qe1 = qs + (ez->max_q + 1);
                if (cnt1 - (j + 1) >= P->min_cnt) split_reg(r, r2, as1 + j + 1 - r->as, qlen, a);
                break;
            } else r->dp_score += ez->score;
            rs = re, qs = qe;
        }
    }

    if (!dropped && qe < qe0 && re < re0) { /* right extension */
        extd2(P, qe0 - qe, qseq0 + qe, re0 - re, tfull + re, bw, P->zdrop, EZ_EXTZ_ONLY, ez);
        if (ez->n_cigar > 0) {
            append_cigar(r, ez->n_cigar, ez->cigar);
            r->dp_score += ez->max;
        }
        re1 = re + (ez->max_t + 1);
        qe1 = qe + (ez->max_q + 1);
    }

    r->rs = rs1, r->re = re1;
    if (rev) r->qs = qlen - qe1, r->qe = qlen - qs1;
    else r->qs = qs1, r->qe = qe1;

    if (r->has_p) update_extra(P, r, qseq0 + qs1, tfull + rs1);
    free(qbuf); free(tbuf);
}

/* ----------------------------------------------------------------- driver */

typedef struct {
    kbo_hit_t *h; int32_t n, m;
    uint32_t *cg; int32_t ncg, mcg;
    kbo_anchor_t *an; int64_t nan, man;
    kbo_chain_t *ch; int32_t nch, mch;
} out_t;

static void map_gene(const ctx_t *C, int g, out_t *O, int keep)
{
    const kbo_params_t *P = C->P;
    const kbo_db_t *db = C->db;
    const asm_idx_t *ai = C->ai;
    const int qlen = db->len[g], n_mv = db->n_mv[g], mid_occ = ai->mid_occ;
    const mm128_t *mv = db->mv[g];
    mm128_t *a, *b;
    int64_t n_a = 0, m_a = 0;
    int i, rep_len = 0, rep_st = 0, rep_en = 0, n_u = 0, n_regs, n_a2;
    uint64_t *u = 0;
    reg_t *regs;
    ez_t ez;
    float chn_pen_gap = (float)(P->chain_gap_scale * 0.01 * P->k), chn_pen_skip = 0.0f;
    int qflt = (n_mv > mid_occ && P->q_occ_frac > 0.0f && mid_occ > 0);

    a = 0;
    /* [mm2:seed.c:mm_seed_mz_flt] + [mm2:seed.c:mm_collect_matches] + [mm2:map.c:collect_seed_hits] */
    for (i = 0; i < n_mv; ++i) {
        int64_t lo, hi, kk;
        uint32_t q_pos = (uint32_t)mv[i].y, q_span = (uint32_t)(mv[i].x & 0xff);
        if (qflt && db->qocc[g][i] > mid_occ && (float)db->qocc[g][i] > (float)n_mv * P->q_occ_frac) continue;
        idx_get(ai, mv[i].x >> 8, &lo, &hi);
        if (hi == lo) continue;
        if (hi - lo > mid_occ) {
            int en = (int)(q_pos >> 1) + 1, st = en - (int)q_span;
            if (st > rep_en) { rep_len += rep_en - rep_st; rep_st = st, rep_en = en; }
            else rep_en = en;
            continue;
        }
        if (n_a + (hi - lo) > m_a) { m_a = (n_a + (hi - lo)) * 2; a = (mm128_t *)realloc(a, (size_t)m_a * sizeof(mm128_t)); }
        for (kk = lo; kk < hi; ++kk) {
            uint64_t ry = ai->mz[kk].y;
            int32_t rpos = (int32_t)((uint32_t)ry >> 1);
            mm128_t *p = &a[n_a++];
            if ((ry & 1) == (q_pos & 1)) {
                p->x = (ry & 0xffffffff00000000ULL) | (uint32_t)rpos;
                p->y = (uint64_t)q_span << 32 | q_pos >> 1;
            } else {
                p->x = 1ULL << 63 | (ry & 0xffffffff00000000ULL) | (uint32_t)rpos;
                p->y = (uint64_t)q_span << 32 | (uint32_t)(qlen - ((int)(q_pos >> 1) + 1 - (int)q_span) - 1);
            }
            if (db->tandem[g][i]) p->y |= SEED_TANDEM;
        }
    }
    rep_len += rep_en - rep_st;
    if (n_a == 0) { free(a); return; }
    { /* [mm2:map.c:collect_seed_hits] radix_sort_128x; equal x ordered by query position */
        typedef struct { uint64_t x; uint64_t yq; mm128_t v; } s_t;
        int64_t k;
        s_t *s = (s_t *)malloc((size_t)n_a * sizeof(s_t));
        for (k = 0; k < n_a; ++k) s[k].x = a[k].x, s[k].yq = (uint32_t)a[k].y, s[k].v = a[k];
        qsort(s, (size_t)n_a, sizeof(s_t), cmp_128xy);
        for (k = 0; k < n_a; ++k) a[k] = s[k].v;
        free(s);
    }
    if (keep) {
        int64_t k;
        if (O->nan + n_a > O->man) { O->man = (O->nan + n_a) * 2; O->an = (kbo_anchor_t *)realloc(O->an, (size_t)O->man * sizeof(kbo_anchor_t)); }
        for (k = 0; k < n_a; ++k) {
            kbo_anchor_t *q = &O->an[O->nan++];
            q->gene = g; q->rev = (int32_t)(a[k].x >> 63); q->rid = (int32_t)(a[k].x << 1 >> 33); q->tpos = (int32_t)a[k].x;
            q->qpos = (int32_t)a[k].y; q->flags = (a[k].y & SEED_TANDEM) ? 1 : 0;
        }
    }
    b = lchain_dp(P, chn_pen_gap, chn_pen_skip, n_a, a, &n_u, &u);
    free(a);
    if (n_u == 0 || b == 0) { free(u); return; }
    regs = gen_regs(db->qhash[g], qlen, n_u, u, b);
    n_regs = n_u;
    for (i = 0, n_a2 = 0; i < n_u; ++i) n_a2 += (int32_t)u[i];
    free(u);
    if (keep) {
        if (O->nch + n_regs > O->mch) { O->mch = (O->nch + n_regs) * 2; O->ch = (kbo_chain_t *)realloc(O->ch, (size_t)O->mch * sizeof(kbo_chain_t)); }
        for (i = 0; i < n_regs; ++i) {
            kbo_chain_t *c = &O->ch[O->nch++];
            c->gene = g; c->score = regs[i].score; c->cnt = regs[i].cnt; c->rev = regs[i].rev; c->rid = regs[i].rid;
            c->rs = regs[i].rs; c->re = regs[i].re; c->qs = regs[i].qs; c->qe = regs[i].qe;
        }
    }
    /* [mm2:map.c:chain_post]: pri_ratio == 0 -> mm_select_sub is a no-op (ref:serotyping/core.py:151) */
    set_parent(P->mask_level, P->mask_len, n_regs, regs, P->a * 2 + P->b);

    /* [mm2:align.c:mm_align_skeleton]; chains are all kept so mm_squeeze_a is the identity */
    memset(&ez, 0, sizeof(ez));
    for (i = 0; i < n_regs; ++i) {
        reg_t r2;
        memset(&r2, 0, sizeof(r2));
        align1(C, g, &regs[i], &r2, n_a2, b, &ez);
        if (r2.cnt > 0) { /* [mm2:align.c:mm_insert_reg] */
            regs = (reg_t *)realloc(regs, (size_t)(n_regs + 1) * sizeof(reg_t));
            memmove(&regs[i + 2], &regs[i + 1], (size_t)(n_regs - i - 1) * sizeof(reg_t));
            regs[i + 1] = r2;
            ++n_regs;
        }
    }
    free(ez.cigar);
    { /* [mm2:hit.c:mm_filter_regs]; spec v1 also drops regions that produced no CIGAR */
        int k;
        for (i = k = 0; i < n_regs; ++i) {
            reg_t *r = &regs[i];
            int flt = 0;
            if (r->cnt < P->min_cnt) flt = 1;
            if (!r->has_p || r->n_cigar == 0) flt = 1;
            else if (r->mlen < P->min_chain_score) flt = 1;
            else if (r->dp_max < P->min_dp_max) flt = 1;
            if (flt) { free(r->cigar); r->cigar = 0; }
            else { if (k < i) regs[k] = regs[i]; ++k; }
        }
        n_regs = k;
    }
    if (n_regs > 1) { /* [mm2:hit.c:mm_hit_sort] */
        krec_t *z = (krec_t *)malloc((size_t)n_regs * sizeof(krec_t));
        reg_t *t = (reg_t *)malloc((size_t)n_regs * sizeof(reg_t));
        for (i = 0; i < n_regs; ++i) z[i].key = (uint64_t)(uint32_t)regs[i].dp_max << 32 | regs[i].hash, z[i].idx = i;
        qsort(z, (size_t)n_regs, sizeof(krec_t), cmp_krec_desc);
        for (i = 0; i < n_regs; ++i) t[i] = regs[z[i].idx];
        memcpy(regs, t, (size_t)n_regs * sizeof(reg_t));
        free(z); free(t);
    }
    if (n_regs > 0) {
        set_parent(P->mask_level, P->mask_len, n_regs, regs, P->a * 2 + P->b); /* [mm2:map.c:align_regs] */
        set_mapq(n_regs, regs, P->min_chain_score, P->a, rep_len);
    }
    for (i = 0; i < n_regs; ++i) {
        reg_t *r = &regs[i];
        kbo_hit_t *h;
        if (O->n == O->m) { O->m = O->m ? O->m << 1 : 256; O->h = (kbo_hit_t *)realloc(O->h, (size_t)O->m * sizeof(kbo_hit_t)); }
        if (O->ncg + r->n_cigar > O->mcg) { O->mcg = (O->ncg + r->n_cigar) * 2 + 64; O->cg = (uint32_t *)realloc(O->cg, (size_t)O->mcg * 4); }
        h = &O->h[O->n++];
        h->gene = g; h->q_start = r->qs; h->q_end = r->qe;
        h->t_ctg = r->rid; h->t_len = ai->ctg_len[r->rid]; h->t_start = r->rs; h->t_end = r->re;
        h->strand = r->rev ? -1 : 1;
        h->score = r->dp_score; h->matches = r->mlen; h->block_len = r->blen;
        h->edit_distance = r->blen - r->mlen + r->n_ambi;
        h->mapq = r->mapq; h->is_primary = r->parent == r->id;
        h->dp_max = r->dp_max; h->chain_score = r->score0; h->chain_cnt = r->cnt;
        h->cigar_off = O->ncg; h->n_cigar = r->n_cigar;
        memcpy(O->cg + O->ncg, r->cigar, (size_t)r->n_cigar * 4);
        O->ncg += r->n_cigar;
        free(r->cigar);
    }
    free(regs); free(b);
}

kbo_result_t *kbo_map_assembly(void *db_, const uint8_t *ctg_seqs, const int64_t *ctg_off, const int32_t *ctg_len, int32_t n_ctg, int32_t keep_stages)
{
    kbo_db_t *db = (kbo_db_t *)db_;
    asm_idx_t ai;
    ctx_t C;
    out_t O;
    kbo_result_t *R = (kbo_result_t *)calloc(1, sizeof(kbo_result_t));
    mm128_v v = {0, 0, 0};
    int32_t c, g;
    memset(&O, 0, sizeof(O));
    ai.n_ctg = n_ctg; ai.ctg_len = ctg_len;
    ai.ctg = (uint8_t **)calloc(n_ctg > 0 ? n_ctg : 1, sizeof(void *));
    for (c = 0; c < n_ctg; ++c) { /* [ref:genome.py:188-189] Index.build over every contig */
        int32_t i, L = ctg_len[c];
        const uint8_t *s = ctg_seqs + ctg_off[c];
        ai.ctg[c] = (uint8_t *)malloc(L > 0 ? (size_t)L : 1);
        for (i = 0; i < L; ++i) ai.ctg[c][i] = nt4(s[i]);
        sketch_nt4(ai.ctg[c], L, db->p.w, db->p.k, (uint32_t)c, &v);
    }
    qsort(v.a, v.n, sizeof(mm128_t), cmp_128xy);
    ai.mz = v.a; ai.n_mz = (int64_t)v.n;
    ai.mid_occ = cal_mid_occ(&ai, &db->p);
    C.P = &db->p; C.db = db; C.ai = &ai;
    for (g = 0; g < db->n_genes; ++g) map_gene(&C, g, &O, keep_stages); /* [ref:serotyping/core.py:154] */
    R->n_hits = O.n; R->hits = O.h; R->n_cigar = O.ncg; R->cigar = O.cg;
    R->mid_occ = ai.mid_occ; R->n_minimizers = ai.n_mz;
    R->n_anchors = O.nan; R->anchors = O.an; R->n_chains = O.nch; R->chains = O.ch;
    for (c = 0; c < n_ctg; ++c) free(ai.ctg[c]);
    free(ai.ctg); free(v.a);
    return R;
}

void kbo_result_free(kbo_result_t *r)
{
    if (!r) return;
    free(r->hits); free(r->cigar); free(r->anchors); free(r->chains); free(r);
}
