/* kb_oracle.h -- CPU oracle for the K/O-locus gene -> contig mapping path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under kaptive_b200/ may include, link or
 * dlopen this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in the third-party wheel
 * `rammappy` (pinned 0.1.3 in /root/reference/uv.lock:830-838, spec >=0.1.3 in
 * pyproject.toml:27-32), a Rust "minimap2-based" mapper whose source is not in
 * /root/reference and which cannot be installed here.  The reference holds no
 * golden vector for it.  This file restates the
 * published minimap2 algorithm (Li 2018; minimap2 2.24-2.28 defaults, no
 * preset) with the overrides the reference applies at its call site
 * (src/kaptive/serotyping/core.py:148-152: best_n=50000, pri_ratio=0.0,
 * do_cigar=True), and is anchored on that call site and on the consumer of
 * the hits (src/kaptive/core/alignment.py:392-474).
 *
 * Stated deviations from minimap2 (see DESIGN.md "Mapping spec v1"):
 *   - no RMQ re-chaining / long-join (minimap2 --no-long-join)
 *   - no high-occurrence seed rescue (minimap2 --occ-dist 0)
 *   - no inversion test / inversion alignment
 *   - no mm_update_dp_max rank adjustment, no mm_est_err divergence
 *   - banded DP treats out-of-band cells as -infinity (ksw2 extrapolates)
 *   - every unstable-sort tie in minimap2 is given a fixed total order here
 *   - logf in the MAPQ formula is the fdlibm polynomial evaluated without FMA
 *     contraction, so that it is bit-reproducible on CPU and GPU
 */
#ifndef KB_ORACLE_H
#define KB_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t k, w;
    int32_t min_cnt, min_chain_score, bw, max_gap, max_chain_skip, max_chain_iter;
    float chain_gap_scale;
    int32_t a, b, q, e, q2, e2, sc_ambi;
    int32_t zdrop, min_dp_max, min_ksw_len;
    int32_t mid_occ;          /* <=0: derive from the indexed assembly like minimap2 */
    int32_t min_mid_occ, max_mid_occ;
    float mid_occ_frac, q_occ_frac, mask_level;
    int32_t mask_len;
    int32_t seed;
    int32_t ext_bw;           /* (int)(bw*1.5+1) */
    int32_t max_sw_cells;     /* DP problems larger than this are treated as z-dropped */
} kbo_params_t;

/* One alignment record: the fields the reference drains per hit
 * (core/alignment.py:414-446). All int32 so numpy can view the array. */
typedef struct {
    int32_t gene;        /* query index (the reference names queries str(i)) */
    int32_t q_start, q_end;
    int32_t t_ctg, t_len, t_start, t_end;
    int32_t strand;      /* +1 / -1 */
    int32_t score;       /* DP score (AS) */
    int32_t matches;     /* mlen */
    int32_t block_len;   /* blen */
    int32_t edit_distance;
    int32_t mapq;
    int32_t is_primary;
    int32_t dp_max;
    int32_t chain_score;
    int32_t chain_cnt;
    int32_t cigar_off, n_cigar; /* into the cigar pool, BAM-encoded len<<4|op */
} kbo_hit_t;

typedef struct {
    int32_t gene;
    int32_t rev, rid, tpos;   /* target strand/contig/last base of k-mer */
    int32_t qpos, flags;      /* query last base (strand-oriented); bit0 tandem */
} kbo_anchor_t;

typedef struct {
    int32_t gene, score, cnt, rev, rid, rs, re, qs, qe;
} kbo_chain_t;

typedef struct {
    int32_t n_hits;
    kbo_hit_t *hits;
    int32_t n_cigar;
    uint32_t *cigar;
    int32_t mid_occ;
    int64_t n_minimizers;     /* minimizers in the assembly index */
    /* stage dumps for stage-level parity tests (filled when keep_stages != 0) */
    int64_t n_anchors;
    kbo_anchor_t *anchors;
    int32_t n_chains;
    kbo_chain_t *chains;
} kbo_result_t;

void kbo_params_default(kbo_params_t *p);

/* gene database (the queries; reference: serotyping/core.py:111-121) */
void *kbo_db_create(const uint8_t *seqs, const int64_t *offsets, const int32_t *lengths,
                    int32_t n_genes, const kbo_params_t *p);
void kbo_db_destroy(void *db);

/* map all genes against one assembly (reference: genome.py:188-189 Index.build
 * + serotyping/core.py:154 map_batch) */
kbo_result_t *kbo_map_assembly(void *db, const uint8_t *ctg_seqs, const int64_t *ctg_off,
                               const int32_t *ctg_len, int32_t n_ctg, int32_t keep_stages);
void kbo_result_free(kbo_result_t *r);

/* minimizer sketch of one sequence; returns count, writes up to cap (hash<<8|span, pos<<1|strand) */
int64_t kbo_sketch(const uint8_t *seq, int32_t len, int32_t w, int32_t k,
                   uint64_t *out_x, uint32_t *out_y, int64_t cap);

/* deterministic helpers exposed for unit tests */
float kbo_logf(float x);
float kbo_log2_fast(float x);
uint32_t kbo_hash32(uint32_t key, uint32_t mask);

#ifdef __cplusplus
}
#endif
#endif
