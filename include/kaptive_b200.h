/* kaptive_b200.h -- C-ABI of libkaptive_b200.so
 *
 * Drop-in boundary for the one hot path of klebgenomics/Kaptive that this
 * library replaces: mapping every K/O-locus reference gene onto the contigs of
 * draft assemblies.  In the reference that path is the `rammappy` module
 * (third-party Rust wheel), reached through exactly these call sites:
 *
 *   rammappy.fasta.parse_fasta_bytes   src/kaptive/core/genome.py:45
 *   rammappy.Index.build               src/kaptive/core/genome.py:188-189
 *   rammappy.align.Aligner(...)        src/kaptive/serotyping/core.py:148-152
 *   Aligner.map_batch(gene_seqs)       src/kaptive/serotyping/core.py:154
 *   hit fields drained per record      src/kaptive/core/alignment.py:409-446
 *
 * Each entry point below names the reference interface it replaces.  Plain
 * pointers and sizes only; no exceptions cross the ABI; every function
 * returns 0 on success or a negative kb_status, and kb_last_error() gives the
 * thread-local message.  Handles are immutable after creation and may be
 * shared between threads; each mapping call uses its own CUDA stream.
 *
 * There is NO CPU fallback: without a CUDA device (or without the library)
 * every compute entry point fails with KB_ERR_CUDA.
 */
#ifndef KAPTIVE_B200_H
#define KAPTIVE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    KB_OK = 0,
    KB_ERR_ARG = -1,       /* bad argument */
    KB_ERR_CUDA = -2,      /* CUDA runtime failure / no device */
    KB_ERR_LIMIT = -3,     /* input exceeds a documented limit */
    KB_ERR_CAPACITY = -4,  /* caller buffer too small */
    KB_ERR_INTERNAL = -5
} kb_status;

/* Mapping parameters: minimap2 defaults with no preset (what
 * Aligner(preset=None) means, serotyping/core.py:148) plus the two overrides
 * the reference applies (best_n=50000, pri_ratio=0.0, core.py:150-151), which
 * together mean "keep and extend every chain". */
typedef struct kb_params {
    int32_t k, w;
    int32_t min_cnt, min_chain_score, bw, max_gap, max_chain_skip, max_chain_iter;
    float chain_gap_scale;
    int32_t a, b, q, e, q2, e2, sc_ambi;
    int32_t zdrop, min_dp_max, min_ksw_len;
    int32_t mid_occ;            /* <=0: derived per assembly as minimap2 does */
    int32_t min_mid_occ, max_mid_occ;
    float mid_occ_frac, q_occ_frac, mask_level;
    int32_t mask_len;
    int32_t seed;
    int32_t ext_bw;
    int32_t max_sw_cells;
} kb_params_t;

void kb_params_default(kb_params_t *p);

const char *kb_last_error(void);
int kb_version(void);
/* number of CUDA devices visible to the library (0 => nothing can run) */
int kb_device_count(void);

/* ---- FASTA ingest: replaces rammappy.fasta.parse_fasta_bytes (genome.py:45) ----
 * Two-phase: count records, then fill caller arrays.  name_off/name_len and
 * seq_off/seq_len index into `data`; sequences that span several lines are
 * compacted IN a caller-provided output buffer `seq_out` (>= n bytes). */
int kb_fasta_count(const uint8_t *data, int64_t n, int64_t *n_records, int64_t *n_seq_bytes);
int kb_fasta_parse(const uint8_t *data, int64_t n, int64_t max_records,
                   int64_t *name_off, int32_t *name_len,
                   uint8_t *seq_out, int64_t seq_cap, int64_t *seq_off, int32_t *seq_len);

/* Batch ingest: n_files FASTA buffers (one assembly each) parsed by n_threads host threads into the layout
 * kb_map_assemblies / kb_batch_create take.  Replaces the reference's per-genome read + parse + copy (core/genome.py:45,
 * core/seq.py:307-325) for batches.  Two-phase like the single-file calls: count per file, prefix-sum in the caller
 * (rec_base / seq_base, n_files + 1 entries), then parse into caller buffers (seq_out may be pinned memory). */
int kb_fasta_ingest_count(const uint8_t *const *data, const int64_t *n, int32_t n_files, int32_t n_threads,
                          int64_t *n_records, int64_t *n_seq_bytes);
int kb_fasta_ingest_parse(const uint8_t *const *data, const int64_t *n, int32_t n_files, int32_t n_threads,
                          const int64_t *rec_base, const int64_t *seq_base, uint8_t *seq_out,
                          int64_t *contig_off, int32_t *contig_len, int32_t *asm_contig_start,
                          int64_t *name_off, int32_t *name_len);

/* Packed ingest: the same FASTA buffers -> 2 bit per base + ambiguity mask, written by the host threads straight into caller
 * (pinned) buffers in the device layout, so that 0.375 B per base cross PCIe and no pack kernel runs (replaces core/genome.py:45 +
 * core/seq.py:307-325 for batches; compressed files are opened by the Python layer with the reference's own rules,
 * core/genome.py:105-106,194-214).  Three passes: kb_fasta_ingest_count_records (records per file) -> kb_fasta_ingest_lengths (contig
 * lengths, names) -> kb_packed_layout -> kb_fasta_ingest_pack.
 * Layout: storage counted in bases; 128 padded bases in front, every contig starts on a multiple of 128 bases, 128 padded bases
 * behind; seq2 word k = bases 16k..16k+15 (2 bits each, A C G T = 0 1 2 3), nmask word k = bases 32k..32k+31 (1 = ambiguous/padding). */
int kb_fasta_ingest_count_records(const uint8_t *const *data, const int64_t *n, int32_t n_files, int32_t n_threads, int64_t *n_records);
int kb_packed_layout(const int32_t *contig_len, int64_t n_contigs, int64_t *contig_soff, int64_t *storage_bases);
int kb_fasta_ingest_lengths(const uint8_t *const *data, const int64_t *n, int32_t n_files, int32_t n_threads, const int64_t *rec_base,
                            int32_t *contig_len, int32_t *asm_contig_start, int64_t *name_off, int32_t *name_len);
int kb_fasta_ingest_pack(const uint8_t *const *data, const int64_t *n, int32_t n_files, int32_t n_threads, const int64_t *rec_base,
                         const int32_t *contig_len, const int64_t *contig_soff, int64_t storage_bases, uint32_t *seq2, uint32_t *nmask,
                         int32_t use_simd);

/* ---- gene index: the query side of map_batch (serotyping/core.py:111-121,154) ----
 * Built once per database; device-resident hash of every gene minimizer. */
typedef struct kb_index kb_index_t;
int kb_index_create(const uint8_t *gene_seqs, const int64_t *offsets, const int32_t *lengths,
                    int32_t n_genes, const kb_params_t *params, int device, kb_index_t **out);
void kb_index_destroy(kb_index_t *idx);
int32_t kb_index_n_genes(const kb_index_t *idx);
int64_t kb_index_n_minimizers(const kb_index_t *idx);
/* flat byte image for a one-off NCCL/MPI broadcast to other ranks */
int64_t kb_index_serialized_size(const kb_index_t *idx);
int kb_index_serialize(const kb_index_t *idx, uint8_t *buf, int64_t cap);
int kb_index_deserialize(const uint8_t *buf, int64_t n, int device, kb_index_t **out);

/* ---- assembly batch: replaces rammappy.Index.build (genome.py:188-189) ----
 * Contigs of n_asm assemblies, concatenated ASCII.  `contig_seqs` may be a
 * host pointer (pageable or pinned) or a device pointer (UVA).  The batch
 * keeps only the 2-bit packed sequence + ambiguity mask in HBM. */
typedef struct kb_batch kb_batch_t;
int kb_batch_create(const uint8_t *contig_seqs, const int64_t *contig_off, const int32_t *contig_len,
                    const int32_t *asm_contig_start /* n_asm+1 */, int32_t n_asm, int device, kb_batch_t **out);
/* The same from host-packed contigs (kb_fasta_ingest_pack): seq2 / nmask are the arrays of a whole ingest call, first_soff the
 * storage offset (bases) of this batch's first contig in them (kb_packed_layout). */
int kb_batch_create_packed(const uint32_t *seq2, const uint32_t *nmask, int64_t first_soff, const int32_t *contig_len,
                           const int32_t *asm_contig_start /* n_asm+1 */, int32_t n_asm, int device, kb_batch_t **out);
/* the batch's packed words copied back to the host (storage_bases / 16 and / 32 words; either pointer may be null) */
int kb_batch_download_packed(const kb_batch_t *b, uint32_t *seq2, uint32_t *nmask, int64_t *storage_bases);
void kb_batch_destroy(kb_batch_t *b);
int32_t kb_batch_n_assemblies(const kb_batch_t *b);
int64_t kb_batch_total_bases(const kb_batch_t *b);
int64_t kb_batch_packed_bytes(const kb_batch_t *b);

/* ---- mapping: replaces Aligner(...).map_batch(gene_seqs) (serotyping/core.py:148-154) ---- */
typedef struct kb_result kb_result_t;

/* One field per array, Alignments-style SoA (core/alignment.py:295-317). */
typedef struct kb_hits {
    int64_t capacity;
    int32_t *asm_id;        /* assembly index within the batch */
    int32_t *gene;          /* query index == int(q_name) (serotyping/core.py:158) */
    int32_t *q_start, *q_end;
    int32_t *t_ctg;         /* contig index within its assembly (-> target_name) */
    int32_t *t_len, *t_start, *t_end;
    int8_t  *strand;        /* +1 / -1 */
    int32_t *score, *matches, *block_len, *edit_distance;
    uint8_t *mapq;
    uint8_t *is_primary;
    int64_t *cigar_off;     /* into the cigar pool */
    int32_t *n_cigar;
} kb_hits_t;

int kb_map_batch(const kb_index_t *idx, const kb_batch_t *batch, kb_result_t **out);
int kb_result_size(const kb_result_t *r, int64_t *n_hits, int64_t *n_cigar);
/* copy device results into caller (host) arrays; hits ordered by (asm, gene, rank) */
int kb_result_fetch(const kb_result_t *r, kb_hits_t *dst, uint32_t *cigar, int64_t cigar_cap);
void kb_result_destroy(kb_result_t *r);

/* per-stage device time of the call that produced `r`, CUDA events on its stream */
enum { KB_STAGE_SCAN = 0, KB_STAGE_SORT, KB_STAGE_CHAIN, KB_STAGE_ALIGN, KB_STAGE_FINAL, KB_STAGE_TOTAL, KB_N_STAGES };
int kb_result_stage_ms(const kb_result_t *r, float *ms /* KB_N_STAGES */);
/* counters: [0] minimizers scanned-out, [1] anchors, [2] query groups, [3] chains, [4] raw hits, [5] kernel launches, [6] DP cells,
 * [7] chains the staged aligner handed to the one-warp aligner, [8] alignments DROPPED because they ran into an internal limit
 * (chain window > 65536 target bases, CIGAR > 8192 operations): callers that need completeness must treat [8] != 0 as an error
 * (the Python layer raises), [9..15] reserved */
int kb_result_counters(const kb_result_t *r, int64_t *c /* 16 */);
/* stage dumps for parity tests (device -> host); arrays of int32 records */
int kb_result_fetch_anchors(const kb_result_t *r, int32_t *out /* n x 7: asm,gene,rev,rid,tpos,qpos,flags */, int64_t cap, int64_t *n);
int kb_result_fetch_chains(const kb_result_t *r, int32_t *out /* n x 10: asm,gene,score,cnt,rev,rid,rs,re,qs,qe */, int64_t cap, int64_t *n);
int kb_result_mid_occ(const kb_result_t *r, int32_t *out /* n_asm */);

/* Returns the idle workspace arenas of `device` (-1: every device) to the CUDA allocator.  A mapping call keeps its scratch
 * arena (tens of GB for large batches) for the life of the process so that steady-state calls never reach the allocator;
 * a long-lived server calls this when it goes idle.  Arenas of calls in flight are untouched. */
int kb_release_workspace(int device);

/* diagnostic: calls / cells of the base-level DP per (kind, path) since the last reset, process-wide on the current
 * device; slot 2 * (5 * kind + path) = calls, + 1 = cells; kind 0 gap fill, 1 end extension, 2 gap fill with z-drop;
 * path 0 band certified, 1 band rejected, 2 register single pass, 3 register tiled, 4 scratch-memory DP. */
int kb_debug_dp_stats(int64_t *out32, int reset);

/* diagnostic / parity tests: runs n independent base-level DP problems (nt4 codes 0..4, concatenated) through ONE of the DP
 * kernels: mode 0 scratch-memory DP (the form closest to oracle/kb_oracle.c:extd2), 1 row-stripe wavefront, 2 packed 16-bit
 * row-stripe wavefront (job alone), 3 certified 64-diagonal band pass, 4 packed 16-bit wavefront with jobs 2i and 2i + 1 as a pair.  flag: KB_EZ_* bits (1 extension only, 2 right-aligned gaps,
 * 4 reversed CIGAR, 8 global without z-drop).  out: n x 8 int32 = score, max, max_t, max_q, zdropped, n_cigar, ran (0: the
 * kernel is not eligible for this problem, 2: band pass not certified), 0; cig: n x cig_stride BAM-encoded operations. */
int kb_debug_dp(const kb_params_t *params, int device, const uint8_t *q, const int64_t *q_off, const int32_t *q_len, const uint8_t *t,
                const int64_t *t_off, const int32_t *t_len, const int32_t *flag, const int32_t *w, const int32_t *zdrop, int32_t n,
                int32_t mode, int32_t *out, uint32_t *cig, int32_t cig_stride);

/* one-call convenience for HOST buffers (the end-to-end path): batch_create +
 * map + fetch + destroy; copies are inside. */
int kb_map_assemblies(const kb_index_t *idx,
                      const uint8_t *contig_seqs, const int64_t *contig_off, const int32_t *contig_len,
                      const int32_t *asm_contig_start, int32_t n_asm,
                      kb_hits_t *dst, int64_t *n_hits, uint32_t *cigar, int64_t cigar_cap, int64_t *n_cigar);

/* the same for host-packed contigs: the end-to-end path of a caller that ingests FASTA with kb_fasta_ingest_pack */
int kb_map_assemblies_packed(const kb_index_t *idx, const uint32_t *seq2, const uint32_t *nmask, const int32_t *contig_len,
                             const int32_t *asm_contig_start, int32_t n_asm,
                             kb_hits_t *dst, int64_t *n_hits, uint32_t *cigar, int64_t cigar_cap, int64_t *n_cigar);

/* ---- minimizer scan only (the roofline kernel), for benchmarking / parity ----
 * Runs the scan kernel over the batch and returns all minimizers of assembly
 * `asm_id` as (hash, contig, pos<<1|strand) triples, unsorted. */
int kb_scan_minimizers(const kb_index_t *idx, const kb_batch_t *batch, int32_t asm_id,
                       uint32_t *hash, int32_t *ctg, uint32_t *pos_strand, int64_t cap, int64_t *n);
/* times `iters` launches of the seeding scan kernel alone; returns mean ms */
int kb_bench_scan(const kb_index_t *idx, const kb_batch_t *batch, int iters, float *mean_ms, int64_t *n_anchors);

/* ---- post-mapping numerics, batched over all items of a batch of assemblies ----
 * Host or device pointers in, host arrays out; bit-exact replacements of the reference's numba kernels. */
const char *kb_post_last_error(void);
/* _extract_ragged_kernel (src/kaptive/core/seq.py:612-668): slices [starts, ends) of parent sequences, reverse-
 * complemented where strands < 0.  out_off / out_len: n entries. */
int kb_post_extract(const uint8_t *seqs, int64_t n_seq_bytes, const int64_t *parent_off, int32_t n_parents, const int32_t *indices,
                    const int32_t *starts, const int32_t *ends, const int8_t *strands, int32_t n, uint8_t *out, int64_t out_cap,
                    int64_t *out_off, int32_t *out_len);
/* _translate_ragged_kernel (src/kaptive/core/seq.py:671-741): table 11, per-item frame, optional stop at first '*'. */
int kb_post_translate(const uint8_t *seqs, int64_t n_seq_bytes, const int64_t *offsets, const int32_t *lengths, const int8_t *frames,
                      int32_t n, int32_t to_stop, uint8_t *out, int64_t out_cap, int64_t *out_off, int32_t *out_len, int64_t *n_out);
/* _batched_banded_gotoh (src/kaptive/core/pairwise.py:395-584), unseeded: banded local Gotoh on BLOSUM62 with traceback.
 * res: n x 8 int32 = score, matches, mismatches, gaps, q_start, q_end, t_start, t_end. */
int kb_post_protein_align(const uint8_t *q, const int64_t *q_off, const int32_t *q_len, const uint8_t *t, const int64_t *t_off,
                          const int32_t *t_len, int32_t n, int32_t k, int32_t gap_open, int32_t gap_extend, int32_t *res);

/* _cull_overlaps_kernel (src/kaptive/core/interval.py:698-751; driven by Alignments.cull_overlaps, core/alignment.py:643-686):
 * greedy overlap cull in the given evaluation order, batched over segments (one segment = the hits of one assembly).  Arrays are
 * concatenated over the segments, seg_off has n_seg + 1 entries, `order` holds indices local to each segment.  kept: 0 / 1 per hit. */
int kb_post_cull_overlaps(const int32_t *order, const int32_t *group1, const int32_t *group2, const int32_t *starts, const int32_t *ends,
                          double max_overlap_fraction, const int64_t *seg_off, int32_t n_seg, uint8_t *kept);
/* _cluster_kernel (src/kaptive/core/interval.py:595-639; driven by Intervals.cluster_spatial :471-493): single-linkage clustering
 * of intervals swept in `order`; cluster ids restart at 0 in every segment. */
int kb_post_cluster(const int32_t *starts, const int32_t *ends, const int32_t *groups, int32_t tolerance, const int32_t *order,
                    const int64_t *seg_off, int32_t n_seg, int32_t *cluster_ids);

/* ---- batched typing of mapped assemblies: the post-mapping part of Serotyper.__call__
 * (src/kaptive/serotyping/core.py:157-486) for a whole batch.  Array logic on host threads, numerics (extract + translate +
 * protein Gotoh of every retained hit) in one device pass over the resident packed batch.  Same tie rules as the reference. */
typedef struct kb_typedb kb_typedb_t;
typedef struct kb_typed kb_typed_t;
const char *kb_type_last_error(void);
/* one entry per database gene: db.genes.lengths, db.gene_locus_indices, db.extra_genes, db.gene_positions, db.gene_intervals.strands;
 * locus_len = db.loci.lengths; translations = db.translations (bytes back to back, trans_len each); id_threshold = db.metadata.id_threshold */
int kb_typedb_create(int32_t n_genes, const int32_t *gene_len, const int32_t *gene_locus, const uint8_t *extra, const int32_t *gene_pos,
                     const int8_t *gene_strand, int32_t n_loci, const int32_t *locus_len, int32_t max_locus_length, const uint8_t *translations,
                     const int32_t *trans_len, double id_threshold, int device, kb_typedb_t **out);
void kb_typedb_destroy(kb_typedb_t *d);
/* serotyping/core.py:163-201: locus_scores (n_asm x n_loci float64) and locus_counts (float32) from hits sorted by (assembly, gene);
 * the caller finishes :199-206 (completeness ** 3, argmax) with numpy, whose float32 power the reference uses */
int kb_type_score(const kb_typedb_t *d, const int32_t *asm_id, const int32_t *gene, const int32_t *q_start, const int32_t *q_end, const int32_t *score,
                  int64_t n_hits, int32_t n_asm, double min_gene_coverage, int32_t n_threads, double *locus_scores, float *locus_counts);
/* serotyping/core.py:209-459 for every assembly of `batch` given its best locus (and that locus' un-penalised score, :471) */
int kb_type_call(const kb_typedb_t *d, const kb_batch_t *batch, const int32_t *asm_id, const int32_t *gene, const int32_t *q_start, const int32_t *q_end,
                 const int32_t *t_ctg, const int32_t *t_len, const int32_t *t_start, const int32_t *t_end, const int8_t *strand, const int32_t *score,
                 const int32_t *matches, const uint8_t *mapq, int64_t n_hits, int32_t n_asm, const int32_t *best_locus, const double *best_score,
                 int32_t max_other_genes, double min_completeness, int32_t allow_below_threshold, int32_t partial_edge_tolerance, int32_t n_threads,
                 kb_typed_t **out);
void kb_typed_destroy(kb_typed_t *r);
/* diagnostic: seconds the last kb_type_call spent in pass 1 (host), building the job list, the device numerics, pass 2 (host) */
void kb_type_debug_times(double *t4);
int kb_typed_sizes(const kb_typed_t *r, int64_t *n_gene_hits, int64_t *n_pieces, int64_t *n_missing);
/* per assembly (offset arrays: n_asm + 1); problems: bit 0 fragmented, 1 unexpected genes, 2 missing genes, 3 novel genes, 4 truncated genes */
int kb_typed_fetch_assemblies(const kb_typed_t *r, double *score, double *completeness, double *pcov, double *length_discrepancy, uint8_t *typeable,
                              uint8_t *problems, int32_t *n_pieces, int64_t *gene_hit_off, int64_t *piece_off, int64_t *missing_off);
/* gene hits in the reference's order (GeneHits after the spurious-hit filter); state: 0 normal, 1 partial, 2 truncated, 3 novel */
int kb_typed_fetch_gene_hits(const kb_typed_t *r, int32_t *gene, int32_t *q_start, int32_t *q_end, int32_t *t_ctg, int32_t *t_start, int32_t *t_end,
                             int8_t *strand, int8_t *state, uint8_t *is_expected, uint8_t *is_inside, uint8_t *is_extra, float *prot_ident,
                             float *coverage);
int kb_typed_fetch_pieces(const kb_typed_t *r, int32_t *ctg, int32_t *start, int32_t *end, int8_t *strand, int32_t *missing);

#ifdef __cplusplus
}
#endif
#endif /* KAPTIVE_B200_H */
