// kb_type.cpp -- batched typing of mapped assemblies: the post-mapping part of
// kaptive.serotyping.Serotyper.__call__ (src/kaptive/serotyping/core.py:157-486) for a whole batch of assemblies at once.
//
//   kb_type_score   core.py:163-201   per assembly and locus: sum of the best query coverage of every expected gene, count of matched
//                                      expected genes (the caller finishes :203-207 with numpy, whose float32 power the reference uses)
//   kb_type_call    core.py:209-459   overlap cull with the best locus' genes prioritised (alignment.py:643-686, interval.py:698-751),
//                                      spatial clustering (interval.py:471-493,595-639), locus pieces, inside / expected / missing genes,
//                                      gene states from the translated hits and their protein alignments, confidence
//
// The array logic runs on host threads, one assembly at a time per thread (hundreds of hits each: the reference spends ~40 ms per
// assembly here in numba wake-ups and Python; this is ~20 us); the numerics -- extract + translate + banded protein Gotoh for
// every retained hit of every assembly -- are ONE device pass over the resident 2-bit batch (kb_post_type_numerics).  Tie rules are
// the reference's: stable lexsorts, first maximum, `>=` where it says `>=`; the uint8 negation of mapq in the cull order wraps.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>
#include "kb_common.cuh"

struct kb_batch;
const KbBatchView *kb_batch_view_internal(const kb_batch *b, int *device);
int kb_post_type_numerics(const KbBatchView &bv, int device, const int32_t *ctg, const int32_t *ts, const int32_t *te, const int8_t *strand,
                          const int8_t *frame, const int32_t *gene, int64_t n, const uint8_t *d_trans, const int64_t *d_trans_off,
                          const int32_t *h_trans_len, int32_t k, int32_t go, int32_t ge, int32_t *prot_len, int32_t *res);
void *kb_type_dev_upload(int device, const void *h, size_t bytes);
void kb_type_dev_free(int device, void *p);
int kb_type_dev_download(int device, void *h, const void *dptr, size_t bytes);

#include <chrono>
static double g_type_times[4] = {0, 0, 0, 0};  // last kb_type_call: pass 1, job list, device numerics, pass 2 + assembly (seconds)
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static thread_local std::string g_type_err;
static int tfail(int code, const std::string &m)
{
    g_type_err = m;
    return code;
}

struct kb_typedb {
    int32_t n_genes = 0, n_loci = 0, device = 0;
    std::vector<int32_t> gene_len, gene_locus, gene_pos, locus_len, trans_len;
    std::vector<int8_t> gene_strand;
    std::vector<uint8_t> extra;
    std::vector<std::vector<int32_t>> expected;  // per locus: its non-extra genes, ascending
    int32_t max_locus_length = 0;
    double id_threshold = 0;
    uint8_t *d_trans = nullptr;
    int64_t *d_trans_off = nullptr;
};

struct kb_typed {  // result of kb_type_call
    int32_t n_asm = 0;
    std::vector<double> score, completeness, pcov, length_discrepancy;
    std::vector<uint8_t> typeable, problems;
    std::vector<int32_t> n_pieces;
    std::vector<int64_t> gh_off, piece_off, miss_off;  // n_asm + 1 each
    // gene hits (after the spurious-hit filter), in the reference's order
    std::vector<int32_t> gene, q_start, q_end, t_ctg, t_start, t_end;
    std::vector<int8_t> strand, state;
    std::vector<uint8_t> is_expected, is_inside, is_extra;
    std::vector<float> prot_ident, coverage;
    std::vector<int32_t> piece_ctg, piece_start, piece_end;
    std::vector<int8_t> piece_strand;
    std::vector<int32_t> missing;
};

template <class F>
static void parallel_for(int64_t n, int n_threads, F f)
{
    if (n_threads < 1) n_threads = 1;
    if ((int64_t)n_threads > n) n_threads = n > 0 ? (int)n : 1;
    std::atomic<int64_t> next{0};
    const int64_t grain = 16;
    auto worker = [&]() {
        for (int64_t i; (i = next.fetch_add(grain)) < n;)
            for (int64_t k = i; k < i + grain && k < n; ++k) f(k);
    };
    if (n_threads == 1) {
        worker();
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(worker);
    for (auto &t : th) t.join();
}

extern "C" {

const char *kb_type_last_error(void) { return g_type_err.c_str(); }

// gene_*: one entry per database gene (db.genes.lengths, db.gene_locus_indices, db.extra_genes, db.gene_positions,
// db.gene_intervals.strands); locus_len: db.loci.lengths; translations: db.translations (concatenated bytes + lengths)
int kb_typedb_create(int32_t n_genes, const int32_t *gene_len, const int32_t *gene_locus, const uint8_t *extra, const int32_t *gene_pos,
                     const int8_t *gene_strand, int32_t n_loci, const int32_t *locus_len, int32_t max_locus_length, const uint8_t *translations,
                     const int32_t *trans_len, double id_threshold, int device, kb_typedb **out)
{
    if (!out || n_genes < 0 || n_loci <= 0 || (n_genes > 0 && (!gene_len || !gene_locus || !extra || !gene_pos || !gene_strand || !trans_len)) || !locus_len)
        return tfail(KB_ERR_ARG, "null argument");
    kb_typedb *d = new kb_typedb();
    d->n_genes = n_genes, d->n_loci = n_loci, d->device = device, d->max_locus_length = max_locus_length, d->id_threshold = id_threshold;
    d->gene_len.assign(gene_len, gene_len + n_genes), d->gene_locus.assign(gene_locus, gene_locus + n_genes);
    d->gene_pos.assign(gene_pos, gene_pos + n_genes), d->gene_strand.assign(gene_strand, gene_strand + n_genes);
    d->extra.assign(extra, extra + n_genes), d->locus_len.assign(locus_len, locus_len + n_loci), d->trans_len.assign(trans_len, trans_len + n_genes);
    d->expected.resize((size_t)n_loci);
    std::vector<int64_t> off((size_t)n_genes + 1, 0);
    for (int32_t g = 0; g < n_genes; ++g) {
        if (gene_locus[g] < 0 || gene_locus[g] >= n_loci) {
            delete d;
            return tfail(KB_ERR_ARG, "gene_locus out of range");
        }
        if (!extra[g]) d->expected[(size_t)gene_locus[g]].push_back(g);
        off[(size_t)g + 1] = off[(size_t)g] + trans_len[g];
    }
    if (device >= 0) {  // device < 0: scoring tables only (kb_type_score is host work); kb_type_call then refuses to run
        d->d_trans = (uint8_t *)kb_type_dev_upload(device, translations, (size_t)off[(size_t)n_genes] + 1);
        d->d_trans_off = (int64_t *)kb_type_dev_upload(device, off.data(), ((size_t)n_genes + 1) * 8);
        if (!d->d_trans || !d->d_trans_off) {
            delete d;
            return tfail(KB_ERR_CUDA, "no CUDA device / upload failed: the typing numerics have no CPU fallback");
        }
    }
    *out = d;
    return KB_OK;
}
void kb_typedb_destroy(kb_typedb *d)
{
    if (!d) return;
    kb_type_dev_free(d->device, d->d_trans), kb_type_dev_free(d->device, d->d_trans_off);
    delete d;
}

// Scoring sums (serotyping/core.py:163-201).  Hits sorted by (assembly, gene, rank) as kb_result_fetch returns them.
// locus_scores: n_asm x n_loci float64 (sum of the best q_cov of every matched expected gene, added in gene order like np.add.at);
// locus_counts: n_asm x n_loci float32 (matched expected genes).  Both zero-filled here.
int kb_type_score(const kb_typedb *d, const int32_t *asm_id, const int32_t *gene, const int32_t *q_start, const int32_t *q_end, const int32_t *score,
                  int64_t n_hits, int32_t n_asm, double min_gene_coverage, int32_t n_threads, double *locus_scores, float *locus_counts)
{
    if (!d || n_asm < 0 || (n_hits > 0 && (!asm_id || !gene || !q_start || !q_end || !score)) || !locus_scores || !locus_counts)
        return tfail(KB_ERR_ARG, "null argument");
    const int64_t nl = d->n_loci;
    std::fill(locus_scores, locus_scores + (int64_t)n_asm * nl, 0.0);
    std::fill(locus_counts, locus_counts + (int64_t)n_asm * nl, 0.0f);
    std::vector<int64_t> seg((size_t)n_asm + 1, 0);
    for (int64_t i = 0; i < n_hits; ++i) {
        if (asm_id[i] < 0 || asm_id[i] >= n_asm || gene[i] < 0 || gene[i] >= d->n_genes || (i && asm_id[i] < asm_id[i - 1]))
            return tfail(KB_ERR_ARG, "hits must be sorted by assembly, with valid assembly / gene indices");
        ++seg[(size_t)asm_id[i] + 1];
    }
    for (int32_t a = 0; a < n_asm; ++a) seg[(size_t)a + 1] += seg[(size_t)a];
    parallel_for(n_asm, n_threads, [&](int64_t a) {
        double *ls = locus_scores + a * nl;
        float *lc = locus_counts + a * nl;
        int64_t i = seg[(size_t)a];
        const int64_t e = seg[(size_t)a + 1];
        while (i < e) {  // one gene at a time (hits of a gene are adjacent; a gene out of order starts a new group, like np.unique would merge:
                         // the mapper's order is by gene, so groups are whole)
            const int32_t g = gene[i];
            double best_cov = -1.0;
            int32_t best_score = 0;
            int64_t j = i;
            for (; j < e && gene[j] == g; ++j) {
                const int32_t ql = d->gene_len[g];
                const double cov = ql > 0 ? (double)(q_end[j] - q_start[j]) / (double)ql : 0.0;
                if (!(cov >= min_gene_coverage)) continue;
                // lexsort((-scores, -q_covs, gene)): highest q_cov, then highest score, then first
                if (cov > best_cov || (cov == best_cov && score[j] > best_score)) best_cov = cov, best_score = score[j];
            }
            if (best_cov >= 0.0 && !d->extra[g]) ls[d->gene_locus[g]] += best_cov, lc[d->gene_locus[g]] += 1.0f;
            i = j;
        }
    });
    return KB_OK;
}

// per-assembly working set of kb_type_call
struct AsmWork {
    std::vector<int32_t> idx;        // culled hits: indices into the call's hit arrays, original order
    std::vector<uint8_t> expected, inside, extra;
    std::vector<int32_t> piece_ctg, piece_start, piece_end;
    std::vector<int8_t> piece_strand;
    std::vector<int32_t> missing;
    double completeness = 1.0, pcov = 0.0, ldisc = NAN, score = 0.0;
    int64_t job0 = 0;  // first numerics job of the assembly
};

int kb_type_call(const kb_typedb *d, const kb_batch *batch, const int32_t *asm_id, const int32_t *gene, const int32_t *q_start, const int32_t *q_end,
                 const int32_t *t_ctg, const int32_t *t_len, const int32_t *t_start, const int32_t *t_end, const int8_t *strand, const int32_t *score,
                 const int32_t *matches, const uint8_t *mapq, int64_t n_hits, int32_t n_asm, const int32_t *best_locus, const double *best_score,
                 int32_t max_other_genes, double min_completeness, int32_t allow_below_threshold, int32_t partial_edge_tolerance, int32_t n_threads,
                 kb_typed **out)
{
    if (!d || !batch || !out || n_asm < 0 || !best_locus ||
        (n_hits > 0 && (!asm_id || !gene || !q_start || !q_end || !t_ctg || !t_len || !t_start || !t_end || !strand || !score || !matches || !mapq)))
        return tfail(KB_ERR_ARG, "null argument");
    if (!d->d_trans) return tfail(KB_ERR_CUDA, "typing database was created without a device: the numerics have no CPU fallback");
    int device = 0;
    const KbBatchView *bv = kb_batch_view_internal(batch, &device);
    if (!bv || bv->n_asm != n_asm) return tfail(KB_ERR_ARG, "batch and hit arrays describe different numbers of assemblies");
    if (device != d->device) return tfail(KB_ERR_ARG, "typing database and batch live on different devices");
    std::vector<int64_t> seg((size_t)n_asm + 1, 0);
    for (int64_t i = 0; i < n_hits; ++i) {
        if (asm_id[i] < 0 || asm_id[i] >= n_asm || gene[i] < 0 || gene[i] >= d->n_genes || (i && asm_id[i] < asm_id[i - 1]))
            return tfail(KB_ERR_ARG, "hits must be sorted by assembly, with valid assembly / gene indices");
        ++seg[(size_t)asm_id[i] + 1];
    }
    for (int32_t a = 0; a < n_asm; ++a) seg[(size_t)a + 1] += seg[(size_t)a];
    std::vector<AsmWork> W((size_t)n_asm);
    // the host copy of the batch's contig table: contig ranges per assembly (batch-global contig index = first + t_ctg)
    std::vector<int32_t> acs((size_t)n_asm + 1);
    if (kb_type_dev_download(device, acs.data(), bv->asm_ctg_start, ((size_t)n_asm + 1) * 4)) return tfail(KB_ERR_CUDA, "D2H of the contig table failed");

    // ------------------------------------------------------------ pass 1: cull, cluster, pieces, inside / missing
    const double tm0 = now_s();
    parallel_for(n_asm, n_threads, [&](int64_t a) {
        AsmWork &w = W[(size_t)a];
        const int64_t lo = seg[(size_t)a], n = seg[(size_t)a + 1] - lo;
        const int32_t bl = best_locus[a];
        w.score = best_score ? best_score[a] : 0.0;
        // ---- Alignments.cull_overlaps(by_query=False, priority_mask, 0.1)
        std::vector<uint8_t> kept((size_t)n, 0);
        if (n < 2) std::fill(kept.begin(), kept.end(), 1);
        else {
            std::vector<int32_t> order((size_t)n);
            std::iota(order.begin(), order.end(), 0);
            auto key_score = [&](int32_t i) { return (double)score[lo + i] + (d->gene_locus[gene[lo + i]] == bl ? 1e9 : 0.0); };
            // np.lexsort((-qualities, -matches, -scores)): -scores first, then -matches, then the WRAPPED uint8 negation of mapq; stable
            std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
                const double sx = -key_score(x), sy = -key_score(y);
                if (sx != sy) return sx < sy;
                const int32_t mx = -matches[lo + x], my = -matches[lo + y];
                if (mx != my) return mx < my;
                return (uint8_t)(0u - mapq[lo + x]) < (uint8_t)(0u - mapq[lo + y]);
            });
            for (int64_t i = 0; i < n; ++i) {
                const int32_t x = order[(size_t)i];
                const int32_t s = t_start[lo + x], e = t_end[lo + x], len = e - s;
                if (len <= 0) continue;
                bool hit = false;
                for (int64_t j = 0; j < i && !hit; ++j) {
                    const int32_t p = order[(size_t)j];
                    if (!kept[(size_t)p] || t_ctg[lo + p] != t_ctg[lo + x]) continue;
                    const int32_t ks = t_start[lo + p], ke = t_end[lo + p];
                    const int32_t ov = std::min(e, ke) - std::max(s, ks);
                    if (ov > 0 && ((double)ov / (double)std::min(len, ke - ks)) > 0.1) hit = true;
                }
                if (!hit) kept[(size_t)x] = 1;
            }
        }
        for (int64_t i = 0; i < n; ++i)
            if (kept[(size_t)i]) w.idx.push_back((int32_t)i);
        const int64_t m = (int64_t)w.idx.size();
        auto G = [&](int64_t k) { return gene[lo + w.idx[(size_t)k]]; };
        auto TS = [&](int64_t k) { return t_start[lo + w.idx[(size_t)k]]; };
        auto TE = [&](int64_t k) { return t_end[lo + w.idx[(size_t)k]]; };
        auto TC = [&](int64_t k) { return t_ctg[lo + w.idx[(size_t)k]]; };
        // ---- cluster_spatial(tolerance = max_locus_length, group_by = contig): order = lexsort((ends, starts, groups)), single linkage
        std::vector<int32_t> piece((size_t)m, 0);
        if (m > 0) {
            std::vector<int32_t> order((size_t)m);
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
                if (TC(x) != TC(y)) return TC(x) < TC(y);
                if (TS(x) != TS(y)) return TS(x) < TS(y);
                return TE(x) < TE(y);
            });
            int32_t cur = 0, cur_e = TE(order[0]), cur_g = TC(order[0]);
            piece[(size_t)order[0]] = 0;
            for (int64_t i = 1; i < m; ++i) {
                const int32_t x = order[(size_t)i];
                if (TC(x) == cur_g && (int64_t)TS(x) <= (int64_t)cur_e + d->max_locus_length) cur_e = std::max(cur_e, TE(x));
                else ++cur, cur_e = TE(x), cur_g = TC(x);
                piece[(size_t)x] = cur;
            }
        }
        w.expected.assign((size_t)m, 0), w.extra.assign((size_t)m, 0), w.inside.assign((size_t)m, 0);
        for (int64_t k = 0; k < m; ++k) {
            w.extra[(size_t)k] = d->extra[G(k)];
            w.expected[(size_t)k] = d->gene_locus[G(k)] == bl && !d->extra[G(k)];
        }
        // primary hit of every expected gene: lexsort((-scores, gene)) -> first of each gene = highest score, then first in order
        std::vector<uint8_t> primary((size_t)m, 0);
        {
            std::vector<int32_t> ex;
            for (int64_t k = 0; k < m; ++k)
                if (w.expected[(size_t)k]) ex.push_back((int32_t)k);
            std::stable_sort(ex.begin(), ex.end(), [&](int32_t x, int32_t y) {
                if (G(x) != G(y)) return G(x) < G(y);
                return score[lo + w.idx[(size_t)x]] > score[lo + w.idx[(size_t)y]];
            });
            for (size_t i = 0; i < ex.size(); ++i)
                if (i == 0 || G(ex[i]) != G(ex[i - 1])) primary[(size_t)ex[i]] = 1;
        }
        // bounding pieces of the clusters that hold expected genes, in ascending cluster id
        std::vector<int32_t> cl;
        for (int64_t k = 0; k < m; ++k)
            if (w.expected[(size_t)k]) cl.push_back(piece[(size_t)k]);
        std::sort(cl.begin(), cl.end());
        cl.erase(std::unique(cl.begin(), cl.end()), cl.end());
        std::vector<double> means;
        std::vector<int32_t> pc, ps, pe;
        std::vector<int8_t> pst;
        for (int32_t c : cl) {
            int32_t first_ctg = -1, smin = INT32_MAX, emax = INT32_MIN;
            int64_t npri = 0, strand_sum = 0;
            double pos_sum = 0;
            for (int64_t k = 0; k < m; ++k) {
                if (piece[(size_t)k] != c) continue;
                if (first_ctg < 0) first_ctg = TC(k);
                if (!primary[(size_t)k]) continue;
                ++npri, smin = std::min(smin, TS(k)), emax = std::max(emax, TE(k));
                pos_sum += (double)d->gene_pos[G(k)];
                strand_sum += (int64_t)strand[lo + w.idx[(size_t)k]] * (int64_t)d->gene_strand[G(k)];
            }
            if (!npri) continue;
            pc.push_back(first_ctg), ps.push_back(smin), pe.push_back(emax), pst.push_back(strand_sum < 0 ? -1 : 1);
            means.push_back(pos_sum / (double)npri);
        }
        for (int64_t k = 0; k < m; ++k)
            for (size_t p = 0; p < pc.size(); ++p)
                if (TC(k) == pc[p] && TS(k) <= pe[p] && TE(k) >= ps[p]) w.inside[(size_t)k] = 1;
        std::vector<int32_t> po(pc.size());
        std::iota(po.begin(), po.end(), 0);
        std::stable_sort(po.begin(), po.end(), [&](int32_t x, int32_t y) { return means[(size_t)x] < means[(size_t)y]; });
        for (int32_t p : po) w.piece_ctg.push_back(pc[(size_t)p]), w.piece_start.push_back(ps[(size_t)p]), w.piece_end.push_back(pe[(size_t)p]), w.piece_strand.push_back(pst[(size_t)p]);
        // missing expected genes, completeness, coverage of the locus
        const std::vector<int32_t> &exp_genes = d->expected[(size_t)bl];
        std::vector<uint8_t> found((size_t)d->n_genes, 0);
        for (int64_t k = 0; k < m; ++k)
            if (w.expected[(size_t)k] && w.inside[(size_t)k]) found[(size_t)G(k)] = 1;
        for (int32_t g : exp_genes)
            if (!found[(size_t)g]) w.missing.push_back(g);
        w.completeness = exp_genes.empty() ? 1.0 : 1.0 - ((double)w.missing.size() / (double)exp_genes.size());
        int64_t assem = 0;
        for (size_t p = 0; p < w.piece_start.size(); ++p) assem += (int64_t)w.piece_end[p] - w.piece_start[p];
        const int32_t ref_len = d->locus_len[(size_t)bl];
        w.pcov = ref_len > 0 ? std::min(100.0, ((double)assem / (double)ref_len) * 100.0) : 0.0;
        w.ldisc = w.piece_start.size() == 1 ? (double)(assem - ref_len) : NAN;
    });

    // ------------------------------------------------------------ device pass: translate + protein alignment of every retained hit
    const double tm1 = now_s();
    int64_t n_jobs = 0;
    for (int32_t a = 0; a < n_asm; ++a) W[(size_t)a].job0 = n_jobs, n_jobs += (int64_t)W[(size_t)a].idx.size();
    std::vector<int32_t> j_ctg((size_t)n_jobs), j_ts((size_t)n_jobs), j_te((size_t)n_jobs), j_gene((size_t)n_jobs), prot_len((size_t)n_jobs), res((size_t)n_jobs * 8);
    std::vector<int8_t> j_strand((size_t)n_jobs), j_frame((size_t)n_jobs);
    parallel_for(n_asm, n_threads, [&](int64_t a) {
        const AsmWork &w = W[(size_t)a];
        const int64_t lo = seg[(size_t)a];
        for (size_t k = 0; k < w.idx.size(); ++k) {
            const int64_t h = lo + w.idx[k], j = w.job0 + (int64_t)k;
            j_ctg[(size_t)j] = acs[(size_t)a] + t_ctg[h], j_ts[(size_t)j] = t_start[h], j_te[(size_t)j] = t_end[h], j_gene[(size_t)j] = gene[h];
            j_strand[(size_t)j] = strand[h];
            j_frame[(size_t)j] = (int8_t)(((-(int64_t)q_start[h]) % 3 + 3) % 3);  // GeneHits.frames: (-q_starts) % 3 (Python modulo)
        }
    });
    const double tm2 = now_s();
    if (int rc = kb_post_type_numerics(*bv, device, j_ctg.data(), j_ts.data(), j_te.data(), j_strand.data(), j_frame.data(), j_gene.data(), n_jobs, d->d_trans,
                                       d->d_trans_off, d->trans_len.data(), 20, 11, 1, prot_len.data(), res.data()))
        return tfail(rc, "typing numerics failed on the device");

    // ------------------------------------------------------------ pass 2: gene states, confidence, problems
    const double tm3 = now_s();
    kb_typed *R = new kb_typed();
    R->n_asm = n_asm;
    R->score.resize((size_t)n_asm), R->completeness.resize((size_t)n_asm), R->pcov.resize((size_t)n_asm), R->length_discrepancy.resize((size_t)n_asm);
    R->typeable.resize((size_t)n_asm), R->problems.resize((size_t)n_asm), R->n_pieces.resize((size_t)n_asm);
    R->gh_off.assign((size_t)n_asm + 1, 0), R->piece_off.assign((size_t)n_asm + 1, 0), R->miss_off.assign((size_t)n_asm + 1, 0);
    std::vector<std::vector<int32_t>> keep((size_t)n_asm);  // positions in W[a].idx that survive the spurious-hit filter
    std::vector<std::vector<int8_t>> states((size_t)n_asm);
    std::vector<std::vector<float>> idents((size_t)n_asm), covs((size_t)n_asm);
    parallel_for(n_asm, n_threads, [&](int64_t a) {
        const AsmWork &w = W[(size_t)a];
        const int64_t lo = seg[(size_t)a], m = (int64_t)w.idx.size();
        bool typeable = !(w.completeness < min_completeness);
        int64_t unexpected = 0;
        bool novel_inside = false, trunc_inside = false, unexp_inside = false, exp_outside = false;
        for (int64_t k = 0; k < m; ++k) {
            const int64_t h = lo + w.idx[(size_t)k], j = w.job0 + k;
            const int32_t g = gene[h], ql = d->gene_len[(size_t)g];
            // is_partial(edge_tolerance): the hit hangs over a contig edge
            const bool fwd = strand[h] == 1;
            const bool pl = t_start[h] <= partial_edge_tolerance && (fwd ? q_start[h] > 0 : q_end[h] < ql);
            const bool pr = t_end[h] >= t_len[h] - partial_edge_tolerance && (fwd ? q_end[h] < ql : q_start[h] > 0);
            const bool partial = pl || pr;
            const double prot_cov = ((double)prot_len[(size_t)j] * 3.0) / (double)ql;
            int8_t st = 0;
            if (partial) st = 1;
            if (!partial && prot_cov < 0.90) st = 2;
            const int32_t *r = &res[(size_t)j * 8];
            const int64_t tot = (int64_t)r[1] + r[2] + r[3];
            const float ident = (float)(tot > 0 ? ((double)r[1] * 100.0) / (double)tot : 0.0);
            // (float32 array) < (Python float): numpy compares in float32
            const bool below = ident < (float)d->id_threshold;
            if (!w.inside[(size_t)k] && below) continue;  // spurious homology outside the locus: dropped
            if (st == 0 && below) st = 3;
            keep[(size_t)a].push_back((int32_t)k), states[(size_t)a].push_back(st), idents[(size_t)a].push_back(ident);
            covs[(size_t)a].push_back((float)std::min(100.0, std::max(0.0, prot_cov * 100.0)));
            const bool in = w.inside[(size_t)k], ex = w.expected[(size_t)k], xt = w.extra[(size_t)k];
            if (in && !ex && !xt) {
                unexp_inside = true;
                if (st != 2) ++unexpected;
            }
            if (in && st == 3) novel_inside = true;
            if (in && (st == 2 || st == 1)) trunc_inside = true;
            if (!in && ex) exp_outside = true;
        }
        if (unexpected > max_other_genes) typeable = false;
        if (!allow_below_threshold && novel_inside) typeable = false;
        uint8_t p = 0;
        if (w.piece_start.size() > 1) p |= 1;
        if (unexp_inside) p |= 2;
        if (w.completeness < 1.0 || exp_outside) p |= 4;
        if (novel_inside) p |= 8;
        if (trunc_inside) p |= 16;
        R->score[(size_t)a] = w.score, R->completeness[(size_t)a] = w.completeness, R->pcov[(size_t)a] = w.pcov, R->length_discrepancy[(size_t)a] = w.ldisc;
        R->typeable[(size_t)a] = typeable, R->problems[(size_t)a] = p, R->n_pieces[(size_t)a] = (int32_t)w.piece_start.size();
    });
    for (int32_t a = 0; a < n_asm; ++a) {
        R->gh_off[(size_t)a + 1] = R->gh_off[(size_t)a] + (int64_t)keep[(size_t)a].size();
        R->piece_off[(size_t)a + 1] = R->piece_off[(size_t)a] + (int64_t)W[(size_t)a].piece_start.size();
        R->miss_off[(size_t)a + 1] = R->miss_off[(size_t)a] + (int64_t)W[(size_t)a].missing.size();
    }
    const size_t ng = (size_t)R->gh_off[(size_t)n_asm], np_ = (size_t)R->piece_off[(size_t)n_asm], nm = (size_t)R->miss_off[(size_t)n_asm];
    R->gene.resize(ng), R->q_start.resize(ng), R->q_end.resize(ng), R->t_ctg.resize(ng), R->t_start.resize(ng), R->t_end.resize(ng);
    R->strand.resize(ng), R->state.resize(ng), R->is_expected.resize(ng), R->is_inside.resize(ng), R->is_extra.resize(ng);
    R->prot_ident.resize(ng), R->coverage.resize(ng);
    R->piece_ctg.resize(np_), R->piece_start.resize(np_), R->piece_end.resize(np_), R->piece_strand.resize(np_), R->missing.resize(nm);
    parallel_for(n_asm, n_threads, [&](int64_t a) {
        const AsmWork &w = W[(size_t)a];
        const int64_t lo = seg[(size_t)a];
        size_t o = (size_t)R->gh_off[(size_t)a];
        for (size_t i = 0; i < keep[(size_t)a].size(); ++i, ++o) {
            const int32_t k = keep[(size_t)a][i];
            const int64_t h = lo + w.idx[(size_t)k];
            R->gene[o] = gene[h], R->q_start[o] = q_start[h], R->q_end[o] = q_end[h], R->t_ctg[o] = t_ctg[h], R->t_start[o] = t_start[h], R->t_end[o] = t_end[h];
            R->strand[o] = strand[h], R->state[o] = states[(size_t)a][i], R->is_expected[o] = w.expected[(size_t)k], R->is_inside[o] = w.inside[(size_t)k];
            R->is_extra[o] = w.extra[(size_t)k], R->prot_ident[o] = idents[(size_t)a][i], R->coverage[o] = covs[(size_t)a][i];
        }
        size_t po = (size_t)R->piece_off[(size_t)a];
        for (size_t i = 0; i < w.piece_start.size(); ++i, ++po)
            R->piece_ctg[po] = w.piece_ctg[i], R->piece_start[po] = w.piece_start[i], R->piece_end[po] = w.piece_end[i], R->piece_strand[po] = w.piece_strand[i];
        std::copy(w.missing.begin(), w.missing.end(), R->missing.begin() + R->miss_off[(size_t)a]);
    });
    g_type_times[0] = tm1 - tm0, g_type_times[1] = tm2 - tm1, g_type_times[2] = tm3 - tm2, g_type_times[3] = now_s() - tm3;
    *out = R;
    return KB_OK;
}
// diagnostic: seconds spent by the last kb_type_call in pass 1 / job list / device numerics / pass 2
void kb_type_debug_times(double *t4) { memcpy(t4, g_type_times, sizeof(g_type_times)); }

void kb_typed_destroy(kb_typed *r) { delete r; }
int kb_typed_sizes(const kb_typed *r, int64_t *n_gene_hits, int64_t *n_pieces, int64_t *n_missing)
{
    if (!r) return tfail(KB_ERR_ARG, "null result");
    if (n_gene_hits) *n_gene_hits = r->gh_off[(size_t)r->n_asm];
    if (n_pieces) *n_pieces = r->piece_off[(size_t)r->n_asm];
    if (n_missing) *n_missing = r->miss_off[(size_t)r->n_asm];
    return KB_OK;
}
// per assembly: n_asm entries each (offset arrays n_asm + 1)
int kb_typed_fetch_assemblies(const kb_typed *r, double *score, double *completeness, double *pcov, double *length_discrepancy, uint8_t *typeable,
                              uint8_t *problems, int32_t *n_pieces, int64_t *gene_hit_off, int64_t *piece_off, int64_t *missing_off)
{
    if (!r) return tfail(KB_ERR_ARG, "null result");
    const size_t n = (size_t)r->n_asm;
#define CP(dst, src, cnt) \
    if (dst) memcpy(dst, src.data(), (cnt) * sizeof(src[0]))
    CP(score, r->score, n); CP(completeness, r->completeness, n); CP(pcov, r->pcov, n); CP(length_discrepancy, r->length_discrepancy, n);
    CP(typeable, r->typeable, n); CP(problems, r->problems, n); CP(n_pieces, r->n_pieces, n);
    CP(gene_hit_off, r->gh_off, n + 1); CP(piece_off, r->piece_off, n + 1); CP(missing_off, r->miss_off, n + 1);
    return KB_OK;
}
int kb_typed_fetch_gene_hits(const kb_typed *r, int32_t *gene, int32_t *q_start, int32_t *q_end, int32_t *t_ctg, int32_t *t_start, int32_t *t_end,
                             int8_t *strand, int8_t *state, uint8_t *is_expected, uint8_t *is_inside, uint8_t *is_extra, float *prot_ident,
                             float *coverage)
{
    if (!r) return tfail(KB_ERR_ARG, "null result");
    const size_t n = r->gene.size();
    CP(gene, r->gene, n); CP(q_start, r->q_start, n); CP(q_end, r->q_end, n); CP(t_ctg, r->t_ctg, n); CP(t_start, r->t_start, n); CP(t_end, r->t_end, n);
    CP(strand, r->strand, n); CP(state, r->state, n); CP(is_expected, r->is_expected, n); CP(is_inside, r->is_inside, n); CP(is_extra, r->is_extra, n);
    CP(prot_ident, r->prot_ident, n); CP(coverage, r->coverage, n);
    return KB_OK;
}
int kb_typed_fetch_pieces(const kb_typed *r, int32_t *ctg, int32_t *start, int32_t *end, int8_t *strand, int32_t *missing)
{
    if (!r) return tfail(KB_ERR_ARG, "null result");
    const size_t n = r->piece_ctg.size();
    CP(ctg, r->piece_ctg, n); CP(start, r->piece_start, n); CP(end, r->piece_end, n); CP(strand, r->piece_strand, n);
    CP(missing, r->missing, r->missing.size());
#undef CP
    return KB_OK;
}

}  // extern "C"
