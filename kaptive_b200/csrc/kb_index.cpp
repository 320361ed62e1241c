// kb_index.cpp -- host-side construction of the gene index and of the batch layout.
//
// The gene index is the query side of the reference's map_batch call
// (src/kaptive/serotyping/core.py:111-121,154): every DB gene is sketched once
// (minimap2 mm_sketch, k=15 w=10) and its minimizers are put in an open-addressing
// hash keyed by the 30-bit minimizer hash.  The scan kernel streams assemblies
// past this table instead of rebuilding a 5 Mb index per assembly.
#include <algorithm>
#include <cstring>
#include "kb_host.h"
#include "kb_scan.cuh"

namespace {
struct HostFetch {
    const uint8_t *s;
    int operator()(int i) const { return s[i]; }
};
struct HostEmit {
    std::vector<uint32_t> *x, *y;
    void operator()(uint32_t hx, uint32_t hy) { x->push_back(hx), y->push_back(hy); }
};
uint32_t x31_hash_decimal(int g)
{
    char tmp[16], name[16];
    int n = 0, x = g;
    if (x == 0) tmp[n++] = '0';
    while (x > 0) tmp[n++] = (char)('0' + x % 10), x /= 10;
    for (int i = 0; i < n; ++i) name[i] = tmp[n - 1 - i];
    name[n] = 0;
    uint32_t h = (uint32_t)name[0];
    for (int i = 1; i < n; ++i) h = (h << 5) - h + (uint32_t)name[i];
    return h;
}
template <class T>
void put(uint8_t *&p, const std::vector<T> &v)
{
    int64_t n = (int64_t)v.size();
    memcpy(p, &n, 8), p += 8;
    if (n) memcpy(p, v.data(), (size_t)n * sizeof(T));
    p += ((size_t)n * sizeof(T) + 7) & ~(size_t)7;
}
template <class T>
bool get(const uint8_t *&p, const uint8_t *end, std::vector<T> &v)
{
    if (p + 8 > end) return false;
    int64_t n;
    memcpy(&n, p, 8), p += 8;
    size_t bytes = ((size_t)n * sizeof(T) + 7) & ~(size_t)7;
    if (n < 0 || p + bytes > end) return false;
    v.resize((size_t)n);
    if (n) memcpy(v.data(), p, (size_t)n * sizeof(T));
    p += bytes;
    return true;
}
template <class T>
int64_t sz(const std::vector<T> &v)
{
    return 8 + (int64_t)(((size_t)v.size() * sizeof(T) + 7) & ~(size_t)7);
}
}  // namespace

std::string KbHostIndex::build(const uint8_t *seqs, const int64_t *off, const int32_t *len, int32_t n, const kb_params_t &pp)
{
    p = pp;
    if (p.k != 15 || p.w != 10) return "only k=15, w=10 (minimap2 defaults, Aligner(preset=None)) are compiled in";
    if (n < 0 || n > KB_MAX_GENES) return "too many genes (limit 32768)";
    n_genes = n;
    gene_len.assign(len, len + n);
    gene_nmin.resize(n), gene_min_off.resize(n), gene_seq_off.resize(n), gene_hash.resize(n);
    int64_t total = 0;
    for (int g = 0; g < n; ++g) {
        if (len[g] < 0 || len[g] >= (1 << 24)) return "gene length out of range";
        gene_seq_off[g] = total, total += len[g];
    }
    gseq_fwd.resize((size_t)total + 16), gseq_rev.resize((size_t)total + 16);
    gm_qpos_z.clear(), gm_qocc.clear(), gm_tandem.clear(), gm_hash.clear();
    for (int g = 0; g < n; ++g) {
        const uint8_t *s = seqs + off[g];
        uint8_t *f = gseq_fwd.data() + gene_seq_off[g], *r = gseq_rev.data() + gene_seq_off[g];
        int L = len[g];
        for (int i = 0; i < L; ++i) {
            f[i] = kb_nt4(s[i]);
            r[L - 1 - i] = f[i] < 4 ? (uint8_t)(3 - f[i]) : (uint8_t)4;
        }
        std::vector<uint32_t> hx, hy;
        HostFetch fetch{f};
        HostEmit emit{&hx, &hy};
        if (L > 0) kb_sketch_slice<10, 15>(L, 0, L, fetch, emit);
        int m = (int)hx.size();
        gene_min_off[g] = (int64_t)gm_hash.size();
        gene_nmin[g] = m;
        std::vector<int> order(m);
        for (int i = 0; i < m; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int a, int b) { return hx[a] != hx[b] ? hx[a] < hx[b] : a < b; });
        std::vector<int32_t> qocc(m);
        for (int st = 0, i = 1; i <= m; ++i)
            if (i == m || hx[order[i]] != hx[order[st]]) {
                for (int j = st; j < i; ++j) qocc[order[j]] = i - st;
                st = i;
            }
        for (int i = 0; i < m; ++i) {
            uint8_t tan = 0;
            if (i > 0 && hx[i] == hx[i - 1]) tan = 1;
            if (i < m - 1 && hx[i] == hx[i + 1]) tan = 1;
            gm_hash.push_back(hx[i]), gm_qpos_z.push_back(hy[i]), gm_qocc.push_back(qocc[i]), gm_tandem.push_back(tan);
            if (qocc[i] > max_qocc) max_qocc = qocc[i];
        }
        uint32_t h = x31_hash_decimal(g);  // the reference names query i str(i) (serotyping/core.py:113)
        h ^= kb_wang_hash((uint32_t)L) + kb_wang_hash((uint32_t)p.seed);
        gene_hash[g] = kb_wang_hash(h);
    }
    // entries sorted by (hash, gene, ordinal)
    int64_t ne = (int64_t)gm_hash.size();
    if (ne > (int64_t)KB_HT_START_MASK) return "too many gene minimizers (limit 8M)";
    std::vector<int64_t> ord((size_t)ne);
    std::vector<int32_t> gene_of((size_t)ne);
    for (int g = 0; g < n; ++g)
        for (int i = 0; i < gene_nmin[g]; ++i) gene_of[(size_t)(gene_min_off[g] + i)] = g;
    for (int64_t i = 0; i < ne; ++i) ord[(size_t)i] = i;
    std::sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b) { return gm_hash[(size_t)a] != gm_hash[(size_t)b] ? gm_hash[(size_t)a] < gm_hash[(size_t)b] : a < b; });
    ent.resize((size_t)ne);
    int64_t n_distinct = 0;
    for (int64_t i = 0; i < ne; ++i) {
        int64_t s = ord[(size_t)i];
        int g = gene_of[(size_t)s];
        KbEntry &e = ent[(size_t)i];
        e.gene = g;
        e.qpos_z = gm_qpos_z[(size_t)s];
        e.mi_flags = (uint32_t)(s - gene_min_off[g]) | ((uint32_t)gm_tandem[(size_t)s] << 31);
        e.qocc = gm_qocc[(size_t)s];
        if (i == 0 || gm_hash[(size_t)s] != gm_hash[(size_t)ord[(size_t)i - 1]]) ++n_distinct;
    }
    uint32_t slots = 1024;
    while ((int64_t)slots < n_distinct * 2) slots <<= 1;
    ht.assign(slots, 0);
    for (int64_t st = 0, i = 1; i <= ne; ++i)
        if (i == ne || gm_hash[(size_t)ord[(size_t)i]] != gm_hash[(size_t)ord[(size_t)st]]) {
            uint32_t h = gm_hash[(size_t)ord[(size_t)st]];
            int64_t cnt = i - st;
            if (cnt > (int64_t)KB_HT_COUNT_MASK) return "a minimizer occurs in more than 2047 gene positions";
            uint32_t slot = h & (slots - 1);
            while (ht[slot]) slot = (slot + 1) & (slots - 1);
            ht[slot] = ((uint64_t)h << KB_HT_KEY_SHIFT) | ((uint64_t)st << KB_HT_START_SHIFT) | (uint64_t)cnt;
            st = i;
        }
    return "";
}

KbIndexView KbHostIndex::host_view() const
{
    KbIndexView v;
    v.p = p, v.n_genes = n_genes, v.n_entries = (int64_t)ent.size();
    v.ht_mask = (uint32_t)ht.size() - 1, v.ht = ht.data(), v.ent = ent.data();
    v.gene_len = gene_len.data(), v.gene_nmin = gene_nmin.data(), v.gene_min_off = gene_min_off.data();
    v.gm_qpos_z = gm_qpos_z.data(), v.gm_qocc = gm_qocc.data(), v.gene_hash = gene_hash.data();
    v.gene_seq_off = gene_seq_off.data(), v.gseq_fwd = gseq_fwd.data(), v.gseq_rev = gseq_rev.data();
    v.bloom = nullptr, v.bloom_mask = 0;
    return v;
}

int64_t KbHostIndex::serialized_size() const
{
    return 16 + (int64_t)sizeof(kb_params_t) + 8 + sz(gene_len) + sz(gene_nmin) + sz(gene_min_off) + sz(gene_seq_off) +
           sz(gm_qpos_z) + sz(gm_qocc) + sz(gm_tandem) + sz(gm_hash) + sz(gene_hash) + sz(gseq_fwd) + sz(gseq_rev) + sz(ent) + sz(ht);
}

void KbHostIndex::serialize(uint8_t *buf) const
{
    uint8_t *q = buf;
    memcpy(q, "KBIDX001", 8), q += 8;
    int64_t psz = sizeof(kb_params_t);
    memcpy(q, &psz, 8), q += 8;
    memcpy(q, &p, sizeof(p)), q += sizeof(p);
    int64_t ng = n_genes | ((int64_t)max_qocc << 32);
    memcpy(q, &ng, 8), q += 8;
    put(q, gene_len), put(q, gene_nmin), put(q, gene_min_off), put(q, gene_seq_off), put(q, gm_qpos_z), put(q, gm_qocc);
    put(q, gm_tandem), put(q, gm_hash), put(q, gene_hash), put(q, gseq_fwd), put(q, gseq_rev), put(q, ent), put(q, ht);
}

std::string KbHostIndex::deserialize(const uint8_t *buf, int64_t n)
{
    const uint8_t *q = buf, *end = buf + n;
    if (n < 32 || memcmp(q, "KBIDX001", 8) != 0) return "not a serialized gene index";
    q += 8;
    int64_t psz;
    memcpy(&psz, q, 8), q += 8;
    if (psz != (int64_t)sizeof(kb_params_t) || q + psz + 8 > end) return "index image was written with a different parameter layout";
    memcpy(&p, q, sizeof(p)), q += sizeof(p);
    int64_t ng;
    memcpy(&ng, q, 8), q += 8;
    n_genes = (int32_t)(ng & 0xffffffff), max_qocc = (int32_t)(ng >> 32);
    bool ok = get(q, end, gene_len) && get(q, end, gene_nmin) && get(q, end, gene_min_off) && get(q, end, gene_seq_off) &&
              get(q, end, gm_qpos_z) && get(q, end, gm_qocc) && get(q, end, gm_tandem) && get(q, end, gm_hash) &&
              get(q, end, gene_hash) && get(q, end, gseq_fwd) && get(q, end, gseq_rev) && get(q, end, ent) && get(q, end, ht);
    if (!ok || (int32_t)gene_len.size() != n_genes) return "truncated gene index image";
    return "";
}

std::string KbHostBatchLayout::build(const int64_t *ctg_off, const int32_t *ctg_len_in, const int32_t *asm_ctg_start_in, int32_t n_asm_in)
{
    (void)ctg_off;
    if (n_asm_in < 0 || n_asm_in > KB_MAX_ASM) return "too many assemblies in one batch (limit 131072)";
    n_asm = n_asm_in;
    asm_ctg_start.assign(asm_ctg_start_in, asm_ctg_start_in + n_asm + 1);
    n_ctg = n_asm ? asm_ctg_start[n_asm] : 0;
    if (asm_ctg_start[0] != 0) return "asm_contig_start[0] must be 0";
    ctg_len.assign(ctg_len_in, ctg_len_in + n_ctg);
    ctg_soff.resize(n_ctg), ctg_asm.resize(n_ctg), ctg_vstart.resize(n_ctg);
    chunk_ctg.clear(), chunk_start.clear();
    int64_t soff = 128;  // one padded group in front so look-back loads of the first contig stay in bounds
    total_bases = 0;
    for (int a = 0; a < n_asm; ++a) {
        if (asm_ctg_start[a + 1] < asm_ctg_start[a]) return "asm_contig_start must be non-decreasing";
        int64_t v = 0;
        for (int c = asm_ctg_start[a]; c < asm_ctg_start[a + 1]; ++c) {
            int32_t L = ctg_len[c];
            if (L < 0) return "negative contig length";
            ctg_asm[c] = a, ctg_soff[c] = soff, ctg_vstart[c] = (int32_t)v;
            v += (int64_t)L + KB_CTG_VGAP;
            if (v > (int64_t)KB_VPOS_MASK) return "assembly too large for one batch entry (limit 128 Mb incl. 8 kb per contig)";
            soff += ((int64_t)L + 127) & ~(int64_t)127;  // 128 bases = 16 B of mask: keeps uint4 loads of both arrays aligned
            total_bases += L;
            for (int32_t s = 0; s < L; s += KB_CHUNK_BASES) chunk_ctg.push_back(c), chunk_start.push_back(s);
        }
    }
    storage_bases = soff + 128;
    return "";
}

void kb_pack_host(const uint8_t *ascii, const int64_t *ctg_off, const KbHostBatchLayout &L, std::vector<uint32_t> &seq2,
                  std::vector<uint32_t> &nmask)
{
    seq2.assign((size_t)(L.storage_bases >> 4), 0u);
    nmask.assign((size_t)(L.storage_bases >> 5), 0xffffffffu);
    for (int c = 0; c < L.n_ctg; ++c) {
        const uint8_t *s = ascii + ctg_off[c];
        int64_t so = L.ctg_soff[c];
        for (int32_t i = 0; i < L.ctg_len[c]; ++i) {
            int code = kb_nt4(s[i]);
            int64_t b = so + i;
            if (code < 4) {
                seq2[(size_t)(b >> 4)] |= (uint32_t)code << (2 * (b & 15));
                nmask[(size_t)(b >> 5)] &= ~(1u << (b & 31));
            }
        }
    }
}
