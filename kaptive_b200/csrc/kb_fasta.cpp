// kb_fasta.cpp -- FASTA ingest, the replacement for rammappy.fasta.parse_fasta_bytes
// (reference call site src/kaptive/core/genome.py:45-46: name -> SeqRecord.id (str),
// sequence -> bytes).  Record name = header up to the first whitespace; sequence lines are
// joined with CR/LF removed; bytes are otherwise passed through unchanged (case preserved).
#include <cstring>
#include <string>
#include "kb_common.cuh"

static thread_local std::string g_fasta_err;

extern "C" int kb_fasta_count(const uint8_t *data, int64_t n, int64_t *n_records, int64_t *n_seq_bytes)
{
    if ((!data && n > 0) || !n_records || !n_seq_bytes) return KB_ERR_ARG;
    int64_t rec = 0, bytes = 0, i = 0;
    bool in_rec = false;
    while (i < n) {
        const uint8_t *nl = (const uint8_t *)memchr(data + i, '\n', (size_t)(n - i));
        int64_t e = nl ? (int64_t)(nl - data) : n;
        if (data[i] == '>') ++rec, in_rec = true;
        else if (in_rec) {
            int64_t len = e - i;
            while (len > 0 && (data[i + len - 1] == '\r' || data[i + len - 1] == ' ' || data[i + len - 1] == '\t')) --len;
            bytes += len;
        }
        i = e + 1;
    }
    *n_records = rec, *n_seq_bytes = bytes;
    return KB_OK;
}

extern "C" int kb_fasta_parse(const uint8_t *data, int64_t n, int64_t max_records, int64_t *name_off, int32_t *name_len,
                              uint8_t *seq_out, int64_t seq_cap, int64_t *seq_off, int32_t *seq_len)
{
    if ((!data && n > 0) || !name_off || !name_len || !seq_off || !seq_len || (!seq_out && seq_cap > 0)) return KB_ERR_ARG;
    int64_t rec = -1, out = 0, i = 0;
    while (i < n) {
        const uint8_t *nl = (const uint8_t *)memchr(data + i, '\n', (size_t)(n - i));
        int64_t e = nl ? (int64_t)(nl - data) : n;
        if (data[i] == '>') {
            if (++rec >= max_records) return KB_ERR_CAPACITY;
            int64_t s = i + 1, t = s;
            while (t < e && data[t] != ' ' && data[t] != '\t' && data[t] != '\r') ++t;
            name_off[rec] = s, name_len[rec] = (int32_t)(t - s);
            seq_off[rec] = out, seq_len[rec] = 0;
        } else if (rec >= 0) {
            int64_t len = e - i;
            while (len > 0 && (data[i + len - 1] == '\r' || data[i + len - 1] == ' ' || data[i + len - 1] == '\t')) --len;
            if (out + len > seq_cap) return KB_ERR_CAPACITY;
            if (len > 0) memcpy(seq_out + out, data + i, (size_t)len);
            out += len;
            if ((int64_t)seq_len[rec] + len > 0x7fffffff) return KB_ERR_LIMIT;
            seq_len[rec] += (int32_t)len;
        }
        i = e + 1;
    }
    return KB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Batch ingest: many FASTA buffers, one assembly each, parsed by a pool of host threads
// straight into the host layout kb_map_assemblies / kb_batch_create take (concatenated sequences, contig offsets and
// lengths, contig range per assembly), so that at > 10^3 assemblies/s the reference's per-genome read + parse + copy
// (core/genome.py:45, core/seq.py:307-325) is not the bottleneck.  Same record semantics as kb_fasta_parse.
#include <atomic>
#include <thread>
#include <vector>

template <class F>
static void kb_parallel_files(int32_t n_files, int32_t n_threads, F f)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_files) n_threads = n_files > 0 ? n_files : 1;
    std::atomic<int32_t> next{0};
    auto worker = [&]() {
        for (int32_t i; (i = next.fetch_add(1)) < n_files;) f(i);
    };
    if (n_threads == 1) {
        worker();
        return;
    }
    std::vector<std::thread> th;
    for (int32_t t = 0; t < n_threads; ++t) th.emplace_back(worker);
    for (auto &t : th) t.join();
}

extern "C" int kb_fasta_ingest_count(const uint8_t *const *data, const int64_t *n, int32_t n_files, int32_t n_threads, int64_t *n_records,
                                     int64_t *n_seq_bytes)
{
    if (n_files < 0 || (n_files > 0 && (!data || !n || !n_records || !n_seq_bytes))) return KB_ERR_ARG;
    std::atomic<int> rc{KB_OK};
    kb_parallel_files(n_files, n_threads, [&](int32_t i) {
        int r = kb_fasta_count(data[i], n[i], &n_records[i], &n_seq_bytes[i]);
        if (r != KB_OK) rc = r;
    });
    return rc;
}

// rec_base / seq_base: exclusive prefix sums of the per-file counts (n_files + 1 entries each).  contig_off are
// offsets into seq_out; name_off index into the file's own buffer.
extern "C" int kb_fasta_ingest_parse(const uint8_t *const *data, const int64_t *n, int32_t n_files, int32_t n_threads, const int64_t *rec_base,
                                     const int64_t *seq_base, uint8_t *seq_out, int64_t *contig_off, int32_t *contig_len,
                                     int32_t *asm_contig_start, int64_t *name_off, int32_t *name_len)
{
    if (n_files < 0 || !rec_base || !seq_base || !asm_contig_start ||
        (n_files > 0 && (!data || !n || !contig_off || !contig_len || !name_off || !name_len || (!seq_out && seq_base[n_files] > 0))))
        return KB_ERR_ARG;
    if (rec_base[n_files] > 0x7fffffff) return KB_ERR_LIMIT;
    std::atomic<int> rc{KB_OK};
    kb_parallel_files(n_files, n_threads, [&](int32_t i) {
        const int64_t r0 = rec_base[i], nr = rec_base[i + 1] - r0, s0 = seq_base[i], ns = seq_base[i + 1] - s0;
        int r = kb_fasta_parse(data[i], n[i], nr, name_off + r0, name_len + r0, seq_out + s0, ns, contig_off + r0, contig_len + r0);
        if (r != KB_OK) rc = r;
        for (int64_t k = 0; k < nr; ++k) contig_off[r0 + k] += s0;  // file-local -> global offsets
        asm_contig_start[i] = (int32_t)r0;
    });
    asm_contig_start[n_files] = (int32_t)rec_base[n_files];
    return rc;
}
