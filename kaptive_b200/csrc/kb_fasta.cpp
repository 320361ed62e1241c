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
