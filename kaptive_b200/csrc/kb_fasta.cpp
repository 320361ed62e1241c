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

// ---------------------------------------------------------------------------------------------------------------
// Packed ingest (SURVEY.md section 8f, row 3 as specified): FASTA bytes -> 2 bit per base + ambiguity mask, written by the host
// threads straight into (pinned) buffers in the layout the device batch uses, so that 0.375 B per base cross PCIe instead of 1 B
// and the device-side pack kernel disappears from the host-buffer path.  Replaces core/genome.py:45 + core/seq.py:307-325.
//
// Layout (kb_packed_layout, the rule of KbHostBatchLayout::build): storage is counted in bases; 128 padded bases in front, every
// contig starts on a multiple of 128 bases (32 B of sequence words, 16 B of mask words), 128 padded bases behind the last one.
// seq2 word k holds bases 16k .. 16k + 15 (base i in bits 2i, 2i + 1; A C G T = 0 1 2 3), nmask word k holds bases 32k .. 32k + 31
// (bit set = ambiguous or padding).
#include <immintrin.h>

extern "C" int kb_packed_layout(const int32_t *contig_len, int64_t n_contigs, int64_t *contig_soff, int64_t *storage_bases)
{
    if (n_contigs < 0 || (n_contigs > 0 && (!contig_len || !contig_soff)) || !storage_bases) return KB_ERR_ARG;
    int64_t soff = 128;
    for (int64_t c = 0; c < n_contigs; ++c) {
        if (contig_len[c] < 0) return KB_ERR_ARG;
        contig_soff[c] = soff;
        soff += ((int64_t)contig_len[c] + 127) & ~(int64_t)127;
    }
    *storage_bases = soff + 128;
    return KB_OK;
}

// records of one FASTA buffer = '>' at the start of a line; memchr for the rare '>' instead of one call per 80-column line
static int64_t kb_fasta_count_records(const uint8_t *data, int64_t n)
{
    int64_t rec = 0, i = 0;
    while (i < n) {
        const uint8_t *g = (const uint8_t *)memchr(data + i, '>', (size_t)(n - i));
        if (!g) break;
        const int64_t p = (int64_t)(g - data);
        if (p == 0 || data[p - 1] == '\n') ++rec;
        i = p + 1;
    }
    return rec;
}
extern "C" int kb_fasta_ingest_count_records(const uint8_t *const *data, const int64_t *n, int32_t n_files, int32_t n_threads, int64_t *n_records)
{
    if (n_files < 0 || (n_files > 0 && (!data || !n || !n_records))) return KB_ERR_ARG;
    kb_parallel_files(n_files, n_threads, [&](int32_t i) { n_records[i] = kb_fasta_count_records(data[i], n[i]); });
    return KB_OK;
}

// record lengths and names of one FASTA buffer (no copy): the pass between counting and packing
static int kb_fasta_lengths(const uint8_t *data, int64_t n, int64_t max_records, int64_t *name_off, int32_t *name_len, int32_t *seq_len)
{
    int64_t rec = -1, i = 0;
    while (i < n) {
        const uint8_t *nl = (const uint8_t *)memchr(data + i, '\n', (size_t)(n - i));
        int64_t e = nl ? (int64_t)(nl - data) : n;
        if (data[i] == '>') {
            if (++rec >= max_records) return KB_ERR_CAPACITY;
            int64_t s = i + 1, t = s;
            while (t < e && data[t] != ' ' && data[t] != '\t' && data[t] != '\r') ++t;
            name_off[rec] = s, name_len[rec] = (int32_t)(t - s), seq_len[rec] = 0;
        } else if (rec >= 0) {
            int64_t len = e - i;
            while (len > 0 && (data[i + len - 1] == '\r' || data[i + len - 1] == ' ' || data[i + len - 1] == '\t')) --len;
            if ((int64_t)seq_len[rec] + len > 0x7fffffff) return KB_ERR_LIMIT;
            seq_len[rec] += (int32_t)len;
        }
        i = e + 1;
    }
    return KB_OK;
}

// appends `n` <= 32 bases (2-bit codes in `codes`, bit i of `valid` set where base i is A/C/G/T/U) at base position b
struct KbBitSink {
    uint64_t *seq, *msk;  // storage of the whole batch viewed as 64-bit words (little endian: consistent with the 32-bit word layout)
    inline void put(int64_t b, uint64_t codes, uint32_t valid, int n)
    {
        if (n < 32) {
            const uint32_t keep = (1u << n) - 1u;
            valid &= keep;
            codes &= (n < 32) ? ((1ull << (2 * n)) - 1ull) : ~0ull;
        }
        const int ss = (int)((2 * b) & 63);
        uint64_t *ps = seq + ((2 * b) >> 6);
        ps[0] |= codes << ss;
        if (ss + 2 * n > 64) ps[1] |= codes >> (64 - ss);  // only when bits really spill: the next word may be another thread's contig
        const int ms = (int)(b & 63);
        uint64_t *pm = msk + (b >> 6);
        pm[0] &= ~((uint64_t)valid << ms);
        if (ms + n > 64) pm[1] &= ~((uint64_t)valid >> (64 - ms));
    }
};

// 32 ASCII bases -> 64 bits of codes (invalid bases 0) + validity bits; generic form
static inline void kb_codes32_generic(const uint8_t *s, int n, uint64_t &codes, uint32_t &valid)
{
    codes = 0, valid = 0;
    for (int i = 0; i < n; ++i) {
        const uint32_t c = kb_nt4(s[i]);
        if (c < 4) codes |= (uint64_t)c << (2 * i), valid |= 1u << i;
    }
}
__attribute__((target("avx2,bmi2"))) static inline void kb_codes32_avx2(const uint8_t *s, uint64_t &codes, uint32_t &valid)
{
    const __m256i x = _mm256_loadu_si256((const __m256i *)s);
    const __m256i u = _mm256_and_si256(x, _mm256_set1_epi8((char)0xdf));
    __m256i v = _mm256_cmpeq_epi8(u, _mm256_set1_epi8('A'));
    v = _mm256_or_si256(v, _mm256_cmpeq_epi8(u, _mm256_set1_epi8('C')));
    v = _mm256_or_si256(v, _mm256_cmpeq_epi8(u, _mm256_set1_epi8('G')));
    v = _mm256_or_si256(v, _mm256_cmpeq_epi8(u, _mm256_set1_epi8('T')));
    v = _mm256_or_si256(v, _mm256_cmpeq_epi8(u, _mm256_set1_epi8('U')));
    valid = (uint32_t)_mm256_movemask_epi8(v);
    // A C G T/U -> 0 1 2 3: ((c >> 1) ^ (c >> 2)) & 3
    const __m256i c1 = _mm256_srli_epi16(x, 1), c2 = _mm256_srli_epi16(x, 2);
    const __m256i cd = _mm256_and_si256(_mm256_and_si256(_mm256_xor_si256(c1, c2), _mm256_set1_epi8(3)), v);
    alignas(32) uint64_t q[4];
    _mm256_store_si256((__m256i *)q, cd);
    const uint64_t m = 0x0303030303030303ull;
    codes = _pext_u64(q[0], m) | _pext_u64(q[1], m) << 16 | _pext_u64(q[2], m) << 32 | _pext_u64(q[3], m) << 48;
}

// 32 bytes -> codes / validity as above plus the positions of '\n' in the vector
__attribute__((target("avx2,bmi2"))) static inline void kb_codes32_nl_avx2(const uint8_t *s, uint64_t &codes, uint32_t &valid, uint32_t &nl)
{
    const __m256i x = _mm256_loadu_si256((const __m256i *)s);
    nl = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, _mm256_set1_epi8('\n')));
    const __m256i u = _mm256_and_si256(x, _mm256_set1_epi8((char)0xdf));
    __m256i v = _mm256_cmpeq_epi8(u, _mm256_set1_epi8('A'));
    v = _mm256_or_si256(v, _mm256_cmpeq_epi8(u, _mm256_set1_epi8('C')));
    v = _mm256_or_si256(v, _mm256_cmpeq_epi8(u, _mm256_set1_epi8('G')));
    v = _mm256_or_si256(v, _mm256_cmpeq_epi8(u, _mm256_set1_epi8('T')));
    v = _mm256_or_si256(v, _mm256_cmpeq_epi8(u, _mm256_set1_epi8('U')));
    valid = (uint32_t)_mm256_movemask_epi8(v);
    const __m256i c1 = _mm256_srli_epi16(x, 1), c2 = _mm256_srli_epi16(x, 2);
    const __m256i cd = _mm256_and_si256(_mm256_and_si256(_mm256_xor_si256(c1, c2), _mm256_set1_epi8(3)), v);
    alignas(32) uint64_t q[4];
    _mm256_store_si256((__m256i *)q, cd);
    const uint64_t m = 0x0303030303030303ull;
    codes = _pext_u64(q[0], m) | _pext_u64(q[1], m) << 16 | _pext_u64(q[2], m) << 32 | _pext_u64(q[3], m) << 48;
}

// portable form: line by line
static int kb_fasta_pack_one_generic(const uint8_t *data, int64_t n, int64_t max_records, const int64_t *soff, KbBitSink sink)
{
    int64_t rec = -1, i = 0, b = 0;
    while (i < n) {
        const uint8_t *nl = (const uint8_t *)memchr(data + i, '\n', (size_t)(n - i));
        int64_t e = nl ? (int64_t)(nl - data) : n;
        if (data[i] == '>') {
            if (++rec >= max_records) return KB_ERR_CAPACITY;
            b = soff[rec];
        } else if (rec >= 0) {
            int64_t len = e - i;
            while (len > 0 && (data[i + len - 1] == '\r' || data[i + len - 1] == ' ' || data[i + len - 1] == '\t')) --len;
            const uint8_t *s = data + i;
            uint64_t codes;
            uint32_t valid;
            for (int64_t k = 0; k < len; k += 32) {
                const int m = len - k < 32 ? (int)(len - k) : 32;
                kb_codes32_generic(s + k, m, codes, valid);
                sink.put(b + k, codes, valid, m);
            }
            b += len;
        }
        i = e + 1;
    }
    return KB_OK;
}
// AVX2 + BMI2 form: the vector that yields the codes also finds the end of the line (no memchr per 80-column line)
__attribute__((target("avx2,bmi2"))) static int kb_fasta_pack_one_fast(const uint8_t *data, int64_t n, int64_t max_records, const int64_t *soff,
                                                                      KbBitSink sink)
{
    int64_t rec = -1, i = 0, b = 0;
    while (i < n) {
        if (data[i] == '>') {  // header line
            const uint8_t *nl = (const uint8_t *)memchr(data + i, '\n', (size_t)(n - i));
            if (++rec >= max_records) return KB_ERR_CAPACITY;
            b = soff[rec];
            i = nl ? (int64_t)(nl - data) + 1 : n;
            continue;
        }
        if (rec < 0) {  // text before the first header is ignored
            const uint8_t *nl = (const uint8_t *)memchr(data + i, '\n', (size_t)(n - i));
            i = nl ? (int64_t)(nl - data) + 1 : n;
            continue;
        }
        // one sequence line starting at i
        for (;;) {
            uint64_t codes;
            uint32_t valid, nlm;
            int64_t len;  // bases of this vector that belong to the line
            bool eol;
            if (i + 32 <= n) {
                kb_codes32_nl_avx2(data + i, codes, valid, nlm);
                eol = nlm != 0;
                len = eol ? (int64_t)__builtin_ctz(nlm) : 32;
            } else {  // the last < 32 bytes of the buffer
                const uint8_t *nl = (const uint8_t *)memchr(data + i, '\n', (size_t)(n - i));
                len = nl ? (int64_t)(nl - (data + i)) : n - i;
                eol = true;
                kb_codes32_generic(data + i, (int)len, codes, valid);
            }
            int64_t keep = len;
            if (eol)  // trailing blanks of the line are not sequence
                while (keep > 0 && (data[i + keep - 1] == '\r' || data[i + keep - 1] == ' ' || data[i + keep - 1] == '\t')) --keep;
            if (keep > 0) sink.put(b, codes, valid, (int)keep);
            b += keep;
            i += eol ? len + 1 : len;
            if (eol || i >= n) break;
        }
    }
    return KB_OK;
}

// pass 2 of the packed ingest: per-record lengths and names (rec_base from kb_fasta_ingest_count's counts)
extern "C" int kb_fasta_ingest_lengths(const uint8_t *const *data, const int64_t *n, int32_t n_files, int32_t n_threads, const int64_t *rec_base,
                                       int32_t *contig_len, int32_t *asm_contig_start, int64_t *name_off, int32_t *name_len)
{
    if (n_files < 0 || !rec_base || !asm_contig_start || (n_files > 0 && (!data || !n || !contig_len || !name_off || !name_len))) return KB_ERR_ARG;
    if (rec_base[n_files] > 0x7fffffff) return KB_ERR_LIMIT;
    std::atomic<int> rc{KB_OK};
    kb_parallel_files(n_files, n_threads, [&](int32_t i) {
        const int64_t r0 = rec_base[i], nr = rec_base[i + 1] - r0;
        int r = kb_fasta_lengths(data[i], n[i], nr, name_off + r0, name_len + r0, contig_len + r0);
        if (r != KB_OK) rc = r;
        asm_contig_start[i] = (int32_t)r0;
    });
    asm_contig_start[n_files] = (int32_t)rec_base[n_files];
    return rc;
}

// pass 3: pack.  contig_soff / storage_bases from kb_packed_layout over ALL contigs of the call; seq2: storage_bases / 16 words,
// nmask: storage_bases / 32 words (both may be pinned memory).  Every word of both arrays is written (padding included).
// use_simd: 1 = AVX2 + BMI2 when the CPU has them, 0 = portable loop (the two are bit-identical; the flag exists for the tests).
extern "C" int kb_fasta_ingest_pack(const uint8_t *const *data, const int64_t *n, int32_t n_files, int32_t n_threads, const int64_t *rec_base,
                                    const int32_t *contig_len, const int64_t *contig_soff, int64_t storage_bases, uint32_t *seq2, uint32_t *nmask,
                                    int32_t use_simd)
{
    if (n_files < 0 || !rec_base || !seq2 || !nmask || storage_bases < 256 || (storage_bases & 127) ||
        (n_files > 0 && (!data || !n || (rec_base[n_files] > 0 && (!contig_len || !contig_soff)))))
        return KB_ERR_ARG;
    const bool fast = use_simd && __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
    const int64_t n_ctg = rec_base[n_files];
    // every file's thread first clears the storage of its own contigs (lead-in of the first file and tail of the last included):
    // sequence words 0, mask words all ones
    auto clear = [&](int64_t b0, int64_t b1) {
        memset(seq2 + (b0 >> 4), 0, (size_t)((b1 - b0) >> 4) * 4);
        memset(nmask + (b0 >> 5), 0xff, (size_t)((b1 - b0) >> 5) * 4);
    };
    std::atomic<int> rc{KB_OK};
    KbBitSink sink{reinterpret_cast<uint64_t *>(seq2), reinterpret_cast<uint64_t *>(nmask)};
    kb_parallel_files(n_files, n_threads, [&](int32_t i) {
        const int64_t r0 = rec_base[i], r1 = rec_base[i + 1];
        // storage range owned by this file: from its first contig (or the very start) to the next file's first contig (or the very end)
        int64_t b0 = r0 < n_ctg ? contig_soff[r0] : storage_bases - 128, b1 = r1 < n_ctg ? contig_soff[r1] : storage_bases;
        if (i == 0) b0 = 0;
        if (i == n_files - 1) b1 = storage_bases;
        if (r0 == r1 && i != 0 && i != n_files - 1) b0 = b1;  // an empty file in the middle owns nothing
        if (b1 > b0) clear(b0, b1);
        int r = fast ? kb_fasta_pack_one_fast(data[i], n[i], r1 - r0, contig_soff + r0, sink)
                     : kb_fasta_pack_one_generic(data[i], n[i], r1 - r0, contig_soff + r0, sink);
        if (r != KB_OK) rc = r;
    });
    if (n_files == 0) clear(0, storage_bases);
    return rc;
}
