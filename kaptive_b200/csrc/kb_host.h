// kb_host.h -- internal host-side objects behind the opaque C-ABI handles.
#pragma once
#include <string>
#include <vector>
#include "kb_common.cuh"

struct KbHostIndex {
    kb_params_t p;
    int32_t n_genes = 0;
    std::vector<int32_t> gene_len, gene_nmin;
    std::vector<int64_t> gene_min_off, gene_seq_off;
    std::vector<uint32_t> gm_qpos_z;
    std::vector<int32_t> gm_qocc;
    std::vector<uint8_t> gm_tandem;
    std::vector<uint32_t> gm_hash;
    std::vector<uint32_t> gene_hash;
    std::vector<uint8_t> gseq_fwd, gseq_rev;
    std::vector<KbEntry> ent;
    std::vector<uint64_t> ht;
    int32_t max_qocc = 0;
    // returns "" or an error message
    std::string build(const uint8_t *seqs, const int64_t *off, const int32_t *len, int32_t n, const kb_params_t &p);
    KbIndexView host_view() const;
    int64_t serialized_size() const;
    void serialize(uint8_t *buf) const;
    std::string deserialize(const uint8_t *buf, int64_t n);
};

struct KbHostBatchLayout {
    int32_t n_asm = 0, n_ctg = 0;
    int64_t total_bases = 0, storage_bases = 0;
    std::vector<int64_t> ctg_soff;
    std::vector<int32_t> ctg_len, ctg_asm, ctg_vstart, asm_ctg_start, chunk_ctg, chunk_start;
    std::string build(const int64_t *ctg_off, const int32_t *ctg_len_in, const int32_t *asm_ctg_start_in, int32_t n_asm_in);
};

// host packer (used by the host emulation and as the reference for the pack kernel's layout)
void kb_pack_host(const uint8_t *ascii, const int64_t *ctg_off, const KbHostBatchLayout &L, std::vector<uint32_t> &seq2,
                  std::vector<uint32_t> &nmask);
