// kb_kernels.h -- launch prototypes shared between the .cu translation units.
#pragma once
#include "kb_common.cuh"
#ifdef __CUDACC__
#include <cuda_runtime.h>

// counters: [0] minimizers, [1] anchors, [2] groups, [3] chains, [4] raw hits, [5] cigar words, [6] error bits, [7] debug dump count, [8] DP cells
#define KB_N_COUNTERS 48  // [0..15] main pipeline, [16..31] auxiliary scans (census, dumps), [32..] staged-alignment queue cursors

// job lists of the staged alignment path (device memory, job_cap entries each)
struct KbStageLists {
    int32_t *band_list, *rows_list, *r16_list, *r16_list2;  // 32-bit band kernel, 32-bit rows kernel, packed rows kernel (queued / sorted)
    uint32_t *r16_key, *r16_key2;
    int32_t *b16_list[3];                                    // packed band kernel, window width K = 1, 2, 4
    uint32_t *b16_key[3];
    void *sort_tmp;
    size_t sort_tmp_bytes;
    int64_t job_cap;
};

void kb_launch_scan(const KbIndexView &ix, const KbBatchView &bt, uint64_t *akey, uint32_t *aval,
                    unsigned long long *counters, int64_t anchor_cap, uint32_t *mz_hash, int32_t *mz_ctg,
                    uint32_t *mz_pos, int64_t mz_cap, int32_t mz_asm, int n_sm, cudaStream_t st);
#endif
