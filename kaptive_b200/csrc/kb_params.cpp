// kb_params.cpp -- default mapping parameters.
// minimap2 defaults with no preset (Aligner(preset=None), src/kaptive/serotyping/core.py:148);
// the reference's two overrides, best_n=50000 and pri_ratio=0.0 (core.py:150-151), mean
// "keep every chain", which is how the chain-selection stage is written, so they have no field here.
#include <climits>
#include <cstring>
#include "kb_common.cuh"

extern "C" void kb_params_default(kb_params_t *p)
{
    memset(p, 0, sizeof(*p));
    p->k = 15, p->w = 10;
    p->min_cnt = 3, p->min_chain_score = 40, p->bw = 500, p->max_gap = 5000;
    p->max_chain_skip = 25, p->max_chain_iter = 5000;
    p->chain_gap_scale = 0.8f;
    p->a = 2, p->b = 4, p->q = 4, p->e = 2, p->q2 = 24, p->e2 = 1, p->sc_ambi = 1;
    p->zdrop = 400, p->min_dp_max = 80, p->min_ksw_len = 200;
    p->mid_occ = 0, p->min_mid_occ = 10, p->max_mid_occ = 1000000;
    p->mid_occ_frac = 2e-4f, p->q_occ_frac = 0.01f, p->mask_level = 0.5f;
    p->mask_len = INT_MAX;
    p->seed = 11;
    p->ext_bw = (int)(500 * 1.5 + 1.);
    p->max_sw_cells = 100000000;  // minimap2 max_sw_mat
}
extern "C" void kbe_params_default(kb_params_t *p) { kb_params_default(p); }
