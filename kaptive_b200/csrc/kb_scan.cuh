// kb_scan.cuh -- minimizer sketch state machine for one lane's slice of a contig.
//
// Restates minimap2's mm_sketch (the algorithm behind rammappy.Index.build,
// reference call site src/kaptive/core/genome.py:188-189) so that it can be
// run on independent slices: the sketch state after base i is a pure function
// of the last w window entries and of l = min(run of unambiguous bases, w+k),
// so a lane that starts KB_SCAN_LOOKBACK = w+k-1 bases early in "silent" mode
// reproduces the sequential state bit for bit ("Scan kernel").
//
// K must be odd (a k-mer can then never equal its reverse complement, which
// removes mm_sketch's `continue` that stalls the window); K <= 15 keeps the
// k-mers and the invertible hash in 32-bit registers.
#pragma once
#include "kb_common.cuh"

template <int W, int K>
struct KbSketchState {
    uint32_t bx[W], by[W];  // window: hash / (pos<<1|strand); KB_MAXU = empty
    uint32_t min_x, min_y;
    int min_pos;
    uint32_t fwd, rev;
    int l;
    KB_HD void reset()
    {
#pragma unroll
        for (int j = 0; j < W; ++j) bx[j] = by[j] = KB_MAXU;
        min_x = min_y = KB_MAXU;
        min_pos = 0;
        fwd = rev = 0;
        l = 0;
    }
};

// One sketch step at window slot U (compile-time), contig position i, base code c (0..3, 4 = ambiguous).
// `live` = this step belongs to the lane (i >= lane start); silent steps only rebuild state.
template <int W, int K, int U, class Emit>
KB_HD void kb_sketch_step(KbSketchState<W, K> &s, int i, int c, bool live, Emit &emit)
{
    static_assert(K & 1, "K must be odd");
    static_assert(K <= 15, "K must fit 30 bits");
    const uint32_t mask = (1u << (2 * K)) - 1u;
    const int shift1 = 2 * (K - 1);
    uint32_t ix = KB_MAXU, iy = KB_MAXU;
    if (c < 4) {
        s.fwd = ((s.fwd << 2) | (uint32_t)c) & mask;
        s.rev = (s.rev >> 2) | ((3u ^ (uint32_t)c) << shift1);
        int z = s.fwd < s.rev ? 0 : 1;
        ++s.l;
        if (s.l >= K) {
            ix = kb_hash32(z ? s.rev : s.fwd, mask);
            iy = ((uint32_t)i << 1) | (uint32_t)z;
        }
    } else s.l = 0;
    s.bx[U] = ix;
    s.by[U] = iy;
    if (s.l == W + K - 1 && s.min_x != KB_MAXU) {  // first full window: equal minima not stored yet
#pragma unroll
        for (int j = U + 1; j < W; ++j)
            if (s.min_x == s.bx[j] && s.by[j] != s.min_y && live) emit(s.bx[j], s.by[j]);
#pragma unroll
        for (int j = 0; j < U; ++j)
            if (s.min_x == s.bx[j] && s.by[j] != s.min_y && live) emit(s.bx[j], s.by[j]);
    }
    if (ix <= s.min_x) {  // new minimum: write the old one
        if (s.l >= W + K && s.min_x != KB_MAXU && live) emit(s.min_x, s.min_y);
        s.min_x = ix, s.min_y = iy, s.min_pos = U;
    } else if (s.min_pos == U) {  // old minimum left the window
        if (s.l >= W + K - 1 && s.min_x != KB_MAXU && live) emit(s.min_x, s.min_y);
        s.min_x = KB_MAXU;
#pragma unroll
        for (int j = U + 1; j < W; ++j)
            if (s.min_x >= s.bx[j]) s.min_x = s.bx[j], s.min_y = s.by[j], s.min_pos = j;
#pragma unroll
        for (int j = 0; j <= U; ++j)
            if (s.min_x >= s.bx[j]) s.min_x = s.bx[j], s.min_y = s.by[j], s.min_pos = j;
        if (s.l >= W + K - 1 && s.min_x != KB_MAXU) {
#pragma unroll
            for (int j = U + 1; j < W; ++j)
                if (s.min_x == s.bx[j] && s.min_y != s.by[j] && live) emit(s.bx[j], s.by[j]);
#pragma unroll
            for (int j = 0; j <= U; ++j)
                if (s.min_x == s.bx[j] && s.min_y != s.by[j] && live) emit(s.bx[j], s.by[j]);
        }
    }
}

template <int W, int K, int U, class Emit, class Fetch>
struct KbSketchUnroll {
    KB_HD static void run(KbSketchState<W, K> &s, int i0, int i_end, int live_from, Fetch &fetch, Emit &emit)
    {
        int i = i0 + U;
        if (i < i_end) kb_sketch_step<W, K, U>(s, i, fetch(i), i >= live_from, emit);
        KbSketchUnroll<W, K, U + 1, Emit, Fetch>::run(s, i0, i_end, live_from, fetch, emit);
    }
};
template <int W, int K, class Emit, class Fetch>
struct KbSketchUnroll<W, K, W, Emit, Fetch> {
    KB_HD static void run(KbSketchState<W, K> &, int, int, int, Fetch &, Emit &) {}
};

// Sketch positions [start, end) of a contig of length ctg_len; emits every minimizer that
// mm_sketch pushes while processing those positions (plus the final push if end == ctg_len).
// fetch(i) returns the base code at contig position i.
template <int W, int K, class Emit, class Fetch>
KB_HD void kb_sketch_slice(int ctg_len, int start, int end, Fetch &fetch, Emit &emit)
{
    KbSketchState<W, K> s;
    s.reset();
    int p0 = start >= KB_SCAN_LOOKBACK ? start - KB_SCAN_LOOKBACK : 0;
    for (int i0 = p0; i0 < end; i0 += W) KbSketchUnroll<W, K, 0, Emit, Fetch>::run(s, i0, end, start, fetch, emit);
    if (end == ctg_len && end > start && s.min_x != KB_MAXU) emit(s.min_x, s.min_y);
}
