// kb_scan.cuh -- minimizer sketch state machine for one lane's slice of a contig.
//
// Restates minimap2's mm_sketch (the algorithm behind rammappy.Index.build,
// reference call site src/kaptive/core/genome.py:188-189) so that it can be
// run on independent slices: the sketch state after base i is a pure function
// of the last w window entries and of l = min(run of unambiguous bases, w+k),
// so a lane that starts KB_SCAN_LOOKBACK = w+k-1 bases early in "silent" mode
// reproduces the sequential state bit for bit ("Scan kernel").
// Used on the host: the gene sketch at index build (kb_index.cpp) and the host
// emulation of the pipeline (tests/host_emul).  The assembly scan on the GPU
// (kb_scan.cu) evaluates a position-parallel restatement of the same rules.
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

// ---------------------------------------------------------------------------------------------------------------
// Fast sketch: the same emissions as kb_sketch_step, without the data-dependent window rescans.
//
// mm_sketch's `min` is always the latest minimal entry of the last W window entries, so it can be maintained as a
// sliding-window minimum: positions are grouped in blocks of W; the window at offset u of the current block is the
// suffix [u+1, W) of the previous block plus the prefix [0, u] of the current one.  Prefix minima are a running
// value, suffix minima are rebuilt once per block (W-1 compare/selects), and a "tie" flag carried with every minimum
// says whether the minimal hash occurs more than once in the window -- the only case in which mm_sketch's
// "identical k-mer" loops can emit anything, so the per-slot loops run only then (rare outside low-complexity DNA).
// Every decision is a select; the only branches are warp-uniform or the rare slow paths.
template <int W, int K>
struct KbFastSketch {
    uint32_t bx[W], by[W];  // the last W window entries (circular, slot = step % W); KB_MAXU = empty
    uint32_t sx[W], sy[W];  // suffix minima of the previous block: s[k] = latest minimal entry of slots k..W-1
    uint32_t sdup;          // bit k: the hash of s[k] occurs more than once in slots k..W-1
    uint32_t px, py;        // prefix minimum of the current block (latest minimal entry of slots 0..u)
    uint32_t pdup;
    uint32_t mx, my;        // mm_sketch's `min` after the previous step
    uint32_t fwd, rev;
    int l;
    KB_HD void reset()
    {
#pragma unroll
        for (int j = 0; j < W; ++j) bx[j] = by[j] = sx[j] = sy[j] = KB_MAXU;
        sdup = pdup = 0;
        px = py = mx = my = KB_MAXU;
        fwd = rev = 0;
        l = 0;
    }
};

// One step at compile-time window slot U.  Fast emission (at most one per step) is returned through (*ex, *ey) with
// the return value true; the rare identical-k-mer emissions go through `slow(x, y)`.
template <int W, int K, int U, class Slow>
KB_HD bool kb_fast_step(KbFastSketch<W, K> &s, int i, int c, bool live, uint32_t *ex, uint32_t *ey, Slow &slow)
{
    const uint32_t mask = (1u << (2 * K)) - 1u;
    const int shift1 = 2 * (K - 1);
    uint32_t ix = KB_MAXU, iy = KB_MAXU;
    if (c < 4) {
        s.fwd = ((s.fwd << 2) | (uint32_t)c) & mask;
        s.rev = (s.rev >> 2) | ((3u ^ (uint32_t)c) << shift1);
        const int z = s.fwd < s.rev ? 0 : 1;
        ++s.l;
        if (s.l >= K) ix = kb_hash32(z ? s.rev : s.fwd, mask), iy = ((uint32_t)i << 1) | (uint32_t)z;
    } else s.l = 0;
    s.bx[U] = ix, s.by[U] = iy;
    const uint32_t omx = s.mx, omy = s.my;  // `min` before this step
    // first full window: mm_sketch emits the entries equal to the old minimum that are not the minimum itself
    if (s.l == W + K - 1 && omx != KB_MAXU && live) {
#pragma unroll
        for (int j = U + 1; j < W; ++j)
            if (omx == s.bx[j] && s.by[j] != omy) slow(s.bx[j], s.by[j]);
#pragma unroll
        for (int j = 0; j < U; ++j)
            if (omx == s.bx[j] && s.by[j] != omy) slow(s.bx[j], s.by[j]);
    }
    // prefix minimum of the current block (the latest entry wins ties)
    if (U == 0) s.px = ix, s.py = iy, s.pdup = 0;
    else if (ix <= s.px) s.pdup = (ix == s.px), s.px = ix, s.py = iy;
    // window minimum = suffix of the previous block (slots U+1..W-1) combined with the prefix (later, so it wins ties)
    uint32_t nx = s.px, ny = s.py, ndup = s.pdup;
    if (U + 1 < W) {
        const uint32_t qx = s.sx[U + 1 < W ? U + 1 : 0], qy = s.sy[U + 1 < W ? U + 1 : 0], qd = (s.sdup >> (U + 1)) & 1u;
        if (qx < nx) nx = qx, ny = qy, ndup = qd;
        else if (qx == nx) ndup = 1;
    }
    bool emit = false;
    if (ix <= omx) {  // new minimum: write the old one
        emit = s.l >= W + K && omx != KB_MAXU;
    } else if ((omy >> 1) == (uint32_t)(i - W)) {  // the old minimum left the window (it is valid here: ix > omx)
        emit = s.l >= W + K - 1;
        if (s.l >= W + K - 1 && nx != KB_MAXU && ndup && live) {  // identical k-mers of the new minimum
#pragma unroll
            for (int j = U + 1; j < W; ++j)
                if (nx == s.bx[j] && ny != s.by[j]) slow(s.bx[j], s.by[j]);
#pragma unroll
            for (int j = 0; j <= U; ++j)
                if (nx == s.bx[j] && ny != s.by[j]) slow(s.bx[j], s.by[j]);
        }
    }
    *ex = omx, *ey = omy;
    s.mx = nx, s.my = ny;
    if (U == W - 1) {  // block complete: rebuild the suffix minima (later entries win ties)
        s.sx[W - 1] = s.bx[W - 1], s.sy[W - 1] = s.by[W - 1];
        uint32_t dup = 0;
#pragma unroll
        for (int k = W - 2; k >= 0; --k) {
            const uint32_t cx = s.bx[k], nxt = s.sx[k + 1];
            const uint32_t dk = (dup >> (k + 1)) & 1u;
            if (cx < nxt) s.sx[k] = cx, s.sy[k] = s.by[k];
            else s.sx[k] = nxt, s.sy[k] = s.sy[k + 1], dup |= ((cx == nxt) ? 1u : dk) << k;
        }
        s.sdup = dup;
    }
    return emit && live;
}

template <int W, int K, int U, class Emit, class Fetch>
struct KbFastUnroll {
    KB_HD static void run(KbFastSketch<W, K> &s, int i0, int i_end, int live_from, Fetch &fetch, Emit &emit)
    {
        const int i = i0 + U;
        if (i < i_end) {
            uint32_t ex, ey;
            if (kb_fast_step<W, K, U>(s, i, fetch(i), i >= live_from, &ex, &ey, emit)) emit(ex, ey);
        }
        KbFastUnroll<W, K, U + 1, Emit, Fetch>::run(s, i0, i_end, live_from, fetch, emit);
    }
};
template <int W, int K, class Emit, class Fetch>
struct KbFastUnroll<W, K, W, Emit, Fetch> {
    KB_HD static void run(KbFastSketch<W, K> &, int, int, int, Fetch &, Emit &) {}
};

// Same contract as kb_sketch_slice, on the fast state machine (used by the host emulation to pin it to the simple one).
template <int W, int K, class Emit, class Fetch>
KB_HD void kb_fast_slice(int ctg_len, int start, int end, Fetch &fetch, Emit &emit)
{
    KbFastSketch<W, K> s;
    s.reset();
    const int p0 = start >= KB_SCAN_LOOKBACK ? start - KB_SCAN_LOOKBACK : 0;
    for (int i0 = p0; i0 < end; i0 += W) KbFastUnroll<W, K, 0, Emit, Fetch>::run(s, i0, end, start, fetch, emit);
    if (end == ctg_len && end > start && s.mx != KB_MAXU) emit(s.mx, s.my);
}
