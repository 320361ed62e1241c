"""The base-level DP kernels one by one (kb_debug_dp): the register-resident forms -- row-stripe wavefront (kb_rows), its packed
16-bit twin with two cells per DPX instruction (kb_rows16), the certified band pass -- against the scratch-memory DP, which is the
statement closest to oracle/kb_oracle.c:extd2 and is itself covered against the oracle by the hit-level parity tests.  Bit-exact:
score, maximum and its cell, z-drop flag, CIGAR.  The packed kernel runs two jobs per warp: mode 2 leaves the second half empty,
mode 4 pairs jobs 2i and 2i + 1, which here differ in kind (left / right extension / global fill), size and band."""

from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EXTZ_ONLY, RIGHT, REV_CIGAR, GLOBAL = 1, 2, 4, 8
NT = np.frombuffer(b"ACGT", np.uint8)


def _mutate(rng, s, sub, indel):
    out = []
    for b in s:
        r = rng.random()
        if r < indel / 2:
            continue
        if r < indel:
            out.append(int(rng.integers(0, 4)))
        out.append(int((b + rng.integers(1, 4)) % 4) if rng.random() < sub else int(b))
    return np.array(out, np.uint8) if out else np.zeros(1, np.uint8)


def make_jobs(seed: int, n: int):
    rng = np.random.default_rng(seed)
    q, t, fl, w, zd = [], [], [], [], []
    for i in range(n):
        kind = i % 3
        regime = rng.integers(0, 6) if i % 10 else 5
        ql = int({0: rng.integers(1, 40), 1: rng.integers(30, 140), 2: rng.integers(100, 400), 3: rng.integers(300, 760),
                  4: rng.integers(1, 760), 5: rng.integers(700, 1300)}[int(regime)])  # 5: the band of the spec (w = 751) binds
        qs = rng.integers(0, 4, size=ql).astype(np.uint8)
        sub = float(rng.choice([0.0, 0.02, 0.08, 0.15, 0.3]))
        ind = float(rng.choice([0.0, 0.005, 0.03]))
        core = _mutate(rng, qs, sub, ind)
        if kind == 2:  # global gap fill: target = diverged copy, sometimes with a long indel
            ts = core
            if rng.random() < 0.3 and len(ts) > 60:
                c = int(rng.integers(10, len(ts) - 10))
                ts = np.concatenate([ts[:c], rng.integers(0, 4, size=int(rng.integers(1, 90))).astype(np.uint8), ts[c:]])
            ts = ts[:1500]
            f, ww, z = GLOBAL, int(rng.choice([751, 30001])), -1
        else:  # end extension: the target window is about twice the query (mm_align1), alignment near the diagonal then random
            tl = max(1, min(2 * ql + int(rng.integers(-3, 4)), 752 if regime < 5 else 1700))
            tail = rng.integers(0, 4, size=tl).astype(np.uint8)
            ts = np.concatenate([core, tail])[:tl]
            if rng.random() < 0.15:  # poor start: z-drop territory
                ts = rng.integers(0, 4, size=tl).astype(np.uint8)
            f = (EXTZ_ONLY | RIGHT | REV_CIGAR) if kind == 0 else EXTZ_ONLY
            ww, z = 751, int(rng.choice([400, 400, 60]))
        if rng.random() < 0.2:  # ambiguous bases on either side
            for arr in (qs, ts):
                k = int(rng.integers(0, 4))
                if k and len(arr):
                    arr[rng.integers(0, len(arr), size=k)] = 4
        q.append(qs), t.append(ts.astype(np.uint8)), fl.append(f), w.append(ww), zd.append(z)
    return q, t, np.array(fl, np.int32), np.array(w, np.int32), np.array(zd, np.int32)


def run(L, params, q, t, fl, w, zd, mode, stride=4096):
    from kaptive_b200._lib import ptr

    ql = np.array([len(x) for x in q], np.int32)
    tl = np.array([len(x) for x in t], np.int32)
    qo = np.concatenate([[0], np.cumsum(ql[:-1])]).astype(np.int64)
    to = np.concatenate([[0], np.cumsum(tl[:-1])]).astype(np.int64)
    qq, tt = np.concatenate(q), np.concatenate(t)
    n = len(q)
    out = np.zeros((n, 8), np.int32)
    cig = np.zeros((n, stride), np.uint32)
    rc = L.kb_debug_dp(C.byref(params), 0, ptr(qq), ptr(qo), ptr(ql), ptr(tt), ptr(to), ptr(tl), ptr(fl), ptr(w), ptr(zd), n, mode, ptr(out),
                       ptr(cig), stride)
    assert rc == 0, rc
    return out, cig


@pytest.mark.parametrize("seed", [1, 2])
def test_register_dp_kernels_equal_the_scratch_dp(seed):
    from kaptive_b200 import _lib

    L = _lib.load()
    P = _lib.default_params()
    q, t, fl, w, zd = make_jobs(seed, 900)
    ref, rcig = run(L, P, q, t, fl, w, zd, 0)
    n_ran = {}
    for mode, name in ((1, "rows"), (2, "rows16"), (4, "rows16 pair"), (3, "band")):
        got, gcig = run(L, P, q, t, fl, w, zd, mode)
        ran = got[:, 6] == 1
        n_ran[name] = int(ran.sum())
        for i in np.nonzero(ran)[0]:
            desc = (name, int(i), len(q[i]), len(t[i]), int(fl[i]), int(w[i]), int(zd[i]))
            if fl[i] & GLOBAL:
                assert got[i, 0] == ref[i, 0], (desc, "score", got[i].tolist(), ref[i].tolist())
            else:
                assert got[i, 1:5].tolist() == ref[i, 1:5].tolist(), (desc, "max / zdrop", got[i].tolist(), ref[i].tolist())
                if not ref[i, 4]:
                    assert got[i, 0] == ref[i, 0], (desc, "score", got[i].tolist(), ref[i].tolist())
            assert got[i, 5] == ref[i, 5], (desc, "n_cigar", got[i].tolist(), ref[i].tolist())
            nc = int(ref[i, 5])
            assert np.array_equal(gcig[i, :nc], rcig[i, :nc]), (desc, "cigar")
    assert n_ran["rows"] > 800 and n_ran["rows16"] > 650 and n_ran["rows16 pair"] > 450 and n_ran["band"] > 50, n_ran


def test_rows16_range_limits_fall_back():
    """A rectangle the packed kernel must refuse (its traceback does not fit half of the warp's scratch) next to one it takes."""
    from kaptive_b200 import _lib

    L = _lib.load()
    P = _lib.default_params()
    rng = np.random.default_rng(3)
    q = [rng.integers(0, 4, size=2500).astype(np.uint8), rng.integers(0, 4, size=50).astype(np.uint8)]
    t = [rng.integers(0, 4, size=2500).astype(np.uint8), rng.integers(0, 4, size=20).astype(np.uint8)]
    fl, w, zd = np.array([EXTZ_ONLY, EXTZ_ONLY], np.int32), np.array([751, 751], np.int32), np.array([400, 400], np.int32)
    got, _ = run(L, P, q, t, fl, w, zd, 2)
    assert got[:, 6].tolist() == [0, 1]


def make_fills(seed: int, n: int):
    """Global gap fills of every size the packed band pass takes (qlen + tlen <= 3000), neighbours unrelated in size and divergence."""
    rng = np.random.default_rng(seed)
    q, t = [], []
    for i in range(n):
        ql = int(rng.choice([rng.integers(33, 80), rng.integers(60, 300), rng.integers(250, 900), rng.integers(800, 1480)]))
        qs = rng.integers(0, 4, size=ql).astype(np.uint8)
        sub = float(rng.choice([0.0, 0.01, 0.04, 0.1, 0.18, 0.3]))
        ind = float(rng.choice([0.0, 0.004, 0.02]))
        ts = _mutate(rng, qs, sub, ind)
        if rng.random() < 0.35 and len(ts) > 80:  # one long indel: what the wider windows are for
            c = int(rng.integers(20, len(ts) - 20))
            g = int(rng.integers(5, 110))
            ts = np.concatenate([ts[:c], rng.integers(0, 4, size=g).astype(np.uint8), ts[c:]]) if rng.random() < 0.5 else np.concatenate([ts[:c], ts[c + g:]])
        ts = ts[:1490]
        if len(ts) < 33:
            ts = np.concatenate([ts, rng.integers(0, 4, size=40).astype(np.uint8)])
        if rng.random() < 0.15:
            qs[rng.integers(0, len(qs), size=2)] = 4  # ambiguous query bases stay in the packed pass
        if rng.random() < 0.05:
            ts[rng.integers(0, len(ts))] = 4          # an ambiguous target base sends the pair to the 32-bit kernel
        q.append(qs), t.append(ts.astype(np.uint8))
    fl = np.full(n, GLOBAL, np.int32)
    return q, t, fl, np.full(n, 30001, np.int32), np.full(n, -1, np.int32)


@pytest.mark.parametrize("seed", [5, 6])
def test_packed_band_pass_equals_the_scratch_dp(seed):
    """kb_band16 (two fills per warp, window of 64 / 128 / 256 diagonals): whatever it certifies is the full DP's alignment, bit for bit."""
    from kaptive_b200 import _lib

    L = _lib.load()
    P = _lib.default_params()
    q, t, fl, w, zd = make_fills(seed, 800)
    ref, rcig = run(L, P, q, t, fl, w, zd, 0)
    legacy, _ = run(L, P, q, t, fl, w, zd, 3)
    n_ok = {}
    for mode, K in ((5, 1), (6, 2), (7, 4)):
        got, gcig = run(L, P, q, t, fl, w, zd, mode)
        ok = got[:, 6] == 1
        n_ok[K] = int(ok.sum())
        for i in np.nonzero(ok)[0]:
            desc = (K, int(i), len(q[i]), len(t[i]))
            assert got[i, 0] == ref[i, 0], (desc, "score", got[i].tolist(), ref[i].tolist())
            assert got[i, 5] == ref[i, 5], (desc, "n_cigar", got[i].tolist(), ref[i].tolist())
            nc = int(ref[i, 5])
            assert np.array_equal(gcig[i, :nc], rcig[i, :nc]), (desc, "cigar")
        # a pass that ran to the end without certifying still holds a valid alignment's score: a lower bound of the optimum
        low = (got[:, 6] == 2) & (got[:, 7] == 1)
        assert np.all(got[low, 0] <= ref[low, 0]), K
        if K == 1:  # the 64-diagonal window certifies (nearly) what the 32-bit pass certifies: its band may sit one diagonal lower
            both = (legacy[:, 6] == 1) & (got[:, 6] != 0)
            assert (got[both, 6] == 1).mean() > 0.97
    assert n_ok[1] > 150 and n_ok[2] > n_ok[1] and n_ok[4] > n_ok[2], n_ok
