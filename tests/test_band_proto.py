"""CPU tier: the certificate of the banded gap-fill pass (kaptive_b200/csrc/kb_align_reg.cuh: kb_global_band /
kb_global_bandK) stated as executable Python (tests/proto_band.py): whenever a band pass accepts itself, its score and
CIGAR equal the full-rectangle DP's, on random, indel-rich, ambiguous and tandem-repeat inputs, for several band widths."""

import random

import pytest

from proto_band import band_dp, full_dp


def _pair(rng, n):
    ts = [rng.randint(0, 3) for _ in range(n)]
    sub_r, ind_r = rng.choice([0, 0.02, 0.05, 0.15, 0.3]), rng.choice([0, 0.01, 0.03, 0.1])
    qs = []
    for c in ts:
        u = rng.random()
        if u < ind_r / 2:
            continue
        if u < ind_r:
            qs.append(rng.randint(0, 3))
        qs.append(rng.randint(0, 3) if rng.random() < sub_r else c)
    if qs and rng.random() < 0.1:
        qs[rng.randrange(len(qs))] = 4
    if rng.random() < 0.1:  # tandem repeats: many co-optimal paths far from the diagonal
        unit = ts[:7]
        ts = (unit * 20)[:n]
        qs = (unit * 20)[: max(3, n - rng.randint(0, 10))]
    return qs, ts


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_certified_band_equals_full_dp(seed):
    rng = random.Random(seed)
    n_cert = n_rej = 0
    for _ in range(120):
        qs, ts = _pair(rng, rng.randint(5, 70))
        if not qs:
            continue
        want_s, want_c = full_dp(qs, ts)
        ok, s, c = band_dp(qs, ts, W=rng.choice([4, 6, 8, 16]))
        if ok:
            n_cert += 1
            assert (s, c) == (want_s, want_c)
        else:
            n_rej += 1
            assert s is None or s <= want_s  # a band pass is a lower bound
    assert n_cert > 40 and n_rej > 5


def test_full_width_band_is_always_certified():
    rng = random.Random(9)
    for _ in range(30):
        qs, ts = _pair(rng, rng.randint(5, 25))
        if not qs:
            continue
        ok, s, c = band_dp(qs, ts, W=64)
        assert ok and (s, c) == full_dp(qs, ts)
