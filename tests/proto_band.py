"""Prototype / executable specification of the certified band DP (kb_global_band in csrc/kb_align_reg.cuh).

Lane-level simulation of the kernel's data movement (W lanes = W cells per anti-diagonal, alternating
shuffle direction) next to a plain full-rectangle DP that follows oracle/kb_oracle.c:extd2 for a global
alignment (EZ_GLOBAL_NO_ZDROP, band never binding).  Used by tests/test_band_proto.py: whenever the band pass
certifies itself, score and CIGAR must equal the full DP's.
"""

from __future__ import annotations

NEG = -0x20000000
A, Bm, Q, E, Q2, E2, AMBI = 2, 4, 4, 2, 24, 1, 1


def gc(l):
    return min(Q + E * l, Q2 + E2 * l)


def sub(ct, cq):
    return -AMBI if (ct > 3 or cq > 3) else (A if ct == cq else -Bm)


def cell(h_up, a1, a2, h_left, b1, b2, h_diag, ct, cq, rb=0):
    a1 = max(h_up - Q, a1) - E
    a2 = max(h_up - Q2, a2) - E2
    b1 = max(h_left - Q, b1) - E
    b2 = max(h_left - Q2, b2) - E2
    z = h_diag + sub(ct, cq)
    d = 0
    if a1 + rb > z: d, z = 1, a1
    if b1 + rb > z: d, z = 2, b1
    if a2 + rb > z: d, z = 3, a2
    if b2 + rb > z: d, z = 4, b2
    hq, hq2 = z - Q - rb, z - Q2 - rb
    d |= (0x08 if a1 > hq else 0) | (0x10 if b1 > hq else 0) | (0x20 if a2 > hq2 else 0) | (0x40 if b2 > hq2 else 0)
    return z, a1, a2, b1, b2, d


def backtrack(get, tlen, qlen):
    """ksw_backtrack from (tlen-1, qlen-1); get(i, j) -> traceback byte."""
    cig = []

    def push(op, n):
        if cig and cig[-1][0] == op: cig[-1][1] += n
        else: cig.append([op, n])

    i, j, state = tlen - 1, qlen - 1, 0
    while i >= 0 and j >= 0:
        tmp = get(i, j)
        if state == 0: state = tmp & 7
        elif not (tmp >> (state + 2) & 1): state = 0
        if state == 0: state = tmp & 7
        if state == 0: push(0, 1); i -= 1; j -= 1
        elif state in (1, 3): push(2, 1); i -= 1
        else: push(1, 1); j -= 1
    if i >= 0: push(2, i + 1)
    if j >= 0: push(1, j + 1)
    return [tuple(c) for c in reversed(cig)]


def full_dp(qs, ts):
    qlen, tlen = len(qs), len(ts)
    H = {}; E1 = {}; E2_ = {}; F1 = {}; F2 = {}; tb = {}
    for r in range(qlen + tlen - 1):
        for t in range(max(0, r - qlen + 1), min(tlen - 1, r) + 1):
            j = r - t
            if t == 0: h_up, a1, a2 = -gc(j + 1), NEG, NEG
            else: h_up, a1, a2 = H[t - 1, j], E1[t - 1, j], E2_[t - 1, j]
            if j == 0: h_left, b1, b2 = -gc(t + 1), NEG, NEG
            else: h_left, b1, b2 = H[t, j - 1], F1[t, j - 1], F2[t, j - 1]
            if t == 0 and j == 0: h_diag = 0
            elif t == 0: h_diag = -gc(j)
            elif j == 0: h_diag = -gc(t)
            else: h_diag = H[t - 1, j - 1]
            H[t, j], E1[t, j], E2_[t, j], F1[t, j], F2[t, j], tb[t, j] = cell(h_up, a1, a2, h_left, b1, b2, h_diag, ts[t], qs[j])
    return H[tlen - 1, qlen - 1], backtrack(lambda i, j: tb[i, j], tlen, qlen)


def band_dp(qs, ts, W=32, min_margin=2):
    """Returns (certified, score, cigar).  W lanes <-> 2W diagonals [dlo, dlo + 2W - 1]."""
    qlen, tlen = len(qs), len(ts)
    d1 = tlen - qlen
    lo_d, hi_d = min(0, d1), max(0, d1)
    margin = (2 * W - 1 - (hi_d - lo_d)) >> 1
    if margin < min_margin: return False, None, None
    dlo = lo_d - margin
    dhi = dlo + 2 * W - 1
    # shifted coordinates: t' = t + 1, j' = j + 1; index 0 is the virtual boundary row / column
    T = lambda rp: (rp + dlo + 1) >> 1
    H1 = [NEG] * W; H2 = [NEG] * W; E1 = [NEG] * W; E2_ = [NEG] * W; F1 = [NEG] * W; F2 = [NEG] * W
    tb = {}
    r_end = tlen + qlen
    for rp in range(0, r_end + 1):
        stepB = (rp + dlo) & 1
        if not stepB:  # A step: up from lane - 1, left own
            upH = [NEG] + H1[:-1]; upE1 = [NEG] + E1[:-1]; upE2 = [NEG] + E2_[:-1]
            lfH, lfF1, lfF2 = H1, F1, F2
        else:  # B step: left from lane + 1, up own
            lfH = H1[1:] + [NEG]; lfF1 = F1[1:] + [NEG]; lfF2 = F2[1:] + [NEG]
            upH, upE1, upE2 = H1, E1, E2_
        nH = [0] * W; nE1 = [0] * W; nE2 = [0] * W; nF1 = [0] * W; nF2 = [0] * W
        for l in range(W):
            tp = T(rp) + l
            jp = rp - tp
            if tp < 0 or jp < 0 or tp > tlen or jp > qlen:
                nH[l] = nE1[l] = nE2[l] = nF1[l] = nF2[l] = NEG  # never read by a valid cell
            elif tp == 0 or jp == 0:
                nH[l] = 0 if (tp == 0 and jp == 0) else -gc(max(tp, jp))
                nE1[l] = nE2[l] = nF1[l] = nF2[l] = NEG
            else:
                nH[l], nE1[l], nE2[l], nF1[l], nF2[l], d = cell(upH[l], upE1[l], upE2[l], lfH[l], lfF1[l], lfF2[l], H2[l],
                                                                 ts[tp - 1], qs[jp - 1])
                tb[rp, l] = d
        H2 = H1; H1, E1, E2_, F1, F2 = nH, nE1, nE2, nF1, nF2
    score = H1[tlen - T(r_end)]
    # certificate: any path that leaves the band scores at most `bound`
    def bound(D, I):
        m = tlen - D
        return NEG if (m < 0 or qlen - I < 0) else A * m - gc(D) - gc(I)
    D_hi = dhi + 1; I_hi = D_hi - d1
    I_lo = 1 - dlo; D_lo = I_lo + d1
    b = max(bound(D_hi, I_hi), bound(D_lo, I_lo))
    if not score > b: return False, score, None

    def get(i, j):
        rp = i + j + 2
        l = i + 1 - T(rp)
        assert 0 <= l < W, "certified path left the band"
        return tb[rp, l]

    return True, score, backtrack(get, tlen, qlen)
