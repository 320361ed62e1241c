"""Prototype: mm_sketch (k=15, w=10) as a per-position, order-free formulation -- checked against the oracle's sequential sketch.

Per position i:  x[i] (hash or MAX), y[i], l[i] = run of unambiguous bases ending at i.
M[i] = rightmost arg-min of x over [i-w+1, i].   om = M[i-1].
  E1: x[i] <= om.x, l[i] >= w+k,  om.x != MAX                  -> emit om
  E2: x[i-w] < M[i].x (strict), l[i] >= w+k-1                  -> emit (x[i-w], y[i-w]) (= om), then every j in [i-w+1, i]
                                                                  with x[j] == M[i].x, j != M[i].pos
  E0: l[i] == w+k-1, om.x != MAX                               -> every j in [i-w+1, i-1] with x[j] == om.x, j != om.pos
  End: i == len-1, M[i].x != MAX                               -> emit M[i]
"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib

K, W = 15, 10
MASK = (1 << 2 * K) - 1
MAX = 0xFFFFFFFF

def hash32(key):
    key = (~key + (key << 21)) & MASK
    key = key ^ key >> 24
    key = ((key + (key << 3)) + (key << 8)) & MASK
    key = key ^ key >> 14
    key = ((key + (key << 2)) + (key << 4)) & MASK
    key = key ^ key >> 28
    return key

NT4 = np.full(256, 4, np.uint8)
for ch, v in zip("ACGTacgt", [0, 1, 2, 3] * 2):
    NT4[ord(ch)] = v
NT4[ord("U")] = NT4[ord("u")] = 3

def sketch_parallel(seq: bytes):
    c = NT4[np.frombuffer(seq, np.uint8)]
    n = len(c)
    if n == 0: return []
    x = [MAX] * n; y = [MAX] * n; l = [0] * n
    fwd = rev = 0; run = 0
    for i in range(n):
        if c[i] < 4:
            fwd = ((fwd << 2) | int(c[i])) & MASK
            rev = (rev >> 2) | ((3 ^ int(c[i])) << (2 * (K - 1)))
            run += 1
            if run >= K:
                z = 0 if fwd < rev else 1
                x[i] = hash32(rev if z else fwd); y[i] = (i << 1) | z
        else:
            run = 0
        l[i] = run
    gx = lambda j: x[j] if j >= 0 else MAX
    def M(i):  # rightmost argmin of window(i)
        if i < 0: return MAX, -1
        bx, bp = MAX, -1
        for j in range(i - W + 1, i + 1):
            if gx(j) <= bx: bx, bp = gx(j), j
        return bx, bp
    out = []
    for i in range(n):
        omx, omp = M(i - 1); mx, mp = M(i)
        if l[i] == W + K - 1 and omx != MAX:
            for j in range(i - W + 1, i):
                if gx(j) == omx and j != omp: out.append((x[j], y[j]))
        if x[i] <= omx:
            if l[i] >= W + K and omx != MAX: out.append((omx, y[omp]))
        elif gx(i - W) < mx:
            assert omp == i - W
            if l[i] >= W + K - 1:
                out.append((omx, y[omp]))
                if mx != MAX:
                    for j in range(i - W + 1, i + 1):
                        if gx(j) == mx and j != mp: out.append((x[j], y[j]))
        if i == n - 1 and mx != MAX: out.append((mx, y[mp]))
    return out

def oracle(seq: bytes):
    ox, oy = oracle_lib.sketch(seq, W, K)
    return [(int(a) >> 8, int(b)) for a, b in zip(ox, oy)]

def rand_seq(rng, n, mode):
    if mode == 0: s = rng.choice(list(b"ACGT"), n)
    elif mode == 1: s = rng.choice(list(b"ACGTN"), n, p=[.24, .24, .24, .24, .04])
    elif mode == 2:  # tandem repeats of short periods + noise
        unit = rng.choice(list(b"ACGT"), rng.integers(1, 9)); s = np.resize(unit, n).copy()
        m = rng.random(n) < 0.02; s[m] = rng.choice(list(b"ACGTN"), m.sum())
    elif mode == 3:  # low complexity two-letter
        s = rng.choice(list(b"AT"), n)
    else:  # blocks: repeats, random, N runs
        parts = []
        while sum(map(len, parts)) < n:
            t = rng.integers(0, 4)
            if t == 0: parts.append(rng.choice(list(b"ACGT"), rng.integers(1, 60)))
            elif t == 1: parts.append(np.resize(rng.choice(list(b"ACGT"), rng.integers(1, 7)), rng.integers(10, 80)))
            elif t == 2: parts.append(np.full(rng.integers(1, 4), ord("N")))
            else:
                p = rng.choice(list(b"ACGT"), rng.integers(8, 30)); comp = {65: 84, 67: 71, 71: 67, 84: 65}
                parts.append(np.concatenate([p, np.array([comp[int(v)] for v in p[::-1]])]))  # inverted repeat
        s = np.concatenate(parts)[:n] if parts else np.zeros(0, np.uint8)
    return bytes(np.asarray(s, np.uint8))

if __name__ == "__main__":
    rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    bad = 0
    for t in range(int(sys.argv[2]) if len(sys.argv) > 2 else 600):
        n = int(rng.choice([0, 1, 14, 15, 23, 24, 25, 26, 40, 100, 300, 700]))
        s = rand_seq(rng, n, t % 5)
        a, b = sorted(sketch_parallel(s)), sorted(oracle(s))
        if a != b:
            bad += 1
            if bad < 4: print("MISMATCH", t, n, s[:120], "\n ours-only", sorted(set(a) - set(b))[:5], "oracle-only", sorted(set(b) - set(a))[:5], len(a), len(b))
    print("mismatches:", bad)
