"""The scan kernel does not run mm_sketch's state machine: it evaluates a per-position restatement of its emission rules
(kaptive_b200/csrc/kb_scan.cu, "position-parallel mm_sketch").  This pins the restatement itself, in plain Python, against the
oracle's sequential mm_sketch on adversarial sequences; the kernel is checked against the same oracle in
tests/test_gpu_parity.py::test_scan_minimizers_adversarial_sequences."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "scripts"))

from proto_sketch_parallel import oracle, rand_seq, sketch_parallel  # noqa: E402


def test_order_free_rules_equal_sequential_sketch():
    rng = np.random.default_rng(11)
    for t in range(300):
        n = int(rng.choice([0, 1, 14, 15, 23, 24, 25, 26, 40, 100, 300]))
        s = rand_seq(rng, n, t % 5)
        assert sorted(sketch_parallel(s)) == sorted(oracle(s)), (t, n, s[:80])


def test_hand_made_corner_cases():
    for s in (b"A" * 200, b"AC" * 100, b"ACG" * 70, b"ACGTACGTAC" * 20, b"N" * 30, b"ACGT" * 6 + b"N" + b"ACGT" * 6,
              b"GATTACAGATTACAGATTACA" * 5 + b"N" * 3 + b"TGTAATCTGTAATCTGTAATC" * 5):
        assert sorted(sketch_parallel(s)) == sorted(oracle(s)), s[:40]
