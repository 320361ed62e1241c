"""kaptive_b200 -- B200-native replacement for the gene -> contig mapping hot path of Kaptive
(the ``rammappy`` calls at ``src/kaptive/core/genome.py:45,188-189`` and
``src/kaptive/serotyping/core.py:148-154``).

* ``kaptive_b200.mapper``  : arrays in / arrays out over the C-ABI (``include/kaptive_b200.h``)
* ``kaptive_b200.shim``    : a module named ``rammappy`` so the unmodified reference runs on the GPU path
* ``kaptive_b200.synth``   : seeded synthetic databases and assemblies for tests and benchmarks
"""

from __future__ import annotations

import sys
from pathlib import Path

__version__ = "0.1.0"


def install_rammappy_shim() -> None:
    """Make ``import rammappy`` resolve to the CUDA-backed drop-in (must run before ``import kaptive.serotyping``)."""
    shim = str(Path(__file__).resolve().parent / "shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
