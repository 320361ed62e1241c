"""Drop-in ``rammappy`` module backed by libkaptive_b200 (CUDA, sm_100a).

Put ``<repo>/kaptive_b200/shim`` in front of ``sys.path`` (or call
``kaptive_b200.install_rammappy_shim()``) and the unmodified reference runs on the GPU mapper:
``import kaptive.serotyping`` imports this module at ``serotyping/core.py:15-16`` and
``core/genome.py:42,186``.  There is no CPU fallback: without the CUDA library these calls raise.
"""

from __future__ import annotations

from . import align, fasta
from ._objects import Preset, Strand
from ._index import Index

__all__ = ["Index", "Preset", "Strand", "align", "fasta"]
__version__ = "0.1.3+kaptive_b200"
