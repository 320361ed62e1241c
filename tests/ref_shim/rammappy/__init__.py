"""TEST INFRASTRUCTURE: a ``rammappy`` module backed by the CPU oracle, used only to generate the golden
end-to-end vectors in this container (which has the reference but no GPU).  Same objects as the product
shim, different engine.  The product never imports this."""

from __future__ import annotations

import sys
from pathlib import Path

_ROOT = Path(__file__).resolve().parents[3]
for _p in (str(_ROOT), str(_ROOT / "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import importlib.util as _ilu

_spec = _ilu.spec_from_file_location("_kb_shim_objects", _ROOT / "kaptive_b200/shim/rammappy/_objects.py")
_objects = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(_objects)
Preset, Strand = _objects.Preset, _objects.Strand

import sys as _s
_s.modules[__name__ + "._objects"] = _objects
from . import align, fasta  # noqa: E402


class Index:
    def __init__(self, names, contigs):
        self.names, self.contigs = names, contigs

    @classmethod
    def build(cls, seqs):
        seqs = list(seqs)
        return cls([bytes(n) for n, _ in seqs], [bytes(s) for _, s in seqs])
