"""``rammappy.Index``: the contigs of one assembly, 2-bit packed and resident on the GPU
(reference: ``Index.build([(id.encode(), seq) ...])`` at core/genome.py:188-189; the result is cached on
the ``GenomeAssembly`` and shared across threads, so it is immutable after ``build``)."""

from __future__ import annotations

import os

from kaptive_b200 import mapper


class Index:
    __slots__ = ("names", "lengths", "batch", "device")

    def __init__(self, names, lengths, batch, device):
        self.names = names
        self.lengths = lengths
        self.batch = batch
        self.device = device

    @classmethod
    def build(cls, seqs) -> "Index":
        seqs = list(seqs)
        device = int(os.environ.get("KAPTIVE_B200_DEVICE", "0"))
        names = [bytes(n) for n, _ in seqs]
        contigs = [bytes(s) for _, s in seqs]
        batch = mapper.AssemblyBatch.from_contigs([contigs], device=device)
        return cls(names, [len(c) for c in contigs], batch, device)

    def __len__(self) -> int:
        return len(self.names)
