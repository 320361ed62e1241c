from __future__ import annotations

import numpy as np

import os

import oracle_lib as ol

from . import _objects as O


def _engine(flat_genes):
    """KB_REF_SHIM_ENGINE=emul: the PRODUCT's device logic compiled for the host (tests/host_emul) instead of the oracle, so that the
    unmodified reference pipeline can run on top of this repo's own mapping code in a container without a GPU."""
    if os.environ.get("KB_REF_SHIM_ENGINE", "oracle") == "emul":
        import emul_lib as el

        return el.EmulIndex(*flat_genes)
    return ol.OracleDB(*flat_genes)

_cache: dict = {}
last_hits = None  # the raw oracle output of the most recent map_batch call (golden generation reads it)


def _flat(seqs):
    lengths = np.array([len(s) for s in seqs], dtype=np.int32)
    offsets = np.zeros(len(seqs), dtype=np.int64)
    if len(seqs) > 1:
        np.cumsum(lengths[:-1].astype(np.int64), out=offsets[1:])
    data = np.frombuffer(b"".join(seqs), dtype=np.uint8) if seqs else np.zeros(0, np.uint8)
    return data, offsets, lengths


class Aligner:
    def __init__(self, index=None, preset=None, do_cigar=True, do_cs=False, do_md=False):
        self.index, self.preset = index, preset
        self.do_cigar, self.do_cs, self.do_md = do_cigar, do_cs, do_md
        self.options = O.Options()

    def map_batch(self, queries):
        global last_hits
        queries = list(queries)
        O.check_supported(self.options, self.do_cigar, self.do_cs, self.do_md, self.preset)
        key = (len(queries), hash(tuple(q[1] for q in queries)))
        if key not in _cache:
            _cache[key] = (os.environ.get("KB_REF_SHIM_ENGINE", "oracle"), _engine(_flat([bytes(q[1]) for q in queries])))
        if _cache[key][0] != os.environ.get("KB_REF_SHIM_ENGINE", "oracle"):
            _cache[key] = (os.environ.get("KB_REF_SHIM_ENGINE", "oracle"), _engine(_flat([bytes(q[1]) for q in queries])))
        r = _cache[key][1].map(*_flat(self.index.contigs))
        last_hits = r
        per = [[] for _ in queries]
        for h in r["hits"]:
            cg = r["cigar"][h["cigar_off"] : h["cigar_off"] + h["n_cigar"]]
            ctg = int(h["t_ctg"])
            per[int(h["gene"])].append(
                O.Hit(self.index.names[ctg], int(h["q_start"]), int(h["q_end"]), len(self.index.contigs[ctg]), int(h["t_start"]),
                      int(h["t_end"]), O.Strand.Forward if h["strand"] > 0 else O.Strand.Reverse, int(h["block_len"]),
                      int(h["matches"]), int(h["edit_distance"]), int(h["score"]), int(h["mapq"]), bool(h["is_primary"]),
                      O.cigar_to_bytes(cg))
            )
        return [iter(x) for x in per]
