"""``rammappy.align.Aligner`` (reference call site serotyping/core.py:148-154)."""

from __future__ import annotations

import threading

from kaptive_b200 import mapper

from ._objects import Hit, Options, Strand, check_supported, cigar_to_bytes

_index_cache: dict = {}
_cache_lock = threading.Lock()


def _gene_index(queries, device: int) -> mapper.GeneIndex:
    """One device-resident gene index per distinct query list (the Serotyper passes the same list every call)."""
    key = (device, len(queries), hash(tuple(q[1] for q in queries)))
    with _cache_lock:
        gi = _index_cache.get(key)
        if gi is None:
            if len(_index_cache) >= 8:
                _index_cache.pop(next(iter(_index_cache)))
            gi = mapper.GeneIndex([bytes(q[1]) for q in queries], device=device)
            _index_cache[key] = gi
    return gi


class Aligner:
    def __init__(self, index=None, preset=None, do_cigar=True, do_cs=False, do_md=False):
        if index is None:
            raise ValueError("Aligner needs an Index")
        self.index = index
        self.preset = preset
        self.do_cigar, self.do_cs, self.do_md = do_cigar, do_cs, do_md
        self.options = Options()

    def map_batch(self, queries):
        """One iterator of hits per query, in query order (zip at core/alignment.py:409)."""
        queries = list(queries)
        check_supported(self.options, self.do_cigar, self.do_cs, self.do_md, self.preset)
        gi = _gene_index(queries, self.index.device)
        res = gi.map(self.index.batch)
        per_query: list[list[Hit]] = [[] for _ in queries]
        h = res.hits
        names, lens = self.index.names, self.index.lengths
        for i in range(len(res)):
            ctg = int(h["t_ctg"][i])
            per_query[int(h["gene"][i])].append(
                Hit(
                    target_name=names[ctg],
                    query_start=int(h["q_start"][i]),
                    query_end=int(h["q_end"][i]),
                    target_len=lens[ctg],
                    target_start=int(h["t_start"][i]),
                    target_end=int(h["t_end"][i]),
                    strand=Strand.Forward if h["strand"][i] > 0 else Strand.Reverse,
                    block_len=int(h["block_len"][i]),
                    matches=int(h["matches"][i]),
                    edit_distance=int(h["edit_distance"][i]),
                    score=int(h["score"][i]),
                    mapq=int(h["mapq"][i]),
                    is_primary=bool(h["is_primary"][i]),
                    cigar=cigar_to_bytes(res.cigar_of(i)),
                )
            )
        return [iter(x) for x in per_query]

    def map(self, query):
        name, seq = query if isinstance(query, tuple) else (b"0", query)
        return self.map_batch([(name, seq)])[0]
