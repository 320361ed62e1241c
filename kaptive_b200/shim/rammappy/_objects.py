"""Plain Python objects of the ``rammappy`` surface the reference touches:
``Preset`` (annotation only, serotyping/core.py:68), aligner options with ``filtering.best_n`` /
``filtering.pri_ratio`` (core.py:149-152), and the per-hit record drained by
``Alignments.from_mapping_iterators`` (core/alignment.py:414-446)."""

from __future__ import annotations

import enum


class Preset(enum.Enum):
    """Name only: the reference stores it and always passes ``preset=None`` (serotyping/core.py:95,148)."""

    MapOnt = "map-ont"
    Asm5 = "asm5"
    Asm10 = "asm10"
    Asm20 = "asm20"
    Sr = "sr"


class Strand(enum.Enum):
    Forward = 1
    Reverse = -1

    def __repr__(self) -> str:  # alignment.py:428 tests `"Forward" in repr(h.strand)`
        return f"Strand.{self.name}"


class Filtering:
    __slots__ = ("best_n", "pri_ratio")

    def __init__(self) -> None:
        self.best_n = 5  # minimap2 defaults; the reference overrides both
        self.pri_ratio = 0.8


class Options:
    __slots__ = ("filtering",)

    def __init__(self) -> None:
        self.filtering = Filtering()


class Hit:
    """One alignment record; field names as read at core/alignment.py:414-446."""

    __slots__ = (
        "target_name", "query_start", "query_end", "target_len", "target_start", "target_end", "strand", "block_len",
        "matches", "edit_distance", "score", "mapq", "is_primary", "is_supplementary", "is_spliced", "divergence",
        "cs", "md", "cigar",
    )  # fmt: skip

    def __init__(self, target_name, query_start, query_end, target_len, target_start, target_end, strand, block_len,
                 matches, edit_distance, score, mapq, is_primary, cigar):
        self.target_name = target_name
        self.query_start = query_start
        self.query_end = query_end
        self.target_len = target_len
        self.target_start = target_start
        self.target_end = target_end
        self.strand = strand
        self.block_len = block_len
        self.matches = matches
        self.edit_distance = edit_distance
        self.score = score
        self.mapq = mapq
        self.is_primary = is_primary
        self.is_supplementary = False
        self.is_spliced = False
        # gap-compressed-free per-base divergence of the aligned block; carried by the reference, never read
        self.divergence = (1.0 - matches / block_len) if block_len > 0 else 0.0
        self.cs = None
        self.md = None
        self.cigar = cigar

    def __repr__(self) -> str:
        return (f"Hit({self.target_name!r}, q={self.query_start}-{self.query_end}, t={self.target_start}-{self.target_end}, "
                f"{self.strand!r}, score={self.score}, mapq={self.mapq})")


def cigar_to_bytes(cig) -> bytes:
    ops = b"MIDNSHP=X"
    return b"".join(b"%d%c" % (int(c) >> 4, ops[int(c) & 0xF]) for c in cig)


def check_supported(options: Options, do_cigar: bool, do_cs: bool, do_md: bool, preset) -> None:
    """Only the configuration the reference uses is implemented (serotyping/core.py:148-152)."""
    if preset is not None:
        raise NotImplementedError("kaptive_b200 implements the no-preset (minimap2 default) configuration only")
    if do_cs or do_md:
        raise NotImplementedError("cs / MD tags are not produced (the reference passes do_cs=False, do_md=False)")
    if not do_cigar:
        raise NotImplementedError("kaptive_b200 always performs base-level alignment (do_cigar=True)")
    f = options.filtering
    if f.pri_ratio != 0.0 or f.best_n < 50000:
        raise NotImplementedError(
            "kaptive_b200 keeps and extends every chain: set options.filtering.best_n >= 50000 and pri_ratio = 0.0 "
            "as kaptive.serotyping.Serotyper does"
        )
