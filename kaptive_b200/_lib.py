"""ctypes loader for ``libkaptive_b200.so`` (the C-ABI declared in ``include/kaptive_b200.h``).

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible, the
compute entry points raise.  The library is built in-tree by ``kaptive_b200/build.py``.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
# KAPTIVE_B200_LIB: another build of the library (kernel experiments); the in-tree build otherwise
SO_PATH = Path(os.environ["KAPTIVE_B200_LIB"]) if os.environ.get("KAPTIVE_B200_LIB") else HERE / "_lib" / "libkaptive_b200.so"

KB_N_STAGES = 6
STAGE_NAMES = ("scan", "sort", "chain", "align", "final", "total")


class KbError(RuntimeError):
    """Raised for any non-zero status crossing the C-ABI."""


class KbParams(C.Structure):
    _fields_ = [
        ("k", C.c_int32), ("w", C.c_int32),
        ("min_cnt", C.c_int32), ("min_chain_score", C.c_int32), ("bw", C.c_int32), ("max_gap", C.c_int32),
        ("max_chain_skip", C.c_int32), ("max_chain_iter", C.c_int32),
        ("chain_gap_scale", C.c_float),
        ("a", C.c_int32), ("b", C.c_int32), ("q", C.c_int32), ("e", C.c_int32), ("q2", C.c_int32), ("e2", C.c_int32),
        ("sc_ambi", C.c_int32),
        ("zdrop", C.c_int32), ("min_dp_max", C.c_int32), ("min_ksw_len", C.c_int32),
        ("mid_occ", C.c_int32), ("min_mid_occ", C.c_int32), ("max_mid_occ", C.c_int32),
        ("mid_occ_frac", C.c_float), ("q_occ_frac", C.c_float), ("mask_level", C.c_float),
        ("mask_len", C.c_int32), ("seed", C.c_int32), ("ext_bw", C.c_int32), ("max_sw_cells", C.c_int32),
    ]  # fmt: skip


HIT_FIELDS = (
    ("asm_id", np.int32), ("gene", np.int32), ("q_start", np.int32), ("q_end", np.int32),
    ("t_ctg", np.int32), ("t_len", np.int32), ("t_start", np.int32), ("t_end", np.int32),
    ("strand", np.int8), ("score", np.int32), ("matches", np.int32), ("block_len", np.int32),
    ("edit_distance", np.int32), ("mapq", np.uint8), ("is_primary", np.uint8),
    ("cigar_off", np.int64), ("n_cigar", np.int32),
)  # fmt: skip


class KbHits(C.Structure):
    _fields_ = [("capacity", C.c_int64)] + [(n, C.c_void_p) for n, _ in HIT_FIELDS]


# every symbol include/kaptive_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
_SYMBOLS = [
    ("kb_params_default", None, [C.POINTER(KbParams)]),
    ("kb_last_error", C.c_char_p, []),
    ("kb_version", C.c_int, []),
    ("kb_device_count", C.c_int, []),
    ("kb_fasta_count", C.c_int, [_P, C.c_int64, _P, _P]),
    ("kb_fasta_parse", C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, C.c_int64, _P, _P]),
    ("kb_fasta_ingest_count", C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P]),
    ("kb_fasta_ingest_parse", C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P]),
    ("kb_fasta_ingest_count_records", C.c_int, [_P, _P, C.c_int32, C.c_int32, _P]),
    ("kb_packed_layout", C.c_int, [_P, C.c_int64, _P, _P]),
    ("kb_fasta_ingest_lengths", C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P]),
    ("kb_fasta_ingest_pack", C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P, C.c_int64, _P, _P, C.c_int32]),
    ("kb_index_create", C.c_int, [_P, _P, _P, C.c_int32, C.POINTER(KbParams), C.c_int, C.POINTER(_P)]),
    ("kb_index_destroy", None, [_P]),
    ("kb_index_n_genes", C.c_int32, [_P]),
    ("kb_index_n_minimizers", C.c_int64, [_P]),
    ("kb_index_serialized_size", C.c_int64, [_P]),
    ("kb_index_serialize", C.c_int, [_P, _P, C.c_int64]),
    ("kb_index_deserialize", C.c_int, [_P, C.c_int64, C.c_int, C.POINTER(_P)]),
    ("kb_batch_create", C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int, C.POINTER(_P)]),
    ("kb_batch_create_packed", C.c_int, [_P, _P, C.c_int64, _P, _P, C.c_int32, C.c_int, C.POINTER(_P)]),
    ("kb_batch_download_packed", C.c_int, [_P, _P, _P, _P]),
    ("kb_batch_destroy", None, [_P]),
    ("kb_batch_n_assemblies", C.c_int32, [_P]),
    ("kb_batch_total_bases", C.c_int64, [_P]),
    ("kb_batch_packed_bytes", C.c_int64, [_P]),
    ("kb_map_batch", C.c_int, [_P, _P, C.POINTER(_P)]),
    ("kb_result_size", C.c_int, [_P, _P, _P]),
    ("kb_result_fetch", C.c_int, [_P, C.POINTER(KbHits), _P, C.c_int64]),
    ("kb_result_destroy", None, [_P]),
    ("kb_result_stage_ms", C.c_int, [_P, _P]),
    ("kb_result_counters", C.c_int, [_P, _P]),
    ("kb_result_fetch_anchors", C.c_int, [_P, _P, C.c_int64, _P]),
    ("kb_result_fetch_chains", C.c_int, [_P, _P, C.c_int64, _P]),
    ("kb_result_mid_occ", C.c_int, [_P, _P]),
    ("kb_release_workspace", C.c_int, [C.c_int]),
    ("kb_debug_dp_stats", C.c_int, [_P, C.c_int]),
    ("kb_debug_dp", C.c_int, [C.POINTER(KbParams), C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, C.c_int32]),
    ("kb_map_assemblies", C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.POINTER(KbHits), _P, _P, C.c_int64, _P]),
    ("kb_map_assemblies_packed", C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.POINTER(KbHits), _P, _P, C.c_int64, _P]),
    ("kb_scan_minimizers", C.c_int, [_P, _P, C.c_int32, _P, _P, _P, C.c_int64, _P]),
    ("kb_bench_scan", C.c_int, [_P, _P, C.c_int, _P, _P]),
    ("kb_type_last_error", C.c_char_p, []),
    ("kb_typedb_create", C.c_int, [C.c_int32, _P, _P, _P, _P, _P, C.c_int32, _P, C.c_int32, _P, _P, C.c_double, C.c_int, C.POINTER(_P)]),
    ("kb_typedb_destroy", None, [_P]),
    ("kb_type_score", C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_double, C.c_int32, _P, _P]),
    ("kb_type_call", C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, _P, _P, C.c_int32, C.c_double,
                               C.c_int32, C.c_int32, C.c_int32, C.POINTER(_P)]),
    ("kb_typed_destroy", None, [_P]),
    ("kb_type_debug_times", None, [_P]),
    ("kb_typed_sizes", C.c_int, [_P, _P, _P, _P]),
    ("kb_typed_fetch_assemblies", C.c_int, [_P] * 11),
    ("kb_typed_fetch_gene_hits", C.c_int, [_P] * 14),
    ("kb_typed_fetch_pieces", C.c_int, [_P] * 6),
    ("kb_post_last_error", C.c_char_p, []),
    ("kb_post_extract", C.c_int, [_P, C.c_int64, _P, C.c_int32, _P, _P, _P, _P, C.c_int32, _P, C.c_int64, _P, _P]),
    ("kb_post_translate", C.c_int, [_P, C.c_int64, _P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int64, _P, _P, _P]),
    ("kb_post_protein_align", C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    ("kb_post_cull_overlaps", C.c_int, [_P, _P, _P, _P, _P, C.c_double, _P, C.c_int32, _P]),
    ("kb_post_cluster", C.c_int, [_P, _P, _P, C.c_int32, _P, _P, C.c_int32, _P]),
]  # fmt: skip

_lib = None


def declared_symbols() -> list[str]:
    return [s[0] for s in _SYMBOLS]


def load() -> C.CDLL:
    """dlopen the library and bind every declared symbol; raises if the build is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not SO_PATH.exists():
        raise KbError(
            f"{SO_PATH} not found: build it with `python -m kaptive_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback for the mapping path."
        )
    L = C.CDLL(str(SO_PATH))
    for name, res, args in _SYMBOLS:
        fn = getattr(L, name)  # AttributeError here means header and library disagree
        fn.restype = res
        if args is not None:
            fn.argtypes = args
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = load().kb_last_error()
        raise KbError(f"libkaptive_b200 status {rc}: {msg.decode() if msg else ''}")


def ptr(a: np.ndarray | None) -> C.c_void_p:
    if a is None:
        return C.c_void_p(0)
    return a.ctypes.data_as(C.c_void_p)


def default_params(**over) -> KbParams:
    p = KbParams()
    load().kb_params_default(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p
