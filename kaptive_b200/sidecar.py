"""Build-once gene index beside the compiled database.

The reference compiles a database once and caches it as ``<keyword>.pkl`` + ``<keyword>.json`` under ``~/.kaptive`` or
``$KAPTIVE_DB_DIR`` (``/root/reference/src/kaptive/db/manager.py:63-73,539-558``).  The device gene index (minimizer hash of every
gene, ~20 MB for a K database) is a pure function of the gene sequences and the mapping parameters, so it is cached the same way:
``<keyword>.kb200idx`` next to the pickle, re-used while its key (SHA-256 of the gene bytes and of the parameter block, library
version) still matches, rebuilt and rewritten otherwise.  The file is the flat image ``kb_index_serialize`` produces (the same
bytes that are broadcast over NCCL to the other ranks) behind a small JSON header."""

from __future__ import annotations

import ctypes as C
import hashlib
import json
import os
from pathlib import Path

import numpy as np

from . import _lib, mapper

MAGIC = b"KB200IDX1\n"


def _key(genes: list[bytes], params) -> dict:
    h = hashlib.sha256()
    for g in genes:
        h.update(len(g).to_bytes(4, "little"))
        h.update(g)
    return {"genes_sha256": h.hexdigest(), "n_genes": len(genes), "params_sha256": hashlib.sha256(bytes(params)).hexdigest(),
            "lib_version": int(_lib.load().kb_version())}


def save(gi: "mapper.GeneIndex", path: str | os.PathLike, genes: list[bytes]) -> None:
    img = gi.serialize()
    head = json.dumps(_key(genes, gi.params)).encode() + b"\n"
    tmp = Path(str(path) + ".tmp")
    with open(tmp, "wb") as fh:
        fh.write(MAGIC + len(head).to_bytes(4, "little") + head)
        fh.write(img.tobytes())
    os.replace(tmp, path)  # atomic: a concurrent reader sees the old file or the new one


def load(path: str | os.PathLike, genes: list[bytes], params=None, device: int = 0) -> "mapper.GeneIndex | None":
    """The cached index, or None when the file is missing, damaged or was built for other genes / parameters / library."""
    params = params or _lib.default_params()
    try:
        with open(path, "rb") as fh:
            if fh.read(len(MAGIC)) != MAGIC:
                return None
            n = int.from_bytes(fh.read(4), "little")
            head = json.loads(fh.read(n))
            if head != _key(genes, params):
                return None
            img = np.frombuffer(fh.read(), dtype=np.uint8)
        gi = mapper.GeneIndex.deserialize(img, device=device)
        gi.params = params
        return gi
    except (OSError, ValueError, _lib.KbError):
        return None


def sidecar_path(db_path: str | os.PathLike) -> Path:
    """``~/.kaptive/kpsc_k.pkl`` -> ``~/.kaptive/kpsc_k.kb200idx``"""
    p = Path(db_path)
    return p.with_suffix(".kb200idx")


def index_for(db_path: str | os.PathLike, genes: list[bytes], params=None, device: int = 0) -> tuple["mapper.GeneIndex", bool]:
    """The gene index of the database compiled at `db_path`: loaded from its sidecar when that is current, else built and cached.
    Returns (index, was_cached).  An unwritable directory is not an error: the index is simply rebuilt next time."""
    params = params or _lib.default_params()
    side = sidecar_path(db_path)
    gi = load(side, genes, params, device)
    if gi is not None:
        return gi, True
    gi = mapper.GeneIndex(genes, params=params, device=device)
    try:
        save(gi, side, genes)
    except OSError:
        pass
    return gi, False
