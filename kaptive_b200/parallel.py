"""Multi-GPU plumbing for the mapping path.

Assemblies are independent units and the gene index is read-only, so the path shards with NO data-path
collective: rank r maps a contiguous block of assemblies.  The only communication is one broadcast of the
serialized gene index at start-up (rank 0 builds it once) and, if the caller wants everything on one rank, a
gather of the small per-assembly hit arrays.  ``torch.distributed`` is plumbing here (NCCL on GPUs, gloo in
the CPU tests); no kernel in this repo communicates.
"""

from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; blocks differ in size by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_bytes(buf: np.ndarray | None, src: int = 0, device=None) -> np.ndarray:
    """Broadcast a uint8 array (the gene index image) from `src` to every rank."""
    import torch
    import torch.distributed as dist

    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    n = torch.tensor([len(buf) if dist.get_rank() == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    if dist.get_rank() == src:
        t = torch.from_numpy(np.ascontiguousarray(buf, dtype=np.uint8)).to(dev)
    else:
        t = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src)
    return t.cpu().numpy()


def gather_hits(hits: dict[str, np.ndarray], asm_offset: int, dst: int = 0) -> dict[str, np.ndarray] | None:
    """Gather per-rank hit SoA dicts on `dst`; assembly ids are shifted to global numbering first."""
    import torch.distributed as dist

    local = {k: v.copy() for k, v in hits.items()}
    if "asm_id" in local:
        local["asm_id"] = local["asm_id"] + np.int32(asm_offset)
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(local, out, dst=dst)
    if out is None:
        return None
    return {k: np.concatenate([o[k] for o in out]) for k in local}


def host_threads(cap: int = 32) -> int:
    """Host threads one process should use: this rank's share of the cores it is allowed to run on (torchrun sets LOCAL_WORLD_SIZE;
    eight ranks of one box that each start a thread per core only get in each other's way), at least 2, at most `cap`."""
    import os

    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    return max(1, min(cap, max(2 if cores > 1 else 1, cores // ranks)))
