"""Builds libkaptive_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""

from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "_lib" / "libkaptive_b200.so"
SOURCES = ["kb_api.cu", "kb_scan.cu", "kb_pipeline.cu", "kb_post.cu", "kb_index.cpp", "kb_params.cpp", "kb_fasta.cpp", "kb_type.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
    "--extended-lambda", "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-Wno-unknown-pragmas", "-shared",
]  # fmt: skip


def needs_build() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    deps = list(CSRC.glob("*")) + [HERE.parent / "include" / "kaptive_b200.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    OUT.parent.mkdir(parents=True, exist_ok=True)
    import os
    import shlex

    extra = shlex.split(os.environ.get("KB_NVCC_EXTRA", ""))  # experiments, e.g. -DKB_ALIGN_MINB=4
    cmd = ["nvcc", *NVCC_FLAGS, *extra, "-o", str(OUT)] + [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
