"""Builds ``lib/libmpr_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m multimodalpromptretrieval_b200.csrc.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OUT = os.path.join(PKG, "lib", "libmpr_b200.so")
SOURCES = ["mpr_abi.cu", "token_cache.cpp"]
HEADERS = ["ptx.cuh", "topk_key.cuh", "scan_topk.cuh", "bank_build.cuh", "merge_topk.cuh", "prompt_gather.cuh", "tail.cuh", "embed_gather.cuh",
           os.path.join(ROOT, "include", "mpr_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _digest() -> str:
    """Hash of everything the library is built from (sources, headers, flags, compiler path).  A fresh checkout gives
    every file the same mtime, so staleness is decided by content, not by timestamps."""
    import hashlib
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc").encode())
    deps = [os.path.join(HERE, s) for s in SOURCES] + [x if os.path.isabs(x) else os.path.join(HERE, x) for x in HEADERS]
    for d in sorted(deps):
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale() -> bool:
    if not os.path.exists(OUT) or not os.path.exists(OUT + ".sha256"):
        return True
    with open(OUT + ".sha256") as f:
        return f.read().strip() != _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + \
          [os.path.join(HERE, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}): {' '.join(cmd)}")
    with open(OUT + ".sha256", "w") as f:
        f.write(_digest() + "\n")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
