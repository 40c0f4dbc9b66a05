"""Build recipe for liboryon_b200.so (hand-written sm_100a CUDA + the C ABI in include/oryon_b200.h).

``python -m oryon_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a GPU.
The library is built IN-TREE (oryon_b200/lib/) so that it travels to the GPU box with the repo
snapshot; it is git-ignored.  It links against cudart only (the driver's cuTensorMapEncodeTiled is
resolved at run time through cudaGetDriverEntryPoint), so it has no torch dependency.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "liboryon_b200.so")
STAMP = os.path.join(LIB_DIR, "liboryon_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    root = os.path.dirname(PKG)
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(root, "include", "oryon_b200.h")]
    for f in files:
        with open(f, "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: liboryon_b200.so cannot be built (there is no non-CUDA fallback)")
    return cand


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into one shared library; returns its path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP) and open(STAMP).read().strip() == fp:
        return LIB_PATH
    cmd = [nvcc_path(), *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", LIB_PATH, *_sources(), "-lcudart"]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building liboryon_b200.so:\n" + proc.stdout[-4000:])
    with open(STAMP, "w") as fh:
        fh.write(fp)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
