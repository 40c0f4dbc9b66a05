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
    for f in files:   # names relative to the package: the fingerprint must not depend on where the tree is checked out
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: liboryon_b200.so cannot be built (there is no non-CUDA fallback)")
    return cand


def _is_current(fp: str) -> bool:
    return os.path.exists(LIB_PATH) and os.path.exists(STAMP) and open(STAMP).read().strip() == fp


def is_current() -> bool:
    """True when the in-tree library was built from the sources as they are now (``oryon_b200._lib.load`` refuses a stale one)."""
    return _is_current(_fingerprint())


def build(force: bool = False, verbose: bool = False, jobs: int = 0) -> str:
    """Compile every .cu under csrc/ (one nvcc per file, in parallel) and link them into one shared library; returns its path.
    Safe under several processes at once (torchrun ranks importing the package with no library yet): the build runs under an
    exclusive file lock, into temporary files, and the finished library is moved into place atomically -- a rank can never
    dlopen a half-written file; the ranks that waited for the lock find the library current and return."""
    import fcntl
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB_DIR, exist_ok=True)
    fp = _fingerprint()
    if not force and _is_current(fp):
        return LIB_PATH
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _is_current(fp):           # another process built it while this one waited
                return LIB_PATH
            nvcc = nvcc_path()
            obj_dir = os.path.join(LIB_DIR, f".obj.{os.getpid()}")
            os.makedirs(obj_dir, exist_ok=True)
            compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
            srcs = _sources()
            objs = [os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o") for src in srcs]

            def one(job):
                src, obj = job
                cmd = [nvcc, *compile_flags, *(["-Xptxas", "-v"] if verbose else []), "-c", "-o", obj, src]
                return src, subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)

            try:
                with ThreadPoolExecutor(jobs or min(len(srcs), os.cpu_count() or 1)) as pool:
                    results = list(pool.map(one, zip(srcs, objs)))
                log = "".join(f"== {os.path.basename(src)}\n{proc.stdout}" for src, proc in results if proc.stdout)
                if verbose or any(proc.returncode for _, proc in results):
                    sys.stderr.write(log)
                bad = [os.path.basename(src) for src, proc in results if proc.returncode]
                if bad:
                    raise RuntimeError(f"nvcc failed on {bad} building liboryon_b200.so:\n" + log[-6000:])
                tmp = LIB_PATH + f".tmp.{os.getpid()}"
                link = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp, *objs, "-lcudart"],
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                if link.returncode != 0:
                    raise RuntimeError("link of liboryon_b200.so failed:\n" + link.stdout[-4000:])
                os.replace(tmp, LIB_PATH)
                with open(STAMP + ".tmp", "w") as fh:
                    fh.write(fp)
                os.replace(STAMP + ".tmp", STAMP)
            finally:
                shutil.rmtree(obj_dir, ignore_errors=True)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
