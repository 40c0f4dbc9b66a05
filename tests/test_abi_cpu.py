"""CPU: the C-ABI library builds, loads and exports every symbol include/oryon_b200.h declares; the
ctypes binding covers the same set; nothing under oryon_b200/ reaches for the oracle or the reference."""
import ctypes
import os
import re

import pytest

from oryon_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "oryon_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oryon_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _header_functions()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/oryon_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding and header disagree"


def test_binding_loads_and_reports_abi_version():
    lib = _lib.load()
    assert lib.oryon_abi_version() == _lib.ABI_VERSION


def test_create_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.OryonError):
        _lib.handle(0)


def test_product_code_never_imports_the_oracle_or_the_reference():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "oryon_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                if re.search(r"^\s*(import|from)\s+(oryon_oracle|backbone_oracle|eval_oracle|stage_oracle|vsd_oracle|oracle|ref_shims)\b", txt, flags=re.M) or "/root/reference" in txt:
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_build_fingerprint_does_not_depend_on_the_checkout_path(tmp_path):
    """The GPU box runs the tree from a scratch path: a fingerprint over absolute file names would call the shipped library stale there
    and rebuild it on every first import (and crash under ncu, which follows the nvcc children)."""
    import shutil
    from oryon_b200 import build as b
    fp = b._fingerprint()
    root = os.path.dirname(os.path.dirname(os.path.abspath(b.__file__)))
    copy = tmp_path / "elsewhere"
    shutil.copytree(os.path.join(root, "oryon_b200", "csrc"), copy / "oryon_b200" / "csrc")
    shutil.copytree(os.path.join(root, "include"), copy / "include")
    old = (b.PKG, b.CSRC)
    try:
        b.PKG, b.CSRC = str(copy / "oryon_b200"), str(copy / "oryon_b200" / "csrc")
        assert b._fingerprint() == fp
    finally:
        b.PKG, b.CSRC = old
