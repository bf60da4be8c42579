"""The C-ABI boundary without a GPU: ``libscg_cuda.so`` loads, exports every symbol ``include/scg.h`` declares (and the
ctypes table of solvcon_b200/_lib.py covers them), compute entries fail loudly -- not silently on a CPU path -- when no
device is present, and nothing under ``solvcon_b200/`` imports the test oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "scg.h")).read()
    return sorted(set(re.findall(r"SCG_API\s+[\w\s\*]+?\b(scg_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    names = _declared()
    assert len(names) > 60
    for so in ("libscg_cuda.so", "libscg_cuda_fma.so"):
        path = os.path.join(ROOT, "solvcon_b200", so)
        if not os.path.exists(path):
            pytest.skip(f"{so} not built")
        lib = ctypes.CDLL(path)
        missing = [n for n in names if not hasattr(lib, n)]
        assert not missing, (so, missing)


def test_ctypes_table_covers_the_header():
    from solvcon_b200 import _lib
    table = set(_lib.SIGNATURES) if hasattr(_lib, "SIGNATURES") else None
    if table is None:
        src = open(os.path.join(ROOT, "solvcon_b200", "_lib.py")).read()
        table = set(re.findall(r'"(scg_\w+)"\s*:', src))
    missing = [n for n in _declared() if n not in table]
    assert not missing, missing


def test_compute_entries_fail_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from solvcon_b200 import _lib
    from solvcon_b200.onedim import Euler1DCore
    with pytest.raises(Exception, match="(?i)cuda|device"):
        Euler1DCore(11, 0.1)
    assert b"" != _lib.load().scg_last_error()


def test_product_package_does_not_import_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "solvcon_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M) or "liboracle" in src or "libsolvcon_ref" in src:
                    bad.append(os.path.relpath(os.path.join(dirpath, f), ROOT))
    assert not bad, bad
