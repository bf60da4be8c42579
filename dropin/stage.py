"""Stage the reference's UNMODIFIED Python package and hot-path tests next to the ``_solvcon`` shim.

``python dropin/stage.py`` (also run by ``__graft_entry__.build()`` when ``/root/reference`` exists) copies, byte for
byte and without editing anything,

* ``/root/reference/solvcon/**/*.py``  ->  ``dropin/_stage/solvcon/``    (the reference's whole Python package)
* the hot-path test files of ``/root/reference/tests`` and the three ``.msh`` fixtures they read
  ->  ``tests/_refsuite/``

Both targets are git-ignored (like ``oracle/_ref``: nothing of the reference enters this repository's history) but not
gpurun-ignored, so they travel to the GPU box, where ``pytest tests -m gpu`` runs the reference's own tests against
``libscg_cuda.so`` through ``dropin/_solvcon`` (``tests/conftest.py`` puts ``dropin`` and ``dropin/_stage`` on
``sys.path`` and marks the staged device tests ``gpu``).  A ``MANIFEST.json`` with the SHA-256 of every source file is
written next to the copies so a reader can check they are verbatim.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = os.environ.get("SCG_REFERENCE", "/root/reference")
STAGE_PKG = os.path.join(HERE, "_stage")
STAGE_TESTS = os.path.join(ROOT, "tests", "_refsuite")

# The reference's tests of the hot path (the classes SURVEY.md 8b lists and the callers of 8f)
TEST_FILES = [
    "test_onedim_euler.py",
    "test_spacetime_grid.py", "test_spacetime_celm_selm.py", "test_spacetime_solution.py",
    "test_spacetime_linear_scalar.py", "test_spacetime_inviscid_burgers.py",
    "test_mesh.py", "test_mesh_2d.py", "test_gmsh.py",
    "test_multidim.py", "test_euler_multidim.py", "test_euler.py",
    "multidim/euler/test_field.py",
    "apps/obsrefl/test_driver.py", "apps/obsrefl/test_analytic.py", "apps/obsrefl/test_session.py",
]
DATA_FILES = ["data/rectangle.msh", "data/mixed_tq.msh", "data/gmsh_triangle.msh"]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def _copy(src, dst, manifest, key):
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(src, dst)
    manifest[key] = _sha(src)


def stage(reference=REFERENCE, quiet=False):
    if not os.path.isdir(os.path.join(reference, "solvcon")):
        if not quiet:
            print(f"dropin/stage.py: {reference} not present; nothing staged", file=sys.stderr)
        return False
    for d in (STAGE_PKG, STAGE_TESTS):
        shutil.rmtree(d, ignore_errors=True)
    manifest = {"reference": reference, "package": {}, "tests": {}}
    pkg = os.path.join(reference, "solvcon")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = os.path.join(base, f)
                rel = os.path.relpath(src, reference)
                _copy(src, os.path.join(STAGE_PKG, rel), manifest["package"], rel)
    for rel in TEST_FILES + DATA_FILES:
        src = os.path.join(reference, "tests", rel)
        _copy(src, os.path.join(STAGE_TESTS, rel), manifest["tests"], "tests/" + rel)
    for d, m in ((STAGE_PKG, manifest), (STAGE_TESTS, manifest)):
        with open(os.path.join(d, "MANIFEST.json"), "w") as f:
            json.dump(m, f, indent=1, sort_keys=True)
    if not quiet:
        print(f"staged {len(manifest['package'])} package files and {len(manifest['tests'])} test files")
    return True


if __name__ == "__main__":
    stage()
