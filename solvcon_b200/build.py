"""Build ``libscg_cuda.so`` (strict, bit-parity) and ``libscg_cuda_fma.so`` (speed variant) in-tree
with nvcc for sm_100a.  Usage: ``python -m solvcon_b200.build [--force] [--no-fma]``."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["scg_core.cu", "scg_st.cu", "scg_e1d.cu", "scg_md.cu", "scg_mesh.cu"]

COMMON = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2", "-shared", "-cudart", "static",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libscg_cuda.so cannot be built")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, fma_variant=True, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "scg.h"))
    variants = [("libscg_cuda.so", ["-fmad=false", "-DSCG_STRICT_FP"])]
    if fma_variant:
        variants.append(("libscg_cuda_fma.so", ["-fmad=true"]))
    built = []
    for name, flags in variants:
        target = os.path.join(HERE, name)
        if not force and not _stale(target, deps):
            continue
        cmd = [_nvcc()] + COMMON + flags + (["-Xptxas", "-v"] if verbose else []) + srcs + ["-o", target]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed building {name}")
        if verbose:
            sys.stderr.write(res.stderr)
        built.append(target)
    return built


if __name__ == "__main__":
    out = build(force="--force" in sys.argv, fma_variant="--no-fma" not in sys.argv, verbose="-v" in sys.argv)
    print("built:", out or "up to date")
