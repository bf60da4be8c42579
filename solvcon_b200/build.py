"""Build ``libscg_cuda.so`` (strict, bit-parity) and ``libscg_cuda_fma.so`` (speed variant) in-tree
with nvcc for sm_100a.  Usage: ``python -m solvcon_b200.build [--force] [--no-fma] [-v]``.

Every source is compiled to an object file (in parallel, cached by mtime under ``csrc/_obj/``) and the
objects are linked into one shared library per variant."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
SOURCES = ["scg_core.cu", "scg_comm.cu", "scg_st.cu", "scg_e1d.cu", "scg_md.cu", "scg_mesh.cu", "scg_hmesh.cpp", "scg_hmesh_dev.cu", "scg_meshgen.cpp"]

# host code keeps the reference's x86-64 floating-point contract too: no FMA contraction
HOSTFLAGS = "-fPIC,-fvisibility=hidden,-O2,-ffp-contract=off,-pthread"
COMPILE = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", HOSTFLAGS]
LINK = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", HOSTFLAGS, "-ldl"]


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


def _run(cmd, what, verbose):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed: {what}")
    if verbose:
        sys.stderr.write(res.stderr)


def build(force=False, fma_variant=True, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "scg.h"))
    variants = [("libscg_cuda.so", "strict", ["-fmad=false", "-DSCG_STRICT_FP"])]
    if fma_variant:
        variants.append(("libscg_cuda_fma.so", "fma", ["-fmad=true"]))
    # experiment variants: SCG_EXTRA_VARIANTS="tag:-DFOO=1 -DBAR=2;tag2:..." -> libscg_cuda_<tag>.so (strict)
    for spec in filter(None, os.environ.get("SCG_EXTRA_VARIANTS", "").split(";")):
        tag, _, fl = spec.partition(":")
        variants.append((f"libscg_cuda_{tag}.so", tag, ["-fmad=false", "-DSCG_STRICT_FP"] + fl.split()))
    jobs, built = [], []
    for name, tag, flags in variants:
        for s in SOURCES:
            src = os.path.join(CSRC, s)
            if not os.path.exists(src):
                continue
            obj = os.path.join(OBJ, f"{os.path.splitext(s)[0]}.{tag}.o")
            if force or _stale(obj, [src] + headers):
                cmd = [_nvcc()] + COMPILE + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
                jobs.append((cmd, f"{s} [{tag}]"))
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(lambda j: _run(j[0], j[1], verbose), jobs))
    for name, tag, flags in variants:
        target = os.path.join(HERE, name)
        objs = [os.path.join(OBJ, f"{os.path.splitext(s)[0]}.{tag}.o") for s in SOURCES
                if os.path.exists(os.path.join(CSRC, s))]
        if force or _stale(target, objs):
            _run([_nvcc()] + LINK + objs + ["-o", target], f"link {name}", verbose)
            built.append(target)
    return built


if __name__ == "__main__":
    out = build(force="--force" in sys.argv, fma_variant="--no-fma" not in sys.argv, verbose="-v" in sys.argv)
    print("built:", out or "up to date")
