#!/usr/bin/env python
"""bench.py -- throughput of the CESE march on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one ``march(chunk)`` call of the solver over the whole mesh/grid (``chunk`` time
steps = ``2*chunk`` half-steps/substeps), the call a user of the reference makes
(``svr.march_alpha2(steps)`` / ``EulerCore.march(steps)``).  Unit everywhere: one *update* = one
solution point (1D) / one body cell (multidim) advanced by one half-step / substep.

* ``value``       device-resident throughput (inputs already in HBM), CUDA events, max over ranks
* ``e2e``         the same metric through the public solver class with HOST arrays: upload +
                  march + download inside the timed region
* ``roofline``    algorithmic bytes per launch / average launch time of the dominant kernel
* ``cpu_baseline`` the reference's own CPU march (oracle/_ref, else the pinned C port) timed on
                  this box's host cores on a bounded sample of the same workload

``--impl reference`` runs only the CPU reference arm and prints its line.  The reference march is single-threaded, so
``value`` is one thread; ``cpu_baseline.all_cores`` adds the summed rate of one independent replica per host thread.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "fp64 CESE cell-updates/s (Euler)"
UNIT = "updates/s"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------- clocks ---

class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        return False

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------- workloads ---

class E1DWorkload:
    """1D Euler CESE (onedim.Euler1DCore), Sod shock tube on ncoord points."""

    name = "e1d_sod"
    dtype = "f64"
    balg = 128  # bytes per update, SURVEY.md 8(d)

    def __init__(self, ncoord, chunk):
        self.ncoord = int(ncoord) | 1
        self.chunk = chunk
        self.dt = 0.8 * 2.0 / (self.ncoord - 1)

    def describe(self, world):
        return {"workload": f"1D Euler CESE (onedim), Sod shock tube, ncoord={self.ncoord} per GPU, "
                            f"march_alpha2({self.chunk}) per step",
                "points_per_gpu": self.ncoord, "chunk_steps": self.chunk,
                "l2_policy": "inputs larger than L2 (%.1f GB state per GPU)" % (self.ncoord * 128 / 1e9),
                "decomposition": "none" if world == 1 else "independent slabs (NRBC ends are local)"}

    def updates_per_step(self):
        # each half step writes (N-3)/2 (+1 on the odd plane) targets
        return (self.ncoord - 3) * self.chunk + self.chunk

    def launches_per_step(self):
        return 2 * self.chunk

    def alg_bytes_per_launch(self):
        return self.balg * self.updates_per_step() / self.launches_per_step()

    def build_device(self):
        from cases import sod_ic
        from solvcon_b200.onedim import Euler1DCore
        self.dev = Euler1DCore(self.ncoord, self.dt)
        sod_ic(self.dev)
        return self.dev

    def resident_begin(self):
        self.dev.resident(True)

    def resident_step(self):
        self.dev.march_alpha2(self.chunk)

    def e2e_step(self):
        # public API with host arrays: upload (coord, gamma, so0, so1, cfl) + march + download
        self.dev.resident(False) if self.dev._resident else None
        self.dev.march_alpha2(self.chunk)

    def e2e_bytes(self):
        n = self.ncoord
        return n * 8 * (1 + 1 + 3 + 3 + 1), n * 8 * (3 + 3 + 1)

    def result_check(self):
        return float(self.dev.so0[:, 0].sum())

    # CPU reference arm on a bounded sample
    def cpu_run(self, budget_s=15.0):
        from oracle import bridge
        from cases import sod_ic
        n = min(self.ncoord, 2_000_001)
        kind = "reference" if bridge.have_ref() else "port"
        cls = bridge.RefEuler1D if bridge.have_ref() else bridge.PortEuler1D
        chk = cls(n, 0.8 * 2.0 / (n - 1))
        sod_ic(chk)
        chk.march(2, 1)
        t0 = time.perf_counter()
        steps = 0
        while True:
            chk.march(2, 2)
            steps += 2
            el = time.perf_counter() - t0
            if el > budget_s or steps >= 64:
                break
        upd = (n - 3) * steps + steps
        return {"value": upd / el, "unit": UNIT, "cores": 1, "kind": kind,
                "sample": f"1D Euler Sod, ncoord={n}, march_alpha2({steps}) in {el:.2f} s, single thread "
                          f"(the reference is single-threaded)"}, upd, el


def make_workload(args):
    if args.workload == "e1d":
        return E1DWorkload(args.size or 100_000_001, args.chunk)
    if args.workload == "sod":
        # BASELINE.json configs[0]: 1e4 points x 1000 steps -- launch-latency bound, reported as a latency figure
        wl = E1DWorkload(args.size or 10_001, args.chunk if args.chunk >= 100 else 1000)
        wl.name = "e1d_sod_config0"
        wl.dt = 1.6e-4
        return wl
    if args.workload in ("st1d", "st1d_burgers"):
        return ST1DWorkload(args.size or 1_000_000_001, args.chunk, eq=1 if args.workload.endswith("burgers") else 0)
    if args.workload in ("md3d", "md2d", "md3d_strong"):
        from solvcon_b200 import bench_md  # noqa
        return bench_md.make(args)
    raise SystemExit(f"unknown workload {args.workload}")


class ST1DWorkload:
    """1D scalar CESE (spacetime.LinearScalarSolver / InviscidBurgersSolver), BASELINE.json configs[1]: a periodic
    grid of ``ncelm`` conservation elements PER GPU (2*ncelm+1 solution points), so0 = 1/2 sin x, dt = dx/2; with
    more than one GPU the ranks are the slabs of one ring (solvcon_b200/decomp1d.py)."""

    dtype = "f64"
    balg = 56  # bytes per update, SURVEY.md 8(d)

    def __init__(self, npoints, chunk, eq=0):
        self.ncelm = max(4, int(npoints) // 2)
        self.chunk = chunk
        self.eq = int(eq)
        self.name = "st1d_burgers" if self.eq else "st1d_linear"
        self.dominant_kernel = "st_half_kernel<%s, alpha 2> (one launch per half step)" % ("burgers" if self.eq else "linear")

    def describe(self, world):
        return {"workload": f"1D scalar CESE (spacetime.{'InviscidBurgers' if self.eq else 'LinearScalar'}Solver), periodic, "
                            f"{2 * self.ncelm + 1} solution points per GPU, so0 = sin(x)/2, dt = dx/2, "
                            f"march_alpha2({self.chunk}) per step",
                "points_per_gpu": 2 * self.ncelm + 1, "chunk_steps": self.chunk,
                "l2_policy": "inputs larger than L2 (%.1f GB state per GPU)" % (self.ncelm * 2 * 32 / 1e9),
                "decomposition": "none" if world == 1 else
                "contiguous slabs on a ring, one per GPU; {so0, so1} of the two odd-plane end points exchanged with "
                "both neighbours by ncclSend/ncclRecv from the library once per step (32 B per rank per step)"}

    def updates_per_step(self):
        return (2 * self.ncelm + 1) * self.chunk

    def launches_per_step(self):
        return 2 * self.chunk

    def alg_bytes_per_launch(self):
        return self.balg * self.updates_per_step() / self.launches_per_step()

    def _ic(self, x, so0, so1):
        so0[2::2] = 0.5 * np.sin(x[2::2])
        so1[2::2] = 0.5 * np.cos(x[2::2])

    def build_device(self):
        from solvcon_b200 import spacetime as st
        world, rank = getattr(self, "world", 1), getattr(self, "rank", 0)
        n = self.ncelm
        # dx = 2^-14 exactly: the reference's Grid fills xloc by cumulative addition (core.cpp:36-43), which at 5e8
        # elements drifts past xmax for a general dx (its own constructor then rejects the grid); with a power-of-two
        # spacing every partial sum is exact.  ~1e5 points per wave length of the initial sine at every size.
        span = n * 2.0 ** -14
        self.dt = 0.5 * span / n
        self.ring = None
        if world > 1:
            # weak scaling: the ring is `world` copies of the span; every rank generates its own slab's coordinates
            # (bit-parity of the decomposition against one global grid is what tests/test_gpu_1d.py checks)
            from solvcon_b200.decomp1d import NcclRing
            xs = st.Grid(rank * span, (rank + 1) * span, n).xcoord.ndarray
            self.ring = NcclRing(self.eq, xs, 0, n, self.dt, rank, world)   # the library's own communicator
            self.dev = self.ring.slab.svr
        else:
            grid = st.Grid(0.0, span, n)
            self.dev = (st.InviscidBurgersSolver if self.eq else st.LinearScalarSolver)(grid, self.dt)
        self._ic(self.dev.grid.xcoord.ndarray, self.dev.so0.ndarray[:, 0], self.dev.so1.ndarray[:, 0])
        return self.dev

    def resident_begin(self):
        self.dev.resident(True)

    def resident_step(self):
        if self.ring is not None:
            self.ring.march(2, self.chunk)
        else:
            self.dev.march_alpha2(self.chunk)

    def e2e_step(self):
        if self.ring is not None:
            self.dev._push(True)
            self.ring.march(2, self.chunk)
            self.dev.sync_to_host()
        else:
            self.dev.march_alpha2(self.chunk)

    def e2e_bytes(self):
        nx = 2 * self.ncelm + 5
        return nx * 8 * 4, nx * 8 * 3

    def result_check(self):
        return float(self.dev.so0.ndarray[2::2, 0].sum())

    def cpu_run(self, budget_s=15.0):
        from oracle import bridge
        n = min(self.ncelm, 2_000_000)
        kind = "reference" if bridge.have_ref() else "port"
        cls = bridge.RefScalar1D if bridge.have_ref() else bridge.PortScalar1D
        span = n * 2.0 ** -14
        chk = cls(self.eq, 0.0, span, n, 0.5 * span / n)
        chk.so0[...] = 0.0
        chk.so1[...] = 0.0
        chk.cfl[...] = 0.0
        self._ic(chk.x, chk.so0, chk.so1)
        chk.march(2, 1)
        t0 = time.perf_counter()
        steps = 0
        while True:
            chk.march(2, 2)
            steps += 2
            el = time.perf_counter() - t0
            if el > budget_s or steps >= 200:
                break
        upd = (2 * n + 1) * steps
        return {"value": upd / el, "unit": UNIT, "cores": 1, "kind": kind,
                "sample": f"1D {'Burgers' if self.eq else 'linear'} scalar, ncelm={n}, march_alpha2({steps}) in {el:.2f} s, "
                          f"single thread (the reference is single-threaded)"}, upd, el


# --------------------------------------------------------------------------------- main ---

def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # the data path's communicator lives in the library (csrc/scg_comm.cu); torch.distributed carries its unique
        # id and keeps the bench's own plumbing (barrier, max over ranks)
        from solvcon_b200 import _lib
        from solvcon_b200.decomp import comm_init
        _lib.init(local)
        comm_init(rank, world, dist=dist)
        return rank, world, local, dist
    return rank, world, local, None


def barrier(dist, local):
    if dist is not None:
        import torch
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()


def max_over_ranks(dist, v, local):
    if dist is None:
        return v
    import torch
    t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, v, local):
    if dist is None:
        return v
    import torch
    t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def run_replica(args):
    """``--impl reference-replica``: one bounded sample of the reference march, printed as a bare JSON object (child of
    :func:`all_core_replicas`)."""
    wl = make_workload(args)
    _, upd, el = wl.cpu_run(budget_s=args.ref_budget)
    print(json.dumps({"upd": upd, "el": el}))


def all_core_replicas(args):
    """The reference's march is single-threaded, so one solver cannot use more than one core; what a host CAN do with all
    its cores is run independent solvers side by side.  Start one replica of the same bounded sample per host thread
    (separate processes), and report the sum of their rates: an upper bound for a perfectly threaded reference, memory
    contention included.  Extra information in ``cpu_baseline.all_cores``; the line's ``value`` stays the real thing."""
    n = args.ref_replicas
    if n < 0:
        n = min(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1), 128)
    if n <= 1:
        return None
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference-replica", "--workload", args.workload,
           "--size", str(args.size), "--ref-budget", str(args.ref_budget)]
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", RANK="0")
    procs = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env) for _ in range(n)]
    rate, done = 0.0, 0
    deadline = time.monotonic() + 4 * args.ref_budget + 180     # one deadline for the whole batch
    for p in procs:
        try:
            out, _ = p.communicate(timeout=max(1.0, deadline - time.monotonic()))
            d = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
            rate += d["upd"] / d["el"]
            done += 1
        except Exception:   # a replica that failed or overran is left out of the sum (and of ``cores``)
            p.kill()
    if not done:
        return None
    return {"value": rate, "unit": UNIT, "cores": done,
            "sample": f"{done} concurrent independent replicas of the same bounded sample, one single-threaded reference "
                      f"march per host thread; sum of the replicas' rates"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    vals, samples = [], None
    # each step is a bounded sample; the whole run stays within a few minutes whatever K and W the caller picks
    args.ref_budget = min(args.ref_budget, max(1.0, 150.0 / max(args.warmup + args.steps, 1)))
    for i in range(args.warmup + args.steps):
        base, upd, el = wl.cpu_run(budget_s=args.ref_budget)
        if i >= args.warmup:
            vals.append((upd, el))
            samples = base
    upd = sum(v[0] for v in vals)
    el = sum(v[1] for v in vals)
    single = upd / el
    samples["value"] = single
    # The reference's march is single-threaded.  "All the host threads it can use" is therefore one independent
    # replica of the same bounded sample per host thread: the line's value is the SUM of their rates (an upper bound
    # for a perfectly threaded reference, memory contention included); the single-thread figure stays next to it.
    allc = all_core_replicas(args)
    value, cores = single, 1
    if allc:
        samples["all_cores"] = allc
        value, cores = allc["value"], allc["cores"]
    cpu = {"value": value, "unit": UNIT, "cores": cores, "kind": samples["kind"],
           "sample": samples["sample"] + (f"; value = sum over {cores} concurrent independent replicas, one per host thread"
                                          if cores > 1 else ""),
           "single_thread": single}
    if allc:
        cpu["all_cores"] = allc
    cfg = wl.describe(world)
    cfg["reference_sample_note"] = ("the reference arm marches a bounded SAMPLE of this configuration (same generator, "
                                    "initial and boundary conditions, at the size named in cpu_baseline.sample), not the "
                                    "full-size mesh: the sample is cache-resident, which favours the reference")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / max(len(vals), 1) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype,
            "data": "synthetic", "config": cfg, "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def measure(wl, steps, warmup, args, rank, world, local, dist, e2e=True, cpu=False):
    """Build the workload and measure it: device-resident arm (CUDA events on the library stream, max over ranks),
    roofline of the dominant kernel, end-to-end arm with host arrays, field health, optional CPU baseline."""
    from solvcon_b200 import _lib
    wl.build_device()
    barrier(dist, local)
    steps_marched = 0
    chunk = wl.chunk

    # ---- device-resident arm
    wl.resident_begin()
    for _ in range(warmup):
        wl.resident_step()
    _lib.sync()
    barrier(dist, local)
    _lib.launch_count(reset=True)
    with ClockSampler(local) as clk:
        with _lib.Timer() as t:
            for _ in range(steps):
                wl.resident_step()
        _lib.sync()
    barrier(dist, local)
    launches = _lib.launch_count()
    steps_marched += (warmup + steps) * chunk
    ms = max_over_ranks(dist, t.ms, local)
    upd_total = sum_over_ranks(dist, float(wl.updates_per_step()), local) * steps
    value = upd_total / (ms * 1e-3)

    # ---- roofline of the dominant kernel (this rank's launches, timed above on the library stream)
    peak, peak_src = measured_peak()
    n_dom = wl.launches_per_step() * steps
    ach = wl.alg_bytes_per_launch() / (t.ms * 1e-3 / n_dom) / 1e9
    roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": getattr(wl, "ncu_traffic", None), "peak_source": peak_src,
            "kernel": getattr(wl, "dominant_kernel", "e1d_half_kernel"),
            "alg_bytes_per_launch": wl.alg_bytes_per_launch(),
            "avg_launch_ms": t.ms / n_dom}
    clocks = clk.summary()
    roof64 = wl.roofline_fp64(t.ms * 1e-3 / n_dom, clocks) if callable(getattr(wl, "roofline_fp64", None)) else None

    # ---- end-to-end arm through the public class with host arrays
    e2e_res = None
    if e2e:
        if hasattr(wl, "e2e_begin"):
            wl.e2e_begin()
        else:
            if hasattr(wl.dev, "materialise"):
                wl.dev.materialise()
            if getattr(wl, "dist_svr", None) is None:
                wl.dev.resident(False)
        wl.e2e_step()
        if hasattr(wl, "e2e_begin"):
            wl._b0, wl._e2e_steps = (wl.dev.h2d_bytes, wl.dev.d2h_bytes), 0     # count the timed steps only
        n_e2e = max(1, min(steps, 3))
        barrier(dist, local)
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            wl.e2e_step()
        _lib.sync()
        barrier(dist, local)
        el = max_over_ranks(dist, time.perf_counter() - t0, local)
        hi, lo = wl.e2e_bytes()
        steps_marched += (1 + n_e2e) * chunk
        e2e_res = {"value": sum_over_ranks(dist, float(wl.updates_per_step()), local) * n_e2e / el, "unit": UNIT,
                   "h2d_bytes_per_step": hi, "d2h_bytes_per_step": lo, "steps": n_e2e,
                   "check": wl.result_check()}
        if hasattr(wl, "e2e_fields_only") and world == 1:
            alt = wl.e2e_fields_only()
            if alt:
                e2e_res["fields_only"] = alt
                steps_marched += (1 + alt["steps"]) * chunk

    # ---- the field every timed step ran on must still be a flow field: finite everywhere, density positive
    check = sum_over_ranks(dist, wl.result_check(), local)
    health = None
    if hasattr(wl, "health"):
        health = wl.health()
        health["cflo_max"] = max_over_ranks(dist, health["cflo_max"], local)
        health["rho_min"] = -max_over_ranks(dist, -health["rho_min"], local)
        health["p_min"] = -max_over_ranks(dist, -health["p_min"], local)
        health["finite"] = bool(max_over_ranks(dist, 0.0 if health["finite"] else 1.0, local) == 0.0)
        health["steps_marched"] = steps_marched
    bad = (not np.isfinite(check)) or (health is not None and (not health["finite"] or not health["rho_min"] > 0.0))

    cpu_res = None
    if cpu:
        cpu_res, _, _ = wl.cpu_run(budget_s=args.ref_budget)
    config = wl.describe(world)
    if health is not None and "mesh" in config:
        config["mesh"].update(health)
    return {"value": value, "ms_per_step": ms / steps, "scaling": "strong" if getattr(wl, "strong", False) else "weak",
            "dtype": wl.dtype, "config": config, "clocks": clocks, "e2e": e2e_res, "gpu_launches": int(launches),
            "roofline": roof, "roofline_fp64": roof64, "cpu_baseline": cpu_res, "check": check, "health": health,
            "bad": bool(bad)}


def run_subs(args, rank, world, local, dist):
    """The other configurations BASELINE.json names, measured in the same run right after the headline workload and
    attached to its JSON line as ``sub_results`` (device-resident arm + roofline; each bounded to a few steps).
    One GPU: configs[1] (1D scalar, 1e9 points, linear and Burgers), the 1D Euler core at 1e8 points, configs[0]
    (Sod tube, 1e4 points x 1000 steps, a latency number: wall time of the whole march), configs[2] (2D, 2e7 cells).
    More than one GPU: configs[3], the strong-scaling case (ONE 1e8-cell mesh cut into one part per rank)."""
    import copy
    import gc
    out = []

    def one(workload, steps, chunk, size=0, e2e=False, extra=None, order=None, tag=None):
        a = copy.copy(args)
        a.workload, a.size, a.chunk = workload, size, chunk
        if order is not None:
            a.order = order
        t0 = time.perf_counter()
        try:
            wl = make_workload(a)
            wl.rank, wl.world, wl.dist, wl.local = rank, world, dist, local
            par = wl.parity() if (rank == 0 and world == 1 and hasattr(wl, "parity") and not args.no_parity and not tag) else None
            r = measure(wl, steps, 3, a, rank, world, local, dist, e2e=e2e, cpu=False)
            entry = {"workload": wl.name, "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"],
                     "steps": steps, "warmup": 3, "scaling": r["scaling"], "n_gpus": world,
                     "roofline": {k: r["roofline"][k] for k in ("achieved", "peak", "frac", "unit", "kernel")},
                     "clocks": r["clocks"], "check": r["check"], "config": r["config"], "gpu_launches": r["gpu_launches"],
                     "bad": r["bad"]}
            if r.get("roofline_fp64"):
                entry["roofline_fp64"] = r["roofline_fp64"]
            if r["e2e"]:
                entry["e2e"] = r["e2e"]
            if par is not None:
                entry["parity"] = par
            if extra:
                entry.update(extra(wl))
            if tag:
                entry["workload"] = tag
            del wl
        except Exception as exc:   # a sub-result that cannot run says so; the headline line is still printed
            entry = {"workload": workload, "error": f"{type(exc).__name__}: {exc}"}
        entry["wall_s"] = round(time.perf_counter() - t0, 1)
        gc.collect()
        out.append(entry)

    if world == 1:
        one("st1d", 4, 5)
        one("st1d_burgers", 4, 5)
        one("e1d", 4, 5)
        one("sod", 3, 1000)
        one("md2d", 6, 5)
        # BASELINE configs[4]'s ordering pair: the same 3D mesh under both cell orderings (L2 hit rate and DRAM bytes
        # per update of the pair come from ncu: profiles/r02e_ordering_ncu.txt)
        for o in ("morton", "rcm"):
            one("md3d", 4, 5, size=8_000_000, order=o, tag=f"md3d_8M_{o}")
    else:
        one("md3d_strong", max(4, min(args.steps, 10)), args.chunk)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=os.environ.get("SCG_BENCH_WORKLOAD", "md3d"))
    ap.add_argument("--size", type=int, default=0, help="points / cells per GPU (0 = workload default)")
    ap.add_argument("--chunk", type=int, default=5, help="time steps per march() call (one bench step)")
    ap.add_argument("--order", default="rcm", help="multidim cell ordering: none | morton | rcm (BASELINE configs[4] compares morton and rcm: on the lattice-generated bench meshes rcm is 4-5 % faster, profiles/README.md)")
    ap.add_argument("--ref-budget", type=float, default=12.0)
    ap.add_argument("--ref-replicas", type=int, default=-1,
                    help="--impl reference: also time this many concurrent replicas (-1 = one per host thread, 0 = off)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity leg (profiling runs)")
    ap.add_argument("--no-subs", action="store_true", help="default workload only: skip the sub_results legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return
    if args.impl == "reference-replica":
        run_replica(args)
        return

    rank, world, local, dist = dist_setup(args.gpus)
    from solvcon_b200 import _lib
    _lib.init(local)
    lib = _lib.load()
    wl = make_workload(args)
    wl.rank, wl.world, wl.dist, wl.local = rank, world, dist, local
    # ---- in-run parity leg (before anything is timed): the workload's own generator, IC and BCs at a size the CPU
    # checker finishes in seconds, device march against oracle/_ref (else the pinned C port), every element compared
    parity = None
    if rank == 0 and not args.no_parity and hasattr(wl, "parity"):
        parity = wl.parity()
        if not parity["bit_equal"] and parity.get("build", "").startswith("strict"):
            print(json.dumps({"error": "parity leg failed", "parity": parity}), file=sys.stderr)
            sys.exit(4)
    # ---- multi-rank parity leg: a small case cut into one part per rank, marched over the real NCCL transport
    # (overlapped exchange), every rank's cells compared on rank 0 with the undecomposed device run
    if world > 1 and not args.no_parity and args.workload in ("md3d", "md3d_strong"):
        from solvcon_b200 import bench_md
        nccl_parity, ok = bench_md.nccl_parity_3d(rank, world, local, dist, args.order, transport="lib")
        torch_parity, ok2 = bench_md.nccl_parity_3d(rank, world, local, dist, args.order, transport="torch")
        ok = ok and ok2
        if rank == 0:
            parity = {"single_rank": parity, "decomposed": nccl_parity, "decomposed_torch_transport": torch_parity}
        if not ok and bool(lib.scg_build_is_strict()):
            if rank == 0:
                print(json.dumps({"error": "multi-rank parity leg failed", "parity": parity}), file=sys.stderr)
            dist.destroy_process_group()
            sys.exit(4)
    res = measure(wl, args.steps, args.warmup, args, rank, world, local, dist, e2e=not args.no_e2e,
                  cpu=(rank == 0 and world == 1 and not args.no_cpu))
    bad = res["bad"]

    # ---- the other named configurations (BASELINE.json configs[0..3]) in the same run, bounded: one entry each
    subs = None
    if not bad and args.workload == "md3d" and not args.no_subs and args.size == 0:
        del wl
        import gc
        gc.collect()
        subs = run_subs(args, rank, world, local, dist)
        bad = bad or any(sr.get("bad") for sr in subs)

    if rank == 0:
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": res["scaling"], "vs_baseline": None, "dtype": res["dtype"], "data": "synthetic",
                "config": res["config"], "clocks": res["clocks"], "e2e": res["e2e"],
                "gpu_launches": res["gpu_launches"], "roofline": res["roofline"], "cpu_baseline": res["cpu_baseline"],
                "strict_fp": bool(lib.scg_build_is_strict()), "parity": parity, "check": res["check"]}
        if res.get("roofline_fp64"):
            line["roofline_fp64"] = res["roofline_fp64"]
        if subs is not None:
            line["sub_results"] = subs
        if bad:
            # a march that left the physical regime (NaN / non-positive density) is not a measurement
            print(json.dumps({"error": "a marched field is not finite and positive", "check": repr(res["check"]),
                              "health": {k: repr(v) for k, v in (res["health"] or {}).items()},
                              "subs": [sr.get("workload") for sr in (subs or []) if sr.get("bad")]}), file=sys.stderr)
        else:
            print(json.dumps(line, allow_nan=False))
        if os.environ.get("SCG_TIMING") == "1":   # diagnostics: per-kernel device times to stderr
            lib.scg_timing_dump()
    if dist is not None:
        dist.destroy_process_group()
    if bad:
        sys.exit(3)


if __name__ == "__main__":
    main()
