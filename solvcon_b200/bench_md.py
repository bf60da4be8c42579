"""Multidim benchmark workloads for bench.py: 3D Euler CESE on the synthetic mixed tet/hex/prism stepped
channel (BASELINE.json configs[3]/[4]) and 2D Euler on the mixed tri/quad forward-facing step (configs[2]).

Everything here is product code (mesh generator -> native StaticMesh builder -> device EulerCore) and nothing here
imports ``oracle/``: the CPU reference arm (``cpu_run``) forwards to ``bench_cpu.py`` at the repo root, which belongs
to bench.py's ``cpu_baseline`` / ``--impl reference`` leg.
"""
from __future__ import annotations

import time

import numpy as np

from . import meshgen
from .mesh import make_mesh
from .multidim import EulerCore

UNIT = "updates/s"


def dims_3d(ncell_target):
    """Box counts (nx, ny, nz, x1, z1) of the stepped channel with ~ncell_target cells, one third
    hexahedra / prisms / tetrahedra by cell count: hex boxes a*c*V, prisms 2(1-a)c*V, tets 6(1-a)(1-c)V
    with a = x1/nx = 2/3, c = z1/nz = 3/4 -> 1.5 V cells."""
    v = ncell_target / 1.5
    u = (v / (4 * 3 * 3)) ** (1.0 / 3.0)  # aspect 4:3:3, divisible by 3 (x) and 4 (z)
    nx = max(3, int(round(4 * u / 3.0)) * 3)
    nz = max(4, int(round(3 * u / 4.0)) * 4)
    ny = max(2, int(round(v / (nx * nz))))
    return nx, ny, nz, 2 * nx // 3, 3 * nz // 4


GAMMA, RHO0, P0, MACH_U = 1.4, 1.4, 1.0, 3.0   # sound speed 1, so the inlet speed IS the Mach number


def pocket_at_rest(so0n_body, clcnd_body, z_level):
    """Initial condition of the stepped channel on top of the uniform Mach-3 stream of ``init_solution``: the gas above
    the step level (``z > z_level``: the tetrahedral block C, which lies in the lee of the solid block) starts AT REST
    at the stream's density and pressure.  Round 1 started it at Mach 3 too; that gas then runs away from the solid
    block's back wall, evacuates the corner (rho -> 5e-4, then negative) and the reference's scheme -- whose only guard
    is the pressure clamp of ``euler_solution.cpp:106-160`` -- goes to NaN after ~105 steps (reproduced on the compiled
    reference).  With the pocket at rest the stream expands round the corner into still gas: rho >= 0.33, p >= 0.37,
    cfl <= 0.64 over 600 steps on the compiled reference at 6e4 and 2.5e5 cells.  Both arms (device, CPU reference)
    call this same function on their own ``so0n`` view, like a reference front end writing its IC through the array
    views (``tests/multidim/euler/test_field.py:60-62``)."""
    m = clcnd_body[:, 2] > z_level
    so0n_body[m, 1:4] = 0.0
    so0n_body[m, 4] = P0 / (GAMMA - 1.0)
    return int(m.sum())


def field_health(svr):
    """(cflo max, min density, min pressure) of the newest solution, from the device (derived fields + reduction)."""
    rho = svr.derived("density")
    prs = svr.derived("pressure")
    cfl = svr.derived("cfl")
    return {"cflo_max": float(cfl.max()), "rho_min": float(rho.min()), "p_min": float(prs.min()),
            "finite": bool(np.isfinite(rho).all() and np.isfinite(prs).all() and np.isfinite(cfl).all())}


def build_case_3d(ncell_target, order="morton", kinds="hpt", log=None):
    t0 = time.perf_counter()
    nx, ny, nz, x1, z1 = dims_3d(ncell_target)
    h = 1.0 / ny
    ndcrd, cltpn, clnds = meshgen.mixed_3d(nx, ny, nz, lx=nx * h, ly=1.0, lz=nz * h, x1=x1, z1=z1, kinds=kinds,
                                           jitter=0.1)
    t1 = time.perf_counter()
    mesh = make_mesh(3, ndcrd, cltpn, clnds)
    t2 = time.perf_counter()
    lo, hi = ndcrd.min(axis=0), ndcrd.max(axis=0)
    inlet, outlet, wall = meshgen.boundary_groups_of(mesh, lo, hi)
    dt = 0.05 * h
    svr = EulerCore(mesh, dt, order=order)
    import os
    if os.environ.get("SCG_MD_FUSED"):
        svr.set_fused(int(os.environ["SCG_MD_FUSED"]))
    t3 = time.perf_counter()
    # IC / BC of SURVEY.md 8d config 4: Mach-3 stream (pocket above the step at rest), inlet / non-reflective outlet /
    # slip walls
    svr.init_solution(GAMMA, RHO0, [MACH_U, 0.0, 0.0], P0)
    npocket = pocket_at_rest(svr.so0n.body, mesh.clcnd.body, z1 * h) if kinds == "hpt" else 0
    svr.add_inlet(inlet, [RHO0, MACH_U, 0.0, 0.0, P0, GAMMA])
    svr.add_nonrefl(outlet)
    svr.add_slipwall(wall)
    svr.resident(True)
    svr.bc_soln()
    svr.bc_dsoln()
    info = {"boxes": [nx, ny, nz], "x1": x1, "z1": z1, "h": h, "dt": dt, "ncell": int(mesh.ncell),
            "cells_at_rest": npocket,
            "nface": int(mesh.nface), "ngstcell": int(mesh.ngstcell),
            "cells_by_type": {"hex": int((cltpn == 5).sum()), "tet": int((cltpn == 6).sum()),
                              "prism": int((cltpn == 7).sum())},
            "setup_s": {"generate": round(t1 - t0, 2), "tables": round(t2 - t1, 2), "device": round(t3 - t2, 2)}}
    if log:
        log(info)
    return mesh, svr, info


def _finish_dist_case(part, dt, order, lo, hi, z_level, local_device, kinds, transport="lib"):
    """Device solver of one partition with the bench's IC / BCs, halo geometry and state exchanged once.
    ``transport``: "lib" = the library's own NCCL communicator (NcclExchange), "torch" = torch.distributed."""
    from .decomp import DistEulerCore, NcclExchange, TorchExchange
    mesh = part.build_mesh()
    svr = DistEulerCore(part, dt, order=order)
    inlet, outlet, wall = part.physical_boundary_groups(lo, hi)
    core = svr.core
    core.resident(False)
    core.init_solution(GAMMA, RHO0, [MACH_U, 0.0, 0.0], P0)
    if kinds == "hpt":
        pocket_at_rest(core.so0n.body, mesh.clcnd.body, z_level)
    core.add_inlet(inlet, [RHO0, MACH_U, 0.0, 0.0, P0, GAMMA])
    core.add_nonrefl(outlet)
    core.add_slipwall(wall)
    core.resident(True)
    core.bc_soln()
    core.bc_dsoln()
    svr.transport = NcclExchange(svr) if transport == "lib" else TorchExchange(svr, f"cuda:{local_device}")
    svr.transport.exchange(1)
    svr.transport.exchange(0)
    return mesh, svr


def _dist_info(part, mesh, svr, extra):
    own = part.owned.astype(bool)
    core = svr.core
    info = {"ncell": int(own.sum()), "ncell_local_mesh": int(mesh.ncell), "halo_send_cells": int(svr.nsend),
            "halo_recv_cells": int(svr.nrecv), "halo_bytes_per_substep": int(svr.nsend * core.halo_width(0) * 8),
            "neighbour_ranks": [int(q) for q in part.neighbours],
            "nface": int(mesh.nface), "ngstcell": int(mesh.ngstcell),
            "cells_by_type": {"hex": int((part.cltpn[own] == 5).sum()), "tet": int((part.cltpn[own] == 6).sum()),
                              "prism": int((part.cltpn[own] == 7).sum())}}
    info.update(extra)
    return info


def build_case_3d_dist(ncell_target, rank, world, local_device, order="morton", kinds="hpt"):
    """Weak-scaling case: the channel is extended in y with the number of ranks; rank r
    owns ny box layers and builds only those plus two halo layers per side (solvcon_b200.decomp)."""
    from .decomp import slab_partition_3d
    t0 = time.perf_counter()
    nx, ny, nz, x1, z1 = dims_3d(ncell_target)
    h = 1.0 / ny
    part = slab_partition_3d(rank, world, nx, ny, nz, h, x1, z1, kinds=kinds, jitter=0.1)
    t1 = time.perf_counter()
    part.build_mesh()
    t2 = time.perf_counter()
    dt = 0.05 * h
    lo = np.array([0.0, 0.0, 0.0])
    hi = np.array([nx * h, world * ny * h, nz * h])
    mesh, svr = _finish_dist_case(part, dt, order, lo, hi, z1 * h, local_device, kinds)
    t3 = time.perf_counter()
    info = _dist_info(part, mesh, svr, {"boxes_per_rank": [nx, ny, nz], "x1": x1, "z1": z1, "h": h, "dt": dt,
                                        "setup_s": {"generate": round(t1 - t0, 2), "tables": round(t2 - t1, 2),
                                                    "device": round(t3 - t2, 2)}})
    return mesh, svr, info


def build_case_3d_strong(ncell_total, rank, world, local_device, order="morton", kinds="hpt"):
    """Strong-scaling case (BASELINE.json configs[3]): ONE stepped channel of ``ncell_total``
    cells cut into ``world`` parts by cost-weighted recursive coordinate bisection of its box lattice
    (``decomp.lattice_rcb``); every rank generates only its own box range plus two layers
    (``decomp.lattice_partition_3d``), orders its cells (Morton / RCM) inside the part, and exchanges the state of
    the cells on its cuts after every substep, overlapped with the interior march."""
    from .decomp import CELL_COST, lattice_partition_3d, lattice_rcb
    t0 = time.perf_counter()
    nx, ny, nz, x1, z1 = dims_3d(ncell_total)
    h = 1.0 / ny
    ranges = lattice_rcb(nx, ny, nz, x1, z1, world, cost=CELL_COST)
    part = lattice_partition_3d(rank, ranges, nx, ny, nz, h, x1, z1, kinds=kinds, jitter=0.1)   # builds the local mesh
    t2 = time.perf_counter()
    dt = 0.05 * h
    lo = np.array([0.0, 0.0, 0.0])
    hi = np.array([nx * h, ny * h, nz * h])
    mesh, svr = _finish_dist_case(part, dt, order, lo, hi, z1 * h, local_device, kinds)
    t3 = time.perf_counter()
    info = _dist_info(part, mesh, svr, {"boxes": [nx, ny, nz], "x1": x1, "z1": z1, "h": h, "dt": dt,
                                        "box_range_of_rank0": [list(r) for r in ranges[0]],
                                        "setup_s": {"generate+tables": round(t2 - t0, 2), "device": round(t3 - t2, 2)}})
    return mesh, svr, info


def nccl_parity_3d(rank, world, local_device, dist, order, ncell_target=120_000, steps=6, transport="lib"):
    """Multi-rank parity leg of bench.py (before anything is timed): a small stepped channel cut by ``lattice_rcb``
    into one part per rank and marched through the REAL transport (``TorchExchange.exchange_split``: NCCL send/recv on
    the communication stream, overlapped with the interior march); every rank's owned cells are gathered on rank 0 and
    compared, bit for bit, with the undecomposed device run of the same case (which the single-rank parity leg and the
    GPU tests hold bit-equal to the compiled reference)."""
    import torch
    from . import _lib
    from .decomp import CELL_COST, lattice_partition_3d, lattice_rcb
    nx, ny, nz, x1, z1 = dims_3d(ncell_target)
    h = 1.0 / ny
    dt = 0.05 * h
    ranges = lattice_rcb(nx, ny, nz, x1, z1, world, cost=CELL_COST)
    part = lattice_partition_3d(rank, ranges, nx, ny, nz, h, x1, z1, jitter=0.1)
    lo, hi = np.array([0.0, 0.0, 0.0]), np.array([nx * h, ny * h, nz * h])
    mesh, svr = _finish_dist_case(part, dt, order, lo, hi, z1 * h, local_device, "hpt", transport)
    svr.march(steps)
    _lib.sync()
    svr.core.materialise(("so0n", "so1n", "cflo", "cflc"))
    svr.core.sync_to_host()
    own = part.owned.astype(bool)
    mine = {"key": part.gid[own]}
    for n in ("so0n", "so1n", "cflo", "cflc"):
        mine[n] = getattr(svr.core, n).body[own].copy()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    result = None
    if rank == 0:
        ndcrd, cltpn, clnds, box = meshgen.mixed_3d(nx, ny, nz, lx=nx * h, ly=ny * h, lz=nz * h, x1=x1, z1=z1, jitter=0.1,
                                                    return_box=True)
        key = meshgen.cell_key(box, ny, nz)
        _, ref, info = build_case_3d(ncell_target, order)
        assert info["ncell"] == len(cltpn)
        ref.march(steps)
        ref.materialise(("so0n", "so1n", "cflo", "cflc"))
        ref.sync_to_host()
        equal, seen, worst = True, np.zeros(len(cltpn), dtype=np.int64), 0.0
        for g in gathered:
            gid = np.searchsorted(key, g["key"])
            seen[gid] += 1
            for n in ("so0n", "so1n", "cflo", "cflc"):
                a, c = g[n], getattr(ref, n).body[gid]
                equal = equal and bool(np.array_equal(a, c))
                worst = max(worst, float(np.abs(a - c).max()) / (float(np.abs(c).max()) + 1e-300))
        result = {"cells": int(len(cltpn)), "parts": world, "steps": steps, "bit_equal": bool(equal and (seen == 1).all()),
                  "max_rel_linf": worst,
                  "transport": ("scg_md_march_dist: ncclSend/ncclRecv issued by the library on its communication "
                                "stream, overlapped with the interior march" if transport == "lib" else
                                "TorchExchange.exchange_split: torch.distributed isend/irecv over NCCL, overlapped"),
                  "against": "undecomposed device march of the same case",
                  "halo_cells_sent_by_rank0": int(svr.nsend)}
    flag = torch.tensor([1 if (result is None or result["bit_equal"]) else 0], device=f"cuda:{local_device}")
    dist.broadcast(flag, src=0)
    return result, bool(flag.item())


class MD3DWorkload:
    name = "md3d"
    dtype = "f64"
    dominant_kernel = "one substep = md_solt + md_solncfl_staged<hex|prism|tet> + md_dsoln_eqo<hex|prism|tet> + boundary kernels"

    def __init__(self, ncell, chunk, order="morton", kinds="hpt", strong=False):
        self.ncell_target = int(ncell)
        self.chunk = chunk
        self.order = order
        self.kinds = kinds
        self.strong = bool(strong)      # True: ncell is the TOTAL, cut into one part per rank (configs[3])
        if self.strong:
            self.name = "md3d_strong"

    def build_device(self):
        world = getattr(self, "world", 1)
        if world > 1 and self.strong:
            self.mesh, self.dist_svr, self.info = build_case_3d_strong(self.ncell_target, self.rank, world, self.local,
                                                                       self.order, self.kinds)
            self.dev = self.dist_svr.core
        elif world > 1:
            # every rank builds its own slab on the host (~0.85 KB of host memory per cell at the peak of the table
            # construction): keep the ranks of one box inside its RAM
            try:
                import psutil
                budget = 0.6 * psutil.virtual_memory().available / world
                cap = int(budget / 850.0)
                if getattr(self, "dist", None) is not None:   # the ranks must agree on the size: take the smallest view
                    import torch
                    t = torch.tensor([cap], dtype=torch.int64, device=f"cuda:{self.local}")
                    self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
                    cap = int(t.item())
                if cap < self.ncell_target:
                    self.size_note = (f"cells per GPU reduced from {self.ncell_target} to {cap}: host RAM "
                                      f"({psutil.virtual_memory().available / 2**30:.0f} GiB available for {world} ranks)")
                    self.ncell_target = max(cap, 1_000_000)
            except ImportError:
                pass
            self.mesh, self.dist_svr, self.info = build_case_3d_dist(self.ncell_target, self.rank, world, self.local,
                                                                     self.order, self.kinds)
            self.dev = self.dist_svr.core
        else:
            self.mesh, self.dev, self.info = build_case_3d(self.ncell_target, self.order, self.kinds)
            self.dist_svr = None
        self.ncell = self.info["ncell"]
        self.balg_substep = self.dev.algorithmic_bytes()
        # measured DRAM bytes per cell-substep from the committed ncu capture, scaled to this mesh
        import json, os
        tp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic_md3d.json")
        self.ncu_traffic = None
        if os.path.exists(tp) and self.kinds == "hpt":
            with open(tp) as f:
                self.ncu_traffic = float(json.load(f)["bytes_per_cell_substep"]) * self.ncell
        return self.dev

    def _planned(self):
        """Mesh numbers of the target size without building anything (the reference arm never builds the device case)."""
        nx, ny, nz, x1, z1 = dims_3d(self.ncell_target)
        nhex, npri, ntet = x1 * ny * z1, 2 * (nx - x1) * ny * z1, 6 * (nx - x1) * ny * (nz - z1)
        ncell = nhex + npri + ntet
        info = {"boxes": [nx, ny, nz], "x1": x1, "z1": z1, "h": 1.0 / ny, "dt": 0.05 / ny, "ncell": ncell,
                "cells_by_type": {"hex": nhex, "tet": ntet, "prism": npri}, "planned_only": True}
        return ncell, info, float(1752 * nhex + 1428 * npri + 1104 * ntet)

    def describe(self, world):
        if hasattr(self, "info"):
            ncell, info, balg = self.ncell, self.info, float(self.balg_substep)
        else:
            ncell, info, balg = self._planned()
        size = (f"{self.ncell_target} cells in total, {ncell} on rank 0" if self.strong and world > 1
                else f"{ncell} cells per GPU")
        d = {"workload": f"3D Euler CESE (multidim EulerCore.march), synthetic mixed tet/hex/prism stepped channel, "
                         f"{size} ({info['cells_by_type']}), Mach-3 inlet (gas above the step at rest) / nonrefl "
                         f"outlet / slip walls, march({self.chunk}) per step, {self.order} cell ordering",
             "cells_per_gpu": ncell, "chunk_steps": self.chunk, "ordering": self.order,
             "alg_bytes_per_cell_substep": balg / ncell,
             "l2_policy": "inputs larger than L2 (%.1f GB streamed per substep)" % (balg / 1e9),
             "decomposition": "none" if world == 1 else
             (f"ONE mesh of {self.ncell_target} cells cut into {world} parts by cost-weighted recursive coordinate "
              f"bisection of the box lattice, {self.order} ordering inside each part; " if self.strong else
              f"y-slabs, one per GPU, domain extended in y with the rank count; ") +
             f"halo exchange of so0n+so1n after every substep by NCCL isend/irecv, overlapped with the interior march "
             f"({info.get('halo_bytes_per_substep', 0)} B sent by rank 0 per substep)",
             "mesh": info}
        if getattr(self, "size_note", None):
            d["size_note"] = self.size_note
        return d

    def updates_per_step(self):
        return 2 * self.chunk * self.ncell

    def launches_per_step(self):
        return 2 * self.chunk  # substeps

    def alg_bytes_per_launch(self):
        return float(self.balg_substep)

    def resident_begin(self):
        self.dev.resident(True)

    def resident_step(self):
        if self.dist_svr is not None:
            self.dist_svr.march(self.chunk)
        else:
            self.dev.march(self.chunk)

    E2E_IN = ("so0n", "so1n", "gamma")      # the host arrays a march() reads: solution, gradient, gamma
    E2E_OUT = ("so0n", "so1n", "cflo")      # what the caller gets back: the new solution + gradient and the CFL table

    def e2e_begin(self):
        """The caller's side of the end-to-end arm: a front end that HOLDS views of its field arrays, like the
        reference's (``tests/multidim/euler/test_field.py:50-83``, ``solvcon/multidim/euler/_field.py:53-54,124-125``).
        Held views are what makes ``EulerCore.march`` move bytes: it uploads every held array it reads and refreshes
        every held array it wrote (solvcon_b200/multidim.py, "coherence")."""
        d = self.dev
        d.materialise(tuple(set(self.E2E_IN + self.E2E_OUT)))
        if self.dist_svr is None:
            d.resident(False)
            self._views = [getattr(d, n).ndarray for n in set(self.E2E_IN + self.E2E_OUT)]
        self._b0 = (d.h2d_bytes, d.d2h_bytes)
        self._e2e_steps = 0

    def e2e_step(self):
        d = self.dev
        if self.dist_svr is not None:
            # the distributed march keeps the field resident; the same arrays travel through the explicit calls
            d.upload(self.E2E_IN)
            self.dist_svr.march(self.chunk)
            d.download(self.E2E_OUT)
        else:
            d.march(self.chunk)
        self._e2e_steps += 1

    def e2e_fields_only(self, steps=2):
        """Second end-to-end pattern, reported next to the full-state one: the reference's own session loop
        (``solvcon/pilot/apps/obsrefl/_session.py:123-140``: march a chunk, then read the CFL range and the overall
        mass through EulerField, i.e. ``cflo`` and ``so0n``) -- the caller holds views of those two arrays only, so
        the gradient never crosses the bus.  Single-GPU arm; returns updates/s and the bytes per step."""
        if self.dist_svr is not None:
            return None
        d = self.dev
        self._views = [d.so0n.ndarray, d.cflo.ndarray]
        d.march(self.chunk)
        b0 = (d.h2d_bytes, d.d2h_bytes)
        t0 = time.perf_counter()
        for _ in range(steps):
            d.march(self.chunk)
            chk = float(self._views[1].max())
        el = time.perf_counter() - t0
        return {"value": self.updates_per_step() * steps / el, "unit": UNIT, "steps": steps,
                "h2d_bytes_per_step": (d.h2d_bytes - b0[0]) // steps, "d2h_bytes_per_step": (d.d2h_bytes - b0[1]) // steps,
                "pattern": "caller holds views of so0n and cflo only (the reference's session loop)", "cflo_max": chk}

    def e2e_bytes(self):
        """Bytes per step actually copied by the library during the end-to-end arm (counted at every transfer)."""
        d = self.dev
        n = max(1, self._e2e_steps)
        return (d.h2d_bytes - self._b0[0]) // n, (d.d2h_bytes - self._b0[1]) // n

    def result_check(self):
        return float(self.dev.so0n.body[:, 0].sum())

    def cpu_run(self, budget_s=15.0):
        import bench_cpu   # repo root, next to bench.py: the only code on the bench path that touches oracle/
        return bench_cpu.cpu_reference_3d(budget_s, self.kinds)

    def parity(self):
        import bench_cpu
        return bench_cpu.parity_3d(self.order, self.kinds)

    def health(self):
        return field_health(self.dev)

    # fp64-pipe instructions one cell-substep executes in the default kernels, from the committed ncu --set full capture
    # of the round (profiles/r02h_md3d_ncu_full_summary.txt; executed DADD + DMUL + DFMA + DSETP of its source page, tools/ncu_hot.py's
    # input, md_solncfl_staged + md_dsoln_eqo, per 6.666e5 cells): hex 4880 + 3540, tet 2681 + 2024, prism 4085 + 2899
    # (round 1: 9080 / 5070 / 7380).  The arithmetic is the reference's, operation for operation (no FMA contraction).
    FP64_INSTR_PER_CELL = {"hex": 8420.0, "tet": 4705.0, "prism": 6984.0}

    def roofline_fp64(self, substep_s, clocks):
        """Second roofline of the 3D march: the fp64 pipe.  B200 issues 64 fp64 lanes per SM per clock (16 per
        sub-partition: one warp instruction every 2 cycles), so a substep of this rank's cells cannot run faster than
        ``sum_type n_type * instr_type / (148 * 64 * f_clk)``; reported at the maximum SM clock and at the median clock
        measured during the timed region (the board is power-capped under this kernel mix)."""
        by = self.info["cells_by_type"]
        lanes = sum(self.FP64_INSTR_PER_CELL[k] * by.get(k, 0) for k in self.FP64_INSTR_PER_CELL)
        if lanes <= 0:
            return None
        sm = 148
        fmax = (clocks.get("sm_max_mhz") or 1965.0) * 1e6
        fmeas = (clocks.get("sm_mhz") or 0.0) * 1e6
        floor_max = lanes / (sm * 64.0 * fmax)
        out = {"bound": "fp64 pipe (64 lanes/clk/SM)", "fp64_instr_per_cell": self.FP64_INSTR_PER_CELL,
               "source": "profiles/r02h_md3d_ncu_full_summary.txt (ncu --set full: executed instructions x fp64-pipe share)",
               "floor_substep_ms_at_max_clock": floor_max * 1e3,
               "ceiling_updates_per_s_at_max_clock": self.ncell / floor_max,
               "frac_of_ceiling_at_max_clock": floor_max / substep_s,
               "hbm_frac_at_fp64_ceiling": (self.balg_substep / floor_max / 1e9) / 6555.5}
        if fmeas > 0:
            floor_meas = lanes / (sm * 64.0 * fmeas)
            out["ceiling_updates_per_s_at_measured_clock"] = self.ncell / floor_meas
            out["frac_of_ceiling_at_measured_clock"] = floor_meas / substep_s
            out["measured_sm_mhz"] = clocks.get("sm_mhz")
        return out



# ------------------------------------------------------------------------------------------ 2D ---

def dims_2d(ncell_target):
    """Box counts (nx, ny) of the forward-facing step domain [0,3]x[0,1] minus [0.6,3]x[0,0.2] (SURVEY.md 8d
    config 3): 3 ny^2 boxes, 16 % removed, a quarter of the rest split into two triangles -> 3.15 ny^2 cells."""
    ny = max(5, int(round((ncell_target / 3.15) ** 0.5 / 5.0)) * 5)
    return 3 * ny, ny


def build_case_2d(ncell_target, order="morton"):
    t0 = time.perf_counter()
    nx, ny = dims_2d(ncell_target)
    h = 1.0 / ny
    ndcrd, cltpn, clnds = meshgen.mixed_2d(nx, ny, lx=3.0, ly=1.0, tri_every=4, jitter=0.3, step=(0.6, 0.2))
    t1 = time.perf_counter()
    mesh = make_mesh(2, ndcrd, cltpn, clnds)
    t2 = time.perf_counter()
    lo, hi = ndcrd.min(axis=0), ndcrd.max(axis=0)
    inlet, outlet, wall = meshgen.boundary_groups(2, mesh.fccnd.body, mesh.fccls.body, mesh.nface, lo, hi)
    dt = 0.05 * h
    svr = EulerCore(mesh, dt, order=order)
    import os
    if os.environ.get("SCG_MD_FUSED"):
        svr.set_fused(int(os.environ["SCG_MD_FUSED"]))
    t3 = time.perf_counter()
    svr.resident(True)
    svr.init_solution(GAMMA, RHO0, [MACH_U, 0.0], P0)
    svr.add_inlet(inlet, [RHO0, MACH_U, 0.0, P0, GAMMA])
    svr.add_nonrefl(outlet)
    svr.add_slipwall(wall)
    svr.bc_soln()
    svr.bc_dsoln()
    info = {"boxes": [nx, ny], "h": h, "dt": dt, "ncell": int(mesh.ncell), "nface": int(mesh.nface),
            "ngstcell": int(mesh.ngstcell),
            "cells_by_type": {"quad": int((cltpn == 3).sum()), "tri": int((cltpn == 4).sum())},
            "setup_s": {"generate": round(t1 - t0, 2), "tables": round(t2 - t1, 2), "device": round(t3 - t2, 2)}}
    return mesh, svr, info


class MD2DWorkload(MD3DWorkload):
    name = "md2d"
    roofline_fp64 = None
    dominant_kernel = "one substep = md_solt + md_solncfl_staged<quad|tri> + md_dsoln_eqo<quad|tri> + boundary kernels"

    def build_device(self):
        if getattr(self, "world", 1) > 1:
            raise SystemExit("md2d: single-GPU workload (the multi-GPU arm is md3d)")
        self.mesh, self.dev, self.info = build_case_2d(self.ncell_target, self.order)
        self.dist_svr = None
        self.ncell = self.info["ncell"]
        self.balg_substep = self.dev.algorithmic_bytes()
        self.ncu_traffic = None
        return self.dev

    def describe(self, world):
        if hasattr(self, "info"):
            ncell, info, balg = self.ncell, self.info, float(self.balg_substep)
        else:
            nx, ny = dims_2d(self.ncell_target)
            kept = nx * ny - (nx - int(round(0.6 / 3.0 * nx))) * int(round(0.2 * ny))
            ncell = kept + kept // 4
            info = {"boxes": [nx, ny], "h": 1.0 / ny, "dt": 0.05 / ny, "ncell": ncell, "planned_only": True,
                    "cells_by_type": {"quad": kept - kept // 4, "tri": 2 * (kept // 4)}}
            balg = float(616 * info["cells_by_type"]["quad"] + 524 * info["cells_by_type"]["tri"])
        return {"workload": f"2D Euler CESE (multidim EulerCore.march), synthetic mixed tri/quad forward-facing step, "
                            f"{ncell} cells ({info['cells_by_type']}), Mach-3 inlet / nonrefl outlet / slip walls, "
                            f"march({self.chunk}) per step, {self.order} cell ordering",
                "cells_per_gpu": ncell, "chunk_steps": self.chunk, "ordering": self.order,
                "alg_bytes_per_cell_substep": balg / ncell,
                "l2_policy": "inputs larger than L2 (%.1f GB streamed per substep)" % (balg / 1e9),
                "decomposition": "none", "mesh": info}

    def cpu_run(self, budget_s=15.0):
        import bench_cpu
        return bench_cpu.cpu_reference_2d(budget_s)

    def parity(self):
        import bench_cpu
        return bench_cpu.parity_2d(self.order)


def make(args):
    if args.workload == "md3d":
        return MD3DWorkload(args.size or 50_000_000, args.chunk, order=getattr(args, "order", "morton"))
    if args.workload == "md3d_strong":
        return MD3DWorkload(args.size or 100_000_000, args.chunk, order=getattr(args, "order", "morton"), strong=True)
    if args.workload == "md2d":
        return MD2DWorkload(args.size or 20_000_000, args.chunk, order=getattr(args, "order", "morton"))
    raise SystemExit(f"workload {args.workload} not available")
