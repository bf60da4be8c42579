"""Golden meshes and constants of the oblique-shock-reflection driver (row f4), produced by the REFERENCE's own driver
source (solvcon/pilot/apps/obsrefl/_driver.py, loaded unmodified from /root/reference) running on top of THIS
repo's drop-in classes in place of ``solvcon.core`` (StaticMesh, SimpleArray; a list-backed PointPad stands in for the
out-of-scope PointPadFp64).  Run in the build container:  python tests/golden/make_obsrefl_golden.py

That the reference's Python driver runs unchanged over solvcon_b200's StaticMesh mirror is itself the drop-in check of
SURVEY.md 8(b); the saved nodes / cells / boundary groups / analytic states then pin solvcon_b200.apps.obsrefl."""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/solvcon/pilot/apps/obsrefl/_driver.py"


class PointPad:
    def __init__(self, ndim):
        self.p = []

    def append(self, x, y):
        self.p.append((float(x), float(y)))

    def __len__(self):
        return len(self.p)

    def x_at(self, i):
        return self.p[i][0]

    def y_at(self, i):
        return self.p[i][1]

    def pack_array(self):
        return types.SimpleNamespace(ndarray=np.array(self.p, dtype=np.float64))


def load_reference_driver():
    from solvcon_b200.mesh import StaticMesh
    core = types.ModuleType("solvcon.core")
    core.StaticMesh, core.PointPadFp64, core.EulerCore = StaticMesh, PointPad, None
    for name in ("solvcon", "solvcon.pilot", "solvcon.pilot.apps", "solvcon.pilot.apps.obsrefl"):
        pkg = types.ModuleType(name)
        pkg.__path__ = []
        sys.modules[name] = pkg
    sys.modules["solvcon.core"] = core
    sys.modules["solvcon"].core = core
    spec = importlib.util.spec_from_file_location("solvcon.pilot.apps.obsrefl._driver", REF)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    drv = load_reference_driver()
    out = {}
    for ct in ("quad", "triangle", "unstructured"):
        msher = drv.ObliqueShockMesher(nx=24, ny=8)
        mh = msher.make_mesh(cell_type=ct)
        ng = mh.ngstcell
        out[ct + "_ndcrd"] = mh.ndcrd.ndarray[mh.ngstnode:].copy()
        out[ct + "_clnds"] = mh.clnds.ndarray[ng:, :5].copy()
        for name in msher.BOUNDARY_NAMES:
            out[ct + "_" + name] = mh.bc(name).facn.ndarray[:, 0].copy()
        print(ct, mh.ncell, mh.nface, [len(mh.bc(n).facn.ndarray) for n in msher.BOUNDARY_NAMES])
    sh = drv.ObliqueShock()
    sh.build_constant()
    out["zone_states"] = np.array(sh.zone_states())
    out["angles"] = np.array([sh.theta, sh.shock_angle, sh.shock_angle2, sh.mach2, sh.mach3])
    sh.mesher = drv.ObliqueShockMesher()
    out["shock_path"] = np.array(sh.shock_path())
    np.savez_compressed(os.path.join(HERE, "obsrefl.npz"), **out)


if __name__ == "__main__":
    main()
