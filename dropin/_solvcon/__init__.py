"""``_solvcon`` -- the reference's extension-module name, served by the B200 device path.

The reference's Python package reaches its C++ core through ONE extension module, ``_solvcon``
(``cpp/solvcon/python/module.cpp:34-61``; ``solvcon/core.py:193-213`` imports a top-level ``_solvcon`` first, the
in-package one second).  Putting this directory on ``sys.path`` ahead of the reference's build makes the reference's
UNMODIFIED Python package -- ``solvcon/core.py``, ``solvcon/spacetime.py``, ``solvcon/onedim/*``,
``solvcon/multidim/euler/_field.py``, ``solvcon/pilot/apps/obsrefl/{_driver,_analytic,_session}.py`` -- and its
unmodified hot-path tests run on ``libscg_cuda.so``:

    PYTHONPATH=dropin:dropin/_stage python -c "import solvcon; print(solvcon.EulerCore)"

(``dropin/_stage/solvcon`` is a verbatim copy of ``/root/reference/solvcon`` made by ``dropin/stage.py``; it is
git-ignored like ``oracle/_ref`` and never edited.)

What is served:

* hot-path classes: ``StaticMesh``, ``StaticMeshBc``, ``EulerCore``, ``GradientElement``,
  ``SimpleArray*``, ``Gmsh``, sub-modules ``onedim`` (``Euler1DCore``) and ``spacetime`` (``Grid``, ``Celm``, ``Selm``,
  ``Kernel``, ``Solver``, ``LinearScalarSolver``, ``InviscidBurgersSolver``) -- the classes of ``solvcon_b200``;
* two point/segment CONTAINERS (``PointPadFp64``, ``SegmentPadFp64``) because the oblique-shock driver and its
  analysis keep their node cloud / shock path in them (``_driver.py:84,140``, ``_analytic.py:168``);
* every other name ``solvcon/core.py`` copies out of the extension (linear algebra, FFT, geometry world, OASIS, the
  profiler, toggles ...; SURVEY.md 2 "OUT OF SCOPE") resolves to an inert placeholder that raises
  ``NotImplementedError`` when used, so ``import solvcon`` works and nothing outside the CESE path pretends to exist.

There is no CPU fallback here either: constructing a solver without the CUDA library or without a device raises.
"""
from solvcon_b200.simplearray import SimpleArray, SimpleArrayFloat64, SimpleArrayInt32  # noqa: F401
from solvcon_b200.mesh import StaticMesh, StaticMeshBc  # noqa: F401
from solvcon_b200.multidim import EulerCore  # noqa: F401
from solvcon_b200.gradient_element import GradientElement  # noqa: F401
from solvcon_b200.gmsh import Gmsh  # noqa: F401

from . import onedim, spacetime  # noqa: F401
from ._pads import PointPadFp64, SegmentPadFp64  # noqa: F401

HAS_PILOT = False      # python/module.cpp: no Qt GUI in this build
METAL_BUILT = False

# The names solvcon/core.py:18-187 loads from the extension that are outside the CESE hot path.
_OUT_OF_SCOPE = frozenset("""
ConcreteBuffer BufferExpander SimpleArrayBool SimpleArrayInt8 SimpleArrayInt16 SimpleArrayInt64 SimpleArrayUint8
SimpleArrayUint16 SimpleArrayUint32 SimpleArrayUint64 SimpleArrayFloat32 SimpleArrayComplex64 SimpleArrayComplex128
SimpleCollectorBool SimpleCollectorInt8 SimpleCollectorInt16 SimpleCollectorInt32 SimpleCollectorInt64
SimpleCollectorUint8 SimpleCollectorUint16 SimpleCollectorUint32 SimpleCollectorUint64 SimpleCollectorFloat32
SimpleCollectorFloat64 SimpleCollectorComplex64 SimpleCollectorComplex128 Plot3d complex64 complex128 StaticGrid1d
StaticGrid2d StaticGrid3d CommandLineInfo ProcessInfo WrapperProfilerStatus wrapper_profiler_status StopWatch
stop_watch CallProfiler call_profiler CallProfilerProbe HierarchicalToggleAccess Toggle ToggleCategory
ToggleSubscription metal_running FourierTransform llt_factorization llt_solve lu_factorization lu_solve lu_inv lu_det
LuFactorizationFloat32 LuFactorizationFloat64 LuFactorizationComplex64 LuFactorizationComplex128 EigenSystem
EigenSystemFloat32 EigenSystemFloat64 EigenSystemComplex64 EigenSystemComplex128 KalmanStateInfoFp32
KalmanStateInfoFp64 KalmanStateInfoComplex64 KalmanStateInfoComplex128 KalmanFilterFp32 KalmanFilterFp64
KalmanFilterComplex64 KalmanFilterComplex128 calc_bernstein_polynomial interpolate_bernstein BoundBox3dFp32
BoundBox3dFp64 Point3dFp32 Point3dFp64 Segment3dFp32 Segment3dFp64 Triangle3dFp32 Triangle3dFp64 Bezier3dFp32
Bezier3dFp64 PointPadFp32 SegmentPadFp32 TrianglePadFp32 TrianglePadFp64 CurvePadFp32 CurvePadFp64 WorldFp32 WorldFp64
PolygonPadFp32 PolygonPadFp64 Polygon3dFp32 Polygon3dFp64 TrapezoidPadFp32 TrapezoidPadFp64
TrapezoidalDecomposerFp32 TrapezoidalDecomposerFp64 ViewTransform2dFp64 OasisDevice OasisRecordRect OasisRecordPoly
timeseries
""".split())   # `timeseries`: the sub-module solvcon/timeseries.py:11-14 loads its log-analysis kernels from


class _OutOfScope:
    """Placeholder for a name of the reference extension that lies outside the CESE hot path."""

    def __init__(self, name):
        self._name = name

    def __getattr__(self, attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _OutOfScope(f"{self._name}.{attr}")

    def __call__(self, *args, **kw):
        raise NotImplementedError(f"_solvcon.{self._name}: outside the CESE hot path served by solvcon_b200")

    def __repr__(self):
        return f"<_solvcon.{self._name}: out of scope>"


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        return _OutOfScope(name)
    raise AttributeError(f"module '_solvcon' has no attribute {name!r}")
