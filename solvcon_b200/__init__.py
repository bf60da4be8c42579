"""solvcon_b200 -- B200-native (sm_100a) device path for solvcon's space-time CESE march.

The package holds only what the hot path needs:

* ``csrc/``        hand-written CUDA kernels + the C-ABI of ``include/scg.h`` -> ``libscg_cuda.so``
* ``onedim``       mirror of ``_solvcon.onedim.Euler1DCore``
* ``spacetime``    mirror of ``_solvcon.spacetime.{Grid,LinearScalarSolver,InviscidBurgersSolver}``
* ``mesh``         mirror of ``_solvcon.StaticMesh`` (host table builder)
* ``multidim``     mirror of ``_solvcon.EulerCore``
* ``gradient_element``  mirror of ``_solvcon.GradientElement`` (host-side probe of one cell)
* ``gmsh``         Gmsh MSH 2.2 reader (``solvcon.Gmsh``)
* ``meshgen``      deterministic synthetic meshes for the named benchmark configurations
* ``decomp`` / ``decomp1d``  domain decomposition and halo exchange (one process per GPU)

The reference's front ends (``solvcon/onedim/euler1d.py``, ``solvcon/onedim/linear_scalar.py``,
``solvcon/multidim/euler/_field.py``, ``solvcon/pilot/apps/obsrefl``) are NOT re-implemented here: they run unmodified
over these classes through ``dropin/_solvcon`` (see ``dropin/stage.py``, ``INTEGRATION.md``).

The names of the reference's extension module are available at the package level too (``EulerCore``, ``StaticMesh``,
``GradientElement``, ``Gmsh``, ...), resolved lazily so that importing the package stays cheap.

There is no CPU fallback: every solver call goes through the CUDA library and raises when it
is missing or no device is present.
"""
from . import _lib  # noqa: F401
from .simplearray import SimpleArray, SimpleArrayFloat64, SimpleArrayInt32  # noqa: F401

__all__ = ["_lib", "SimpleArray", "SimpleArrayFloat64", "SimpleArrayInt32", "EulerCore", "StaticMesh", "StaticMeshBc",
           "GradientElement", "Gmsh", "Euler1DCore"]

_LAZY = {"EulerCore": "multidim", "StaticMesh": "mesh", "StaticMeshBc": "mesh", "GradientElement": "gradient_element",
         "Gmsh": "gmsh", "Euler1DCore": "onedim"}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        return getattr(importlib.import_module("." + _LAZY[name], __name__), name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
