"""solvcon_b200 -- B200-native (sm_100a) device path for solvcon's space-time CESE march.

The package holds only what the hot path needs:

* ``csrc/``        hand-written CUDA kernels + the C-ABI of ``include/scg.h`` -> ``libscg_cuda.so``
* ``onedim``       mirror of ``_solvcon.onedim.Euler1DCore``
* ``spacetime``    mirror of ``_solvcon.spacetime.{Grid,LinearScalarSolver,InviscidBurgersSolver}``
* ``mesh``         mirror of ``_solvcon.StaticMesh`` (host table builder)
* ``multidim``     mirror of ``_solvcon.EulerCore``
* ``meshgen``      deterministic synthetic meshes for the named benchmark configurations
* ``decomp``       slab / RCB domain decomposition and halo exchange (one process per GPU)

There is no CPU fallback: every solver call goes through the CUDA library and raises when it
is missing or no device is present.
"""
from . import _lib  # noqa: F401
from .simplearray import SimpleArray, SimpleArrayFloat64, SimpleArrayInt32  # noqa: F401

__all__ = ["_lib", "SimpleArray", "SimpleArrayFloat64", "SimpleArrayInt32"]
