"""Derived fields of ``Euler1DCore`` (density ... entropy; Euler1DCore.hpp:106-168) against the compiled reference's own
accessors on the same conserved state.  CPU only: the properties are host expressions over the host twin of ``so0``, so
they are exercised on an instance whose arrays are set directly (no device object is created)."""
import numpy as np
import pytest

from cases import sod_ic


@pytest.fixture(scope="module")
def marched(bridge):
    if not bridge.have_ref():
        pytest.skip("compiled reference not available")
    chk = bridge.RefEuler1D(401, 4e-3)
    sod_ic(chk)
    chk.march(2, 30)
    return chk


@pytest.mark.parametrize("name", ["density", "velocity", "pressure", "temperature", "internal_energy", "entropy"])
def test_derived_field_matches_reference(marched, bridge, name):
    from solvcon_b200.onedim import Euler1DCore
    host = Euler1DCore.__new__(Euler1DCore)          # host expressions only; __del__ tolerates the missing handle
    host._so0, host._gamma = marched.so0.copy(), marched.gamma.copy()
    ours = getattr(host, name)
    ref = marched.derived(name)
    assert ours.shape == ref.shape
    if name == "entropy":
        # std::pow against numpy's pow: the same libm on this image, but allow the last bit
        np.testing.assert_allclose(ours, ref, rtol=4e-16, atol=0)
    else:
        assert np.array_equal(ours, ref), f"{name}: max abs diff {np.abs(ours - ref).max()}"
    assert np.isfinite(ref).all() and np.ptp(ref) > 0
