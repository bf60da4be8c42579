"""``_solvcon.onedim`` (``cpp/solvcon/onedim/pymod/wrap_onedim.cpp:38-150``)."""
from solvcon_b200.onedim import Euler1DCore  # noqa: F401
