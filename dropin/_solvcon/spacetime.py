"""``_solvcon.spacetime`` (``cpp/solvcon/spacetime/pymod/wrap_spacetime.{hpp,cpp}``)."""
from solvcon_b200.spacetime import (Celm, Grid, InviscidBurgersSolver, Kernel, LinearScalarSolver, Selm,  # noqa: F401
                                    Solver, SolverElementIterator)

# the reference binds one cursor class per solver kind; here one class serves all three and prints the kind's name
LinearScalarCelm = InviscidBurgersCelm = Celm
LinearScalarSelm = InviscidBurgersSelm = Selm
