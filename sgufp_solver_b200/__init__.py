"""sgufp_solver_b200 — B200-native scenario-cut evaluation and DD longest path for SGUFP.

Only the hot path of var-nan/SGUFP_Solver: `GuroSolver.solveSubProblem` and the
cut application of `RelaxedDDNew` / `RestrictedDDNew`, behind the C ABI of include/sgufp_b200.h.
"""
from . import instances  # noqa: F401
from .solver import FEASIBILITY, OPTIMALITY, Cut, GuroSolver, getKey  # noqa: F401
