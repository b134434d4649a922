"""Drop-in for the reference's `IMEX_PDE_solver_class` module: `from IMEX_PDE_solver_class import IMEXPDE` in the unchanged
run / sweep scripts (IMEX_PDE_solver_run*.py) resolves to the CUDA-backed mirror when this directory is first on sys.path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200.imex_pde import IMEXPDE, solve_many  # noqa: E402,F401
