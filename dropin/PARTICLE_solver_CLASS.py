"""Drop-in for the reference's PARTICLE_solver_CLASS module.

Put this directory first on PYTHONPATH and the unchanged drivers' `from PARTICLE_solver_CLASS import
ParticleSystem` (e.g. PARTICLE_solver_BIOLOGY_EXCLUSION.py:12) picks up the B200 implementation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200.particle_system import ParticleSystem, PhiloxRNG  # noqa: E402,F401
