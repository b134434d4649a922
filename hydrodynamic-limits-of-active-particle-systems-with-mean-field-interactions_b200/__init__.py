"""B200-native active-exclusion-process stepper (hot path of the reference's ParticleSystem).

Layout (only what the path needs):
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/aps.h)
  capi.py          ctypes binding of the C ABI
  particle_system.py   host-side mirror of PARTICLE_solver_CLASS.ParticleSystem
  launcher.py      ensemble / sweep launcher (replaces the drivers' serial loops)
  build.py         in-tree nvcc build
The directory name contains hyphens, so it is imported through the alias package `aps_b200`
at the repo root.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
