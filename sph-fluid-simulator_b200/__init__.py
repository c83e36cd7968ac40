"""B200-native SPH simulation step (the SPHSystem::update() hot path of lijenicol/SPH-Fluid-Simulator).

Layout:
  csrc/     CUDA kernels (sm_100a) and the C-ABI implementation (include/sph_b200.h)
  host/     C++ mirror of the reference's SPHSystem class over the C-ABI
  binding.py  ctypes binding used by the tests, bench.py and the multi-GPU slab driver
  slab.py     one-process-per-GPU slab decomposition (torch.distributed plumbing)

The directory name is not a valid Python identifier; import it with
    importlib.import_module("sph-fluid-simulator_b200")
or through the `sph_b200` alias module at the repository root.
"""
from .binding import (  # noqa: F401
    ORDER_DEVICE, ORDER_HASH16, ORDER_ID, NO_PARTICLE, TABLE_SIZE, Derived, Settings, Sim, SphError, Stats,
    System, build_library, default_settings, derive, load_library, scaled_settings, scene_block, scene_block_slice, scene_cube,
)
