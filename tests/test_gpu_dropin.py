"""GPU: the reference's OWN dispatcher updateParticles(..., onGPU=true) (src/sph.cpp:277-290) linked
with the product's updateParticlesGPU drop-in (host/updateParticlesGPU_dropin.cpp) instead of the
reference kernel file, stepped side by side with updateParticles(..., onGPU=false).
Needs oracle/_ref/libsph_ref_dropin.so (prebuilt in the build container; travels with the snapshot)."""
import os

import numpy as np
import pytest

from conftest import ROOT, assert_fields_close, by_id, load_golden

pytestmark = pytest.mark.gpu

DROPIN = os.path.join(ROOT, "oracle", "_ref", "libsph_ref_dropin.so")


@pytest.fixture(scope="module")
def dropin(sph):
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/libsph_ref_dropin.so not built")
    from oracle.pyoracle import Reference
    return Reference(DROPIN)


def test_reference_dispatcher_runs_the_b200_step(dropin):
    from oracle.pyoracle import DEFAULT_SETTINGS
    g = load_golden("cube20_step200.npz")
    p, v = g["pos0"], g["vel0"]
    ids = np.arange(p.shape[0], dtype=np.uint32)
    for step in range(5):  # step-locked: both legs start every step from the CPU leg's state
        cpu = dropin.step(DEFAULT_SETTINGS, 0.003, p, v, ids, on_gpu=False, transforms=True)
        gpu = dropin.step(DEFAULT_SETTINGS, 0.003, p, v, ids, on_gpu=True, transforms=True)
        # the array comes back sorted by start-of-step hash16 from both legs
        assert np.array_equal(gpu["hash"], cpu["hash"])
        assert np.array_equal(np.sort(gpu["id"]), np.sort(cpu["id"]))
        assert_fields_close(by_id(gpu), by_id(cpu), f"dispatcher step {step}")
        # transforms line up with the particle rows
        assert np.array_equal(gpu["transforms"][:, 12:15].view(np.uint32), gpu["pos"].view(np.uint32))
        p, v, ids = cpu["pos"], cpu["vel"], cpu["id"]


def test_dispatcher_handles_growth_and_new_settings(dropin):
    """Particle count and settings may change between calls (GUI RESET, src/Tester.cpp:169-173)."""
    rng = np.random.default_rng(0)
    for n, s7 in ((500, (0.02, 1000.0, 1.0, 1.04, 0.15, -9.8, 0.2)), (4000, (0.02, 1000.0, 1.0, 3.5, 0.2, -9.8, 1.0)),
                  (100, (0.05, 800.0, 2.0, 1.0, 0.1, -9.8, 1.0))):
        p = rng.uniform([-2, 0.2, -2], [2, 3, 2], (n, 3)).astype(np.float32)
        v = rng.normal(0, 0.5, (n, 3)).astype(np.float32)
        cpu = dropin.step(s7, 0.003, p, v, on_gpu=False)
        gpu = dropin.step(s7, 0.003, p, v, on_gpu=True)
        assert np.array_equal(gpu["hash"], cpu["hash"])
        assert_fields_close(by_id(gpu), by_id(cpu), f"n={n}", gas_constant=s7[2])
