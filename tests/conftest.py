"""Shared fixtures. `-m "not gpu"` runs on a CPU-only box; `-m gpu` needs a B200.

The oracle (oracle/) is test infrastructure: it is loaded here, never by the product.
Nothing in this directory reads /root/reference at run time; the reference is only present
inside oracle/_ref/libsph_ref.so when that prebuilt file exists.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle, build
    build(ref=False)
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle.pyoracle import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libsph_ref.so not built (needs /root/reference at build time)")
    return Reference()


@pytest.fixture(scope="session")
def sph():
    import sph_b200
    if not os.path.exists(sph_b200.binding.LIB_PATH):
        sph_b200.build_library()
    sph_b200.load_library()
    return sph_b200


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name)))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_bit_equal(a, b, what=""):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if not np.array_equal(bits(a), bits(b)):
        bad = np.nonzero(bits(a).reshape(a.shape[0], -1) != bits(b).reshape(b.shape[0], -1))[0]
        raise AssertionError(f"{what}: {len(np.unique(bad))} rows differ bitwise, first row {bad[0]}: "
                             f"{a[bad[0]]} vs {b[bad[0]]}")


# Single-step tolerances from identical state (the reference against itself with
# only the summation order changed gives max rel d(rho) 1.2e-7, d(F) 2.1e-5, |dx| 2.4e-7).
RHO_RTOL = 1e-5
POS_ATOL = 1e-5
VEL_ATOL = 1e-4   # v = v + (F/rho + g) dt inherits the force's relative error times |F/rho| dt
FORCE_RTOL = 1e-3  # relative to max(|F_i|, median |F|)


def assert_fields_close(got, want, what="", gas_constant=1.0):
    """got / want: dicts with pos, vel, force, density, pressure in the SAME row order."""
    d = np.abs(got["density"] - want["density"]) / np.abs(want["density"])
    assert d.max() <= RHO_RTOL, f"{what}: density rel err {d.max():.3e}"
    # pressure = k (rho - rho0): its absolute error is k times the density's
    pe = np.abs(got["pressure"] - want["pressure"])
    bound = abs(gas_constant) * RHO_RTOL * np.abs(want["density"]) * 1.01 + 4 * np.spacing(np.abs(want["pressure"]))
    assert (pe <= bound).all(), f"{what}: pressure abs err {pe.max():.3e}"
    fn = np.linalg.norm(want["force"], axis=1)
    scale = np.maximum(fn, np.median(fn) if fn.size else 0.0)
    scale = np.maximum(scale, 1e-30)
    fe = np.linalg.norm(got["force"] - want["force"], axis=1) / scale
    assert fe.max() <= FORCE_RTOL, f"{what}: force rel err {fe.max():.3e}"
    pe = np.abs(got["pos"] - want["pos"]).max()
    assert pe <= POS_ATOL, f"{what}: position abs err {pe:.3e}"
    ve = np.abs(got["vel"] - want["vel"]).max()
    assert ve <= VEL_ATOL, f"{what}: velocity abs err {ve:.3e}"


def by_id(d):
    """Reorder every per-particle array of a step dict into particle-id order."""
    inv = np.argsort(d["id"], kind="stable")
    out = {}
    for k, v in d.items():
        out[k] = v[inv] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == inv.shape[0] else v
    return out
