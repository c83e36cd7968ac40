"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares;
host-side pieces that need no GPU (settings derivation, scene generators) match the golden
fixtures; and the product fails loudly — no fallback — when no CUDA device is usable."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

from conftest import ROOT, assert_bit_equal, load_golden


def declared_functions():
    names = []
    for path in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
        names += re.findall(r"\b(sph_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("sph_create", "sph_destroy", "sph_upload", "sph_step", "sph_download", "sph_read_positions",
                 "sph_write_transforms", "sph_update_particles_aos", "sph_last_error", "sph_system_create",
                 "sph_system_update", "sph_system_reset", "sph_system_start"):
        assert must in names


def test_library_exports_every_declared_symbol(sph):
    lib = ctypes.CDLL(sph.binding.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_library_is_built_for_sm_100a(sph):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", sph.binding.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_settings_derivation_known_answers(sph):
    g = load_golden("settings_kat.npz")
    names = [str(n) for n in g["names"]]
    m = dict(poly6="poly6", spikyGrad="spiky_grad", spikyLap="spiky_lap", h2="h2", selfDens="self_dens",
             massPoly6Product="mass_poly6", sphereScale="sphere_scale")
    for row_in, row_out in zip(g["inputs"], g["outputs"]):
        s = sph.default_settings()
        s.mass, s.rest_density, s.gas_constant, s.viscosity, s.h, s.g, s.tension = [float(v) for v in row_in]
        d = sph.derive(s)
        for n, want in zip(names, row_out):
            if n in m:
                assert np.float32(getattr(d, m[n])).view(np.uint32) == np.float32(want).view(np.uint32), n


def test_default_settings_are_the_reference_defaults(sph):
    s = sph.default_settings()
    assert s.as_tuple7() == pytest.approx((0.02, 1000.0, 1.0, 1.04, 0.15, -9.8, 0.2))
    assert (s.dt, s.box_half_width, s.elasticity) == pytest.approx((0.003, 8.0, 0.5))
    assert s.wall_offset == pytest.approx(1e-4)


@pytest.mark.parametrize("w", [1, 2, 15])
def test_scene_cube_is_init_particles(sph, w):
    g = load_golden(f"init_cube_w{w}.npz")
    pos, vel = sph.scene_cube(w, 0.15)
    assert_bit_equal(pos, g["pos"], "scene_cube")
    assert not vel.any()


def test_scene_block_matches_oracle_generator(sph, oracle):
    a = sph.scene_block(5, 6, 7, 0.08, (-7.8, 0.125, -0.3), 0.075, 99)
    b = oracle.init_block(5, 6, 7, 0.08, (-7.8, 0.125, -0.3), 0.075, 99)
    assert_bit_equal(a[0], b[0], "scene_block")
    assert sph.scene_block(0, 3, 3, 0.1, (0, 0, 0), 0.1)[0].shape == (0, 3)


def test_rand_jump_ahead_reproduces_glibc_rand(sph):
    """The device scene generators rebuild glibc's rand() stream chunk by chunk from jumped states
    (x^n mod x^31 - x^28 - 1 over Z/2^32); this host-only self-test compares with rand() itself."""
    lib = sph.load_library()
    for seed, ndraws, chunk in ((1024, 2_000_000, 3072), (1, 50_000, 7), (0, 10_000, 1000), (4_000_000_000, 20_000, 31),
                                (99, 400_000, 3 * 4096)):
        bad = ctypes.c_uint64(1)
        assert lib.sph_selftest_glibc_rand(seed, ndraws, chunk, ctypes.byref(bad)) == 0
        assert bad.value == 0, (seed, ndraws, chunk, bad.value)


def test_no_fallback_without_a_gpu(sph):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(sph.SphError) as e:
        sph.Sim(sph.default_settings(), capacity=16)
    assert "cuda" in str(e.value).lower()
    with pytest.raises(sph.SphError):
        sph.System(3)


def test_null_and_bad_arguments_do_not_crash(sph):
    lib = sph.load_library()
    assert lib.sph_destroy(None) != 0
    assert lib.sph_step(None, ctypes.c_float(0.003), 1) != 0
    assert lib.sph_count(None) == 0
    assert lib.sph_last_error(None) is not None
    out = ctypes.c_void_p()
    assert lib.sph_create(None, 10, 0, ctypes.byref(out)) != 0 and not out.value
    s = sph.default_settings()
    s.h = 0.0
    assert lib.sph_create(ctypes.byref(s), 10, 0, ctypes.byref(out)) != 0
    assert b"h must be" in lib.sph_last_error(None)


def test_header_is_plain_c99_and_cxx11(tmp_path):
    """The boundary is a C ABI: the header must compile as C (not only as C++) with warnings on."""
    import shutil
    import subprocess
    header = os.path.join(ROOT, "include", "sph_b200.h")
    for compiler, std, ext in (("gcc", "-std=c99", "c"), ("g++", "-std=c++11", "cpp")):
        exe = "/usr/bin/" + compiler if os.path.exists("/usr/bin/" + compiler) else shutil.which(compiler)
        src = tmp_path / f"hdr.{ext}"
        src.write_text(f'#include "{header}"\nint main(void) {{ sph_settings s; (void)s; return SPH_OK; }}\n')
        subprocess.run([exe, std, "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", str(src)], check=True)


def test_c_example_links_against_the_library(sph, tmp_path):
    """examples/step_cube.c is the smallest host of the C-ABI; without a GPU it must stop with the
    library's error message, not fall back to anything."""
    import shutil
    import subprocess
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    libdir = os.path.dirname(sph.binding.LIB_PATH)
    exe = tmp_path / "step_cube"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-O2", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "step_cube.c"), "-L", libdir, "-lsph_b200", f"-Wl,-rpath,{libdir}",
                    "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    import torch
    if torch.cuda.is_available():
        assert out.returncode == 0 and "step 1000" in out.stdout, out.stderr
    else:
        assert out.returncode == 1 and "sph_create" in out.stderr and "failed" in out.stderr
