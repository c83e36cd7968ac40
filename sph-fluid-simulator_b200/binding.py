"""ctypes binding of the C-ABI in include/sph_b200.h (lib/libsph_b200.so).

This is plumbing for the Python-side tests, bench and multi-GPU driver: every call goes straight
to the CUDA library. There is no fallback of any kind — if the library is missing or no CUDA
device is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libsph_b200.so")

SPH_OK = 0
ORDER_DEVICE, ORDER_ID, ORDER_HASH16 = 0, 1, 2
TABLE_SIZE = 262144
NO_PARTICLE = 0xFFFFFFFF


class SphError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"sph_b200 error {code}: {message}")
        self.code = code


class Settings(C.Structure):
    """struct sph_settings."""
    _fields_ = [(k, C.c_float) for k in (
        "mass", "rest_density", "gas_constant", "viscosity", "h", "g", "tension",
        "dt", "box_half_width", "elasticity", "wall_offset")]

    def as_tuple7(self):
        return (self.mass, self.rest_density, self.gas_constant, self.viscosity, self.h, self.g, self.tension)


class Derived(C.Structure):
    """struct sph_derived."""
    _fields_ = [(k, C.c_float) for k in (
        "poly6", "spiky_grad", "spiky_lap", "h2", "self_dens", "mass_poly6", "sphere_scale")]


class Stats(C.Structure):
    """struct sph_stats."""
    _fields_ = [("count", C.c_uint64), ("steps", C.c_uint64), ("grid_origin", C.c_int32 * 3),
                ("grid_dim", C.c_int32 * 3), ("grid_cells", C.c_uint64), ("clamped", C.c_uint64),
                ("nan_count", C.c_uint64), ("mean_density", C.c_double), ("max_density", C.c_double),
                ("kinetic_energy", C.c_double), ("deferred_density", C.c_uint64), ("deferred_forces", C.c_uint64),
                ("nlist_rows", C.c_uint64)]


_lib = None


def build_library() -> None:
    """Compile lib/libsph_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-s", "-C", PKG_DIR], check=True)


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with `make -C {PKG_DIR}` (or __graft_entry__.build()). "
            "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    fp, u16p, u32p, u64p = (C.POINTER(t) for t in (C.c_float, C.c_uint16, C.c_uint32, C.c_uint64))
    sp, hp = C.POINTER(Settings), C.c_void_p
    sig = {
        "sph_settings_default": ([sp], C.c_int),
        "sph_settings_derive": ([sp, C.POINTER(Derived)], C.c_int),
        "sph_create": ([sp, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)], C.c_int),
        "sph_destroy": ([hp], C.c_int),
        "sph_set_settings": ([hp, sp], C.c_int),
        "sph_last_error": ([hp], C.c_char_p),
        "sph_upload": ([hp, C.c_uint64, fp, fp, u32p], C.c_int),
        "sph_upload_device": ([hp, C.c_uint64, C.c_void_p, C.c_void_p], C.c_int),
        "sph_download": ([hp, C.c_int, fp, fp, fp, fp, fp, u16p, u32p], C.c_int),
        "sph_read_positions": ([hp, fp], C.c_int),
        "sph_write_transforms": ([hp, fp], C.c_int),
        "sph_read_positions_device": ([hp, C.c_void_p], C.c_int),
        "sph_write_transforms_device": ([hp, C.c_void_p], C.c_int),
        "sph_count": ([hp], C.c_uint64),
        "sph_capacity": ([hp], C.c_uint64),
        "sph_step": ([hp, C.c_float, C.c_int], C.c_int),
        "sph_sync": ([hp], C.c_int),
        "sph_neighbor_search": ([hp, C.c_int], C.c_int),
        "sph_update_particles_aos": ([hp, C.c_void_p, fp, C.c_uint64, C.c_float], C.c_int),
        "sph_hash_table": ([hp, u32p], C.c_int),
        "sph_neighbor_lists": ([hp, u32p, u64p, u32p, C.c_uint64, u32p], C.c_int),
        "sph_get_stats": ([hp, C.POINTER(Stats)], C.c_int),
        "sph_candidate_count": ([hp, u64p, u64p], C.c_int),
        "sph_enable_pass_timing": ([hp, C.c_int], C.c_int),
        "sph_pass_times": ([hp, fp, u64p], C.c_int),
        "sph_launch_count": ([hp], C.c_uint64),
        "sph_selftest_division": ([hp, C.c_uint64, C.c_uint64, u64p], C.c_int),
        "sph_selftest_packed_dist2": ([hp, C.c_uint64, C.c_uint64, u64p], C.c_int),
        "sph_stream": ([hp], C.c_void_p),
        "sph_scene_cube": ([C.c_int, C.c_float, fp, fp], C.c_int),
        "sph_scene_block": ([C.c_int] * 3 + [C.c_float] * 5 + [C.c_uint, fp, fp], C.c_int),
        "sph_scene_cube_device": ([hp, C.c_int], C.c_int),
        "sph_scene_block_device": ([hp] + [C.c_int] * 3 + [C.c_float] * 4 + [C.c_uint, C.c_int, C.c_int], C.c_int),
        "sph_selftest_glibc_rand": ([C.c_uint, C.c_uint64, C.c_uint64, u64p], C.c_int),
        "sph_set_reset_point": ([hp], C.c_int),
        "sph_reset": ([hp], C.c_int),
        "sph_slab_enable": ([hp, C.c_int], C.c_int),
        "sph_slab_owned": ([hp], C.c_uint64),
        "sph_slab_count": ([hp, C.POINTER(C.c_int32), C.c_int, u64p], C.c_int),
        "sph_slab_pack": ([hp, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_void_p, u64p], C.c_int),
        "sph_slab_append": ([hp, C.c_void_p, C.c_uint64, C.c_int], C.c_int),
        "sph_slab_pack_halo": ([hp, C.c_int32, C.c_int, C.c_void_p, C.c_uint64, u64p], C.c_int),
        "sph_slab_step_density": ([hp], C.c_int),
        "sph_slab_pack_halo_density": ([hp, C.c_int, C.c_void_p], C.c_int),
        "sph_slab_set_ghost_density": ([hp, C.c_int, C.c_void_p, C.c_uint64], C.c_int),
        "sph_slab_step_forces": ([hp, C.c_float], C.c_int),
        "sph_slab_download_owned": ([hp, fp, fp, u32p, C.c_uint64, u64p], C.c_int),
        "sph_slab_xcell_histogram": ([hp, C.c_int32, C.c_uint32, u64p], C.c_int),
        "sph_scene_block_slice": ([C.c_int] * 3 + [C.c_float] * 5 + [C.c_uint, C.c_int, C.c_int, fp, fp, u32p], C.c_int),
        "sph_slab_fast_begin": ([hp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_void_p, C.c_void_p], C.c_int),
        "sph_slab_fast_arrivals": ([hp, C.c_void_p, C.c_void_p, C.c_uint64], C.c_int),
        "sph_slab_fast_halo": ([hp, C.c_int32, C.c_int32, C.c_uint64, C.c_void_p, C.c_void_p], C.c_int),
        "sph_slab_fast_ghosts": ([hp, C.c_void_p, C.c_void_p, C.c_uint64], C.c_int),
        "sph_slab_fast_pack_density": ([hp, C.c_uint64, C.c_void_p, C.c_void_p], C.c_int),
        "sph_slab_fast_set_ghost_density": ([hp, C.c_void_p, C.c_void_p, C.c_uint64], C.c_int),
        "sph_slab_p2p_create": ([hp, C.c_uint64, C.c_uint64, C.c_void_p], C.c_int),
        "sph_slab_p2p_connect": ([hp, C.c_int, C.c_void_p], C.c_int),
        "sph_slab_p2p_begin": ([hp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint64], C.c_int),
        "sph_slab_p2p_arrivals": ([hp], C.c_int),
        "sph_slab_p2p_density": ([hp], C.c_int),
        "sph_system_create": ([C.c_int, sp, C.c_int, C.c_int, C.POINTER(C.c_void_p)], C.c_int),
        "sph_system_destroy": ([hp], C.c_int),
        "sph_system_last_error": ([hp], C.c_char_p),
        "sph_system_start": ([hp], C.c_int),
        "sph_system_update": ([hp, C.c_float], C.c_int),
        "sph_system_reset": ([hp], C.c_int),
        "sph_system_count": ([hp], C.c_uint64),
        "sph_system_handle": ([hp], C.c_void_p),
        "sph_system_positions": ([hp, fp], C.c_int),
        "sph_system_model_matrices": ([hp, fp], C.c_int),
        "sph_system_download": ([hp, fp, fp], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def default_settings(**overrides) -> Settings:
    s = Settings()
    load_library().sph_settings_default(C.byref(s))
    for k, v in overrides.items():
        setattr(s, k, v)
    return s


def scaled_settings(h: float) -> Settings:
    """Geometric scaling of the shipped defaults to a kernel radius h:
    mass = 0.02*(h/0.15)^3 keeps selfDens, dt = 0.003*(h/0.15)."""
    s = default_settings()
    k = np.float32(h) / np.float32(0.15)
    s.h = h
    s.mass = float(np.float32(0.02) * k * k * k)
    s.dt = float(np.float32(0.003) * k)
    return s


def derive(s: Settings) -> Derived:
    d = Derived()
    load_library().sph_settings_derive(C.byref(s), C.byref(d))
    return d


def scene_cube(width: int, h: float):
    n = width ** 3
    pos = np.empty((n, 3), np.float32)
    vel = np.empty((n, 3), np.float32)
    rc = load_library().sph_scene_cube(width, C.c_float(h), _ptr(pos, C.c_float), _ptr(vel, C.c_float))
    if rc:
        raise SphError(rc, "sph_scene_cube")
    return pos, vel


def scene_block(nx, ny, nz, sep, origin, h, seed=1024):
    n = nx * ny * nz
    pos = np.empty((n, 3), np.float32)
    vel = np.empty((n, 3), np.float32)
    rc = load_library().sph_scene_block(nx, ny, nz, C.c_float(sep), C.c_float(origin[0]), C.c_float(origin[1]),
                                        C.c_float(origin[2]), C.c_float(h), seed, _ptr(pos, C.c_float),
                                        _ptr(vel, C.c_float))
    if rc:
        raise SphError(rc, "sph_scene_block")
    return pos, vel


def scene_block_slice(nx, ny, nz, sep, origin, h, seed, i0, i1):
    n = (i1 - i0) * ny * nz
    pos = np.empty((n, 3), np.float32)
    vel = np.empty((n, 3), np.float32)
    ids = np.empty(n, np.uint32)
    rc = load_library().sph_scene_block_slice(nx, ny, nz, C.c_float(sep), C.c_float(origin[0]), C.c_float(origin[1]),
                                              C.c_float(origin[2]), C.c_float(h), seed, i0, i1, _ptr(pos, C.c_float),
                                              _ptr(vel, C.c_float), _ptr(ids, C.c_uint32))
    if rc:
        raise SphError(rc, "sph_scene_block_slice")
    return pos, vel, ids


class Sim:
    """One sph_handle: persistent particle state on one GPU."""

    def __init__(self, settings: Settings | None = None, capacity: int = 1 << 20, device: int = 0):
        self.lib = load_library()
        self.settings = settings if settings is not None else default_settings()
        self._h = C.c_void_p()
        rc = self.lib.sph_create(C.byref(self.settings), capacity, device, C.byref(self._h))
        if rc:
            raise SphError(rc, self.lib.sph_last_error(None).decode())

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, rc):
        if rc:
            raise SphError(rc, self.lib.sph_last_error(self._h).decode())

    def close(self):
        if self._h:
            self.lib.sph_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def count(self) -> int:
        return int(self.lib.sph_count(self._h))

    @property
    def stream(self) -> int:
        return int(self.lib.sph_stream(self._h) or 0)

    # -- state ------------------------------------------------------------------------------
    def set_settings(self, s: Settings):
        self._ck(self.lib.sph_set_settings(self._h, C.byref(s)))
        self.settings = s

    def upload(self, pos, vel, ids=None):
        pos = np.ascontiguousarray(pos, np.float32)
        vel = np.ascontiguousarray(vel, np.float32)
        ids = None if ids is None else np.ascontiguousarray(ids, np.uint32)
        self._ck(self.lib.sph_upload(self._h, pos.shape[0], _ptr(pos, C.c_float), _ptr(vel, C.c_float),
                                     _ptr(ids, C.c_uint32)))

    def scene_cube_device(self, width):
        """initParticles on the device (bit-identical to scene_cube + upload)."""
        self._ck(self.lib.sph_scene_cube_device(self._h, width))

    def scene_block_device(self, nx, ny, nz, sep, origin, seed=1024, i0=0, i1=None):
        """The dam-break block (rows with lattice x-index in [i0, i1)) generated on the device."""
        self._ck(self.lib.sph_scene_block_device(self._h, nx, ny, nz, C.c_float(sep), C.c_float(origin[0]), C.c_float(origin[1]),
                                                 C.c_float(origin[2]), seed, i0, nx if i1 is None else i1))

    def set_reset_point(self):
        self._ck(self.lib.sph_set_reset_point(self._h))

    def reset(self):
        self._ck(self.lib.sph_reset(self._h))

    def upload_device(self, n, dev_pos_ptr, dev_vel_ptr):
        self._ck(self.lib.sph_upload_device(self._h, n, C.c_void_p(dev_pos_ptr), C.c_void_p(dev_vel_ptr)))

    def download(self, order=ORDER_ID, fields=("pos", "vel", "force", "density", "pressure", "hash", "id")):
        n = self.count
        out = {}
        if "pos" in fields: out["pos"] = np.empty((n, 3), np.float32)
        if "vel" in fields: out["vel"] = np.empty((n, 3), np.float32)
        if "force" in fields: out["force"] = np.empty((n, 3), np.float32)
        if "density" in fields: out["density"] = np.empty(n, np.float32)
        if "pressure" in fields: out["pressure"] = np.empty(n, np.float32)
        if "hash" in fields: out["hash"] = np.empty(n, np.uint16)
        if "id" in fields: out["id"] = np.empty(n, np.uint32)
        self._ck(self.lib.sph_download(
            self._h, order, _ptr(out.get("pos"), C.c_float), _ptr(out.get("vel"), C.c_float),
            _ptr(out.get("force"), C.c_float), _ptr(out.get("density"), C.c_float),
            _ptr(out.get("pressure"), C.c_float), _ptr(out.get("hash"), C.c_uint16),
            _ptr(out.get("id"), C.c_uint32)))
        return out

    def read_positions(self):
        out = np.empty((self.count, 4), np.float32)
        self._ck(self.lib.sph_read_positions(self._h, _ptr(out, C.c_float)))
        return out

    def write_transforms(self):
        out = np.empty((self.count, 16), np.float32)
        self._ck(self.lib.sph_write_transforms(self._h, _ptr(out, C.c_float)))
        return out

    # -- stepping ---------------------------------------------------------------------------
    def step(self, nsteps=1, dt=0.0):
        self._ck(self.lib.sph_step(self._h, C.c_float(dt), nsteps))

    def sync(self):
        self._ck(self.lib.sph_sync(self._h))

    def neighbor_search(self, repeats=1):
        self._ck(self.lib.sph_neighbor_search(self._h, repeats))

    def update_particles_aos(self, particles: np.ndarray, dt=0.0, transforms=True):
        """particles: (n, 15) uint32 view of reference Particle rows (60 bytes each), in place."""
        assert particles.dtype == np.uint32 and particles.shape[1] == 15 and particles.flags.c_contiguous
        n = particles.shape[0]
        mats = np.empty((n, 16), np.float32) if transforms else None
        self._ck(self.lib.sph_update_particles_aos(self._h, particles.ctypes.data_as(C.c_void_p),
                                                   _ptr(mats, C.c_float), n, C.c_float(dt)))
        return mats

    # -- parity / diagnostics ---------------------------------------------------------------
    def hash_table(self):
        out = np.empty(TABLE_SIZE, np.uint32)
        self._ck(self.lib.sph_hash_table(self._h, _ptr(out, C.c_uint32)))
        return out

    def neighbor_lists(self):
        """Returns (ids, counts, offsets, list) in device-row order; list holds neighbour ids."""
        n = self.count
        counts = np.empty(n, np.uint32)
        offsets = np.empty(n + 1, np.uint64)
        ids = np.empty(n, np.uint32)
        self._ck(self.lib.sph_neighbor_lists(self._h, _ptr(counts, C.c_uint32), _ptr(offsets, C.c_uint64), None, 0,
                                             _ptr(ids, C.c_uint32)))
        total = int(offsets[n])
        lst = np.empty(max(total, 1), np.uint32)
        self._ck(self.lib.sph_neighbor_lists(self._h, _ptr(counts, C.c_uint32), _ptr(offsets, C.c_uint64),
                                             _ptr(lst, C.c_uint32), max(total, 1), _ptr(ids, C.c_uint32)))
        return ids, counts, offsets, lst[:total]

    def stats(self) -> Stats:
        st = Stats()
        self._ck(self.lib.sph_get_stats(self._h, C.byref(st)))
        return st

    def candidates_mean(self) -> float:
        """Mean number of candidate rows per owned row in the last step (rows of the 27 cells, self included)."""
        c, r = C.c_uint64(0), C.c_uint64(0)
        self._ck(self.lib.sph_candidate_count(self._h, C.byref(c), C.byref(r)))
        return c.value / max(r.value, 1)

    def enable_pass_timing(self, on=True):
        self._ck(self.lib.sph_enable_pass_timing(self._h, int(on)))

    def pass_times(self):
        """Mean ms per step of each pass over the window since the last call, and the window size."""
        out = np.empty(4, np.float32)
        nsteps = C.c_uint64(0)
        self._ck(self.lib.sph_pass_times(self._h, _ptr(out, C.c_float), C.byref(nsteps)))
        d = dict(zip(("grid", "density", "forces", "integrate"), out.tolist()))
        d["steps"] = int(nsteps.value)
        return d

    def selftest_division(self, n=1 << 22, seed=1) -> int:
        bad = C.c_uint64(0)
        self._ck(self.lib.sph_selftest_division(self._h, n, seed, C.byref(bad)))
        return int(bad.value)

    def selftest_packed_dist2(self, n=1 << 22, seed=1) -> int:
        bad = C.c_uint64(0)
        self._ck(self.lib.sph_selftest_packed_dist2(self._h, n, seed, C.byref(bad)))
        return int(bad.value)

    @property
    def launch_count(self) -> int:
        return int(self.lib.sph_launch_count(self._h))


class System:
    """class SPHSystem through the C-ABI (sph_system_*): the reference's simulation object."""

    def __init__(self, cube_width: int, settings: Settings | None = None, run_on_gpu: bool = True, device: int = 0):
        self.lib = load_library()
        self.settings = settings if settings is not None else default_settings()
        self._s = C.c_void_p()
        rc = self.lib.sph_system_create(cube_width, C.byref(self.settings), int(run_on_gpu), device, C.byref(self._s))
        if rc:
            raise SphError(rc, self.lib.sph_system_last_error(None).decode())

    def _ck(self, rc):
        if rc:
            raise SphError(rc, self.lib.sph_system_last_error(self._s).decode())

    def close(self):
        if self._s:
            self.lib.sph_system_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def particleCount(self) -> int:
        return int(self.lib.sph_system_count(self._s))

    def startSimulation(self):
        self._ck(self.lib.sph_system_start(self._s))

    def update(self, dt: float):
        self._ck(self.lib.sph_system_update(self._s, C.c_float(dt)))

    def reset(self):
        self._ck(self.lib.sph_system_reset(self._s))

    def positions(self):
        out = np.empty((self.particleCount, 4), np.float32)
        self._ck(self.lib.sph_system_positions(self._s, _ptr(out, C.c_float)))
        return out

    def model_matrices(self):
        out = np.empty((self.particleCount, 16), np.float32)
        self._ck(self.lib.sph_system_model_matrices(self._s, _ptr(out, C.c_float)))
        return out

    def download(self):
        n = self.particleCount
        pos = np.empty((n, 3), np.float32)
        vel = np.empty((n, 3), np.float32)
        self._ck(self.lib.sph_system_download(self._s, _ptr(pos, C.c_float), _ptr(vel, C.c_float)))
        return pos, vel
