"""TEST INFRASTRUCTURE — ctypes loaders for the parity oracle. Never imported by the product.

Two libraries:
  * ``Oracle``    — oracle/build/libsph_oracle.so, the plain-C restatement (sph_oracle.c);
  * ``Reference`` — oracle/_ref/libsph_ref.so, the UNMODIFIED reference sources compiled headless
                    (ref_harness.cpp). Built only where /root/reference exists; the prebuilt file
                    travels to the GPU box with the snapshot.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "build", "libsph_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsph_ref.so")

DEFAULT_SETTINGS = (0.02, 1000.0, 1.0, 1.04, 0.15, -9.8, 0.2)  # reference src/Tester.cpp:90
TABLE_SIZE = 262144
NO_PARTICLE = 0xFFFFFFFF


def build(ref: bool = True) -> None:
    """Compile the oracle (and, where /root/reference exists, the reference harness)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class OracleSettings(C.Structure):
    _fields_ = [(k, C.c_float) for k in (
        "mass", "restDensity", "gasConstant", "viscosity", "h", "g", "tension",
        "poly6", "spikyGrad", "spikyLap", "h2", "selfDens", "massPoly6Product", "sphereScale")]


class Oracle:
    """The plain-C restatement."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        fp, u16p, u32p, u64p, ip = (C.POINTER(t) for t in (C.c_float, C.c_uint16, C.c_uint32, C.c_uint64, C.c_int))
        sp = C.POINTER(OracleSettings)
        L.oracle_make_settings.argtypes = [C.c_float] * 7 + [sp]
        L.oracle_get_cell.argtypes = [fp, C.c_float, ip]
        L.oracle_get_hash.argtypes = [ip]
        L.oracle_get_hash.restype = C.c_uint32
        L.oracle_hashes.argtypes = [C.c_uint64, fp, C.c_float, u16p]
        L.oracle_sort_order.argtypes = [C.c_uint64, u16p, u32p]
        L.oracle_neighbor_table.argtypes = [C.c_uint64, u16p, u32p]
        L.oracle_density_pressure.argtypes = [C.c_uint64, fp, u16p, u32p, sp, fp, fp]
        L.oracle_forces.argtypes = [C.c_uint64, fp, fp, fp, fp, u16p, u32p, sp, fp]
        L.oracle_integrate.argtypes = [C.c_uint64, fp, fp, fp, fp, sp, C.c_float, fp]
        L.oracle_step.argtypes = [C.c_uint64, sp, C.c_float, fp, fp, u32p, u32p, fp, fp, fp, u16p, fp]
        L.oracle_step.restype = C.c_int
        L.oracle_neighbor_lists.argtypes = [C.c_uint64, fp, u16p, u32p, sp, u32p, u32p, u64p, u32p]
        L.oracle_neighbor_lists.restype = C.c_uint64
        L.oracle_init_cube.argtypes = [C.c_int, sp, fp, fp]
        L.oracle_init_block.argtypes = [C.c_int] * 3 + [C.c_float] * 5 + [C.c_uint, fp, fp]
        L.oracle_time_steps.argtypes = [C.c_uint64, sp, C.c_float, C.c_int, C.c_int, fp, fp]
        L.oracle_time_steps.restype = C.c_double
        L.oracle_num_threads.restype = C.c_int

    # -- settings / hashing -------------------------------------------------------------
    def settings(self, s7=DEFAULT_SETTINGS) -> OracleSettings:
        out = OracleSettings()
        self.lib.oracle_make_settings(*[C.c_float(v) for v in s7], C.byref(out))
        return out

    def get_cell(self, p, h):
        a = _f32(p)
        out = (C.c_int * 3)()
        self.lib.oracle_get_cell(_p(a, C.c_float), C.c_float(h), out)
        return tuple(out)

    def get_hash(self, cell) -> int:
        c = (C.c_int * 3)(*[int(v) for v in cell])
        return int(self.lib.oracle_get_hash(c))

    def hashes(self, pos, h):
        pos = _f32(pos)
        n = pos.shape[0]
        out = np.empty(n, np.uint16)
        self.lib.oracle_hashes(n, _p(pos, C.c_float), C.c_float(h), _p(out, C.c_uint16))
        return out

    def sort_order(self, hash16):
        hash16 = np.ascontiguousarray(hash16, np.uint16)
        out = np.empty(hash16.shape[0], np.uint32)
        self.lib.oracle_sort_order(hash16.shape[0], _p(hash16, C.c_uint16), _p(out, C.c_uint32))
        return out

    def neighbor_table(self, sorted_hash):
        sorted_hash = np.ascontiguousarray(sorted_hash, np.uint16)
        out = np.empty(TABLE_SIZE, np.uint32)
        self.lib.oracle_neighbor_table(sorted_hash.shape[0], _p(sorted_hash, C.c_uint16), _p(out, C.c_uint32))
        return out

    # -- one step -----------------------------------------------------------------------
    def step(self, s: OracleSettings, dt, pos, vel, ids=None, order=None, transforms=False):
        """Returns a dict of post-step arrays in the post-sort order."""
        pos = _f32(pos).copy()
        vel = _f32(vel).copy()
        n = pos.shape[0]
        ids = np.arange(n, dtype=np.uint32) if ids is None else np.ascontiguousarray(ids, np.uint32).copy()
        force = np.empty((n, 3), np.float32)
        dens = np.empty(n, np.float32)
        pres = np.empty(n, np.float32)
        h16 = np.empty(n, np.uint16)
        tr = np.empty((n, 16), np.float32) if transforms else None
        if order is not None:
            order = np.ascontiguousarray(order, np.uint32)
        rc = self.lib.oracle_step(n, C.byref(s), C.c_float(dt), _p(pos, C.c_float), _p(vel, C.c_float),
                                  _p(ids, C.c_uint32), _p(order, C.c_uint32), _p(force, C.c_float),
                                  _p(dens, C.c_float), _p(pres, C.c_float), _p(h16, C.c_uint16),
                                  _p(tr, C.c_float))
        if rc != 0:
            raise ValueError("oracle_step: supplied order is not hash-sorted")
        return dict(pos=pos, vel=vel, id=ids, force=force, density=dens, pressure=pres, hash=h16,
                    transforms=tr)

    def neighbor_lists(self, s: OracleSettings, pos):
        """Neighbour multisets from an UNSORTED state: returns (order, counts, cand, offsets, list)
        where list holds sorted-array indices; order maps sorted slot -> input index."""
        pos = _f32(pos)
        n = pos.shape[0]
        h16 = self.hashes(pos, s.h)
        order = self.sort_order(h16)
        spos = np.ascontiguousarray(pos[order])
        sh = np.ascontiguousarray(h16[order])
        table = self.neighbor_table(sh)
        counts = np.empty(n, np.uint32)
        cand = np.empty(n, np.uint32)
        args = (n, _p(spos, C.c_float), _p(sh, C.c_uint16), _p(table, C.c_uint32), C.byref(s))
        self.lib.oracle_neighbor_lists(*args, _p(counts, C.c_uint32), _p(cand, C.c_uint32), None, None)
        offsets = np.zeros(n + 1, np.uint64)
        np.cumsum(counts, out=offsets[1:])
        lst = np.empty(int(offsets[-1]), np.uint32)
        self.lib.oracle_neighbor_lists(*args, _p(counts, C.c_uint32), None, _p(offsets, C.c_uint64),
                                       _p(lst, C.c_uint32))
        return order, counts, cand, offsets, lst

    # -- scenes -------------------------------------------------------------------------
    def init_cube(self, width, s: OracleSettings):
        n = width ** 3
        pos = np.empty((n, 3), np.float32)
        vel = np.empty((n, 3), np.float32)
        self.lib.oracle_init_cube(width, C.byref(s), _p(pos, C.c_float), _p(vel, C.c_float))
        return pos, vel

    def init_block(self, nx, ny, nz, sep, origin, h, seed=1024):
        n = nx * ny * nz
        pos = np.empty((n, 3), np.float32)
        vel = np.empty((n, 3), np.float32)
        self.lib.oracle_init_block(nx, ny, nz, C.c_float(sep), C.c_float(origin[0]), C.c_float(origin[1]),
                                   C.c_float(origin[2]), C.c_float(h), seed, _p(pos, C.c_float),
                                   _p(vel, C.c_float))
        return pos, vel

    def time_steps(self, s, dt, warmup, steps, pos, vel):
        pos = _f32(pos).copy()
        vel = _f32(vel).copy()
        sec = self.lib.oracle_time_steps(pos.shape[0], C.byref(s), C.c_float(dt), warmup, steps,
                                         _p(pos, C.c_float), _p(vel, C.c_float))
        return float(sec), pos, vel

    def num_threads(self) -> int:
        return int(self.lib.oracle_num_threads())


class Reference:
    """The unmodified reference step, compiled headless (oracle/_ref/libsph_ref.so)."""

    GPU_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_float)

    @staticmethod
    def available(path: str = REF_SO) -> bool:
        return os.path.exists(path)

    def __init__(self, path: str = REF_SO):
        L = self.lib = C.CDLL(path)
        fp, u16p, u32p, ip = (C.POINTER(t) for t in (C.c_float, C.c_uint16, C.c_uint32, C.c_int))
        L.ref_sizeof_particle.restype = C.c_int
        L.ref_sizeof_settings.restype = C.c_int
        L.ref_hardware_concurrency.restype = C.c_int
        L.ref_table_size.restype = C.c_uint32
        L.ref_make_settings.argtypes = [fp, fp, fp]
        L.ref_get_hash.argtypes = [C.c_int] * 3
        L.ref_get_hash.restype = C.c_uint32
        L.ref_get_cell.argtypes = [C.c_float] * 4 + [ip]
        L.ref_neighbor_table.argtypes = [C.c_uint64, u16p, u32p]
        L.ref_init_cube.argtypes = [C.c_int, fp, fp, fp, fp]
        L.ref_class_run.argtypes = [C.c_int, fp, C.c_int, C.c_int, C.c_int, fp, fp, u32p]
        L.ref_step.argtypes = [C.c_uint64, fp, C.c_float, C.c_int, C.c_int, fp, fp, u32p, fp, fp, fp, u16p, fp]
        L.ref_time_steps.argtypes = [C.c_uint64, fp, C.c_float, C.c_int, C.c_int, fp, fp]
        L.ref_time_steps.restype = C.c_double
        L.ref_set_gpu_hook.argtypes = [C.c_void_p]

    @staticmethod
    def _s7(s7):
        return (C.c_float * 7)(*s7)

    def make_settings(self, s7=DEFAULT_SETTINGS):
        out = np.empty(13, np.float32)
        scale = np.empty(16, np.float32)
        self.lib.ref_make_settings(self._s7(s7), _p(out, C.c_float), _p(scale, C.c_float))
        names = ("poly6", "spikyGrad", "spikyLap", "gasConstant", "mass", "h2", "selfDens", "restDensity",
                 "viscosity", "h", "g", "tension", "massPoly6Product")
        d = dict(zip(names, out))
        d["sphereScale"] = scale
        return d

    def get_hash(self, cell) -> int:
        return int(self.lib.ref_get_hash(int(cell[0]), int(cell[1]), int(cell[2])))

    def get_cell(self, p, h):
        out = (C.c_int * 3)()
        self.lib.ref_get_cell(C.c_float(p[0]), C.c_float(p[1]), C.c_float(p[2]), C.c_float(h), out)
        return tuple(out)

    def neighbor_table(self, sorted_hash):
        sorted_hash = np.ascontiguousarray(sorted_hash, np.uint16)
        out = np.empty(TABLE_SIZE, np.uint32)
        self.lib.ref_neighbor_table(sorted_hash.shape[0], _p(sorted_hash, C.c_uint16), _p(out, C.c_uint32))
        return out

    def init_cube(self, width, s7=DEFAULT_SETTINGS):
        n = width ** 3
        pos = np.empty((n, 3), np.float32)
        vel = np.empty((n, 3), np.float32)
        self.lib.ref_init_cube(width, self._s7(s7), _p(pos, C.c_float), _p(vel, C.c_float), None)
        return pos, vel

    def class_run(self, width, s7, nsteps, start=True, reset_after=False):
        n = width ** 3
        pos = np.empty((n, 3), np.float32)
        vel = np.empty((n, 3), np.float32)
        ids = np.empty(n, np.uint32)
        self.lib.ref_class_run(width, self._s7(s7), nsteps, int(start), int(reset_after),
                               _p(pos, C.c_float), _p(vel, C.c_float), _p(ids, C.c_uint32))
        return pos, vel, ids

    def step(self, s7, dt, pos, vel, ids=None, nsteps=1, on_gpu=False, transforms=False):
        pos = _f32(pos).copy()
        vel = _f32(vel).copy()
        n = pos.shape[0]
        ids = np.arange(n, dtype=np.uint32) if ids is None else np.ascontiguousarray(ids, np.uint32).copy()
        force = np.empty((n, 3), np.float32)
        dens = np.empty(n, np.float32)
        pres = np.empty(n, np.float32)
        h16 = np.empty(n, np.uint16)
        tr = np.empty((n, 16), np.float32) if transforms else None
        self.lib.ref_step(n, self._s7(s7), C.c_float(dt), nsteps, int(on_gpu), _p(pos, C.c_float),
                          _p(vel, C.c_float), _p(ids, C.c_uint32), _p(force, C.c_float), _p(dens, C.c_float),
                          _p(pres, C.c_float), _p(h16, C.c_uint16), _p(tr, C.c_float))
        return dict(pos=pos, vel=vel, id=ids, force=force, density=dens, pressure=pres, hash=h16,
                    transforms=tr)

    def time_steps(self, s7, dt, warmup, steps, pos, vel):
        pos = _f32(pos).copy()
        vel = _f32(vel).copy()
        sec = self.lib.ref_time_steps(pos.shape[0], self._s7(s7), C.c_float(dt), warmup, steps,
                                      _p(pos, C.c_float), _p(vel, C.c_float))
        return float(sec), pos, vel

    def hardware_concurrency(self) -> int:
        return int(self.lib.ref_hardware_concurrency())

    def set_gpu_hook(self, fn_ptr) -> None:
        self.lib.ref_set_gpu_hook(fn_ptr)
