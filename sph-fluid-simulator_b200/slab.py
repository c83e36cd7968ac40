"""Slab decomposition of the SPH step across the GPUs of one box: one process per GPU,
`torch.distributed` for the plumbing (NCCL over NVLink; gloo with host staging for tests), the
device-side work in the CUDA library (csrc/sph_slab.cuh through the sph_slab_* C-ABI).

The reference has no multi-GPU path: this is new work, validated by particle id
against the single-GPU path and the oracle (tests/test_slab_gloo.py, tests/test_gpu_slab.py).

Scheme (1-D slabs along x, cut at cell boundaries; rank r owns cell.x in [cuts[r], cuts[r+1])):

    migrate   every rank sends the particles whose cell.x left its slab straight to their owner
              (all-to-all: counts, then rows of 32 bytes); last step's ghosts are dropped
    halo      the owned particles of the slab's first / last x-cell go to the left / right
              neighbour as ghosts (positions + velocities)
    density   grid build over owned + ghosts, density and neighbour lists of the owned particles
    halo rho  the densities of the same boundary particles follow, 4 bytes each, in the same order
    forces    forces + integration of the owned particles

Only the halo legs involve neighbours; the all-to-all carries nothing between non-adjacent ranks
unless the cuts were just rebalanced. Cells are ordered by particle id inside the library, so the
result does not depend on the order rows arrive in.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

INT_MIN, INT_MAX = -2 ** 31, 2 ** 31 - 1
ROW_FLOATS = 8  # (x, y, z, id bits), (vx, vy, vz, 0)


# ------------------------------------------------------------------------------------------------
# engine: the per-rank device primitives
# ------------------------------------------------------------------------------------------------
class GpuEngine:
    """The CUDA library behind the slab primitives. Buffers are torch CUDA tensors allocated on the
    library's stream so that NCCL collectives issued by torch are ordered with the kernels."""

    def __init__(self, sim, device_index: int):
        self.sim = sim
        self.lib = sim.lib
        self.device = torch.device("cuda", device_index)
        self.stream = torch.cuda.ExternalStream(sim.stream, device=self.device)
        self._ck(self.lib.sph_slab_enable(sim.handle, 1))

    def _ck(self, rc):
        if rc:
            raise RuntimeError(self.lib.sph_last_error(self.sim.handle).decode())

    @staticmethod
    def _cuts(cuts):
        return (C.c_int32 * len(cuts))(*[int(c) for c in cuts])

    def empty_rows(self, n):
        with torch.cuda.stream(self.stream):
            return torch.empty((int(n), ROW_FLOATS), dtype=torch.float32, device=self.device)

    def empty_floats(self, n):
        with torch.cuda.stream(self.stream):
            return torch.empty(int(n), dtype=torch.float32, device=self.device)

    @property
    def owned(self) -> int:
        return int(self.lib.sph_slab_owned(self.sim.handle))

    def count(self, cuts):
        world = len(cuts) - 1
        out = (C.c_uint64 * world)()
        self._ck(self.lib.sph_slab_count(self.sim.handle, self._cuts(cuts), world, out))
        return np.array(out[:], dtype=np.int64)

    def pack(self, cuts, rank, offsets, total):
        world = len(cuts) - 1
        buf = self.empty_rows(max(total, 1))
        off = (C.c_uint64 * world)(*[int(o) for o in offsets])
        self._ck(self.lib.sph_slab_pack(self.sim.handle, self._cuts(cuts), world, rank, C.c_void_p(buf.data_ptr()), off))
        return buf[:total]

    def append(self, rows, kind):
        rows = rows.contiguous()
        self._ck(self.lib.sph_slab_append(self.sim.handle, C.c_void_p(rows.data_ptr()), rows.shape[0], kind))

    def pack_halo(self, cell_x, side, capacity):
        buf = self.empty_rows(max(capacity, 1))
        n = C.c_uint64(0)
        self._ck(self.lib.sph_slab_pack_halo(self.sim.handle, int(cell_x), side, C.c_void_p(buf.data_ptr()), capacity,
                                             C.byref(n)))
        return buf[:int(n.value)]

    def step_density(self):
        self._ck(self.lib.sph_slab_step_density(self.sim.handle))

    def pack_halo_density(self, side, n):
        buf = self.empty_floats(max(n, 1))
        self._ck(self.lib.sph_slab_pack_halo_density(self.sim.handle, side, C.c_void_p(buf.data_ptr())))
        return buf[:n]

    def set_ghost_density(self, side, values):
        values = values.contiguous()
        self._ck(self.lib.sph_slab_set_ghost_density(self.sim.handle, side, C.c_void_p(values.data_ptr()), values.shape[0]))

    def step_forces(self, dt):
        self._ck(self.lib.sph_slab_step_forces(self.sim.handle, C.c_float(dt)))

    def xcell_histogram(self, x_lo, nbins):
        out = (C.c_uint64 * nbins)()
        self._ck(self.lib.sph_slab_xcell_histogram(self.sim.handle, int(x_lo), nbins, out))
        return np.array(out[:], dtype=np.int64)

    def halo_capacity(self):
        return int(self.lib.sph_capacity(self.sim.handle))

    def sync(self):
        self.sim.sync()

    # -- peer-memory path: no transport library in the step at all -----------------------------------
    def p2p_create(self, halo_rows, migrant_rows) -> bytes:
        buf = (C.c_ubyte * 64)()
        self._ck(self.lib.sph_slab_p2p_create(self.sim.handle, halo_rows, migrant_rows, buf))
        return bytes(buf)

    def p2p_connect(self, side, handle: bytes):
        buf = (C.c_ubyte * 64)(*handle)
        self._ck(self.lib.sph_slab_p2p_connect(self.sim.handle, side, buf))

    def p2p_step(self, lo, hi, lo_prev, hi_next, dt):
        h, L = self.sim.handle, self.lib
        self._ck(L.sph_slab_p2p_begin(h, int(lo), int(hi), int(lo_prev), int(hi_next)))
        self._ck(L.sph_slab_p2p_arrivals(h))
        self._ck(L.sph_slab_p2p_halo(h, int(lo), int(hi)))
        self._ck(L.sph_slab_p2p_ghosts(h))
        self._ck(L.sph_slab_step_density(h))
        self._ck(L.sph_slab_p2p_density(h))
        self._ck(L.sph_slab_step_forces(h, C.c_float(dt)))

    # -- sync-free path: fixed-size messages, tensors may be None (no neighbour on that side) --------
    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def fast_begin(self, lo, hi, lo_prev, hi_next, cap, send_l, send_r):
        self._ck(self.lib.sph_slab_fast_begin(self.sim.handle, int(lo), int(hi), int(lo_prev), int(hi_next), cap,
                                              self._p(send_l), self._p(send_r)))

    def fast_arrivals(self, recv_l, recv_r, cap):
        self._ck(self.lib.sph_slab_fast_arrivals(self.sim.handle, self._p(recv_l), self._p(recv_r), cap))

    def fast_halo(self, lo, hi, cap, send_l, send_r):
        self._ck(self.lib.sph_slab_fast_halo(self.sim.handle, int(lo), int(hi), cap, self._p(send_l), self._p(send_r)))

    def fast_ghosts(self, recv_l, recv_r, cap):
        self._ck(self.lib.sph_slab_fast_ghosts(self.sim.handle, self._p(recv_l), self._p(recv_r), cap))

    def fast_pack_density(self, cap, send_l, send_r):
        self._ck(self.lib.sph_slab_fast_pack_density(self.sim.handle, cap, self._p(send_l), self._p(send_r)))

    def fast_set_ghost_density(self, recv_l, recv_r, cap):
        self._ck(self.lib.sph_slab_fast_set_ghost_density(self.sim.handle, self._p(recv_l), self._p(recv_r), cap))


# ------------------------------------------------------------------------------------------------
# cuts
# ------------------------------------------------------------------------------------------------
def choose_cuts(hist: np.ndarray, x_lo: int, world: int) -> list:
    """Balanced cuts from a global histogram of cell.x (bins x_lo, x_lo+1, ...): slab k starts at
    the first cell where the cumulative count reaches k/world of the total. Every slab gets at
    least one x-cell of the occupied range so that halos always come from the adjacent rank."""
    hist = np.asarray(hist, dtype=np.int64)
    occupied = np.nonzero(hist)[0]
    cuts = [INT_MIN]
    if world > 1:
        if occupied.size == 0:
            first, last = 0, len(hist) - 1
        else:
            first, last = int(occupied[0]), int(occupied[-1])
        if last - first + 1 < world:
            raise ValueError(f"the particles span {last - first + 1} x-cells: cannot cut {world} slabs")
        cum = np.cumsum(hist)
        total = int(cum[-1])
        prev = first
        for k in range(1, world):
            target = total * k / world
            c = int(np.searchsorted(cum, target, side="left")) + 1  # first cell of slab k
            c = max(c, prev + 1)                   # at least one cell in slab k-1
            c = min(c, last - (world - 1 - k))     # leave one cell for each later slab
            cuts.append(x_lo + c)
            prev = c
    cuts.append(INT_MAX)
    return cuts


def owner_of(cuts, cell_x: np.ndarray) -> np.ndarray:
    inner = np.asarray(cuts[1:-1], dtype=np.int64)
    return np.searchsorted(inner, np.asarray(cell_x, dtype=np.int64), side="right")


# ------------------------------------------------------------------------------------------------
# driver
# ------------------------------------------------------------------------------------------------
class SlabDriver:
    """Per-rank step orchestration; identical code drives the CUDA engine (product) and the
    oracle-backed CPU engine of the tests."""

    def __init__(self, engine, rank: int, world: int, x_lo: int, nbins: int, group=None):
        self.e = engine
        self.rank, self.world = rank, world
        self.x_lo, self.nbins = int(x_lo), int(nbins)
        self.group = group
        self.backend = dist.get_backend(group) if world > 1 else "none"
        self.cuts = [INT_MIN] + [INT_MAX] * world if world == 1 else None
        self.stats = {"migrated_rows": 0, "halo_rows": 0, "steps": 0}
        # SLAB_PROFILE=1: synchronise after every phase and accumulate wall-clock per phase
        self.profile = os.environ.get("SLAB_PROFILE") == "1"
        self.phase_s = {}
        self._t = None

    # -- communication helpers ---------------------------------------------------------------
    def _comm_device(self, t):
        return t.cpu() if self.backend == "gloo" and t.is_cuda else t

    def _alltoall_counts(self, counts):
        send = torch.tensor(counts, dtype=torch.int64)
        if self.backend == "nccl":
            send = send.to(self.e.device)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        return [int(v) for v in recv.cpu().tolist()]

    def _alltoallv(self, send, send_counts, width):
        """send: [sum(send_counts), width] float32 rows grouped by destination rank."""
        recv_counts = self._alltoall_counts(send_counts)
        total = sum(recv_counts)
        src = self._comm_device(send.reshape(-1))
        out = torch.empty(total * width, dtype=torch.float32, device=src.device)
        dist.all_to_all_single(out, src, [c * width for c in recv_counts], [c * width for c in send_counts],
                               group=self.group)
        if out.device != send.device:
            out = out.to(send.device)
        return out.reshape(total, width) if width > 1 else out, recv_counts

    def _in_stream(self):
        s = getattr(self.e, "stream", None)
        return torch.cuda.stream(s) if s is not None else _NullCtx()

    def _mark(self, name):
        if not self.profile:
            return
        import time
        if hasattr(self.e, "sync"):
            self.e.sync()
        now = time.perf_counter()
        if name is not None and self._t is not None:
            self.phase_s[name] = self.phase_s.get(name, 0.0) + (now - self._t)
        self._t = now

    # -- cuts --------------------------------------------------------------------------------
    def rebalance(self):
        """Recompute balanced cuts from the global cell.x histogram (all ranks get the same cuts)."""
        if self.world == 1:
            return self.cuts
        with self._in_stream():
            hist = torch.from_numpy(self.e.xcell_histogram(self.x_lo, self.nbins))
            if self.backend == "nccl":
                hist = hist.to(self.e.device)
            dist.all_reduce(hist, group=self.group)
            self.cuts = choose_cuts(hist.cpu().numpy(), self.x_lo, self.world)
        return self.cuts

    # -- one step ----------------------------------------------------------------------------
    def step(self, dt: float = 0.0):
        e, r, w = self.e, self.rank, self.world
        if self.cuts is None:
            self.rebalance()
        with self._in_stream():
            self._mark(None)
            if w > 1:
                # 1. migration (also drops last step's ghosts)
                counts = e.count(self.cuts)
                self._mark("count")
                counts[r] = 0
                offsets = np.concatenate([[0], np.cumsum(counts)[:-1]])
                sendbuf = e.pack(self.cuts, r, offsets, int(counts.sum()))
                self._mark("pack")
                arrivals, _ = self._alltoallv(sendbuf, counts.tolist(), ROW_FLOATS)
                e.append(arrivals, 0)
                self._mark("migrate_exchange")
                self.stats["migrated_rows"] += int(counts.sum())
                # 2. halo: first x-cell of the slab -> left neighbour, last x-cell -> right neighbour
                cap = e.halo_capacity()
                left = e.pack_halo(self.cuts[r], 0, cap) if r > 0 else e.empty_rows(0)
                right = e.pack_halo(self.cuts[r + 1] - 1, 1, cap) if r < w - 1 else e.empty_rows(0)
                self._mark("pack_halo")
                hcounts = [0] * w
                if r > 0:
                    hcounts[r - 1] = left.shape[0]
                if r < w - 1:
                    hcounts[r + 1] = right.shape[0]
                ghosts, gcounts = self._alltoallv(torch.cat([left, right]), hcounts, ROW_FLOATS)
                n_from_left = gcounts[r - 1] if r > 0 else 0
                e.append(ghosts[:n_from_left], 1)
                e.append(ghosts[n_from_left:], 2)
                self.stats["halo_rows"] += left.shape[0] + right.shape[0]
                self._mark("halo_exchange")
            else:
                # single rank: still go through pack so stale rows are handled uniformly
                counts = e.count(self.cuts)
                e.pack(self.cuts, 0, [0], 0)
                hcounts, gcounts, n_from_left = [0], [0], 0
            # 3. density over owned + ghosts
            e.step_density()
            self._mark("density")
            if w > 1:
                # 4. densities of the boundary particles follow their positions
                dl = e.pack_halo_density(0, hcounts[r - 1] if r > 0 else 0)
                dr = e.pack_halo_density(1, hcounts[r + 1] if r < w - 1 else 0)
                rho, _ = self._alltoallv(torch.cat([dl, dr]), hcounts, 1)
                e.set_ghost_density(0, rho[:n_from_left])
                e.set_ghost_density(1, rho[n_from_left:])
                self._mark("rho_exchange")
            # 5. forces + integration of the owned particles
            e.step_forces(dt)
            self._mark("forces")
        self.stats["steps"] += 1


    # -- sync-free step --------------------------------------------------------------------------
    def setup_fast(self, halo_rows: int, migrant_rows: int | None = None):
        """Allocate the fixed-size message buffers of the sync-free path. halo_rows must exceed the
        population of a boundary x-cell layer, migrant_rows the particles crossing a cut per step."""
        e, r, w = self.e, self.rank, self.world
        self.fast_H = int(halo_rows)
        self.fast_M = int(migrant_rows if migrant_rows is not None else max(4096, halo_rows // 8))
        has = (r > 0, r < w - 1)
        with self._in_stream():
            mk = lambda n: [e.empty_rows(n) if has[k] else None for k in range(2)]
            mkf = lambda n: [e.empty_floats(n) if has[k] else None for k in range(2)]
            self.f_mig_s, self.f_mig_r = mk(self.fast_M), mk(self.fast_M)
            self.f_halo_s, self.f_halo_r = mk(self.fast_H), mk(self.fast_H)
            self.f_rho_s, self.f_rho_r = mkf(self.fast_H), mkf(self.fast_H)

    def suggest_halo_rows(self, slack: float = 3.0) -> int:
        """A message capacity from the global cell.x histogram: `slack` times the fullest layer
        adjacent to any cut (call after rebalance(); collective)."""
        with self._in_stream():
            hist = torch.from_numpy(self.e.xcell_histogram(self.x_lo, self.nbins))
            if self.backend == "nccl":
                hist = hist.to(self.e.device)
            if self.world > 1:
                dist.all_reduce(hist, group=self.group)
            hist = hist.cpu().numpy()
        worst = 1
        for c in self.cuts[1:-1]:
            b = c - self.x_lo
            worst = max(worst, int(hist[max(b - 2, 0):b + 2].max()))
        return int(max(4096, slack * worst))

    # -- peer-memory step ------------------------------------------------------------------------
    def setup_p2p(self, halo_rows: int, migrant_rows: int | None = None):
        """Create the mailboxes and map the neighbours' (CUDA IPC). Collective; once per run."""
        r, w = self.rank, self.world
        mig = int(migrant_rows if migrant_rows is not None else max(4096, halo_rows // 8))
        mine = self.e.p2p_create(int(halo_rows), mig)
        t = torch.tensor(list(mine), dtype=torch.uint8)
        if self.backend == "nccl":
            t = t.to(self.e.device)
        allh = [torch.empty_like(t) for _ in range(w)]
        if w > 1:
            dist.all_gather(allh, t, group=self.group)
        else:
            allh = [t]
        if r > 0:
            self.e.p2p_connect(0, bytes(allh[r - 1].cpu().tolist()))
        if r < w - 1:
            self.e.p2p_connect(1, bytes(allh[r + 1].cpu().tolist()))
        if w > 1:
            dist.barrier(group=self.group)
        self.p2p_ready = True
        self.fast_H = int(halo_rows)

    def step_p2p(self, dt: float = 0.0):
        """One step whose exchanges are stores into the neighbours' mailboxes: no collective calls,
        no host synchronisation. All ranks must call it in lockstep; cuts must be unchanged since the
        last general step()."""
        r, w, c = self.rank, self.world, self.cuts
        self.e.p2p_step(c[r], c[r + 1], c[r - 1] if r > 0 else INT_MIN, c[r + 2] if r < w - 1 else INT_MAX, dt)
        self.stats["steps"] += 1

    def _p2p(self, send, recv):
        """Exchange fixed-size messages with the adjacent ranks (send/recv = [left, right])."""
        r, w = self.rank, self.world
        ops, staged = [], []
        for side, peer in ((0, r - 1), (1, r + 1)):
            if peer < 0 or peer >= w:
                continue
            s_t, r_t = send[side], recv[side]
            if self.backend == "gloo" and s_t.is_cuda:  # tests: two ranks on one GPU, host staging
                s_c, r_c = s_t.cpu(), torch.empty(r_t.shape, dtype=r_t.dtype)
                staged.append((r_t, r_c))
                s_t, r_t = s_c, r_c
            ops.append(dist.P2POp(dist.isend, s_t, peer, self.group))
            ops.append(dist.P2POp(dist.irecv, r_t, peer, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()  # NCCL: orders the library's stream after the transfer, no host wait
        for dst, src in staged:
            dst.copy_(src)

    def step_fast(self, dt: float = 0.0):
        """One step with no host synchronisation except picking up last step's row count. Valid
        while the cuts are unchanged since the last general step()."""
        e, r, w = self.e, self.rank, self.world
        c = self.cuts
        lo, hi = c[r], c[r + 1]
        lo_prev = c[r - 1] if r > 0 else INT_MIN
        hi_next = c[r + 2] if r < w - 1 else INT_MAX
        with self._in_stream():
            self._mark(None)
            e.fast_begin(lo, hi, lo_prev, hi_next, self.fast_M, self.f_mig_s[0], self.f_mig_s[1])
            self._mark("f_begin")
            self._p2p(self.f_mig_s, self.f_mig_r)
            self._mark("f_migrate_exchange")
            e.fast_arrivals(self.f_mig_r[0], self.f_mig_r[1], self.fast_M)
            e.fast_halo(lo, hi, self.fast_H, self.f_halo_s[0], self.f_halo_s[1])
            self._mark("f_halo_pack")
            self._p2p(self.f_halo_s, self.f_halo_r)
            self._mark("f_halo_exchange")
            e.fast_ghosts(self.f_halo_r[0], self.f_halo_r[1], self.fast_H)
            e.step_density()
            self._mark("f_density")
            e.fast_pack_density(self.fast_H, self.f_rho_s[0], self.f_rho_s[1])
            self._p2p(self.f_rho_s, self.f_rho_r)
            e.fast_set_ghost_density(self.f_rho_r[0], self.f_rho_r[1], self.fast_H)
            self._mark("f_rho_exchange")
            e.step_forces(dt)
            self._mark("f_forces")
        self.stats["steps"] += 1


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


# ------------------------------------------------------------------------------------------------
# convenience: build a driver for the CUDA engine
# ------------------------------------------------------------------------------------------------
def x_cell_range(settings):
    """Histogram range that covers the walled part of the domain with slack (ends are clamped)."""
    half = int(np.ceil(settings.box_half_width / settings.h)) + 4
    return -half, 2 * half + 1


def make_gpu_driver(settings, capacity, device_index, rank, world, group=None):
    import importlib
    S = importlib.import_module("sph-fluid-simulator_b200")
    sim = S.Sim(settings, capacity=capacity, device=device_index)
    x_lo, nbins = x_cell_range(settings)
    return SlabDriver(GpuEngine(sim, device_index), rank, world, x_lo, nbins, group), sim


def gather_owned(sim, fields=("pos", "vel", "density", "force", "hash")):
    """Owned rows of this rank (device order) with their ids; ghost rows are filtered out."""
    import importlib
    S = importlib.import_module("sph-fluid-simulator_b200")
    d = sim.download(S.ORDER_DEVICE, fields=tuple(fields) + ("id",))
    keep = (d["id"] & 0x80000000) == 0
    return {k: v[keep] for k, v in d.items()}


# ------------------------------------------------------------------------------------------------
# bench: weak scaling (config 3), called from bench.py
# ------------------------------------------------------------------------------------------------
def bench_weak_scaling(args, scene_fn, METRIC, UNIT, ClockSampler, peaks):
    import importlib
    import json
    S = importlib.import_module("sph-fluid-simulator_b200")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scene = scene_fn(world)
    s = S.scaled_settings(scene["h"])
    nx, ny, nz = scene["dims"]

    # Weak-scaling base: the same per-GPU workload (one 61 x 256 x 512 block) on ONE GPU, measured in
    # this job on rank 0 with the same settle / warm-up / step counts, so that the scaling series has
    # a like-for-like N = 1 point (bench.py --gpus 1 runs config 1, the 1 M dam break, instead).
    base = None
    if world > 1 and not getattr(args, "no_weak_base", False):
        if rank == 0:
            b = scene_fn(1)
            bx, by, bz = b["dims"]
            bpos, bvel, bids = S.scene_block_slice(bx, by, bz, b["sep"], b["origin"], b["h"], b["seed"], 0, bx)
            bdrv, bsim = make_gpu_driver(s, int(bpos.shape[0] * 1.25) + (1 << 20), local, 0, 1)
            bsim.upload(bpos, bvel, bids)
            for _ in range(args.settle + max(args.warmup, 3)):
                bdrv.step(s.dt)
            bsim.sync()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(bdrv.e.stream):
                t0.record(bdrv.e.stream)
            for _ in range(args.steps):
                bdrv.step(s.dt)
            with torch.cuda.stream(bdrv.e.stream):
                t1.record(bdrv.e.stream)
            bsim.sync()
            bms = t0.elapsed_time(t1) / args.steps
            base = {"n_gpus": 1, "particles": int(bpos.shape[0]), "ms_per_step": bms,
                    "value": bpos.shape[0] / (bms * 1e-3), "workload": b["name"]}
            bsim.close()
            del bpos, bvel, bids, bdrv, bsim
            torch.cuda.empty_cache()
        dist.barrier()

    per = nx // world
    i0, i1 = rank * per, (rank + 1) * per if rank < world - 1 else nx
    pos, vel, ids = S.scene_block_slice(nx, ny, nz, scene["sep"], scene["origin"], scene["h"], scene["seed"], i0, i1)
    n_local, n_total = pos.shape[0], nx * ny * nz
    driver, sim = make_gpu_driver(s, int(n_local * 1.5) + (1 << 20), local, rank, world)
    sim.upload(pos, vel, ids)
    del pos, vel, ids
    driver.rebalance()

    transport = os.environ.get("SPH_SLAB_TRANSPORT", "p2p")  # p2p (peer-memory mailboxes) | nccl | general

    def run(steps, rebalance_every=100):
        # General (synchronous, all-to-all) step right after every change of cuts; in between, the
        # sync-free step: stores into the neighbours' mailboxes (p2p) or fixed-size NCCL messages.
        for k in range(steps):
            if k % rebalance_every == 0 or transport == "general":
                if world > 1 and k % rebalance_every == 0:
                    driver.rebalance()
                driver.step(s.dt)
                if transport == "p2p" and not getattr(driver, "p2p_ready", False):
                    driver.setup_p2p(driver.suggest_halo_rows(slack=2.0))
                elif transport == "nccl":
                    driver.setup_fast(driver.suggest_halo_rows())
            elif transport == "p2p":
                driver.step_p2p(s.dt)
            else:
                driver.step_fast(s.dt)

    run(args.settle)
    if world > 1:
        driver.rebalance()
    run(max(args.warmup, 3))
    sim.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    # rank 0 samples its own GPU; one poller per job keeps NVML out of the other ranks' launch paths
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    launches0 = sim.launch_count
    stream = driver.e.stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        flush.zero_()
        e0.record(stream)
    run(args.steps)
    with torch.cuda.stream(stream):
        e1.record(stream)
    sim.sync()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
    st = sim.stats()
    owned = torch.tensor([int(st.count)], dtype=torch.int64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        owned_all = [torch.zeros_like(owned) for _ in range(world)]
        dist.all_gather(owned_all, owned)
        owned_list = [int(o.item()) for o in owned_all]
    else:
        owned_list = [int(owned.item())]
    clk = clocks.stop() if clocks else None

    # ---- end to end: every step uploads the rank's rows from pinned host memory, steps through the
    # slab driver (general path: an upload resets the resident state) and reads positions and
    # velocities back to the host.
    e2e = None
    if args.e2e_steps > 0:
        cap_rows = int(sim.lib.sph_capacity(sim.handle))
        hp = torch.empty((cap_rows, 3), dtype=torch.float32).pin_memory()
        hv = torch.empty((cap_rows, 3), dtype=torch.float32).pin_memory()
        hi = torch.empty(cap_rows, dtype=torch.int32).pin_memory()
        fp, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        pp, pv, pi_ = C.cast(hp.data_ptr(), fp), C.cast(hv.data_ptr(), fp), C.cast(hi.data_ptr(), u32p)
        cnt = C.c_uint64(0)

        def read_back():
            rc = sim.lib.sph_slab_download_owned(sim.handle, pp, pv, pi_, cap_rows, C.byref(cnt))
            if rc:
                raise RuntimeError(sim.lib.sph_last_error(sim.handle).decode())
            return int(cnt.value)

        m = read_back()
        k_e2e = max(2, min(args.e2e_steps, 5))
        t_e2e = 0.0
        iter_ms = []
        import time
        e2e_warm = 2  # untimed iterations: first-touch of the pinned buffers, lazy transport set-up
        for k in range(k_e2e + e2e_warm):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            rc = sim.lib.sph_upload(sim.handle, m, pp, pv, pi_)   # pinned host rows -> device
            if rc:
                raise RuntimeError(sim.lib.sph_last_error(sim.handle).decode())
            if driver.profile:
                sim.sync(); t1 = time.perf_counter()
            driver.step(s.dt)
            if driver.profile:
                sim.sync(); t2 = time.perf_counter()
            m = read_back()                                       # owned rows -> pinned host
            if driver.profile:
                print(f"[rank {rank}] e2e iteration {k}: upload {1e3 * (t1 - t0):.2f} ms, step {1e3 * (t2 - t1):.2f} ms, "
                      f"download {1e3 * (time.perf_counter() - t2):.2f} ms", file=sys.stderr, flush=True)
            if world > 1:
                dist.barrier()
            iter_ms.append(1e3 * (time.perf_counter() - t0))
            if k >= e2e_warm:
                t_e2e += time.perf_counter() - t0
        tt = torch.tensor([t_e2e], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": n_total * k_e2e / float(tt.item()), "unit": UNIT, "steps": k_e2e,
               "h2d_bytes_per_step": 28 * n_total, "d2h_bytes_per_step": 28 * n_total,
               "iteration_ms_rank0": [round(v, 2) for v in iter_ms],
               "call": "per rank: sph_upload(pos, vel, id in pinned host memory) -> slab step (general path) -> "
                       "sph_slab_download_owned(pos, vel, id into pinned host memory)"}
    if driver.profile:
        print(f"[rank {rank}] phase ms/step:", {k: round(1e3 * v / max(driver.stats['steps'], 1), 3) for k, v in driver.phase_s.items()},
              file=sys.stderr, flush=True)
    total_s = float(ms.item()) * 1e-3
    value = n_total * args.steps / total_s
    if rank == 0:
        peak, peak_src = peaks()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
            "scaling": scene.get("scaling", "weak"), "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": scene["name"], "particles": n_total, "particles_per_gpu": owned_list,
                       "h": scene["h"], "dt": s.dt, "lattice": [nx, ny, nz], "settle_steps": args.settle,
                       "l2": f"state per GPU ({n_total // world} particles, ~{292 * (n_total // world) / 1e9:.2f} GB touched per step) is larger than the 126 MB L2; "
                             "L2 flushed once before the timed region",
                       "decomposition": "x slabs; per step: migrants, 1-cell ghost halo and halo densities go to the adjacent "
                                        "ranks with device-resident counts and no host sync; transport = " + transport +
                                        " (p2p: pack kernels store into the neighbour's CUDA-IPC mailbox over NVLink; "
                                        "nccl: fixed-size send/recv); general all-to-all step + rebalance every 100 steps",
                       "halo_message_rows": getattr(driver, "fast_H", None),
                       "single_gpu_same_workload": base,
                       "rank0_mean_density": st.mean_density, "rank0_grid_dim": list(st.grid_dim),
                       "migrated_rows_rank0": driver.stats["migrated_rows"], "halo_rows_rank0": driver.stats["halo_rows"],
                       "phase_ms_per_step_rank0": {k: 1e3 * v / max(driver.stats["steps"], 1) for k, v in driver.phase_s.items()}},
            "clocks": clk,
            "e2e": e2e,
            "gpu_launches": int(sim.launch_count - launches0),
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": 292.0 * value / world / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": 292.0 * value / world / 1e9 / peak, "traffic": None,
                         "peak_source": peak_src, "note": "per-GPU algorithmic 292 B/particle-step over the step time"},
            "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
