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
        self.phase_events = None
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

    PHASES = ("rows_out", "rows_in", "grid+density", "halo_density", "forces")

    def p2p_step(self, lo, hi, lo_prev, hi_next, dt, migrant_rows=0):
        """One peer-mailbox step: 13 kernels, no host synchronisation. With self.phase_events set (a list),
        CUDA events are recorded on the library's stream between the phases (no synchronisation either)."""
        h, L = self.sim.handle, self.lib
        ev = None
        if self.phase_events is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(self.PHASES) + 1)]
            self.phase_events.append(ev)
            ev[0].record(self.stream)
        self._ck(L.sph_slab_p2p_begin(h, int(lo), int(hi), int(lo_prev), int(hi_next), int(migrant_rows)))
        if ev: ev[1].record(self.stream)
        self._ck(L.sph_slab_p2p_arrivals(h))
        if ev: ev[2].record(self.stream)
        self._ck(L.sph_slab_step_density(h))
        if ev: ev[3].record(self.stream)
        self._ck(L.sph_slab_p2p_density(h))
        if ev: ev[4].record(self.stream)
        self._ck(L.sph_slab_step_forces(h, C.c_float(dt)))
        if ev: ev[5].record(self.stream)

    def phase_ms(self):
        """Mean milliseconds per step of each phase over the steps recorded in self.phase_events (synchronises)."""
        evs = self.phase_events or []
        if not evs:
            return {}
        self.sim.sync()
        acc = [0.0] * len(self.PHASES)
        for ev in evs:
            for k in range(len(self.PHASES)):
                acc[k] += ev[k].elapsed_time(ev[k + 1])
        return {name: acc[k] / len(evs) for k, name in enumerate(self.PHASES)}

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
        self.last_imbalance = 0.0
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

    def global_histogram(self):
        """All-reduced histogram of cell.x over owned rows (host array; synchronises; collective)."""
        with self._in_stream():
            hist = torch.from_numpy(self.e.xcell_histogram(self.x_lo, self.nbins))
            if self.world > 1:
                if self.backend == "nccl":
                    hist = hist.to(self.e.device)
                dist.all_reduce(hist, group=self.group)
            return hist.cpu().numpy()

    def slab_counts(self, hist, cuts=None):
        """Particles per slab under `cuts` from a global histogram."""
        cuts = self.cuts if cuts is None else cuts
        edges = [0] + [int(min(max(c - self.x_lo, 0), len(hist))) for c in cuts[1:-1]] + [len(hist)]
        return np.array([int(hist[edges[k]:edges[k + 1]].sum()) for k in range(self.world)], dtype=np.int64)

    def rebalance_incremental(self, threshold: float = 0.05):
        """Move every interior cut by at most ONE cell towards its balanced position, and only when the
        fullest slab exceeds the mean by more than `threshold`. A one-cell move is something the sync-free
        steps handle by themselves (the rows of the layer that changed hands are ordinary migrants to the
        adjacent rank), so no general all-to-all step is needed afterwards. Returns True when cuts moved.
        Collective; one host synchronisation."""
        if self.world == 1:
            return False
        hist = self.global_histogram()
        counts = self.slab_counts(hist)
        mean = counts.sum() / self.world
        self.last_imbalance = float(counts.max() / mean - 1.0) if mean > 0 else 0.0
        if self.last_imbalance <= threshold:
            return False
        target = choose_cuts(hist, self.x_lo, self.world)
        new = list(self.cuts)
        for k in range(1, self.world):
            step = int(np.sign(target[k] - self.cuts[k]))
            new[k] = self.cuts[k] + step
        for k in range(1, self.world):  # keep at least one x-cell per slab
            lo = (new[k - 1] + 1) if k > 1 else INT_MIN
            new[k] = max(new[k], lo)
        moved = new != list(self.cuts)
        self.cuts = new
        return moved

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

    def step_p2p(self, dt: float = 0.0, migrant_rows: int = 0):
        """One step whose exchanges are stores into the neighbours' mailboxes: no collective calls,
        no host synchronisation. All ranks must call it in lockstep; since the last step the cuts may have
        moved by one cell each (pass migrant_rows = a whole layer then; 0 = the mailbox's capacity)."""
        r, w, c = self.rank, self.world, self.cuts
        self.e.p2p_step(c[r], c[r + 1], c[r - 1] if r > 0 else INT_MIN, c[r + 2] if r < w - 1 else INT_MAX, dt, migrant_rows)
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
# steady-state runner + in-job parity check + bench (configs 2 and 3), called from bench.py
# ------------------------------------------------------------------------------------------------
class SlabRunner:
    """Steps a SlabDriver in steady state. The step counter is persistent across run() calls: the one
    general (all-to-all) step that delivers the initial rows happens once, at step 0; afterwards every step
    is sync-free. Every `check_every` steps the global histogram is looked at (one host synchronisation);
    cuts only move when the fullest slab exceeds the mean by `threshold`, by one cell per cut, which the
    sync-free steps absorb as ordinary migration (transport "general" keeps the round-1 behaviour)."""

    def __init__(self, driver, dt, transport="p2p", check_every=200, threshold=0.05, halo_slack=1.5):
        self.d, self.dt, self.transport = driver, dt, transport
        self.check_every, self.threshold, self.halo_slack = check_every, threshold, halo_slack
        self.k = 0
        self.general_steps = self.cut_moves = self.balance_checks = 0
        self.layer_rows = 0

    def _first_step(self):
        d = self.d
        d.rebalance()
        d.step(self.dt)
        self.general_steps += 1
        if d.world > 1 or self.transport != "general":
            rows = d.suggest_halo_rows(slack=self.halo_slack)
            self.layer_rows = rows
            if self.transport == "p2p":
                d.setup_p2p(rows, migrant_rows=rows)  # a whole layer may change hands when a cut moves
            elif self.transport == "nccl":
                d.setup_fast(rows, migrant_rows=rows)

    def run(self, steps):
        d = self.d
        for _ in range(steps):
            if self.k == 0:
                self._first_step()
            elif self.transport == "general":
                if self.k % self.check_every == 0:
                    d.rebalance()
                d.step(self.dt)
                self.general_steps += 1
            else:
                moved = False
                if self.k % self.check_every == 0 and d.world > 1:
                    self.balance_checks += 1
                    moved = bool(d.rebalance_incremental(self.threshold))
                    self.cut_moves += moved
                if self.transport == "p2p":
                    # migrant messages: a whole layer right after a cut moved, an eighth of one otherwise (the
                    # appended message regions are rows every kernel of the step has to look at)
                    d.step_p2p(self.dt, 0 if moved else max(4096, self.layer_rows // 8))
                else:
                    d.step_fast(self.dt)
            self.k += 1


def slab_parity_check(S, local, rank, world, transport="p2p", steps=4):
    """Cross-GPU evidence inside the bench job: the 1 M dam break (config 1) settled on every rank's own GPU,
    then `steps` steps taken twice from that state — on this GPU alone, and slab-split over all ranks with
    the exchange going through the transport under test — and each rank compares the rows it owns at the
    end, bit for bit, with its single-GPU result. Returns a short verdict string (same on all ranks)."""
    h = 0.075
    sep = h * 16.0 / 15.0
    nx, ny, nz = 64, 80, 196
    s = S.scaled_settings(h)
    pos, vel = S.scene_block(nx, ny, nz, sep, ((h - 8.0) + sep, h * 5.0 / 3.0, -nz * sep / 2.0), h, 1024)
    n = pos.shape[0]
    one = S.Sim(s, capacity=n, device=local)
    one.upload(pos, vel)
    one.step(450)  # the lattice starts without neighbours: let it collapse into an interacting state first
    st = one.download(S.ORDER_ID, fields=("pos", "vel"))
    one.upload(st["pos"], st["vel"])
    one.step(steps)
    want = one.download(S.ORDER_ID, fields=("pos", "vel", "density"))
    interacting = float(one.stats().mean_density)
    one.close()

    x_lo, nbins = x_cell_range(s)
    cx = np.trunc(st["pos"][:, 0] / np.float32(h)).astype(np.int64)  # getCell: fp32 divide, truncation
    hist = np.bincount(np.clip(cx - x_lo, 0, nbins - 1), minlength=nbins)
    cuts = choose_cuts(hist, x_lo, world)
    mine = owner_of(cuts, cx) == rank
    ids = np.arange(n, dtype=np.uint32)
    drv, sim = make_gpu_driver(s, int(mine.sum() * 1.6) + (1 << 19), local, rank, world)
    sim.upload(st["pos"][mine], st["vel"][mine], ids[mine])
    drv.cuts = cuts
    drv.step(s.dt)  # general step: delivers nothing new here, builds the first halos
    if world > 1 and transport in ("p2p", "nccl"):
        rows = drv.suggest_halo_rows(slack=1.5)
        if transport == "p2p":
            drv.setup_p2p(rows)
        else:
            drv.setup_fast(rows)
    for _ in range(steps - 1):
        if world == 1 or transport == "general":
            drv.step(s.dt)
        elif transport == "p2p":
            drv.step_p2p(s.dt)
        else:
            drv.step_fast(s.dt)
    sim.sync()
    got = gather_owned(sim, fields=("pos", "vel", "density"))
    i = got["id"]
    bad = 0
    for k in ("pos", "vel", "density"):
        a, b = got[k].view(np.uint32), want[k][i].view(np.uint32)
        bad += int((a != b).reshape(len(i), -1).any(axis=1).sum())
    sim.close()
    t = torch.tensor([bad, len(i)], dtype=torch.int64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t)
    bad, owned = int(t[0].item()), int(t[1].item())
    if owned != n:
        return f"MISMATCH: {owned} rows owned in total, {n} expected"
    if interacting <= 9.3:
        return "INCONCLUSIVE: the check state is not interacting"
    if bad:
        return f"MISMATCH: {bad} rows differ from the single-GPU run"
    return (f"bit-identical ({n} particles, {steps} steps, {world} GPUs, transport {transport}: pos, vel, density of every "
            f"owned row equal to the single-GPU run on the same GPU; mean density {interacting:.3f})")


def bench_weak_scaling(args, scene_fn, METRIC, UNIT, ClockSampler, peaks, cpu_sample=None):
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
    transport = os.environ.get("SPH_SLAB_TRANSPORT", "p2p")  # p2p (peer-memory mailboxes) | nccl | general
    warmup = max(args.warmup, 3)

    # Cross-GPU correctness first: a bench line without it would only show that the step is fast.
    parity = None
    if world > 1 and not getattr(args, "no_slab_parity", False):
        parity = slab_parity_check(S, local, rank, world, transport)
        torch.cuda.empty_cache()
        dist.barrier()

    # Scaling base: the same per-GPU workload (for config 3, one 61 x 256 x 512 block) on ONE GPU, measured in
    # this job on rank 0 with the same settle / warm-up / step counts and the same library calls, so that the
    # series has a like-for-like N = 1 point (bench.py --gpus 1 reports config 1 and carries this same
    # measurement as config.weak_scaling_base).
    base = None
    if world > 1 and not getattr(args, "no_weak_base", False):
        if rank == 0:
            base = single_gpu_base(S, args, scene_fn, s, local, warmup)
        dist.barrier()

    per = nx // world
    i0, i1 = rank * per, (rank + 1) * per if rank < world - 1 else nx
    n_local, n_total = (i1 - i0) * ny * nz, nx * ny * nz
    driver, sim = make_gpu_driver(s, int(n_local * 1.5) + (1 << 20), local, rank, world)
    # this rank's x-range of the lattice, generated on its GPU (bit-identical to the host generator, whose
    # serial rand() loop over the whole 64 M lattice costs every rank seconds)
    sim.scene_block_device(nx, ny, nz, scene["sep"], scene["origin"], scene["seed"], i0, i1)
    runner = SlabRunner(driver, s.dt, transport)
    runner.run(args.settle)
    sim.sync()
    # rank 0 samples its own GPU; one poller per job keeps NVML out of the other ranks' launch paths. The
    # first sample is taken here, before the warm-up steps (see bench.ClockSampler: a query disturbs the
    # launches that follow it).
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    runner.run(warmup)
    sim.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    launches0 = sim.launch_count
    general0 = runner.general_steps
    stream = driver.e.stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        flush.zero_()
        e0.record(stream)
    runner.run(args.steps)
    with torch.cuda.stream(stream):
        e1.record(stream)
    sim.sync()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
    launches = int(sim.launch_count - launches0)
    general_timed = runner.general_steps - general0
    # The phase table comes from a second, untimed stretch of the same steps (as the N = 1 line takes its pass shares
    # from a separate loop): six event records per step inside the timed region are six more things in the stream.
    phase_ms = {}
    if transport == "p2p" and world > 1:
        driver.e.phase_events = []  # CUDA events between the phases of every step (no synchronisation)
        runner.run(min(args.steps, 20))
        sim.sync()
        phase_ms = driver.e.phase_ms()
        driver.e.phase_events = None
    st = sim.stats()
    owned = torch.tensor([int(st.count)], dtype=torch.int64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        owned_all = [torch.zeros_like(owned) for _ in range(world)]
        dist.all_gather(owned_all, owned)
        owned_list = [int(o.item()) for o in owned_all]
        ph = torch.tensor([phase_ms.get(k, 0.0) for k in GpuEngine.PHASES], dtype=torch.float64, device=f"cuda:{local}")
        ph_max = ph.clone()
        dist.all_reduce(ph_max, op=dist.ReduceOp.MAX)
    else:
        owned_list = [int(owned.item())]
        ph_max = None
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
        k_e2e = max(2, min(args.e2e_steps, 8))
        t_e2e = 0.0
        iter_ms = []
        import time
        e2e_warm = 3  # untimed iterations: first-touch of the pinned buffers, lazy transport set-up
        for k in range(k_e2e + e2e_warm):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            rc = sim.lib.sph_upload(sim.handle, m, pp, pv, pi_)   # pinned host rows -> device
            if rc:
                raise RuntimeError(sim.lib.sph_last_error(sim.handle).decode())
            driver.step(s.dt)
            m = read_back()                                       # owned rows -> pinned host
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
               "call": "per rank: sph_upload(pos, vel, id in pinned host memory) -> slab step (general path: an upload "
                       "resets the resident state, so the all-to-all step delivers and rebuilds the halos) -> "
                       "sph_slab_download_owned(pos, vel, id into pinned host memory)"}
    total_s = float(ms.item()) * 1e-3
    value = n_total * args.steps / total_s

    # Config 2 (16 M dam break, strong scaling) as a sub-record, when this line is the weak-scaling one.
    strong = None
    strong_scene = getattr(args, "strong_scene", None)
    if strong_scene is not None and scene.get("scaling", "weak") == "weak" and not getattr(args, "no_strong_subrecord", False):
        sim.close()  # free this rank's 8 M-particle state first (driver / runner are only read for their counters below)
        torch.cuda.empty_cache()
        strong = strong_scaling_subrecord(S, args, strong_scene, local, rank, world, transport, warmup)

    if rank == 0:
        peak, peak_src = peaks()
        if base:
            base["step_hbm_frac"] = base["step_achieved_gbs"] / peak
        cpu = cpu_sample(scene, s) if (cpu_sample and not args.no_cpu_baseline) else None
        ms_step = 1e3 * total_s / args.steps
        phases = None
        if ph_max is not None:
            phases = {"rank0": {k: round(phase_ms.get(k, 0.0), 4) for k in GpuEngine.PHASES},
                      "max_over_ranks": {k: round(float(ph_max[j].item()), 4) for j, k in enumerate(GpuEngine.PHASES)}}
            slow = max(phases["max_over_ranks"], key=phases["max_over_ranks"].get)
            phases["limiting_phase"] = slow
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": scene.get("scaling", "weak"), "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": scene["name"], "particles": n_total, "particles_per_gpu": owned_list,
                       "h": scene["h"], "dt": s.dt, "lattice": [nx, ny, nz], "settle_steps": args.settle,
                       "l2": f"state per GPU ({n_total // world} particles, ~{292 * (n_total // world) / 1e9:.2f} GB touched per step) is larger than the 126 MB L2; "
                             "L2 flushed once before the timed region",
                       "decomposition": "x slabs; per step: migrants, 1-cell ghost halo and halo densities go to the adjacent "
                                        "ranks with device-resident counts and no host sync; transport = " + transport +
                                        " (p2p: pack kernels store into the neighbour's CUDA-IPC mailbox over NVLink and publish "
                                        "count + flag themselves; nccl: fixed-size send/recv). One general all-to-all step at "
                                        "step 0 only; cuts are looked at every 200 steps and move by one cell when the fullest "
                                        "slab exceeds the mean by 5 %",
                       "general_steps_in_timed_region": general_timed,
                       "balance_checks": runner.balance_checks, "cut_moves": runner.cut_moves,
                       "imbalance_at_last_check": driver.last_imbalance,
                       "halo_message_rows": getattr(driver, "fast_H", None),
                       "single_gpu_same_workload": base,
                       "scaling_efficiency_vs_same_workload": (value / world / base["value"]) if base else None,
                       "slab_parity": parity,
                       "strong_scaling_16M": strong,
                       "rank0_mean_density": st.mean_density, "rank0_grid_dim": list(st.grid_dim),
                       "phase_ms_per_step": phases},
            "clocks": clk,
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": 292.0 * value / world / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": 292.0 * value / world / 1e9 / peak, "traffic": None,
                         "peak_source": peak_src, "note": "per-GPU algorithmic 292 B/particle-step over the step time"},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def strong_scaling_subrecord(S, args, scene, local, rank, world, transport, warmup):
    """Config 2 (the 16 M dam break, the same field whatever the number of GPUs) through the same slab runner,
    timed the same way, as a sub-record of the N > 1 lines: strong scaling, so that it reaches a driver box."""
    s = S.scaled_settings(scene["h"])
    nx, ny, nz = scene["dims"]
    per = nx // world
    i0, i1 = rank * per, (rank + 1) * per if rank < world - 1 else nx
    n_local, n_total = (i1 - i0) * ny * nz, nx * ny * nz
    driver, sim = make_gpu_driver(s, int(n_local * 1.6) + (1 << 20), local, rank, world)
    sim.scene_block_device(nx, ny, nz, scene["sep"], scene["origin"], scene["seed"], i0, i1)
    runner = SlabRunner(driver, s.dt, transport, check_every=100)
    runner.run(args.settle + warmup)
    sim.sync()
    if world > 1:
        dist.barrier()
    stream = driver.e.stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    general0 = runner.general_steps
    e0.record(stream)
    runner.run(args.steps)
    e1.record(stream)
    sim.sync()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
    owned = torch.tensor([int(sim.stats().count)], dtype=torch.int64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        owned_all = [torch.zeros_like(owned) for _ in range(world)]
        dist.all_gather(owned_all, owned)
        owned_list = [int(o.item()) for o in owned_all]
    else:
        owned_list = [int(owned.item())]
    sim.close()
    torch.cuda.empty_cache()
    t = float(ms.item()) * 1e-3
    return {"workload": scene["name"], "scaling": "strong", "particles": n_total, "particles_per_gpu": owned_list,
            "h": scene["h"], "dt": s.dt, "lattice": [nx, ny, nz], "settle_steps": args.settle, "steps": args.steps,
            "ms_per_step": 1e3 * t / args.steps, "value": n_total * args.steps / t,
            "general_steps_in_timed_region": runner.general_steps - general0, "cut_moves": runner.cut_moves,
            "imbalance_at_last_check": driver.last_imbalance}


def single_gpu_base(S, args, scene_fn, s, local, warmup):
    """The per-GPU workload of the scaling series on one GPU through the resident step (sph_step: CUDA-graph
    replay, as a single-GPU user would run it): settle, warm up, then `steps` steps between two events."""
    b = scene_fn(1)
    bx, by, bz = b["dims"]
    n = bx * by * bz
    sim = S.Sim(s, capacity=n, device=local)
    sim.scene_block_device(bx, by, bz, b["sep"], b["origin"], b["seed"])
    sim.step(args.settle + warmup)
    sim.sync()
    stream = torch.cuda.ExternalStream(sim.stream, device=torch.device("cuda", local))
    reps = []
    for _ in range(3):  # the timed region of the series, three times: the median is the base, all three are reported
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        sim.step(args.steps)
        t1.record(stream)
        sim.sync()
        reps.append(t0.elapsed_time(t1) / args.steps)
    ms = sorted(reps)[1]
    sim.enable_pass_timing(True)
    sim.step(min(args.steps, 20))
    passes = sim.pass_times()
    sim.enable_pass_timing(False)
    st = sim.stats()
    out = {"n_gpus": 1, "particles": n, "ms_per_step": ms, "ms_per_step_repeats": reps, "value": n / (ms * 1e-3), "workload": b["name"],
           "pass_ms": {k: v for k, v in passes.items() if k != "steps"}, "mean_density": st.mean_density,
           "step_achieved_gbs": 292.0 * n / (ms * 1e-3) / 1e9}
    sim.close()
    torch.cuda.empty_cache()
    return out
