"""TEST-ONLY engine for sph-fluid-simulator_b200/slab.py: the slab primitives on numpy arrays with
the ORACLE doing the physics. It lets the multi-rank driver logic (cuts, migration, halo and
density-halo exchange) run under gloo on a CPU-only box. The product engine is slab.GpuEngine."""
import ctypes as C

import numpy as np
import torch

GHOST = np.uint32(0x80000000)
ID_MASK = np.uint32(0x7FFFFFFF)


class CpuOracleEngine:
    device = torch.device("cpu")

    def __init__(self, oracle, oset, capacity=1 << 22):
        self.O, self.s = oracle, oset
        self.pos = np.zeros((0, 3), np.float32)
        self.vel = np.zeros((0, 3), np.float32)
        self.idw = np.zeros(0, np.uint32)  # id | ghost bit
        self.capacity = capacity
        self.halo_rows = [np.zeros(0, np.int64), np.zeros(0, np.int64)]
        self.ghost_batch = [(0, 0), (0, 0)]
        self.rho = self.force = self.hash = None

    # -- helpers ---------------------------------------------------------------------------
    def upload(self, pos, vel, ids):
        self.pos, self.vel = np.array(pos, np.float32), np.array(vel, np.float32)
        self.idw = np.array(ids, np.uint32)

    def cell_x(self):
        return np.array([self.O.get_cell(p, self.s.h)[0] for p in self.pos], np.int64) if len(self.pos) else np.zeros(0, np.int64)

    def _rows(self, sel):
        out = np.zeros((len(sel), 8), np.float32)
        out[:, 0:3] = self.pos[sel]
        out[:, 3] = (self.idw[sel] & ID_MASK).view(np.float32)
        out[:, 4:7] = self.vel[sel]
        return torch.from_numpy(out)

    def empty_rows(self, n):
        return torch.empty((int(n), 8), dtype=torch.float32)

    def empty_floats(self, n):
        return torch.empty(int(n), dtype=torch.float32)

    @property
    def owned(self):
        return int(((self.idw & GHOST) == 0).sum())

    def halo_capacity(self):
        return self.capacity

    # -- primitives (same contract as slab.GpuEngine) ------------------------------------------
    def _owner(self, cuts):
        inner = np.asarray(cuts[1:-1], np.int64)
        return np.searchsorted(inner, self.cell_x(), side="right")

    def count(self, cuts):
        live = (self.idw & GHOST) == 0
        return np.bincount(self._owner(cuts)[live], minlength=len(cuts) - 1).astype(np.int64)

    def pack(self, cuts, rank, offsets, total):
        live = (self.idw & GHOST) == 0
        owner = self._owner(cuts)
        order = [np.nonzero(live & (owner == r))[0] for r in range(len(cuts) - 1) if r != rank]
        sel = np.concatenate(order) if order else np.zeros(0, np.int64)
        buf = self._rows(sel)
        keep = live & (owner == rank)
        self.pos, self.vel, self.idw = self.pos[keep], self.vel[keep], self.idw[keep]
        assert buf.shape[0] == total
        return buf

    def append(self, rows, kind):
        rows = rows.numpy()
        n0 = len(self.pos)
        ids = rows[:, 3].copy().view(np.uint32) & ID_MASK
        if kind:
            ids = ids | GHOST
            self.ghost_batch[kind - 1] = (n0, rows.shape[0])
        self.pos = np.concatenate([self.pos, rows[:, 0:3]]).astype(np.float32)
        self.vel = np.concatenate([self.vel, rows[:, 4:7]]).astype(np.float32)
        self.idw = np.concatenate([self.idw, ids]).astype(np.uint32)

    def pack_halo(self, cell_x, side, capacity):
        sel = np.nonzero(((self.idw & GHOST) == 0) & (self.cell_x() == cell_x))[0]
        self.halo_rows[side] = sel
        return self._rows(sel)

    def step_density(self):
        O, s = self.O, self.s
        h16 = O.hashes(self.pos, s.h)
        order = O.sort_order(h16)
        self._order = order
        sp, sh = np.ascontiguousarray(self.pos[order]), np.ascontiguousarray(h16[order])
        self._table = O.neighbor_table(sh)
        n = len(sp)
        dens, pres = np.empty(n, np.float32), np.empty(n, np.float32)
        fp, u16p, u32p = (C.POINTER(t) for t in (C.c_float, C.c_uint16, C.c_uint32))
        O.lib.oracle_density_pressure(n, sp.ctypes.data_as(fp), sh.ctypes.data_as(u16p),
                                      self._table.ctypes.data_as(u32p), C.byref(s), dens.ctypes.data_as(fp),
                                      pres.ctypes.data_as(fp))
        self.rho = np.empty(n, np.float32)
        self.rho[order] = dens  # back to row order
        self.hash = h16

    def pack_halo_density(self, side, n):
        assert n == len(self.halo_rows[side])
        return torch.from_numpy(self.rho[self.halo_rows[side]].copy())

    def set_ghost_density(self, side, values):
        first, n = self.ghost_batch[side]
        assert values.shape[0] == n
        self.rho[first:first + n] = values.numpy()

    def step_forces(self, dt):
        O, s = self.O, self.s
        order = self._order
        n = len(order)
        sp, sv = np.ascontiguousarray(self.pos[order]), np.ascontiguousarray(self.vel[order])
        sh = np.ascontiguousarray(self.hash[order])
        dens = np.ascontiguousarray(self.rho[order])
        pres = (np.float32(s.gasConstant) * (dens - np.float32(s.restDensity))).astype(np.float32)
        force = np.empty((n, 3), np.float32)
        fp, u16p, u32p = (C.POINTER(t) for t in (C.c_float, C.c_uint16, C.c_uint32))
        O.lib.oracle_forces(n, sp.ctypes.data_as(fp), sv.ctypes.data_as(fp), dens.ctypes.data_as(fp),
                            pres.ctypes.data_as(fp), sh.ctypes.data_as(u16p), self._table.ctypes.data_as(u32p),
                            C.byref(s), force.ctypes.data_as(fp))
        O.lib.oracle_integrate(n, sp.ctypes.data_as(fp), sv.ctypes.data_as(fp), force.ctypes.data_as(fp),
                               dens.ctypes.data_as(fp), C.byref(s), C.c_float(dt), None)
        owned = (self.idw[order] & GHOST) == 0
        rows = order[owned]
        self.pos[rows], self.vel[rows] = sp[owned], sv[owned]  # ghosts are not integrated
        self.force = np.zeros((n, 3), np.float32)
        self.force[order] = force

    # -- sync-free path (same contract as slab.GpuEngine.fast_*): fixed-size messages whose unused
    #    rows carry the dropped-row pattern 0xFFFFFFFF in every word --------------------------------
    DROP = np.uint32(0xFFFFFFFF)

    def _fixed_rows(self, sel, cap):
        assert len(sel) <= cap, "message overflow"
        out = np.full((cap, 8), np.float32(np.nan))
        out = out.view(np.uint32)
        out[:] = self.DROP
        if len(sel):
            out[:len(sel)] = self._rows(sel).numpy().view(np.uint32)
        return torch.from_numpy(out.view(np.float32))

    def fast_begin(self, lo, hi, lo_prev, hi_next, cap, send_l, send_r):
        live = (self.idw & GHOST) == 0
        cx = self.cell_x()
        go_l = live & (cx < lo) if send_l is not None else np.zeros(len(cx), bool)
        go_r = live & (cx >= hi) if send_r is not None else np.zeros(len(cx), bool)
        assert not (go_l & (cx < lo_prev)).any() and not (go_r & (cx >= hi_next)).any(), "non-adjacent migrant"
        if send_l is not None:
            send_l.copy_(self._fixed_rows(np.nonzero(go_l)[0], cap))
        if send_r is not None:
            send_r.copy_(self._fixed_rows(np.nonzero(go_r)[0], cap))
        keep = live & ~go_l & ~go_r
        self.pos, self.vel, self.idw = self.pos[keep], self.vel[keep], self.idw[keep]

    def _append_fixed(self, msg, ghost_side):
        rows = msg.numpy().view(np.uint32)
        valid = rows[:, 3] != self.DROP
        self.append(torch.from_numpy(rows[valid].view(np.float32).copy()), 0 if ghost_side is None else ghost_side + 1)

    def fast_arrivals(self, recv_l, recv_r, cap):
        for m in (recv_l, recv_r):
            if m is not None:
                self._append_fixed(m, None)

    def fast_halo(self, lo, hi, cap, send_l, send_r):
        live = (self.idw & GHOST) == 0
        cx = self.cell_x()
        for side, (buf, cell) in enumerate(((send_l, lo), (send_r, hi - 1))):
            if buf is None:
                self.halo_rows[side] = np.zeros(0, np.int64)
                continue
            sel = np.nonzero(live & (cx == cell))[0]
            self.halo_rows[side] = sel
            buf.copy_(self._fixed_rows(sel, cap))

    def fast_ghosts(self, recv_l, recv_r, cap):
        self.ghost_batch = [(len(self.pos), 0), (len(self.pos), 0)]
        for side, m in enumerate((recv_l, recv_r)):
            if m is not None:
                self._append_fixed(m, side)

    def fast_pack_density(self, cap, send_l, send_r):
        for side, buf in enumerate((send_l, send_r)):
            if buf is None:
                continue
            out = np.full(cap, self.DROP, np.uint32)
            sel = self.halo_rows[side]
            out[:len(sel)] = self.rho[sel].view(np.uint32)
            buf.copy_(torch.from_numpy(out.view(np.float32)))

    def fast_set_ghost_density(self, recv_l, recv_r, cap):
        for side, m in enumerate((recv_l, recv_r)):
            if m is None:
                continue
            first, n = self.ghost_batch[side]
            vals = m.numpy()
            assert (vals[:n].view(np.uint32) != self.DROP).all() and (vals[n:].view(np.uint32) == self.DROP).all()
            self.rho[first:first + n] = vals[:n]

    def xcell_histogram(self, x_lo, nbins):
        live = (self.idw & GHOST) == 0
        b = np.clip(self.cell_x()[live] - x_lo, 0, nbins - 1)
        return np.bincount(b, minlength=nbins).astype(np.int64)

    def owned_state(self):
        live = (self.idw & GHOST) == 0
        return dict(id=self.idw[live], pos=self.pos[live], vel=self.vel[live], density=self.rho[live],
                    force=self.force[live], hash=self.hash[live])
