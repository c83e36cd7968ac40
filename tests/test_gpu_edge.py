"""GPU: edge cases and size-independent properties of the CUDA path (through the C-ABI).

Edge cases the reference code has (it has no tests of its own): empty and tiny systems, particles
that start outside the walls, negative coordinates and the double-width cell 0, coincident particles
(NaN forces, exactly as the reference produces them), a grid that has to be clipped, capacity and
state errors. Properties at the full 1 M size of BASELINE.json config 1: determinism, independence
of the upload order, sortedness, id conservation, hash-table consistency.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_bit_equal, assert_fields_close, by_id

pytestmark = pytest.mark.gpu


def step_both(sph, oracle, s, pos, vel, steps=1):
    sim = sph.Sim(s, capacity=max(len(pos), 1))
    sim.upload(pos, vel)
    sim.step(steps)
    got = sim.download(sph.ORDER_ID)
    sim.close()
    os_ = oracle.settings(s.as_tuple7())
    p, v, ids = pos, vel, np.arange(len(pos), dtype=np.uint32)
    for _ in range(steps):
        o = oracle.step(os_, s.dt, p, v, ids)
        p, v, ids = o["pos"], o["vel"], o["id"]
    return got, by_id(o)


def test_empty_and_tiny_systems(sph, oracle):
    s = sph.default_settings()
    sim = sph.Sim(s, capacity=8)
    sim.upload(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    sim.step(3)
    assert sim.count == 0 and sim.download(sph.ORDER_DEVICE, fields=("pos",))["pos"].shape == (0, 3)
    sim.close()
    for n in (1, 2, 3):
        pos = np.array([[0.0, 1.0, 0.0], [0.05, 1.0, 0.0], [0.0, 1.1, 0.05]], np.float32)[:n]
        vel = np.zeros((n, 3), np.float32)
        got, want = step_both(sph, oracle, s, pos, vel, steps=2)
        assert np.array_equal(got["hash"], want["hash"])
        assert_fields_close(got, want, f"n={n}")


def test_particles_outside_walls_negative_coordinates_and_cell_zero(sph, oracle):
    """The shipped cube overflows the box for widths >= 60; x,z < 0 truncate toward zero so cell 0
    spans (-h, h); a floor-penetrating start is mirrored (src/sph.cpp:153-176)."""
    rng = np.random.default_rng(11)
    s = sph.default_settings()
    pos = np.concatenate([
        rng.uniform([-12, -1, -12], [12, 6, 12], (3000, 3)),       # beyond every wall and below the floor
        rng.uniform([-0.16, 0.1, -0.16], [0.16, 0.5, 0.16], (600, 3)),  # straddling cell 0 on x and z
    ]).astype(np.float32)
    vel = rng.normal(0, 1.0, pos.shape).astype(np.float32)
    got, want = step_both(sph, oracle, s, pos, vel, steps=1)
    assert np.array_equal(got["hash"], want["hash"])
    assert_fields_close(got, want, "outside walls")
    assert (got["pos"][:, 1] >= s.h - 1e-6).all(), "floor reflection"


def test_coincident_particles_give_nan_like_the_reference(sph, oracle):
    s = sph.default_settings()
    pos = np.array([[0.5, 1.0, 0.5], [0.5, 1.0, 0.5], [0.55, 1.0, 0.5], [3.0, 2.0, 3.0]], np.float32)
    vel = np.zeros_like(pos)
    got, want = step_both(sph, oracle, s, pos, vel)
    assert np.isnan(want["force"][:2]).all(), "normalize(0) in the reference (src/sph.cpp:111)"
    assert np.isnan(got["force"][:2]).all() and np.isnan(got["pos"][:2]).all()
    assert np.array_equal(np.isnan(got["force"]), np.isnan(want["force"]))
    ok = ~np.isnan(want["pos"]).any(axis=1)
    assert np.abs(got["pos"][ok] - want["pos"][ok]).max() < 1e-5
    assert np.array_equal(got["density"].view(np.uint32), want["density"].view(np.uint32)) or \
        np.abs(got["density"] - want["density"]).max() < 1e-4


def test_grid_clipping_keeps_results_exact():
    """With a tiny dense-grid allocation the box is clipped and outliers are clamped into the edge
    layer: candidates become a superset, results must not change. Runs in a subprocess because the
    allocation is fixed at sph_create from SPH_B200_MAX_CELLS."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import sph_b200 as S
from conftest import load_golden
g = load_golden("cube20_step200.npz")
s = S.default_settings()
pos, vel = g["pos0"].copy(), g["vel0"].copy()
pos[:40, 1] += np.linspace(50, 4000, 40).astype(np.float32)   # a few particles far up: y has no ceiling
sim = S.Sim(s, capacity=len(pos)); sim.upload(pos, vel); sim.step(2)
out = sim.download(S.ORDER_ID); st = sim.stats()
np.savez(sys.argv[1], clamped=st.clamped, cells=st.grid_cells, **out)
''' % (ROOT, os.path.join(ROOT, "tests"))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        outs = []
        for cells in ("67108864", "20000"):  # room for the whole 22 x 26 700 x 22 box / far too small
            env = dict(os.environ)
            env["SPH_B200_MAX_CELLS"] = cells
            f = os.path.join(d, f"o{cells}.npz")
            subprocess.run([sys.executable, "-c", code, f], check=True, env=env)
            outs.append(dict(np.load(f)))
    full, clipped = outs
    # 20 000 cells clip all three axes of the 39 x 26 669 x 39 box: thousands of particles are clamped
    assert int(full["clamped"]) == 0 and int(clipped["clamped"]) > 1000 and int(clipped["cells"]) <= 20000
    assert np.array_equal(clipped["hash"], full["hash"])
    # Same neighbour sets, but merged edge cells change the summation order: tolerance, not bits.
    assert (np.abs(clipped["density"] - full["density"]) / full["density"]).max() < 1e-6
    assert_fields_close(clipped, full, "clipped grid")


def test_box_saturated_on_two_axes_is_clamped_not_overflowed(sph):
    """Positions at +/-inf (or beyond 2^31 h) on two axes saturate getCell on both: the grid plan must clip
    the box instead of overflowing its cell count, the outliers are clamped into the edge layer
    and counted, and everybody else's results do not change."""
    rng = np.random.default_rng(3)
    s = sph.default_settings()
    pos = rng.uniform([-2, 0.2, -2], [2, 2.5, 2], (4000, 3)).astype(np.float32)
    vel = rng.normal(0, 0.3, pos.shape).astype(np.float32)
    ids = np.arange(len(pos), dtype=np.uint32)
    wild = pos.copy()
    wild[0] = [np.inf, np.inf, 1.0]
    wild[1] = [-3.0e38, 3.0e38, 0.5]
    wild[2] = [3.0e38, 1.0, -np.inf]
    outs = []
    for p, keep in ((wild, slice(None)), (pos, slice(3, None))):
        sim = sph.Sim(s, capacity=len(pos))
        sim.upload(p[keep], vel[keep], ids[keep])
        sim.step(2)
        st = sim.stats()
        d = sim.download(sph.ORDER_DEVICE)
        outs.append((st, {k: v[np.argsort(d["id"])] for k, v in d.items()}))
        sim.close()
    (st_w, w), (st_c, c) = outs
    assert st_w.nan_count >= 2 and st_c.nan_count == 0  # the +/-inf rows (3e38 is finite and gets mirrored by the walls)
    assert 0 < st_w.grid_cells <= 1 << 30 and min(st_w.grid_dim) >= 3 and st_w.clamped >= 3
    # the clipped grid clamps (nearly) everybody into merged edge cells: same neighbour sets, another summation
    # order — tolerance, not bits (as in test_grid_clipping_keeps_results_exact)
    assert np.array_equal(w["hash"][3:], c["hash"])
    assert_fields_close({k: v[3:] for k, v in w.items()}, c, "finite particles next to saturated outliers")


def test_ids_with_the_ghost_bit_are_refused(sph):
    """Bit 31 of an id is the library's ghost marker: such ids are refused at upload (ids must also be
    unique — include/sph_b200.h — which is the caller's contract)."""
    sim = sph.Sim(sph.default_settings(), capacity=8)
    pos = np.full((2, 3), 0.5, np.float32)
    with pytest.raises(sph.SphError) as e:
        sim.upload(pos, np.zeros_like(pos), ids=np.array([1, 0x80000005], np.uint32))
    assert e.value.code == 1  # SPH_ERR_INVALID
    sim.close()


@pytest.fixture(params=["64", "0"], ids=["tiles", "warp-per-row"])
def clump_cell(request):
    """SPH_B200_CLUMP_CELL for the handles the test creates: 64 (the default) — deferred rows of crowded cells are
    served an 8-row tile at a time; 0 — every deferred row gets a warp of its own (heavy tail)."""
    old = os.environ.get("SPH_B200_CLUMP_CELL")
    os.environ["SPH_B200_CLUMP_CELL"] = request.param
    yield request.param
    if old is None:
        del os.environ["SPH_B200_CLUMP_CELL"]
    else:
        os.environ["SPH_B200_CLUMP_CELL"] = old


def test_clump_takes_the_warp_cooperative_kernels(sph, oracle, clump_cell):
    """A clump (hundreds of neighbours per particle, what the reference's fluid collapses into after
    ~2000 steps) overflows the per-particle neighbour list and the per-run budget of the one-thread
    kernels: those particles are deferred to the heavy kernels (one warp per row, or — rows of crowded
    cells — a tile of rows on the candidates they share). Same neighbour multisets, same tolerances, and
    the same bits every time, for either class."""
    rng = np.random.default_rng(5)
    s = sph.default_settings()
    d = rng.normal(size=(1500, 3))
    d *= (0.25 * rng.uniform(0, 1, (1500, 1)) ** (1 / 3)) / np.linalg.norm(d, axis=1, keepdims=True)
    pos = np.concatenate([d + [1.0, 1.0, 1.0], rng.uniform([-3, 0.2, -3], [3, 3, 3], (2500, 3))]).astype(np.float32)
    vel = rng.normal(0, 0.5, pos.shape).astype(np.float32)
    os_ = oracle.settings(s.as_tuple7())
    want = by_id(oracle.step(os_, s.dt, pos, vel))
    runs = []
    for _ in range(2):
        sim = sph.Sim(s, capacity=len(pos))
        sim.upload(pos, vel)
        ids, counts, offs, lst = sim.neighbor_lists()
        sim.step(1)
        st = sim.stats()
        runs.append((sim.download(sph.ORDER_ID), counts[np.argsort(ids)]))
        sim.close()
    got, counts = runs[0]
    assert st.deferred_density >= 1400 and st.deferred_forces >= 1400 and st.nlist_rows < counts.max()
    order, ocounts, _, _, _ = oracle.neighbor_lists(os_, pos)
    assert np.array_equal(counts, ocounts[np.argsort(order)]) and counts.max() > 300
    assert np.array_equal(got["hash"], want["hash"])
    rel = np.abs(got["density"] - want["density"]) / want["density"]
    assert rel.max() <= 1e-5, f"density rel err {rel.max():.3e}"
    fn = np.linalg.norm(want["force"], axis=1)
    fe = np.linalg.norm(got["force"] - want["force"], axis=1) / np.maximum(fn, np.median(fn[:1500]))
    assert fe.max() <= 1e-3, f"force rel err {fe.max():.3e}"
    for k in ("pos", "vel", "force", "density"):
        assert_bit_equal(runs[1][0][k], got[k], f"second run: {k}")


_HASH_M = (73856093, 19349663, 83492791)


def _hash16(c):
    return ((c[0] * _HASH_M[0]) ^ (c[1] * _HASH_M[1]) ^ (c[2] * _HASH_M[2])) & 0xFFFF


def _bucket_hashes(c):
    """hash16 of the 27 buckets the reference walks for a particle in cell c (src/sph.cpp:40-44)."""
    return [_hash16((c[0] + x, c[1] + y, c[2] + z)) for x in (-1, 0, 1) for y in (-1, 0, 1) for z in (-1, 0, 1)]


def _multiplicity(buckets, cj):
    """How often the reference accepts a neighbour of cell cj: once per bucket with cj's hash;
    1 unless two of the 27 bucket hashes collide."""
    return buckets.count(_hash16(cj)) if len(set(buckets)) < 27 else 1


def _density_in_documented_order(pos, ids_by_cell, i, h, h2, mp, self_dens, multiplicities=True):
    """nine runs (x offset outer, z offset inner), each the cells y-1, y, y+1 of a column in
    ascending y, rows of a cell by ascending id. One-thread kernel ("seq"): dens = (float)((double)dens + mp * t^3)
    per accepted row, in that order. Clump rows ("tiled", tiled phase of the heavy kernel): four double-precision
    partial sums, part s taking the accepted rows whose rank in their run is s mod 4, combined as
    (p0 + p2) + (p1 + p3) and rounded to float once. One warp per row ("warp"): 32 double-precision partial sums by
    rank mod 32, combined by an xor butterfly, rounded once. Returns the three results, the neighbour count, the
    longest run and the own cell's size."""
    f = np.float32
    c = tuple(int(v) for v in np.trunc(pos[i] / f(h)).astype(np.int64))
    buckets = _bucket_hashes(c)
    dens, cnt, longest = f(0), 0, 0
    part = [np.float64(0)] * 4
    lanes = np.zeros(32, np.float64)  # one warp per row: lane (rank in run) % 32, then a butterfly over the lanes
    for ox in (-1, 0, 1):
        for oz in (-1, 0, 1):
            run = 0
            for oy in (-1, 0, 1):
                cj = (c[0] + ox, c[1] + oy, c[2] + oz)
                m = _multiplicity(buckets, cj) if multiplicities else 1
                for j in ids_by_cell.get(cj, ()):
                    rank = run
                    run += 1
                    if j == i:
                        continue
                    d = pos[j] - pos[i]
                    d2 = f(f(f(d[0] * d[0]) + f(d[1] * d[1])) + f(d[2] * d[2]))
                    if d2 < h2:
                        t = np.float64(f(h2 - d2))
                        term = np.float64(mp) * ((t * t) * t)
                        for _ in range(m):  # only rows without collisions take the one-thread path: m is 1 there
                            dens = f(np.float64(dens) + term)
                        part[rank % 4] = part[rank % 4] + np.float64(m) * term
                        lanes[rank % 32] = lanes[rank % 32] + np.float64(m) * term
                        cnt += m
            longest = max(longest, run)
    tiled = f(f((part[0] + part[2]) + (part[1] + part[3])) + self_dens)
    for o in (16, 8, 4, 2, 1):
        lanes = lanes + lanes[np.arange(32) ^ o]
    warp = f(f(lanes[0]) + self_dens)
    return {"seq": f(dens + self_dens), "tiled": tiled, "warp": warp}, cnt, longest, len(ids_by_cell[c])


def _force_in_documented_order(pos, vel, rho, ids_by_cell, i, K):
    """The force terms of src/sph.cpp:110-121 in float32, operation by operation, accumulated in the documented
    orders (see _density_in_documented_order): the one-thread kernel's running sum, the tiled phase's four
    partial sums by rank in the run combined as (p0 + p2) + (p1 + p3), and the 32 per-lane sums of a row that has a
    warp of its own, combined by the butterfly. rho: density of every particle (by id).
    K: dict of float32 constants h, h2, mass, gas, rest, visc_mass, spiky_grad, spiky_lap."""
    f = np.float32
    c = tuple(int(v) for v in np.trunc(pos[i] / K["h"]).astype(np.int64))
    buckets = _bucket_hashes(c)
    pres_i = f(K["gas"] * f(rho[i] - K["rest"]))
    seq = np.zeros(3, f)
    part = np.zeros((4, 3), f)
    lanes = np.zeros((32, 3), f)

    def add(acc, t):
        for a in range(3):
            acc[a] = f(acc[a] + t[a])

    with np.errstate(all="ignore"):
        for ox in (-1, 0, 1):
            for oz in (-1, 0, 1):
                run = 0
                for oy in (-1, 0, 1):
                    cj = (c[0] + ox, c[1] + oy, c[2] + oz)
                    m = _multiplicity(buckets, cj)
                    for j in ids_by_cell.get(cj, ()):
                        rank = run
                        run += 1
                        if j == i:
                            continue
                        d = pos[j] - pos[i]
                        d2 = f(f(f(d[0] * d[0]) + f(d[1] * d[1])) + f(d[2] * d[2]))
                        if not d2 < K["h2"]:
                            continue
                        dist = np.sqrt(d2, dtype=f)
                        inv = f(f(1.0) / dist)
                        psum = f(pres_i + f(K["gas"] * f(rho[j] - K["rest"])))
                        den = f(f(2.0) * rho[j])
                        hd = f(K["h"] - dist)
                        w2 = f(hd * hd)
                        P = np.zeros(3, f)
                        V = np.zeros(3, f)
                        for a in range(3):
                            n = f(d[a] * inv)
                            P[a] = f(f(f(f(f(f(-n) * K["mass"]) * psum) / den) * K["spiky_grad"]) * w2)
                            u = f(vel[j][a] - vel[i][a])
                            V[a] = f(f(f(K["visc_mass"] * f(u / rho[j])) * K["spiky_lap"]) * hd)
                        for acc in (seq, part[rank % 4], lanes[rank % 32]):
                            for _ in range(m):  # a neighbour accepted m times adds its two terms m times, back to back
                                add(acc, P)
                                add(acc, V)
        tiled = np.array([f(f(part[0][a] + part[2][a]) + f(part[1][a] + part[3][a])) for a in range(3)], f)
        for o in (16, 8, 4, 2, 1):
            lanes = (lanes + lanes[np.arange(32) ^ o]).astype(f)
    return {"seq": seq, "tiled": tiled, "warp": lanes[0].copy()}


def _integrate_as_the_reference_does(p, v, F, rho, s):
    """src/sph.cpp:138-179 in float32: a = F / rho + (0, g, 0); v += a dt; x += v dt; then the five wall tests in
    order, each on the already-updated value."""
    f = np.float32
    h, dt, box, el, off = f(s.h), f(s.dt), f(s.box_half_width), f(s.elasticity), f(s.wall_offset)
    p, v = p.astype(f).copy(), v.astype(f).copy()
    g3 = (f(0), f(s.g), f(0))
    for a in range(3):
        acc = f(f(F[a] / rho) + g3[a])
        v[a] = f(v[a] + f(acc * dt))
        p[a] = f(p[a] + f(v[a] * dt))
    hmb, bmh = f(h - box), f(-h + box)
    if p[1] < h:
        p[1] = f(f(-p[1] + f(f(2) * h)) + off)
        v[1] = f(-v[1] * el)
    for a in (0, 2):
        if p[a] < hmb:
            p[a] = f(f(-p[a] + f(f(2) * hmb)) + off)
            v[a] = f(-v[a] * el)
        if p[a] > bmh:
            p[a] = f(f(-p[a] + f(f(2) * f(-hmb))) - off)
            v[a] = f(-v[a] * el)
    return p, v


def test_sums_are_taken_in_the_documented_order(sph):
    """Bit-exact check of the summation order itself, against a numpy restatement: for rows of the one-thread
    kernel, for clump rows (tiled phase of the heavy kernel) and for rows that get a warp of their own (deferred,
    own cell below the clump threshold), hash-collision neighbourhoods — whose neighbours count once per colliding
    bucket — included."""
    rng = np.random.default_rng(5)
    s = sph.default_settings()
    d = rng.normal(size=(1500, 3))
    d *= (0.25 * rng.uniform(0, 1, (1500, 1)) ** (1 / 3)) / np.linalg.norm(d, axis=1, keepdims=True)
    pos = np.concatenate([d + [1.0, 1.0, 1.0], rng.uniform([-3, 0.2, -3], [3, 3, 3], (2500, 3))]).astype(np.float32)
    vel = rng.normal(0, 0.5, pos.shape).astype(np.float32)
    sim = sph.Sim(s, capacity=len(pos))
    sim.upload(pos, vel)
    sim.step(1)
    out = sim.download(sph.ORDER_ID, fields=("density", "force", "pos", "vel"))
    st = sim.stats()
    sim.close()
    assert st.deferred_density >= 1400
    dv = sph.derive(s)
    checked = {"light": 0, "clump": 0, "warp": 0, "collision": 0}
    rows = list(rng.choice(1500, 60, replace=False)) + list(1500 + rng.choice(2500, 60, replace=False))
    _check_rows_bit_for_bit(rows, pos, vel, out, st, s, dv, checked)
    assert checked["light"] >= 40 and checked["clump"] >= 10 and checked["warp"] >= 5, checked

    # The dense golden cube holds the rows in whose neighbourhood two of the reference's 27 buckets share a hash16:
    # 16 of them have neighbours that the reference therefore counts twice.
    from conftest import load_golden
    g = load_golden("cube20_step200.npz")
    pos, vel = g["pos0"], g["vel0"]
    sim = sph.Sim(s, capacity=len(pos))
    sim.upload(pos, vel)
    sim.step(1)
    out = sim.download(sph.ORDER_ID, fields=("density", "force", "pos", "vel"))
    st = sim.stats()
    sim.close()
    cells = np.trunc(pos / np.float32(s.h)).astype(np.int64)
    collision_rows = [j for j, c in enumerate(map(tuple, cells)) if len(set(_bucket_hashes(c))) < 27]
    checked = {"light": 0, "clump": 0, "warp": 0, "collision": 0, "counted_twice": 0}
    _check_rows_bit_for_bit(collision_rows + list(rng.choice(len(pos), 20, replace=False)), pos, vel, out, st, s, dv, checked)
    assert checked["collision"] >= 60 and checked["counted_twice"] >= 16 and checked["light"] >= 10, checked


def _check_rows_bit_for_bit(rows, pos, vel, out, st, s, dv, checked):
    """Density, force, position and velocity of the given rows after one step from (pos, vel): the GPU's bits (out)
    against the numpy restatement in the order of the kernel class each row falls into."""
    f = np.float32
    h, h2, mp = f(s.h), f(dv.h2), f(f(s.mass) * f(dv.poly6))
    K = dict(h=h, h2=h2, mass=f(s.mass), gas=f(s.gas_constant), rest=f(s.rest_density), visc_mass=f(f(s.viscosity) * f(s.mass)),
             spiky_grad=f(dv.spiky_grad), spiky_lap=f(dv.spiky_lap))
    got, got_force = out["density"], out["force"]
    cells = np.trunc(pos / h).astype(np.int64)
    ids_by_cell = {}
    for j, c in enumerate(map(tuple, cells)):
        ids_by_cell.setdefault(c, []).append(j)
    for i in rows:
        dens, cnt, longest, own = _density_in_documented_order(pos, ids_by_cell, i, h, h2, mp, f(dv.self_dens))
        collision = len(set(_bucket_hashes(tuple(cells[i])))) < 27  # such a row may count some neighbours several times
        checked["collision"] += collision
        if "counted_twice" in checked:
            unique = _density_in_documented_order(pos, ids_by_cell, i, h, h2, mp, f(dv.self_dens), multiplicities=False)[1]
            checked["counted_twice"] += cnt > unique
        deferred = collision or longest > 96 or cnt > st.nlist_rows
        # which kernel sums this row: the density pass, then the force pass (a deferred row whose list fits is
        # back with the one-thread force kernel, which reads the list the heavy kernel wrote in walk order)
        dclass = "seq" if not deferred else ("tiled" if own >= 64 else "warp")
        fclass = "tiled" if dclass == "tiled" else ("warp" if cnt > st.nlist_rows else "seq")
        checked[{"seq": "light", "tiled": "clump", "warp": "warp"}[dclass]] += 1
        want = dens[dclass]
        assert got[i].view(np.uint32) == want.view(np.uint32), (i, dclass, cnt, float(got[i]), float(want))
        # the force sums in the same orders, from the densities the GPU computed (compared bit for bit above for
        # this row; its neighbours' enter as they are)
        fwant = _force_in_documented_order(pos, vel, got, ids_by_cell, i, K)[fclass]
        assert np.array_equal(got_force[i].view(np.uint32), fwant.view(np.uint32)), (i, dclass, fclass, got_force[i], fwant)
        # ... and from there the integration and the walls: the whole step of this row, bit for bit
        pw, vw = _integrate_as_the_reference_does(pos[i], vel[i], fwant, got[i], s)
        assert np.array_equal(out["pos"][i].view(np.uint32), pw.view(np.uint32)), (i, out["pos"][i], pw)
        assert np.array_equal(out["vel"][i].view(np.uint32), vw.view(np.uint32)), (i, out["vel"][i], vw)


def test_captured_steps_replay_the_same_bits():
    """sph_step replays CUDA graphs of 1 and 16 captured steps; SPH_B200_GRAPH=0 launches every kernel
    from the host. 37 = 2 x 16 + 5 steps, then a settings change (new key), an upload and 3 more."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import sph_b200 as S
from conftest import load_golden
g = load_golden("cube20_step200.npz")
s = S.default_settings()
sim = S.Sim(s, capacity=len(g["pos0"])); sim.upload(g["pos0"], g["vel0"])
sim.step(37)
a = sim.download(S.ORDER_ID)
s.viscosity = 1.5; sim.set_settings(s); sim.step(1); sim.step(2)
b = sim.download(S.ORDER_ID)
sim.upload(g["pos0"][:5000], g["vel0"][:5000]); sim.step(3)
c = sim.download(S.ORDER_ID)
st = sim.stats()
np.savez(sys.argv[1], steps=st.steps, launches=sim.launch_count, **{f"{n}_{k}": v for n, d in (("a", a), ("b", b), ("c", c)) for k, v in d.items()})
''' % (ROOT, os.path.join(ROOT, "tests"))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        outs = []
        for graph in ("1", "0"):
            env = dict(os.environ)
            env["SPH_B200_GRAPH"] = graph
            f = os.path.join(d, f"g{graph}.npz")
            subprocess.run([sys.executable, "-c", code, f], check=True, env=env)
            outs.append(dict(np.load(f)))
    assert int(outs[0]["steps"]) == int(outs[1]["steps"]) == 3
    assert int(outs[0]["launches"]) == int(outs[1]["launches"])
    for k in outs[1]:
        if k not in ("steps", "launches"):
            assert_bit_equal(outs[0][k], outs[1][k], f"graph replay vs host launches: {k}")


def test_force_kernel_variants_return_the_same_bits():
    """SPH_B200_FORCES_CFG: default = force terms with x, y packed as fp32x2; 7 = the scalar form; 6 stages
    each block's neighbourhoods in shared memory (kept as a measured alternative). Same lists,
    same order, the same rounded operations: the same bits."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import sph_b200 as S
from conftest import load_golden
g = load_golden("cube20_step200.npz")
sim = S.Sim(S.default_settings(), capacity=len(g["pos0"])); sim.upload(g["pos0"], g["vel0"]); sim.step(5)
np.savez(sys.argv[1], **sim.download(S.ORDER_ID))
''' % (ROOT, os.path.join(ROOT, "tests"))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        outs = []
        for cfg in ("2", "6", "7"):
            env = dict(os.environ)
            env["SPH_B200_FORCES_CFG"] = cfg
            f = os.path.join(d, f"f{cfg}.npz")
            subprocess.run([sys.executable, "-c", code, f], check=True, env=env)
            outs.append(dict(np.load(f)))
    for k in ("pos", "vel", "force", "density"):
        assert_bit_equal(outs[1][k], outs[0][k], f"tile-staged forces: {k}")
        assert_bit_equal(outs[2][k], outs[0][k], f"scalar force terms: {k}")


def test_device_scene_generators_and_reset_point(sph):
    """initParticles / the dam-break block generated on the device are bit-identical to the host generators
    (and hence to the reference's initParticles: init_cube_w15.npz), per x-range too; sph_reset restores the
    reset point device to device."""
    from conftest import load_golden
    g = load_golden("init_cube_w15.npz")
    s = sph.default_settings()
    sim = sph.Sim(s, capacity=3375)
    sim.scene_cube_device(15)
    d = sim.download(sph.ORDER_ID, fields=("pos", "vel"))
    assert_bit_equal(d["pos"], g["pos"], "device initParticles")
    assert not d["vel"].any()
    sim.set_reset_point()
    sim.step(40)
    moved = sim.download(sph.ORDER_ID, fields=("pos",))["pos"]
    assert np.abs(moved - g["pos"]).max() > 1e-3 and sim.stats().steps == 40
    sim.reset()
    assert sim.stats().steps == 0
    assert_bit_equal(sim.download(sph.ORDER_ID, fields=("pos",))["pos"], g["pos"], "after sph_reset")
    sim.step(40)  # and the run repeats itself bit for bit
    assert_bit_equal(sim.download(sph.ORDER_ID, fields=("pos",))["pos"], moved, "second run from the reset point")
    sim.close()
    with pytest.raises(sph.SphError):
        fresh = sph.Sim(s, capacity=8)
        try:
            fresh.reset()  # no reset point
        finally:
            fresh.close()
    # the dam-break block: whole, and as three x-ranges with their ids (h = 0.075: the config-1 recipe)
    h = 0.075
    s2 = sph.scaled_settings(h)
    nx, ny, nz, sep, org = 37, 23, 41, h * 16.0 / 15.0, (-7.8, 0.125, -1.6)
    want_pos, _ = sph.scene_block(nx, ny, nz, sep, org, h, 77)
    sim = sph.Sim(s2, capacity=nx * ny * nz)
    sim.scene_block_device(nx, ny, nz, sep, org, seed=77)
    assert_bit_equal(sim.download(sph.ORDER_ID, fields=("pos",))["pos"], want_pos, "device block")
    for i0, i1 in ((0, 11), (11, 12), (12, 37)):
        hp, hv, hid = sph.scene_block_slice(nx, ny, nz, sep, org, h, 77, i0, i1)
        sim.scene_block_device(nx, ny, nz, sep, org, seed=77, i0=i0, i1=i1)
        got = sim.download(sph.ORDER_DEVICE, fields=("pos", "id"))
        assert np.array_equal(got["id"], hid)
        assert_bit_equal(got["pos"], hp, f"device block rows [{i0}, {i1})")
    sim.close()


def test_density_kernel_variants_return_the_same_bits():
    """SPH_B200_DENSITY_CFG: 0 = staged row walk (default), 10 = the round-1 kernel, 50 = the TMA-staged walk
    (neighbourhoods brought into shared memory by cp.async.bulk under an mbarrier, double-buffered; kept as a
    measured alternative). Same neighbour lists in the same order, the same sequence of rounded
    additions: same bits, on a dense cube with hash-collision cells and on a state with a clump."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import sph_b200 as S
from conftest import load_golden
g = load_golden("cube20_step200.npz")
out = {}
sim = S.Sim(S.default_settings(), capacity=len(g["pos0"])); sim.upload(g["pos0"], g["vel0"]); sim.step(5)
out.update({"a_" + k: v for k, v in sim.download(S.ORDER_ID).items()})
rng = np.random.default_rng(5)
d = rng.normal(size=(1500, 3)); d *= (0.25 * rng.uniform(0, 1, (1500, 1)) ** (1 / 3)) / np.linalg.norm(d, axis=1, keepdims=True)
pos = np.concatenate([d + [1.0, 1.0, 1.0], rng.uniform([-3, 0.2, -3], [3, 3, 3], (2500, 3))]).astype(np.float32)
sim.upload(pos, np.zeros_like(pos)); sim.step(2)
out.update({"b_" + k: v for k, v in sim.download(S.ORDER_ID).items()})
np.savez(sys.argv[1], **out)
''' % (ROOT, os.path.join(ROOT, "tests"))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        outs = {}
        for cfg in ("0", "10", "50"):
            env = dict(os.environ)
            env["SPH_B200_DENSITY_CFG"] = cfg
            f = os.path.join(d, f"d{cfg}.npz")
            subprocess.run([sys.executable, "-c", code, f], check=True, env=env)
            outs[cfg] = dict(np.load(f))
    for cfg in ("10", "50"):
        for k in outs["0"]:
            if k.split("_")[1] in ("pos", "vel", "force", "density"):
                assert_bit_equal(outs[cfg][k], outs["0"][k], f"density cfg {cfg} vs default: {k}")


def test_errors_are_reported_not_thrown(sph):
    s = sph.default_settings()
    sim = sph.Sim(s, capacity=10)
    lib = sph.load_library()
    with pytest.raises(sph.SphError) as e:
        sim.step(1)
    assert e.value.code == 3  # SPH_ERR_STATE: step before upload
    with pytest.raises(sph.SphError) as e:
        sim.upload(np.zeros((11, 3), np.float32), np.zeros((11, 3), np.float32))
    assert e.value.code == 4  # SPH_ERR_CAPACITY
    sim.upload(np.random.default_rng(0).uniform(0.2, 1, (10, 3)).astype(np.float32), np.zeros((10, 3), np.float32))
    with pytest.raises(sph.SphError):
        sim.download(sph.ORDER_ID, fields=("force",))  # undefined before the first step
    sim.step(1)
    sim.upload(np.zeros((4, 3), np.float32) + 0.5, np.zeros((4, 3), np.float32), ids=np.array([7, 8, 9, 10], np.uint32))
    sim.step(1)
    with pytest.raises(sph.SphError):
        sim.download(sph.ORDER_ID)  # ids are not 0..n-1
    assert sim.download(sph.ORDER_DEVICE)["id"].tolist() == [7, 8, 9, 10]
    out = C.c_void_p()
    assert lib.sph_create(C.byref(s), 10, 99, C.byref(out)) != 0 and b"device" in lib.sph_last_error(None)
    sim.close()


def test_settings_change_and_reupload(sph, oracle):
    s = sph.default_settings()
    pos, vel = sph.scene_cube(8, s.h)
    sim = sph.Sim(s, capacity=2 * len(pos))
    sim.upload(pos, vel)
    sim.step(5)
    s2 = sph.default_settings(viscosity=3.5, h=0.2, gas_constant=2.0)
    sim.set_settings(s2)
    cur = sim.download(sph.ORDER_ID, fields=("pos", "vel"))
    sim.step(1)
    got = sim.download(sph.ORDER_ID)
    want = by_id(oracle.step(oracle.settings(s2.as_tuple7()), s2.dt, cur["pos"], cur["vel"]))
    assert np.array_equal(got["hash"], want["hash"])
    assert_fields_close(got, want, "after set_settings", gas_constant=2.0)
    sim.close()


@pytest.fixture(scope="module")
def million(sph):
    """BASELINE.json config 1: the 1 003 520-particle dam break, settled 200 steps."""
    h = 0.075
    s = sph.scaled_settings(h)
    sep = h * 16.0 / 15.0
    pos, vel = sph.scene_block(64, 80, 196, sep, ((h - 8.0) + sep, h * 5.0 / 3.0, -98 * sep), h, 1024)
    sim = sph.Sim(s, capacity=len(pos))
    sim.upload(pos, vel)
    sim.step(200)
    state = sim.download(sph.ORDER_ID, fields=("pos", "vel"))
    sim.close()
    return s, state


def test_full_size_determinism_and_upload_order_independence(sph, million):
    s, st = million
    n = len(st["pos"])
    perm = np.random.default_rng(5).permutation(n)
    outs = []
    for order in (None, None, perm):
        sim = sph.Sim(s, capacity=n)
        if order is None:
            sim.upload(st["pos"], st["vel"])
        else:
            sim.upload(st["pos"][order], st["vel"][order], ids=order.astype(np.uint32))
        sim.step(3)
        outs.append(sim.download(sph.ORDER_ID))
        sim.close()
    for k in ("pos", "vel", "force", "density"):
        assert_bit_equal(outs[1][k], outs[0][k], f"run-to-run {k}")
        assert_bit_equal(outs[2][k], outs[0][k], f"permuted upload {k}")  # cells are ordered by particle id


def test_full_size_sortedness_ids_and_hash_table(sph, oracle, million):
    s, st = million
    n = len(st["pos"])
    sim = sph.Sim(s, capacity=n)
    sim.upload(st["pos"], st["vel"])
    sim.step(1)
    dev = sim.download(sph.ORDER_DEVICE, fields=("id", "hash", "density"))
    table = sim.hash_table()
    h16 = sim.download(sph.ORDER_HASH16, fields=("hash", "id"))
    grid = sim.stats()
    sim.close()
    # every particle exactly once
    assert np.array_equal(np.sort(dev["id"]), np.arange(n, dtype=np.uint32))
    # start-of-step hashes are the oracle's, bit for bit, for all 1 M particles
    want_hash = oracle.hashes(st["pos"], s.h)
    assert np.array_equal(dev["hash"], want_hash[dev["id"]])
    # device order is sorted by dense-grid cell (x slowest, z, y fastest) of the start-of-step position
    cells = np.array([np.trunc(st["pos"][dev["id"], a] / np.float32(s.h)) for a in range(3)], np.int64)
    ox, oy, oz = grid.grid_origin
    nx, ny, nz = grid.grid_dim
    lin = ((cells[0] - ox) * nz + (cells[2] - oz)) * ny + (cells[1] - oy)
    assert (np.diff(lin) >= 0).all(), "rows are not cell-sorted"
    # the reference's hash -> first-index table, rebuilt from the hash16-ordered read-out
    assert (np.diff(h16["hash"].astype(np.int64)) >= 0).all()
    assert np.array_equal(table, oracle.neighbor_table(h16["hash"]))
    assert int((table != 0xFFFFFFFF).sum()) == len(np.unique(want_hash)) and (table[65536:] == 0xFFFFFFFF).all()
    assert np.isfinite(dev["density"]).all() and dev["density"].min() >= 9.28
