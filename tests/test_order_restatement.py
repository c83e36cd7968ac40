"""CPU: the numpy restatement of the library's summation orders (tests/test_gpu_edge.py uses it to check the CUDA
path bit for bit) is itself checked here against the oracle: same neighbour counts, and densities within the
tolerance that a change of summation order is allowed. This pins the checker on a box without a
GPU: what the GPU test then proves is the ORDER, on top of physics already proven to be the reference's."""
import numpy as np

from conftest import by_id
from test_gpu_edge import _density_in_documented_order, _force_in_documented_order, _integrate_as_the_reference_does


def test_restated_sums_are_the_reference_physics(oracle):
    rng = np.random.default_rng(5)
    s = oracle.settings()
    d = rng.normal(size=(1500, 3))
    d *= (0.25 * rng.uniform(0, 1, (1500, 1)) ** (1 / 3)) / np.linalg.norm(d, axis=1, keepdims=True)
    pos = np.concatenate([d + [1.0, 1.0, 1.0], rng.uniform([-3, 0.2, -3], [3, 3, 3], (2500, 3))]).astype(np.float32)
    vel = rng.normal(0, 0.5, pos.shape).astype(np.float32)
    want = by_id(oracle.step(s, 0.003, pos, vel))
    order, ocounts, _, _, _ = oracle.neighbor_lists(s, pos)
    counts = ocounts[np.argsort(order)]
    f = np.float32
    h, h2, mp = f(s.h), f(s.h2), f(s.massPoly6Product)
    cells = np.trunc(pos / h).astype(np.int64)
    ids_by_cell = {}
    for j, c in enumerate(map(tuple, cells)):
        ids_by_cell.setdefault(c, []).append(j)
    M = (73856093, 19349663, 83492791)
    K = dict(h=h, h2=h2, mass=f(s.mass), gas=f(s.gasConstant), rest=f(s.restDensity), visc_mass=f(f(s.viscosity) * f(s.mass)),
             spiky_grad=f(s.spikyGrad), spiky_lap=f(s.spikyLap))
    fmed = float(np.median(np.linalg.norm(want["force"][:1500], axis=1)))
    class Step:  # the settings the integration needs, under the product's field names
        h, dt, g, box_half_width, elasticity, wall_offset = s.h, 0.003, s.g, 8.0, 0.5, 0.0001
    checked = 0
    for i in list(rng.choice(1500, 25, replace=False)) + list(1500 + rng.choice(2500, 25, replace=False)):
        c = tuple(cells[i])
        hs = [((c[0] + x) * M[0] ^ (c[1] + y) * M[1] ^ (c[2] + z) * M[2]) & 0xFFFF
              for x in (-1, 0, 1) for y in (-1, 0, 1) for z in (-1, 0, 1)]
        if len(set(hs)) < 27:
            continue  # hash-collision neighbourhood: the reference counts some neighbours twice
        dens, cnt, longest, own = _density_in_documented_order(pos, ids_by_cell, i, h, h2, mp, f(s.selfDens))
        assert cnt == counts[i], (i, cnt, counts[i])
        for got in dens.values():
            assert abs(float(got) - float(want["density"][i])) <= 1e-5 * float(want["density"][i]), (i, got, want["density"][i])
        forces = _force_in_documented_order(pos, vel, want["density"], ids_by_cell, i, K)
        fw = want["force"][i]
        scale = max(float(np.linalg.norm(fw)), fmed)
        for gotf in forces.values():
            assert float(np.linalg.norm(gotf - fw)) <= 1e-3 * scale, (i, gotf, fw)
        # integration + walls from the ORACLE's force and density: the restatement must then give the oracle's row
        pw, vw = _integrate_as_the_reference_does(pos[i], vel[i], fw, want["density"][i], Step)
        assert np.array_equal(pw.view(np.uint32), want["pos"][i].view(np.uint32)), (i, pw, want["pos"][i])
        assert np.array_equal(vw.view(np.uint32), want["vel"][i].view(np.uint32)), (i, vw, want["vel"][i])
        checked += 1
    assert checked >= 40
