"""CPU: the numpy restatement of the library's step (tests/test_gpu_edge.py uses it to check the CUDA path bit for
bit, in the library's documented summation orders) is itself checked here against the oracle and against the
reference's own outputs in the golden fixtures: neighbour counts exact (multiplicities of hash-collision
neighbourhoods included), density and force within the tolerance that a change of summation order is allowed
(SURVEY.md App. B), integration and walls bit for bit. This pins the checker on a box without a GPU: what the GPU test
then proves is the ORDER, on top of physics already proven to be the reference's."""
import numpy as np

from conftest import by_id, load_golden
from test_gpu_edge import (_bucket_hashes, _density_in_documented_order, _force_in_documented_order,
                           _integrate_as_the_reference_does)


class _Step:  # the settings the integration needs, under the product's field names (shipped defaults)
    h, dt, g, box_half_width, elasticity, wall_offset = 0.15, 0.003, -9.8, 8.0, 0.5, 0.0001


def _check(rows, pos, vel, want, counts, s):
    f = np.float32
    h, h2, mp = f(s.h), f(s.h2), f(s.massPoly6Product)
    K = dict(h=h, h2=h2, mass=f(s.mass), gas=f(s.gasConstant), rest=f(s.restDensity), visc_mass=f(f(s.viscosity) * f(s.mass)),
             spiky_grad=f(s.spikyGrad), spiky_lap=f(s.spikyLap))
    cells = np.trunc(pos / h).astype(np.int64)
    ids_by_cell = {}
    for j, c in enumerate(map(tuple, cells)):
        ids_by_cell.setdefault(c, []).append(j)
    fmed = float(np.median(np.linalg.norm(want["force"], axis=1)))
    counted_twice = 0
    for i in rows:
        dens, cnt, longest, own = _density_in_documented_order(pos, ids_by_cell, i, h, h2, mp, f(s.selfDens))
        unique = _density_in_documented_order(pos, ids_by_cell, i, h, h2, mp, f(s.selfDens), multiplicities=False)[1]
        counted_twice += cnt > unique
        if counts is not None:
            assert cnt == counts[i], (i, cnt, counts[i])
        for got in dens.values():
            assert abs(float(got) - float(want["density"][i])) <= 1e-5 * float(want["density"][i]), (i, got, want["density"][i])
        forces = _force_in_documented_order(pos, vel, want["density"], ids_by_cell, i, K)
        fw = want["force"][i]
        scale = max(float(np.linalg.norm(fw)), fmed)
        for gotf in forces.values():
            assert float(np.linalg.norm(gotf - fw)) <= 1e-3 * scale, (i, gotf, fw)
        # integration + walls from the REFERENCE's force and density: the restatement must then give its row
        pw, vw = _integrate_as_the_reference_does(pos[i], vel[i], fw, want["density"][i], _Step)
        assert np.array_equal(pw.view(np.uint32), want["pos"][i].view(np.uint32)), (i, pw, want["pos"][i])
        assert np.array_equal(vw.view(np.uint32), want["vel"][i].view(np.uint32)), (i, vw, want["vel"][i])
    return counted_twice


def test_restated_step_is_the_reference_physics_clump_scene(oracle):
    rng = np.random.default_rng(5)
    s = oracle.settings()
    d = rng.normal(size=(1500, 3))
    d *= (0.25 * rng.uniform(0, 1, (1500, 1)) ** (1 / 3)) / np.linalg.norm(d, axis=1, keepdims=True)
    pos = np.concatenate([d + [1.0, 1.0, 1.0], rng.uniform([-3, 0.2, -3], [3, 3, 3], (2500, 3))]).astype(np.float32)
    vel = rng.normal(0, 0.5, pos.shape).astype(np.float32)
    want = by_id(oracle.step(s, 0.003, pos, vel))
    order, ocounts, _, _, _ = oracle.neighbor_lists(s, pos)
    rows = list(rng.choice(1500, 25, replace=False)) + list(1500 + rng.choice(2500, 25, replace=False))
    _check(rows, pos, vel, want, ocounts[np.argsort(order)], s)


def test_restated_step_is_the_reference_physics_double_counts(oracle):
    """The dense golden cube: 76 rows sit in neighbourhoods where two of the 27 buckets share a hash16, 16 of them have
    neighbours the reference counts twice. Checked against the REFERENCE's own outputs stored in the fixture."""
    g = load_golden("cube20_step200.npz")
    s = oracle.settings()
    pos, vel = g["pos0"], g["vel0"]
    inv = np.argsort(g["id1"])
    want = {"density": g["density1"][inv], "force": g["force1"][inv], "pos": g["pos1"][inv], "vel": g["vel1"][inv]}
    order, ocounts, _, _, _ = oracle.neighbor_lists(s, pos)
    cells = np.trunc(pos / np.float32(s.h)).astype(np.int64)
    collision_rows = [j for j, c in enumerate(map(tuple, cells)) if len(set(_bucket_hashes(c))) < 27]
    assert len(collision_rows) == 76
    rng = np.random.default_rng(7)
    twice = _check(collision_rows + list(rng.choice(len(pos), 20, replace=False)), pos, vel, want, ocounts[np.argsort(order)], s)
    assert twice == 16
