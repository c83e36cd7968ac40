"""CPU: the plain-C oracle (oracle/sph_oracle.c) against the golden fixtures minted from the
reference build (tests/golden/make_golden.py). This is what pins the oracle on a box that has no
/root/reference."""
import numpy as np
import pytest

from conftest import assert_bit_equal, assert_fields_close, by_id, load_golden

STATES = ["cube15_step0.npz", "cube15_step100.npz", "cube15_step300.npz", "cube20_step200.npz"]


def test_settings_known_answers(oracle):
    g = load_golden("settings_kat.npz")
    names = [str(n) for n in g["names"]]
    for row_in, row_out in zip(g["inputs"], g["outputs"]):
        s = oracle.settings(tuple(float(v) for v in row_in))
        got = dict(poly6=s.poly6, spikyGrad=s.spikyGrad, spikyLap=s.spikyLap, gasConstant=s.gasConstant,
                   mass=s.mass, h2=s.h2, selfDens=s.selfDens, restDensity=s.restDensity, viscosity=s.viscosity,
                   h=s.h, g=s.g, tension=s.tension, massPoly6Product=s.massPoly6Product, sphereScale=s.sphereScale)
        for n, want in zip(names, row_out):
            assert np.float32(got[n]).view(np.uint32) == np.float32(want).view(np.uint32), (n, got[n], want)
    # the shipped defaults, as hex
    s = oracle.settings()
    assert np.float32(s.poly6).view(np.uint32) == 0x4C1B75D1
    assert np.float32(s.h2).view(np.uint32) == 0x3CB851EC


def test_cell_and_hash_known_answers(oracle):
    g = load_golden("cell_hash_kat.npz")
    for a, h in enumerate(g["h_values"]):
        h16 = oracle.hashes(g["pos"], float(h))
        assert np.array_equal(h16, (g["hashes"][a] & 0xFFFF).astype(np.uint16))
        for i in range(0, g["pos"].shape[0], 37):
            assert oracle.get_cell(g["pos"][i], float(h)) == tuple(g["cells"][a, i])
    for c, want in zip(g["icell"], g["ihash"]):
        assert oracle.get_hash(c) == int(want)
    assert int(g["ihash"].max()) < 262144


@pytest.mark.parametrize("w", [1, 2, 15])
def test_init_particles(oracle, w):
    g = load_golden(f"init_cube_w{w}.npz")
    pos, vel = oracle.init_cube(w, oracle.settings())
    assert_bit_equal(pos, g["pos"], "initParticles positions")
    assert_bit_equal(vel, g["vel"], "initParticles velocities")


@pytest.mark.parametrize("name", STATES)
def test_step_bit_exact_with_reference_order(oracle, name):
    """Fed the reference's own post-sort order, every field must match the reference bit for bit."""
    g = load_golden(name)
    s = oracle.settings(tuple(float(v) for v in g["settings"]))
    o = oracle.step(s, float(g["dt"]), g["pos0"], g["vel0"], order=g["id1"], transforms=True)
    assert np.array_equal(o["id"], g["id1"]) and np.array_equal(o["hash"], g["hash1"])
    for k in ("pos", "vel", "force", "density", "pressure", "transforms"):
        assert_bit_equal(o[k], g[k + "1"], f"{name}:{k}")


@pytest.mark.parametrize("name", STATES)
def test_step_with_own_stable_order(oracle, name):
    g = load_golden(name)
    s = oracle.settings(tuple(float(v) for v in g["settings"]))
    o = oracle.step(s, float(g["dt"]), g["pos0"], g["vel0"])
    assert np.array_equal(o["hash"], g["hash1"])  # sorted hash column is order-independent
    assert np.array_equal(oracle.neighbor_table(o["hash"]), g["table1"])
    want = by_id({k[:-1]: g[k] for k in ("pos1", "vel1", "force1", "density1", "pressure1", "id1")})
    assert_fields_close(by_id(o), want, name)


def test_neighbor_table_edge_cases(oracle):
    t = oracle.neighbor_table(np.zeros(0, np.uint16))
    assert (t == 0xFFFFFFFF).all() and t.shape == (262144,)
    t = oracle.neighbor_table(np.array([3, 3, 3, 9, 65535, 65535], np.uint16))
    assert t[3] == 0 and t[9] == 3 and t[65535] == 4 and (np.delete(t, [3, 9, 65535]) == 0xFFFFFFFF).all()


def test_free_running_steps(oracle):
    g = load_golden("cube20_step200.npz")
    s = oracle.settings(tuple(float(v) for v in g["settings"]))
    p, v, ids = g["pos0"], g["vel0"], np.arange(g["pos0"].shape[0], dtype=np.uint32)
    for _ in range(int(g["free_steps"])):
        o = oracle.step(s, float(g["dt"]), p, v, ids)
        p, v, ids = o["pos"], o["vel"], o["id"]
    inv, ginv = np.argsort(ids), np.argsort(g["idN"])
    assert np.abs(p[inv] - g["posN"][ginv]).max() < 2e-5


def test_multiplicity_fixture_has_double_counts(oracle):
    g = load_golden("cube20_step200.npz")
    s = oracle.settings(tuple(float(v) for v in g["settings"]))
    order, counts, cand, offs, lst = oracle.neighbor_lists(s, g["pos0"])
    repeats = sum(len(np.unique(lst[int(offs[k]):int(offs[k + 1])])) != counts[k] for k in range(len(counts)))
    assert repeats >= 10 and counts.mean() > 4


def test_class_surface_semantics(oracle):
    """update() is a no-op until startSimulation(), dt is forced to 0.003, reset re-seeds."""
    g = load_golden("class_surface.npz")
    s = oracle.settings()
    pos, vel = oracle.init_cube(15, s)
    assert_bit_equal(pos, g["pos_not_started"], "not started")
    assert_bit_equal(pos, g["pos_after_reset"], "reset")
    p, v, ids = pos, vel, np.arange(pos.shape[0], dtype=np.uint32)
    for _ in range(20):
        o = oracle.step(s, 0.003, p, v, ids)
        p, v, ids = o["pos"], o["vel"], o["id"]
    inv = np.argsort(ids)
    assert np.abs(p[inv] - g["pos_after20"]).max() <= 1e-6
