"""CPU: live pin of the plain-C oracle to the UNMODIFIED reference code (oracle/_ref). Skipped where
the reference harness was not built (it needs /root/reference at build time; the prebuilt .so
travels with the snapshot, the sources do not)."""
import numpy as np

from conftest import assert_bit_equal, load_golden


def test_settings_and_layout(oracle, reference):
    assert reference.lib.ref_sizeof_particle() == 60   # src/Particle.h
    assert reference.lib.ref_table_size() == 262144
    r = reference.make_settings()
    s = oracle.settings()
    for k in ("poly6", "spikyGrad", "spikyLap", "h2", "selfDens", "massPoly6Product"):
        assert np.float32(getattr(s, k)).view(np.uint32) == np.float32(r[k]).view(np.uint32), k
    assert r["sphereScale"][0] == s.sphereScale and r["sphereScale"][15] == 1.0


def test_cell_hash_random(oracle, reference):
    rng = np.random.default_rng(5)
    for _ in range(2000):
        p = rng.uniform(-20, 20, 3).astype(np.float32)
        h = float(rng.choice([0.15, 0.075, 0.02]))
        c = reference.get_cell(p, h)
        assert oracle.get_cell(p, h) == c
        assert oracle.get_hash(c) == reference.get_hash(c)


def test_steps_bit_exact_when_replaying_reference_order(oracle, reference):
    from oracle.pyoracle import DEFAULT_SETTINGS
    g = load_golden("cube20_step200.npz")
    s = oracle.settings()
    p, v, ids = g["pos0"], g["vel0"], np.arange(g["pos0"].shape[0], dtype=np.uint32)
    for step in range(12):
        r = reference.step(DEFAULT_SETTINGS, 0.003, p, v, ids)
        inv = np.empty_like(ids)
        inv[ids] = np.arange(len(ids), dtype=np.uint32)
        o = oracle.step(s, 0.003, p, v, ids, order=inv[r["id"]])
        for k in ("pos", "vel", "force", "density", "pressure"):
            assert_bit_equal(o[k], r[k], f"step {step} {k}")
        assert np.array_equal(o["hash"], r["hash"])
        assert np.array_equal(oracle.neighbor_table(o["hash"]), reference.neighbor_table(r["hash"]))
        p, v, ids = r["pos"], r["vel"], r["id"]


def test_init_cube_matches(oracle, reference):
    for w in (3, 7):
        pr, vr = reference.init_cube(w)
        po, vo = oracle.init_cube(w, oracle.settings())
        assert_bit_equal(po, pr, "init")
