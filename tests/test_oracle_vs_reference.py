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


def _replay(oracle, reference, s7, dt, p, v, steps, what):
    """Step both from the same state, the oracle replaying the reference's post-sort order: every field
    must be bit-identical (NaNs included: same positions, same payload-insensitive pattern)."""
    s = oracle.settings(s7)
    ids = np.arange(p.shape[0], dtype=np.uint32)
    for step in range(steps):
        r = reference.step(s7, dt, p, v, ids)
        inv = np.empty_like(ids)
        inv[ids] = np.arange(len(ids), dtype=np.uint32)
        o = oracle.step(s, dt, p, v, ids, order=inv[r["id"]])
        assert np.array_equal(o["hash"], r["hash"]), f"{what} step {step} hash"
        for k in ("pos", "vel", "force", "density", "pressure"):
            nan_o, nan_r = np.isnan(o[k]), np.isnan(r[k])
            assert np.array_equal(nan_o, nan_r), f"{what} step {step} {k}: NaN pattern"
            assert_bit_equal(np.where(nan_o, 0, o[k]).astype(np.float32), np.where(nan_r, 0, r[k]).astype(np.float32),
                             f"{what} step {step} {k}")
        p, v, ids = r["pos"], r["vel"], r["id"]
    return r


def test_scaled_settings_block_bit_exact(oracle, reference):
    """The bench recipe: h = 0.075, mass and dt scaled with it, on a block that starts
    compressed (separation 0.8 h) so that every particle has neighbours from the first step."""
    h = np.float32(0.075)
    k = h / np.float32(0.15)
    s7 = (float(np.float32(0.02) * k * k * k), 1000.0, 1.0, 1.04, float(h), -9.8, 0.2)
    dt = float(np.float32(0.003) * k)
    g = np.stack(np.meshgrid(np.arange(14), np.arange(10), np.arange(12), indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(3)
    p = (g * (0.8 * float(h)) + [-7.9, 0.1, -0.4] + rng.uniform(-0.004, 0.004, g.shape)).astype(np.float32)
    v = rng.normal(0, 0.3, p.shape).astype(np.float32)
    r = _replay(oracle, reference, s7, dt, p, v, 6, "scaled block")
    assert r["density"].mean() > 10.5, "the block must be interacting (selfDens alone is 9.284)"


def test_adversarial_positions_bit_exact(oracle, reference):
    """Outside every wall, below the floor, negative coordinates around the double-width cell 0, and two
    coincident particles (NaN forces through normalize(0), src/sph.cpp:111)."""
    from oracle.pyoracle import DEFAULT_SETTINGS
    rng = np.random.default_rng(17)
    p = np.concatenate([
        rng.uniform([-12, -1, -12], [12, 6, 12], (1500, 3)),
        rng.uniform([-0.16, 0.1, -0.16], [0.16, 0.5, 0.16], (500, 3)),
        [[0.5, 1.0, 0.5], [0.5, 1.0, 0.5], [0.55, 1.0, 0.5]],
    ]).astype(np.float32)
    v = rng.normal(0, 1.0, p.shape).astype(np.float32)
    r = _replay(oracle, reference, DEFAULT_SETTINGS, 0.003, p, v, 3, "adversarial")
    assert np.isnan(r["pos"]).any(axis=1).sum() == 2, "the coincident pair went NaN in the first step and stays NaN"
