"""Mint the golden fixtures in this directory from the UNMODIFIED reference step.

Run in the build container (needs /root/reference to have been compiled into
oracle/_ref/libsph_ref.so by `make -C oracle ref`):

    python tests/golden/make_golden.py

The reference ships no tests, golden vectors or fixtures of its own, so these are
produced by executing its code headless (oracle/ref_harness.cpp). Every array in the .npz files
comes out of the reference functions named next to it; nothing is computed by the restatement or by
the CUDA path. The tests then check BOTH the plain-C oracle and the CUDA path against them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle.pyoracle import DEFAULT_SETTINGS, Reference  # noqa: E402

R = Reference()


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def settings_kat():
    """SPHSettings::SPHSettings (src/SPHSystem.cpp:8-26) for a few constructor inputs."""
    rows_in, rows_out = [], []
    cases = [DEFAULT_SETTINGS,
             (0.02, 1000.0, 1.0, 3.5, 0.15, -9.8, 1.0),            # GUI RESET defaults, src/Tester.cpp:148-173
             (0.0025, 1000.0, 1.0, 1.04, 0.075, -9.8, 0.2),        # h = 0.075 scaling (1 M dam break)
             (4.7407406e-05, 1000.0, 1.0, 1.04, 0.02, -9.8, 0.2),  # h = 0.02 scaling (64 M weak scaling)
             (1.0, 2000.0, 5.0, 5.0, 1.0, -9.8, 0.2),
             (0.001, 0.001, 0.001, 0.001, 0.001, -9.8, 0.2)]
    names = None
    for c in cases:
        d = R.make_settings(c)
        names = [k for k in d if k != "sphereScale"]
        rows_in.append(np.array(c, np.float32))
        rows_out.append(np.array([d[k] for k in names] + [d["sphereScale"][0]], np.float32))
    save("settings_kat.npz", inputs=np.stack(rows_in), outputs=np.stack(rows_out),
         names=np.array(names + ["sphereScale"]))


def cell_hash_kat():
    """getCell / getHash (src/neighborTable.cpp:5-17) on adversarial inputs."""
    rng = np.random.default_rng(20231016)
    h_values = np.array([0.15, 0.075, 0.03, 0.02, 1.0, 0.001], np.float32)
    pts = [rng.uniform(-9, 9, (4000, 3)), rng.uniform(-1e-3, 1e-3, (500, 3)), rng.uniform(-800, 800, (500, 3))]
    # exact multiples of h, their float neighbours, values straddling zero
    for h in h_values:
        k = rng.integers(-60, 60, (300, 3)).astype(np.float32)
        base = (k * h).astype(np.float32)
        pts += [base, np.nextafter(base, np.float32(np.inf)), np.nextafter(base, np.float32(-np.inf))]
    pts.append(np.array([[0.0, -0.0, 0.0], [1e-45, -1e-45, 0.0], [-0.1499999, 0.1499999, 0.15],
                         [0.15, -0.15, 0.3], [-7.85, 0.15, 7.85], [7.8501, 0.1501, -7.8501]], np.float32))
    pos = np.concatenate(pts).astype(np.float32)
    cells = np.empty((len(h_values), pos.shape[0], 3), np.int32)
    hashes = np.empty((len(h_values), pos.shape[0]), np.uint32)
    for a, h in enumerate(h_values):
        for i, p in enumerate(pos):
            c = R.get_cell(p, float(h))
            cells[a, i] = c
            hashes[a, i] = R.get_hash(c)
    # raw hash on integer cells including negatives and large magnitudes
    icell = np.concatenate([rng.integers(-120, 120, (3000, 3)), rng.integers(-2 ** 20, 2 ** 20, (1000, 3)),
                            np.array([[0, 0, 0], [-1, -1, -1], [1, 1, 1], [-1, 0, 1], [2 ** 30, -2 ** 30, 12345]])]
                           ).astype(np.int32)
    ihash = np.array([R.get_hash(c) for c in icell], np.uint32)
    save("cell_hash_kat.npz", h_values=h_values, pos=pos, cells=cells, hashes=hashes, icell=icell, ihash=ihash)


def state_dump(name, s7, dt, pos, vel, free_steps=0):
    """One updateParticles(onGPU=false) step (src/sph.cpp:277-290) from (pos, vel): the state before,
    every Particle field after (in the reference's own post-sort order, with the ids that rode
    through std::sort), the transforms, and the neighbour table of the sorted hashes."""
    n = pos.shape[0]
    ids = np.arange(n, dtype=np.uint32)
    r = R.step(s7, dt, pos, vel, ids, nsteps=1, transforms=True)
    table = R.neighbor_table(r["hash"])
    out = dict(settings=np.array(s7, np.float32), dt=np.float32(dt), pos0=pos, vel0=vel,
               pos1=r["pos"], vel1=r["vel"], force1=r["force"], density1=r["density"],
               pressure1=r["pressure"], hash1=r["hash"], id1=r["id"], transforms1=r["transforms"],
               table1=table)
    if free_steps:
        rr = R.step(s7, dt, pos, vel, ids, nsteps=free_steps)
        out.update(free_steps=np.int32(free_steps), posN=rr["pos"], velN=rr["vel"], idN=rr["id"],
                   densityN=rr["density"])
    save(name, **out)


def advance(s7, dt, pos, vel, nsteps):
    r = R.step(s7, dt, pos, vel, None, nsteps=nsteps)
    inv = np.argsort(r["id"])  # back to id order so fixtures do not depend on the sort
    return r["pos"][inv], r["vel"][inv]


def main():
    settings_kat()
    cell_hash_kat()

    # initParticles (src/SPHSystem.cpp:76-108): the shipped cube and two more widths.
    for w in (1, 2, 15):
        pos, vel = R.init_cube(w, DEFAULT_SETTINGS)
        save(f"init_cube_w{w}.npz", pos=pos, vel=vel)

    # Repo-default cube (config 0): before/after one step at steps 0, 100 and 300.
    pos, vel = R.init_cube(15, DEFAULT_SETTINGS)
    done = 0
    for at in (0, 100, 300):
        pos, vel = advance(DEFAULT_SETTINGS, 0.003, pos, vel, at - done) if at > done else (pos, vel)
        done = at
        state_dump(f"cube15_step{at}.npz", DEFAULT_SETTINGS, 0.003, pos, vel, free_steps=10 if at == 0 else 0)

    # Dense state with hash-collision double counts: 20^3 cube after 200 steps (mean ~5.5 neighbours,
    # 16 particles whose neighbour multiset has repeats).
    pos, vel = R.init_cube(20, DEFAULT_SETTINGS)
    pos, vel = advance(DEFAULT_SETTINGS, 0.003, pos, vel, 200)
    state_dump("cube20_step200.npz", DEFAULT_SETTINGS, 0.003, pos, vel, free_steps=10)

    # SPHSystem class surface as Tester drives it (src/Tester.cpp:110-125, 210-221).
    # Rows are returned in particle-id order (id = the index initParticles gave the particle).
    p_idle, _, _ = R.class_run(15, DEFAULT_SETTINGS, 5, start=False)
    p_run, v_run, i_run = R.class_run(15, DEFAULT_SETTINGS, 20, start=True)
    inv = np.argsort(i_run)
    p_reset, v_reset, _ = R.class_run(15, DEFAULT_SETTINGS, 20, start=True, reset_after=True)
    save("class_surface.npz", pos_not_started=p_idle, pos_after20=p_run[inv], vel_after20=v_run[inv],
         pos_after_reset=p_reset, vel_after_reset=v_reset)


if __name__ == "__main__":
    main()
