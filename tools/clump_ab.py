"""Collapsed regime, same state for every variant: the 1 M dam break is stepped to step N once (default kernels), the
state is saved, and each SPH_B200_CLUMP_CELL setting steps that state with per-pass timing.

    python tools/clump_ab.py [--at 5000] [--state /tmp/clump_state.npz] [--cells 0 64] [--steps 10]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sph_b200 as S

ap = argparse.ArgumentParser()
ap.add_argument("--at", type=int, default=5000)
ap.add_argument("--state", default="/tmp/clump_state.npz")
ap.add_argument("--cells", type=int, nargs="+", default=[0, 64])
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()

h = 0.075
s = S.scaled_settings(h)
sep = h * 16.0 / 15.0
if not os.path.exists(args.state):
    pos, vel = S.scene_block(64, 80, 196, sep, ((h - 8.0) + sep, h * 5.0 / 3.0, -98 * sep), h, 1024)
    sim = S.Sim(s, capacity=len(pos)); sim.upload(pos, vel)
    sim.step(args.at); sim.sync()
    d = sim.download(S.ORDER_ID, fields=("pos", "vel"))
    np.savez(args.state, pos=d["pos"], vel=d["vel"])
    sim.close()
st0 = np.load(args.state)
ref = None
for cell in args.cells:
    os.environ["SPH_B200_CLUMP_CELL"] = str(cell)
    sim = S.Sim(s, capacity=len(st0["pos"])); sim.upload(st0["pos"], st0["vel"])
    sim.step(1)
    out = sim.download(S.ORDER_ID, fields=("density", "force"))
    sim.step(2)
    sim.enable_pass_timing(True); sim.step(args.steps); pt = sim.pass_times(); sim.enable_pass_timing(False)
    st = sim.stats()
    rel = None
    if ref is None:
        ref = out
    else:
        rel = float(np.max(np.abs(out["density"] - ref["density"]) / ref["density"]))
    print(json.dumps({"clump_cell": cell, "pass_ms": {k: round(v, 4) for k, v in pt.items() if k != "steps"},
                      "candidates_mean": round(sim.candidates_mean(), 1), "deferred": [int(st.deferred_density), int(st.deferred_forces)],
                      "mean_density": round(st.mean_density, 3), "max_density": round(st.max_density, 1),
                      "density_max_rel_diff_vs_first_after_1_step": rel}), flush=True)
    sim.close()
