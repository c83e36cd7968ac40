"""A/B of the density-pass variants (SPH_B200_DENSITY_CFG) on the settled bench states: per-pass CUDA-event
times and a bitwise comparison of the results against the round-1 one-row-per-thread kernel (cfg 10).

    python tools/sweep_density.py [--cfgs 10 0 1 2 3] [--steps 20] [--big]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sph_b200 as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfgs", type=int, nargs="+", default=[10, 0, 1, 2, 3])
ap.add_argument("--forces", type=int, nargs="+", default=[2])
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--settle", type=int, default=600)
ap.add_argument("--big", action="store_true", help="also the 8 M block of config 3 (61 x 256 x 512, h = 0.02)")
args = ap.parse_args()

scenes = [("dam-break-1M", 0.075, (64, 80, 196))]
if args.big:
    scenes.append(("weak-8M-block", 0.02, (61, 256, 512)))

for name, h, (nx, ny, nz) in scenes:
    sep = h * 16.0 / 15.0
    s = S.scaled_settings(h)
    pos, vel = S.scene_block(nx, ny, nz, sep, ((h - 8.0) + sep, h * 5.0 / 3.0, -nz * sep / 2.0), h, 1024)
    n = pos.shape[0]
    base = S.Sim(s, capacity=n)
    base.upload(pos, vel)
    base.step(args.settle)
    state = base.download(S.ORDER_ID, fields=("pos", "vel"))
    base.close()
    ref = None
    for fcfg in args.forces:
        for cfg in args.cfgs:
            os.environ["SPH_B200_DENSITY_CFG"] = str(cfg)
            os.environ["SPH_B200_FORCES_CFG"] = str(fcfg)
            sim = S.Sim(s, capacity=n)
            sim.upload(state["pos"], state["vel"])
            sim.step(3)
            out = sim.download(S.ORDER_ID, fields=("pos", "vel", "density", "force"))
            sim.enable_pass_timing(True)
            sim.step(args.steps)
            pt = sim.pass_times()
            sim.enable_pass_timing(False)
            st = sim.stats()
            same = None
            if ref is None:
                ref = out
            else:
                same = all(np.array_equal(out[k].view(np.uint32), ref[k].view(np.uint32)) for k in out)
            print(json.dumps({"scene": name, "n": n, "density_cfg": cfg, "forces_cfg": fcfg,
                              "pass_ms": {k: round(v, 4) for k, v in pt.items() if k != "steps"},
                              "step_ms": round(sum(v for k, v in pt.items() if k != "steps"), 4),
                              "bits_equal_to_first": same, "deferred_density": int(st.deferred_density),
                              "mean_density": round(st.mean_density, 4)}), flush=True)
            sim.close()
