"""Short driver for ncu captures: settle the 1M dam break (or a custom lattice), then run a few steps.

    ncu --set full --clock-control none --import-source on -k regex:k_forces -s <skip> -c 1 -o gpurun_out/prof \
        python tools/profile_step.py --settle 600 --steps 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sph_b200 as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--settle", type=int, default=600)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--h", type=float, default=0.075)
ap.add_argument("--dims", type=int, nargs=3, default=[64, 80, 196])
args = ap.parse_args()

h = args.h
sep = h * 16.0 / 15.0
nx, ny, nz = args.dims
s = S.scaled_settings(h)
pos, vel = S.scene_block(nx, ny, nz, sep, ((h - 8.0) + sep, h * 5.0 / 3.0, -nz * sep / 2.0), h, 1024)
sim = S.Sim(s, capacity=pos.shape[0])
sim.upload(pos, vel)
sim.step(args.settle)
sim.sync()
sim.enable_pass_timing(True)
sim.step(args.steps)
sim.sync()
print("pass ms/step", sim.pass_times())
st = sim.stats()
print("particles", pos.shape[0], "mean density", st.mean_density, "grid", list(st.grid_dim), "nan", st.nan_count)
