"""Long-run behaviour of the 1 M dam break: per-window step time, neighbour statistics, NaN count."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sph_b200 as S
h = 0.075
s = S.scaled_settings(h)
sep = h * 16.0 / 15.0
pos, vel = S.scene_block(64, 80, 196, sep, ((h - 8.0) + sep, h * 5.0 / 3.0, -98 * sep), h, 1024)
sim = S.Sim(s, capacity=len(pos)); sim.upload(pos, vel)
stream = torch.cuda.ExternalStream(sim.stream)
win = int(sys.argv[1]) if len(sys.argv) > 1 else 500
nwin = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for w in range(nwin):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream); sim.enable_pass_timing(True); sim.step(win); e1.record(stream)
    sim.sync()
    pt = sim.pass_times(); sim.enable_pass_timing(False)
    st = sim.stats()
    print(f"steps {(w+1)*win:6d}  ms/step {e0.elapsed_time(e1)/win:.4f}  grid {pt['grid']:.3f} dens {pt['density']:.3f} forces {pt['forces']:.3f}  "
          f"mean rho {st.mean_density:.2f} max rho {st.max_density:.0f} KE {st.kinetic_energy:.1f} nan {st.nan_count} grid {list(st.grid_dim)} deferred d/f {st.deferred_density}/{st.deferred_forces} rows {st.nlist_rows}", flush=True)
