"""Step rate of small systems (config 0 = the shipped 15^3 cube and a few larger cubes): launch-bound regime."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sph_b200 as S

s = S.default_settings()
for w in (15, 30, 40, 59):
    pos, vel = S.scene_cube(w, s.h)
    sim = S.Sim(s, capacity=len(pos)); sim.upload(pos, vel)
    sim.step(300); sim.sync()
    t = time.perf_counter(); sim.step(3000); sim.sync(); dt = time.perf_counter() - t
    st = sim.stats()
    print(f"W={w} N={len(pos)} us/step {dt / 3000 * 1e6:.1f} steps/s {3000 / dt:.0f} Mps/s {len(pos) * 3000 / dt / 1e6:.1f} mean rho {st.mean_density:.2f} nan {st.nan_count}", flush=True)
    sim.close()
