"""A/B of an environment switch of the library (read at sph_create) on the resident step: the default cube
(config 0), the 1 M dam break (config 1) and the 8 M block (config 3 base), same state, bits compared.

    python tools/ab_env.py SPH_B200_PDL 0 1 [--big]
"""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sph_b200 as S
import bench as B

var, values = sys.argv[1], [a for a in sys.argv[2:] if not a.startswith("--")]
big = "--big" in sys.argv


def run(name, make, settle, steps, reps=5):
    out = {}
    for v in values:
        os.environ[var] = v
        sim = make()
        sim.step(settle); sim.sync()
        stream = torch.cuda.ExternalStream(sim.stream)
        ms = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); sim.step(steps); e1.record(stream); sim.sync()
            ms.append(e0.elapsed_time(e1) / steps)
        sim.enable_pass_timing(True); sim.step(20); pt = sim.pass_times(); sim.enable_pass_timing(False)
        d = sim.download(S.ORDER_ID, fields=("pos", "vel", "density"))
        dig = hashlib.sha1(d["pos"].tobytes() + d["vel"].tobytes() + d["density"].tobytes()).hexdigest()[:16]
        out[v] = dict(ms_per_step=sorted(ms)[len(ms) // 2], all=[round(m, 5) for m in ms], digest=dig,
                      pass_ms={k: round(x, 4) for k, x in pt.items() if k != "steps"})
        sim.close()
    digs = {o["digest"] for o in out.values()}
    print(json.dumps({"workload": name, "var": var, "results": out, "bit_identical": len(digs) == 1}), flush=True)


def cube():
    s = S.default_settings()
    pos, vel = S.scene_cube(15, s.h)
    sim = S.Sim(s, capacity=len(pos)); sim.upload(pos, vel)
    return sim


def block(b):
    def make():
        s = S.scaled_settings(b["h"])
        nx, ny, nz = b["dims"]
        sim = S.Sim(s, capacity=int(nx * ny * nz * 1.02) + 1024)
        sim.scene_block_device(nx, ny, nz, b["sep"], b["origin"], b["seed"])
        return sim
    return make


run("default-cube-15", cube, 300, 3200)
run("dam-break-1M", block(B.dam_break_1m()), 600, 160)
if big:
    run("weak-scaling-8M-x1", block(B.weak_scaling_block(1)), 600, 48, reps=3)
