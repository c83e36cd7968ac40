"""Secondary same-box baseline: the reference's OWN CUDA step
(src/kernels/sphGPU.cu, compiled unmodified for sm_100a into oracle/_ref/libsph_refgpu.so) timed on
the state it is given: the settled bench state (--state file.npz with pos, vel: what bench.py passes,
so that it is the SAME interacting state the product is timed on), or the initial lattice of the
bench scene. TEST/BENCH INFRASTRUCTURE: run in a subprocess with a timeout — that code has no overflow
guard on its 32-neighbour lists (MAX_NEIGHBORS, src/kernels/sphGPU.cu:12) and off-by-one bounds, so it is
refused states whose largest neighbour count reaches 32, and its results are not compared with anything.

    python tools/bench_reference_cuda.py [--state settled.npz --neighbours-max 25] [--dims 64 80 196] [--h 0.075] [--steps 10]
Prints one JSON line.
"""
import argparse, ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sph_b200 as S

ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, nargs=3, default=[64, 80, 196])
ap.add_argument("--h", type=float, default=0.075)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--state", default=None, help=".npz with pos, vel (n x 3 float32)")
ap.add_argument("--neighbours-max", type=int, default=0, help="largest neighbour count of that state (refused at >= 32)")
args = ap.parse_args()
if args.state and args.neighbours_max >= 32:
    print(json.dumps({"unavailable": f"state has up to {args.neighbours_max} neighbours per particle: the reference kernel's "
                                     "32-entry lists would overflow (src/kernels/sphGPU.cu:103-105)"}))
    sys.exit(0)
lib_path = os.path.join(ROOT, "oracle", "_ref", "libsph_refgpu.so")
if not os.path.exists(lib_path):
    print(json.dumps({"unavailable": "oracle/_ref/libsph_refgpu.so not built"}))
    sys.exit(0)
L = C.CDLL(lib_path)
fp = C.POINTER(C.c_float)
L.refgpu_time_steps.argtypes = [C.c_uint64, fp, C.c_float, C.c_int, C.c_int, fp, fp]
L.refgpu_time_steps.restype = C.c_double
h = args.h
s = S.scaled_settings(h)
sep = h * 16.0 / 15.0
nx, ny, nz = args.dims
if args.state:
    st = np.load(args.state)
    pos, vel = np.ascontiguousarray(st["pos"], np.float32), np.ascontiguousarray(st["vel"], np.float32)
    what = f"the settled bench state ({pos.shape[0]} particles, up to {args.neighbours_max} neighbours each)"
else:
    pos, vel = S.scene_block(nx, ny, nz, sep, ((h - 8.0) + sep, h * 5.0 / 3.0, -nz * sep / 2.0), h, 1024)
    what = f"the initial {nx}x{ny}x{nz} lattice (no neighbours yet)"
n = pos.shape[0]
s7 = (C.c_float * 7)(*s.as_tuple7())
sec = L.refgpu_time_steps(n, s7, C.c_float(s.dt), 1, args.steps, pos.ctypes.data_as(fp), vel.ctypes.data_as(fp))
print(json.dumps({"kind": "reference CUDA path (src/kernels/sphGPU.cu, unmodified, sm_100a)", "particles": n,
                  "steps": args.steps, "ms_per_step": 1e3 * sec / args.steps, "value": n * args.steps / sec,
                  "unit": "particle-steps/s",
                  "sample": f"{args.steps} calls of updateParticlesGPU on {what}; "
                            "host AoS in, host AoS + mat4 out every call, as that function does (compare with e2e)",
                  "note": "not result-compatible with the CPU oracle (box 10, elasticity 1, 32-neighbour cap); timing only"}))
