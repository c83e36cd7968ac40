"""Secondary same-box baseline: the reference's OWN CUDA step
(src/kernels/sphGPU.cu, compiled unmodified for sm_100a into oracle/_ref/libsph_refgpu.so) timed on
the sparse initial lattice of the bench scene. TEST/BENCH INFRASTRUCTURE: run in a subprocess with a
timeout — that code has no overflow guard on its 32-neighbour lists and off-by-one bounds, so it is
only ever given states with few neighbours, and its results are not compared with anything.

    python tools/bench_reference_cuda.py [--dims 64 80 196] [--h 0.075] [--steps 3]
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
args = ap.parse_args()
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
pos, vel = S.scene_block(nx, ny, nz, sep, ((h - 8.0) + sep, h * 5.0 / 3.0, -nz * sep / 2.0), h, 1024)
n = pos.shape[0]
s7 = (C.c_float * 7)(*s.as_tuple7())
sec = L.refgpu_time_steps(n, s7, C.c_float(s.dt), 1, args.steps, pos.ctypes.data_as(fp), vel.ctypes.data_as(fp))
print(json.dumps({"kind": "reference CUDA path (src/kernels/sphGPU.cu, unmodified, sm_100a)", "particles": n,
                  "steps": args.steps, "ms_per_step": 1e3 * sec / args.steps, "value": n * args.steps / sec,
                  "unit": "particle-steps/s",
                  "sample": f"{args.steps} calls of updateParticlesGPU on the initial {nx}x{ny}x{nz} lattice (no neighbours yet); "
                            "host AoS in, host AoS + mat4 out every call, as that function does",
                  "note": "not result-compatible with the CPU oracle (box 10, elasticity 1, 32-neighbour cap); timing only"}))
