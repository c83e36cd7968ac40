#!/usr/bin/env python
"""bench.py — particle-steps/s of the SPH step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one SPHSystem::update()-equivalent pass (neighbour search -> density/pressure ->
forces -> integration) over every particle of the workload.

Workloads:
  N = 1  config 1: 1 003 520-particle dam break (h = 0.075, lattice 64 x 80 x 196), scaled
         default settings, fixed dt = 0.0015, settled for --settle steps before timing.
  N > 1  config 3: weak scaling, 61 x 256 x 512 = 7 995 392 particles per GPU (h = 0.02),
         slab-partitioned along x with ghost exchange (sph-fluid-simulator_b200/slab.py).

Timing: CUDA events on the library's stream around each step, L2 flushed (256 MiB memset) between
timed steps, max over ranks. `value` has the state resident in HBM; `e2e` goes through the
stateless drop-in call (sph_update_particles_aos: 60-byte Particle rows in pinned host memory up,
Particle rows + mat4 transforms down, every step). `--impl reference` times the reference's own
CPU step (oracle/_ref, or the oracle port when that is not built) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec"
UNIT = "particle-steps/s"
B_ALG_STEP = 292.0    # algorithmic bytes per particle-step, whole step
B_ALG_FORCES = 140.0  # forces (R pos 16 + vel 16 + (rho, p) 8 + W force 16) + integration (84), one fused kernel
B_ALG_DENSITY = 24.0  # density pass: R pos 16 + W (rho, p) 8
B_ALG_GRID = 128.0    # hash + sort + cell ranges (passes 1-4)
B_ALG_INTEGRATE = 0.0  # fused into the forces kernel


def dam_break_1m():
    """Config 1."""
    h = 0.075
    sep = h * 16.0 / 15.0
    nx, ny, nz = 64, 80, 196
    origin = ((h - 8.0) + sep, h * 5.0 / 3.0, -nz * sep / 2.0)
    return dict(name="dam-break-1M", h=h, sep=sep, dims=(nx, ny, nz), origin=origin, seed=1024)


def weak_scaling_block(g):
    """Config 3: per-GPU block 61 x 256 x 512, G GPUs => nx = 61 G."""
    h = 0.02
    sep = h * 16.0 / 15.0
    nx, ny, nz = 61 * g, 256, 512
    origin = ((h - 8.0) + sep, h * 5.0 / 3.0, -nz * sep / 2.0)
    return dict(name=f"weak-scaling-8M-per-gpu-x{g}", h=h, sep=sep, dims=(nx, ny, nz), origin=origin, seed=1024)


def dam_break_16m(g):
    """Config 2: 16 M dam break, the same field for any number of GPUs (strong scaling)."""
    h = 0.03
    sep = h * 16.0 / 15.0
    nx, ny, nz = 160, 256, 392
    origin = ((h - 8.0) + sep, h * 5.0 / 3.0, -nz * sep / 2.0)
    return dict(name="dam-break-16M", h=h, sep=sep, dims=(nx, ny, nz), origin=origin, seed=1024, scaling="strong")


def measured_traffic(kernel):
    """DRAM bytes per launch of a kernel from the committed ncu --set full capture of this workload
    (profiles/r01_traffic.json, written by tools/ncu_traffic.py), or None."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f).get(kernel)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled during the timed region: NVML from a thread every 100 ms
    (a sample costs microseconds, so even a quarter-second region gets several), nvidia-smi -lms as the
    fallback when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.rows = []      # nvidia-smi text rows
        self.samples = []   # (sm MHz, reasons bitmask) from NVML
        self.proc = None
        self.index = index
        self.nvml = None
        self.dev = None
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            if os.environ.get("SPH_CLOCK_SAMPLER", "nvml") != "nvml":
                raise ImportError("nvidia-smi sampler requested")
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.dev = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:  # noqa: BLE001 - older torch without .uuid, or CUDA_VISIBLE_DEVICES remapping unknown
                self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:  # noqa: BLE001
            self.nvml = None

    def _sample(self):
        try:
            mhz = float(self.nvml.nvmlDeviceGetClockInfo(self.dev, self.nvml.NVML_CLOCK_SM))
            try:
                bits = int(self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.dev))
            except AttributeError:
                bits = int(self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev))
            self.samples.append((mhz, bits))
        except Exception:  # noqa: BLE001
            pass

    def _loop(self):
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(0.1)

    def start(self):
        if self.nvml:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml:
            self._sample()  # the region has just ended: the GPU is still at its load clocks
            self._stop.set()
            if self._thread:
                self._thread.join(timeout=1.0)
            sm = [m for m, _ in self.samples]
            bits = 0
            for _, b in self.samples:
                bits |= b
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(k for k, v in self.BITS.items() if bits & v), "samples": len(sm), "source": "nvml"}
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nme in enumerate(self.NAMES):
                if len(r) > 2 + k and r[2 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def reference_step_fn():
    """(kind, cores, time_steps(s7, dt, warmup, steps, pos, vel) -> (seconds, pos, vel))."""
    from oracle.pyoracle import Oracle, Reference, build
    if Reference.available():
        R = Reference()
        return "reference", R.hardware_concurrency(), R.time_steps
    build(ref=False)
    O = Oracle()

    def run(s7, dt, warmup, steps, pos, vel):
        return O.time_steps(O.settings(s7), dt, warmup, steps, pos, vel)
    return "port", O.num_threads(), run


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU step on the host cores, same metric. Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import sph_b200 as S
    kind, cores, time_steps = reference_step_fn()
    scene = dam_break_1m() if args.gpus == 1 else weak_scaling_block(args.gpus)
    s = S.scaled_settings(scene["h"])
    s7 = s.as_tuple7()
    nx, ny, nz = scene["dims"]
    # Bounded sample: a z-slice of the same lattice (the fluid spans the full z width; the x-y cross
    # section is kept), sized so that settle + warm-up + timed steps fit the budget. The rate is
    # probed on a thin slice first.
    pos, vel = S.scene_block(nx, ny, 8, scene["sep"], scene["origin"], scene["h"], scene["seed"])
    sec, _, _ = time_steps(s7, s.dt, 1, 2, pos, vel)
    rate = pos.shape[0] * 2 / sec
    total_steps = args.settle_reference + args.warmup + args.steps
    n_budget = rate * args.reference_budget_s / max(total_steps, 1)
    n_full = nx * ny * nz
    snx = nx
    snz = int(max(4, min(nz, n_budget // (nx * ny))))
    origin = (scene["origin"][0], scene["origin"][1], -snz * scene["sep"] / 2.0)
    pos, vel = S.scene_block(snx, ny, snz, scene["sep"], origin, scene["h"], scene["seed"])
    n = pos.shape[0]
    _, pos, vel = time_steps(s7, s.dt, args.settle_reference, 0, pos, vel)
    sec, pos, vel = time_steps(s7, s.dt, args.warmup, args.steps, pos, vel)
    value = n * args.steps / sec
    sample = (f"{snx}x{ny}x{snz} = {n} particle slice of the {scene['name']} lattice ({n_full} particles), "
              f"settled {args.settle_reference} steps on the CPU, then {args.steps} timed steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": scene["name"], "particles_per_step": n, "h": scene["h"], "dt": s.dt,
                   "note": "reference updateParticlesCPU on host threads; bounded sample, see cpu_baseline.sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_single_gpu(args):
    import torch
    import sph_b200 as S

    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    scene = dam_break_1m()
    s = S.scaled_settings(scene["h"])
    nx, ny, nz = scene["dims"]
    pos, vel = S.scene_block(nx, ny, nz, scene["sep"], scene["origin"], scene["h"], scene["seed"])
    n = pos.shape[0]
    sim = S.Sim(s, capacity=n, device=dev)
    sim.upload(pos, vel)
    stream = torch.cuda.ExternalStream(sim.stream, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{dev}")

    sim.step(args.settle)  # the lattice starts with zero neighbours (sep > h): let it collapse first
    sim.sync()
    settled = sim.download(S.ORDER_ID, fields=("pos", "vel"))
    st0 = sim.stats()

    for _ in range(max(args.warmup, 3)):
        sim.step(1)
    sim.sync()

    clocks = ClockSampler(dev)
    clocks.start()
    sim.enable_pass_timing(True)
    launches0 = sim.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize(dev)
    with torch.cuda.stream(stream):
        for a, b in ev:
            flush.zero_()  # evict the previous step's lines from L2 (outside the event pair)
            a.record(stream)
            sim.step(1)
            b.record(stream)
    sim.sync()
    torch.cuda.synchronize(dev)
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    passes = sim.pass_times()
    sim.enable_pass_timing(False)
    launches = sim.launch_count - launches0
    clk = clocks.stop()
    total_s = float(step_ms.sum()) * 1e-3
    value = n * args.steps / total_s
    st1 = sim.stats()

    # Unflushed steady state (state stays in the 126 MB L2 between steps), for information.
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        sim.step(args.steps)
        e1.record(stream)
    sim.sync()
    warm_value = n * args.steps / (e0.elapsed_time(e1) * 1e-3)

    # Neighbour statistics of the timed state (mean / max accepted neighbours).
    _, counts, _, _ = sim.neighbor_lists()

    # ---- end to end through the stateless drop-in call, host buffers, copies inside the timing ----
    rows_t = torch.zeros((n, 15), dtype=torch.int32).pin_memory()
    mats_t = torch.zeros((n, 16), dtype=torch.float32).pin_memory()
    rows = rows_t.numpy().view(np.uint32)
    cur = sim.download(S.ORDER_ID, fields=("pos", "vel"))
    rows[:, 0:3] = cur["pos"].view(np.uint32)
    rows[:, 3:6] = cur["vel"].view(np.uint32)
    import ctypes as C
    lib = S.load_library()
    fptr = mats_t.numpy().ctypes.data_as(C.POINTER(C.c_float))
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    # Secondary figure: the simulation stays resident (the SPHSystem class surface) and every step
    # only reads back what the renderer draws — model matrices (64 B/particle) or float4 positions.
    resident = {}
    for name, fn, floats in (("mat4_transforms", lib.sph_write_transforms, 16), ("float4_positions", lib.sph_read_positions, 4)):
        for k in range(2 + e2e_steps):
            if k == 2:
                t0 = time.perf_counter()
            sim.step(1)
            rc = fn(sim.handle, fptr)
            assert rc == 0, lib.sph_last_error(sim.handle)
        resident[name] = {"value": n * e2e_steps / (time.perf_counter() - t0), "unit": UNIT,
                          "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4 * floats * n}
    for _ in range(2):
        rc = lib.sph_update_particles_aos(sim.handle, rows.ctypes.data_as(C.c_void_p), fptr, n, C.c_float(s.dt))
        assert rc == 0, lib.sph_last_error(sim.handle)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rc = lib.sph_update_particles_aos(sim.handle, rows.ctypes.data_as(C.c_void_p), fptr, n, C.c_float(s.dt))
        assert rc == 0, lib.sph_last_error(sim.handle)
    e2e_s = time.perf_counter() - t0
    e2e_value = n * e2e_steps / e2e_s

    # ---- roofline of the dominant kernel (live CUDA-event pass times over the timed region) ----
    peak, peak_src = peaks()
    pass_bytes = {"grid": B_ALG_GRID, "density": B_ALG_DENSITY, "forces": B_ALG_FORCES, "integrate": B_ALG_INTEGRATE}
    dominant = max(("grid", "density", "forces"), key=lambda k: passes[k])
    dom_name = {"forces": "k_forces_integrate", "density": "k_density", "grid": "grid build (5 kernels)", "integrate": "-"}[dominant]
    achieved = pass_bytes[dominant] * n / (passes[dominant] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": measured_traffic(dom_name), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": pass_bytes[dominant] * n,
                "launch_ms": passes[dominant], "pass_ms": {k: passes[k] for k in pass_bytes},
                "step_achieved_gbs": B_ALG_STEP * value / 1e9, "step_frac": B_ALG_STEP * value / 1e9 / peak}

    # ---- CPU baseline on a bounded sample: the same settled state, a few full-size steps ----
    cpu = None
    if not args.no_cpu_baseline:
        kind, cores, time_steps = reference_step_fn()
        cs = args.cpu_steps
        sec, _, _ = time_steps(s.as_tuple7(), s.dt, 1, cs, settled["pos"], settled["vel"])
        cpu = {"value": n * cs / sec, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{cs} steps of the full {n}-particle settled state (after {args.settle} GPU steps), "
                         f"{sec:.1f} s of wall time"}

    # ---- secondary baseline: the reference's own CUDA path on this GPU (subprocess, bounded) ----
    ref_cuda = None
    if not args.no_cpu_baseline:
        try:
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_reference_cuda.py")],
                                 capture_output=True, text=True, timeout=180)
            rows = [l for l in out.stdout.splitlines() if l.startswith("{")]
            ref_cuda = json.loads(rows[-1]) if rows else {"unavailable": (out.stderr or "no output")[-200:]}
        except Exception as e:  # noqa: BLE001 - a crash of that code must not take the bench down
            ref_cuda = {"unavailable": repr(e)[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": float(step_ms.mean()), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": scene["name"], "particles": n, "h": scene["h"], "dt": s.dt, "lattice": [nx, ny, nz],
                   "settle_steps": args.settle, "l2": "flushed between timed steps (256 MiB memset outside the event pair)",
                   "neighbours_mean": float(counts.mean()), "neighbours_max": int(counts.max()),
                   "mean_density": st1.mean_density, "grid_dim": list(st1.grid_dim), "grid_cells": int(st1.grid_cells),
                   "nan_count": int(st1.nan_count), "value_l2_warm": warm_value,
                   "ms_per_step_p50": float(np.median(step_ms)), "ms_per_step_max": float(step_ms.max())},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 60 * n, "d2h_bytes_per_step": 124 * n,
                "steps": e2e_steps, "call": "sph_update_particles_aos (60-byte Particle rows + mat4 transforms, pinned host memory)",
                "resident_simulation_with_renderer_readout": resident},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "reference_cuda_baseline": ref_cuda,
    }
    print(json.dumps(line), flush=True)
    sim.close()


def run_multi_gpu(args):
    slab = __import__("importlib").import_module("sph-fluid-simulator_b200.slab")
    scene_fn = dam_break_16m if args.workload == "dam-break-16M" else weak_scaling_block
    slab.bench_weak_scaling(args, scene_fn, METRIC, UNIT, ClockSampler, peaks)


def main():
    # NCCL prints its version banner to stdout at VERSION/INFO level; keep stdout to the one JSON line.
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN") and not os.environ.get("SPH_KEEP_NCCL_DEBUG"):
        os.environ.pop("NCCL_DEBUG")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--settle", type=int, default=600, help="untimed steps that let the lattice collapse before timing")
    ap.add_argument("--settle-reference", type=int, default=300)
    ap.add_argument("--reference-budget-s", type=float, default=150.0)
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-weak-base", action="store_true", help="skip the single-GPU run of the per-GPU workload at N>1")
    ap.add_argument("--workload", default="auto", choices=["auto", "dam-break-1M", "weak", "dam-break-16M"],
                    help="auto: config 1 at N=1, config 3 (weak scaling, 8 M particles per GPU) at N>1; "
                         "dam-break-16M: config 2 through the slab driver (strong scaling, any N)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus == 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1 and args.workload not in ("weak", "dam-break-16M"):
        return run_single_gpu(args)
    return run_multi_gpu(args)


if __name__ == "__main__":
    main()
