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
CPU step (oracle/_ref, or the oracle port when that is not built) on the host cores; that arm
imports nothing of the product (the scene comes from the oracle's generator, which the tests prove
bit-identical to the product's).

Scaling series: the N = 1 line is config 1 (the configuration the metric is quoted on); the like-for-like
single-GPU point of the weak-scaling series (config 3, one 8 M block) is measured in the same N = 1 run and
carried as config.weak_scaling_base, and every N > 1 line carries the same measurement taken in its own job
(config.single_gpu_same_workload) together with config.scaling_efficiency_vs_same_workload.
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


def scaled_settings7(h):
    """SURVEY.md 8(d) scaling of the shipped defaults, as sph_b200.scaled_settings computes it (float32
    arithmetic), without importing the product: (mass, restDensity, gasConstant, viscosity, h, g, tension), dt."""
    k = np.float32(h) / np.float32(0.15)
    mass = float(np.float32(0.02) * k * k * k)
    dt = float(np.float32(0.003) * k)
    return (mass, 1000.0, 1.0, 1.04, float(np.float32(h)), -9.8, 0.2), dt


def source_hash():
    """Hash of the CUDA sources: a committed ncu traffic figure is only quoted for the code it was taken on."""
    import hashlib
    hsh = hashlib.sha256()
    d = os.path.join(ROOT, "sph-fluid-simulator_b200", "csrc")
    for f in sorted(os.listdir(d)):
        with open(os.path.join(d, f), "rb") as fh:
            hsh.update(fh.read())
    return hsh.hexdigest()[:16]


def measured_traffic(kernel):
    """DRAM bytes per launch of a kernel from the committed ncu --set full capture of this workload
    (profiles/r01_traffic.json, written by tools/ncu_traffic.py), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        t = json.load(f)
    if t.get("source_hash") != source_hash():
        return None  # captured on other kernels: stale, not quoted
    return t.get("kernels", {}).get(kernel)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons for the timed region. An NVML query is NOT free for the GPU it asks about:
    measured in round 2 on a two-GPU run, every query stalls that rank's kernel launches for ~10 ms (polling
    every 100 ms cost the sampled rank 0.9 ms per step of a 20-step region, every 20 ms from a separate
    process 5 ms per step). So: one sample when the region starts (before the first timed launch), one the
    moment it ends (the GPU is still at its load clocks), and in between only one every 0.5 s — a long region
    is sampled throughout, a 25 ms one is not disturbed. The after-effect of a query outlasts the call (the
    stall shows up in the steps launched after it returned), so the first sample is taken before the warm-up
    steps, not between them and the timed region. nvidia-smi -lms as the fallback without the NVML binding."""
    PERIOD_S = 0.5
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
        self.off = os.environ.get("SPH_CLOCK_SAMPLER") == "off"  # diagnostic runs only: no clocks key
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
        while not self._stop.wait(self.PERIOD_S):
            self._sample()

    def start(self):
        """Call BEFORE the warm-up steps: the first sample is taken here, then 30 ms pass, so that its
        after-effect on the GPU is over when the (untimed) warm-up steps bring the GPU back under load."""
        if self.off:
            return
        if self.nvml:
            self._sample()
            time.sleep(0.03)
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "500"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.off:
            return None
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
                    "reasons": sorted(k for k, v in self.BITS.items() if bits & v), "samples": len(sm),
                    "source": f"nvml: at the start and the end of the timed region and every {self.PERIOD_S} s in between"}
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
def reference_sample(scene, time_steps, settle, warmup, steps, budget_s):
    """The reference CPU step on a bounded sample of `scene`: a z-slice of the same lattice (the fluid spans
    the full z width; the x-y cross section is kept), sized so that settle + warm-up + timed steps fit the
    budget; the rate is probed on a thin slice first. Nothing of the product is imported: the lattice comes
    from the oracle's generator (bit-identical to the product's, tests/test_abi.py)."""
    from oracle.pyoracle import Oracle
    O = Oracle()
    s7, dt = scaled_settings7(scene["h"])
    nx, ny, nz = scene["dims"]
    pos, vel = O.init_block(nx, ny, 8, scene["sep"], scene["origin"], scene["h"], scene["seed"])
    sec, _, _ = time_steps(s7, dt, 1, 2, pos, vel)
    rate = pos.shape[0] * 2 / sec
    total_steps = settle + warmup + steps
    n_budget = rate * budget_s / max(total_steps, 1)
    snz = int(max(4, min(nz, n_budget // (nx * ny))))
    origin = (scene["origin"][0], scene["origin"][1], -snz * scene["sep"] / 2.0)
    pos, vel = O.init_block(nx, ny, snz, scene["sep"], origin, scene["h"], scene["seed"])
    n = pos.shape[0]
    _, pos, vel = time_steps(s7, dt, settle, 0, pos, vel)
    sec, pos, vel = time_steps(s7, dt, warmup, steps, pos, vel)
    sample = (f"{nx}x{ny}x{snz} = {n} particle slice of the {scene['name']} lattice ({nx * ny * nz} particles), "
              f"settled {settle} steps on the CPU, then {steps} timed steps ({sec:.1f} s)")
    return n, sec, dt, sample


def run_reference(args):
    """The reference's own CPU step on the host cores, same metric. Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    kind, cores, time_steps = reference_step_fn()
    scene = dam_break_1m() if args.gpus == 1 else weak_scaling_block(args.gpus)
    nx, ny, nz = scene["dims"]
    n, sec, dt, sample = reference_sample(scene, time_steps, args.settle, args.warmup, args.steps, args.reference_budget_s)
    value = n * args.steps / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": scene["name"], "particles": nx * ny * nz, "h": scene["h"], "dt": dt, "lattice": [nx, ny, nz],
                   "settle_steps": args.settle, "sample_particles": n,
                   "note": "reference updateParticlesCPU on host threads; a particle-step costs the same whatever the slice "
                           "(the per-particle work is set by the local density, which the slice keeps), see cpu_baseline.sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.4: 148 SMs x 128 fp32 lanes x 2 (FMA) x 1.965 GHz


def work_model(candidates_mean, neighbours_mean):
    """SURVEY.md 8(d) FP32 work model, flop per particle-step: 9 per candidate pair in each of the two
    neighbour passes, +6 per accepted neighbour in the density pass, +45 in the force pass."""
    return {"density": 9.0 * candidates_mean + 6.0 * neighbours_mean,
            "forces": 9.0 * candidates_mean + 45.0 * neighbours_mean,
            "step": 18.0 * candidates_mean + 51.0 * neighbours_mean}


def other_configs(S, torch, dev, peak):
    """BASELINE.json configs 0 and 4 as sub-records of the N = 1 line (so that they reach a driver box):
    the shipped default cube (launch-bound: one captured graph per step) and neighbour search only
    (hash + counting sort by cell + cell ranges) on uniform-random fields, 128 B/particle algorithmic."""
    out = {}
    s = S.default_settings()
    sim = S.Sim(s, capacity=3375, device=dev)
    sim.scene_cube_device(15)  # src/Tester.cpp:90-91: SPHSystem(15, SPHSettings(0.02, 1000, 1, 1.04, 0.15, -9.8, 0.2))
    sim.step(300)
    sim.sync()
    t0 = time.perf_counter()
    sim.step(3000)
    sim.sync()
    dt = time.perf_counter() - t0
    st = sim.stats()
    out["config0_default_cube"] = {"particles": 3375, "steps": 3000, "steps_per_second": 3000 / dt, "us_per_step": 1e6 * dt / 3000,
                                   "value": 3375 * 3000 / dt, "unit": UNIT, "mean_density": st.mean_density,
                                   "call": "sph_step(h, 0.003, 3000): captured CUDA graphs of 16 steps, wall clock incl. launches"}
    sim.close()
    rows = []
    for m in (1, 16, 100):
        n = m * 1000000
        h = float((4096.0 * 4.0 / n) ** (1.0 / 3.0))  # ~4 particles per cell in the 16 x 16 x 16 box
        g = torch.Generator(device=f"cuda:{dev}")
        g.manual_seed(1024)
        lo = torch.tensor([h - 8, h, h - 8], device=f"cuda:{dev}")
        hi = torch.tensor([8 - h, 16.0, 8 - h], device=f"cuda:{dev}")
        p4 = torch.zeros((n, 4), dtype=torch.float32, device=f"cuda:{dev}")
        p4[:, :3] = lo + (hi - lo) * torch.rand((n, 3), generator=g, device=f"cuda:{dev}")
        v4 = torch.zeros((n, 4), dtype=torch.float32, device=f"cuda:{dev}")
        torch.cuda.synchronize(dev)
        sim = S.Sim(S.scaled_settings(h), capacity=n, device=dev)
        sim.upload_device(n, p4.data_ptr(), v4.data_ptr())
        del p4, v4
        sim.neighbor_search(3)
        sim.sync()
        stream = torch.cuda.ExternalStream(sim.stream, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sim.neighbor_search(10)
        e1.record(stream)
        sim.sync()
        ms = e0.elapsed_time(e1) / 10
        rows.append({"particles": n, "h": h, "grid_cells": int(sim.stats().grid_cells), "ms": ms, "particles_per_s": n / (ms * 1e-3),
                     "algorithmic_gbs": 128.0 * n / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": 128.0 * n / (ms * 1e-3) / 1e9 / peak})
        sim.close()
        torch.cuda.empty_cache()
    out["config4_neighbour_search_only"] = {"field": "uniform random in the box, ~4 particles per cell, device-generated (seed 1024)",
                                            "call": "sph_neighbor_search(h, 10): 5 kernels per build", "sizes": rows}
    return out


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

    warmup = max(args.warmup, 3)
    clocks = ClockSampler(dev)
    clocks.start()
    for _ in range(warmup):
        sim.step(1)
    sim.sync()

    # ---- the timed region: K steps through sph_step (the call a resident simulation makes: one captured
    # CUDA graph per step), one event pair per step on the library's stream, L2 flushed between steps ----
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
    launches = sim.launch_count - launches0
    clk = clocks.stop()
    total_s = float(step_ms.sum()) * 1e-3
    value = n * args.steps / total_s
    st1 = sim.stats()

    # Per-pass shares from a separate short loop with per-pass events (host launches instead of the graph
    # replay, L2 flushed the same way): the passes' SHARES apply to the timed steps, their sum is a little larger.
    sim.enable_pass_timing(True)
    with torch.cuda.stream(stream):
        for _ in range(min(args.steps, 20)):
            flush.zero_()
            sim.step(1)
    passes = sim.pass_times()
    sim.enable_pass_timing(False)

    # Unflushed steady state (state stays in the 126 MB L2 between steps), for information.
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        sim.step(args.steps)
        e1.record(stream)
    sim.sync()
    warm_value = n * args.steps / (e0.elapsed_time(e1) * 1e-3)

    # Work statistics of the timed state: candidates looked at and neighbours accepted per particle.
    cand_mean = sim.candidates_mean()
    _, counts, _, _ = sim.neighbor_lists()
    nb_mean = float(counts.mean())

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
    sim.close()

    # ---- roofline of the dominant kernel: both roofs, from live CUDA-event pass times ----
    peak, peak_src = peaks()
    flop = work_model(cand_mean, nb_mean)
    pass_bytes = {"grid": B_ALG_GRID, "density": B_ALG_DENSITY, "forces": B_ALG_FORCES}
    dominant = max(("grid", "density", "forces"), key=lambda k: passes[k])
    dom_name = {"forces": "k_forces_integrate", "density": "k_density_staged", "grid": "grid build (5 kernels)"}[dominant]
    per_pass = {}
    for k in ("grid", "density", "forces"):
        t = passes[k] * 1e-3
        hbm = pass_bytes[k] * n / t / 1e9
        row = {"ms": passes[k], "hbm_achieved_gbs": hbm, "hbm_frac": hbm / peak}
        if k in flop:
            row["fp32_achieved_tflops"] = flop[k] * n / t / 1e12
            row["fp32_frac"] = row["fp32_achieved_tflops"] / FP32_PEAK_TFLOPS
        row["bound"] = "hbm" if row["hbm_frac"] >= row.get("fp32_frac", 0.0) else "fp32-issue"
        per_pass[k] = row
    dom = per_pass[dominant]
    if dom["bound"] == "hbm":
        achieved, unit, rpeak = dom["hbm_achieved_gbs"], "GB/s", peak
    else:
        achieved, unit, rpeak = dom["fp32_achieved_tflops"], "TFLOP/s", FP32_PEAK_TFLOPS
    roofline = {"bound": dom["bound"], "kernel": dom_name, "achieved": achieved, "peak": rpeak, "unit": unit,
                "frac": achieved / rpeak, "traffic": measured_traffic(dom_name),
                "peak_source": peak_src if dom["bound"] == "hbm" else
                "148 SMs x 128 fp32 lanes x 2 (FMA) x 1.965 GHz; the step's arithmetic is unfused by contract, so 0.5 is its ceiling",
                "algorithmic_bytes_per_launch": pass_bytes[dominant] * n,
                "model_flop_per_launch": flop.get(dominant, 0.0) * n,
                "launch_ms": passes[dominant], "per_pass": per_pass,
                "work_model": "SURVEY.md 8(d): 9 flop per candidate pair and pass, +6 (density) / +45 (forces) per accepted "
                              "neighbour; candidates_mean and neighbours_mean are in config",
                "issue_note": "ncu (profiles/): the neighbour passes are bound by instruction issue and the L1 data pipe "
                              "(density: issue 73 %, L1 wavefronts 80 % of peak), not by HBM",
                "step_achieved_gbs": B_ALG_STEP * value / 1e9, "step_frac": B_ALG_STEP * value / 1e9 / peak,
                "step_fp32_tflops": flop["step"] * value / 1e12, "step_fp32_frac": flop["step"] * value / 1e12 / FP32_PEAK_TFLOPS}

    # ---- the weak-scaling series' single-GPU point (config 3 at G = 1: one 8 M block), same run ----
    weak_base = None
    if not args.no_weak_base:
        slab = __import__("importlib").import_module("sph-fluid-simulator_b200.slab")
        s3 = S.scaled_settings(weak_scaling_block(1)["h"])
        weak_base = slab.single_gpu_base(S, args, weak_scaling_block, s3, dev, warmup)
        weak_base["step_hbm_frac"] = weak_base["step_achieved_gbs"] / peak

    others = None if args.no_other_configs else other_configs(S, torch, dev, peak)

    # ---- CPU baseline on a bounded sample: the same settled state, a few full-size steps ----
    cpu = None
    if not args.no_cpu_baseline:
        kind, cores, time_steps = reference_step_fn()
        cs = args.cpu_steps
        sec, _, _ = time_steps(s.as_tuple7(), s.dt, 1, cs, settled["pos"], settled["vel"])
        cpu = {"value": n * cs / sec, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{cs} steps of the full {n}-particle settled state (after {args.settle} GPU steps), "
                         f"{sec:.1f} s of wall time"}

    # ---- secondary baseline: the reference's own CUDA path on this GPU, on the SAME settled state
    # (neighbours_max stays below its 32-entry lists there), in a subprocess with a timeout ----
    ref_cuda = None
    if not args.no_cpu_baseline:
        import tempfile
        try:
            with tempfile.TemporaryDirectory() as d:
                f = os.path.join(d, "settled.npz")
                np.savez(f, pos=settled["pos"], vel=settled["vel"])
                out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_reference_cuda.py"), "--state", f,
                                      "--h", str(scene["h"]), "--steps", "10", "--neighbours-max", str(int(counts.max()))],
                                     capture_output=True, text=True, timeout=240)
            rws = [l for l in out.stdout.splitlines() if l.startswith("{")]
            ref_cuda = json.loads(rws[-1]) if rws else {"unavailable": (out.stderr or "no output")[-200:]}
        except Exception as e:  # noqa: BLE001 - a crash of that code must not take the bench down
            ref_cuda = {"unavailable": repr(e)[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": warmup,
        "ms_per_step": float(step_ms.mean()), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": scene["name"], "particles": n, "h": scene["h"], "dt": s.dt, "lattice": [nx, ny, nz],
                   "settle_steps": args.settle, "l2": "flushed between timed steps (256 MiB memset outside the event pair)",
                   "steps_per_second": 1e3 / float(step_ms.mean()),
                   "neighbours_mean": nb_mean, "neighbours_max": int(counts.max()), "candidates_mean": cand_mean,
                   "mean_density": st1.mean_density, "grid_dim": list(st1.grid_dim), "grid_cells": int(st1.grid_cells),
                   "nan_count": int(st1.nan_count), "value_l2_warm": warm_value,
                   "ms_per_step_p50": float(np.median(step_ms)), "ms_per_step_max": float(step_ms.max()),
                   "timed_call": "sph_step(h, dt, 1) per step: one captured CUDA graph (9 kernels) replayed per call",
                   "weak_scaling_base": weak_base,
                   "other_configs": others,
                   "reference_cuda_baseline": ref_cuda},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 60 * n, "d2h_bytes_per_step": 124 * n,
                "steps": e2e_steps, "call": "sph_update_particles_aos (60-byte Particle rows + mat4 transforms, pinned host memory)",
                "resident_simulation_with_renderer_readout": resident},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def cpu_sample_for(args):
    """cpu_baseline of an N > 1 line: the reference CPU step on a bounded slice of that line's lattice."""
    def sample(scene, s):
        kind, cores, time_steps = reference_step_fn()
        n, sec, dt, text = reference_sample(scene, time_steps, min(args.settle, 100), 1, 4, args.multi_cpu_budget_s)
        return {"value": n * 4 / sec, "unit": UNIT, "cores": cores, "kind": kind, "sample": text}
    return sample


def run_multi_gpu(args):
    slab = __import__("importlib").import_module("sph-fluid-simulator_b200.slab")
    scene_fn = dam_break_16m if args.workload == "dam-break-16M" else weak_scaling_block
    args.strong_scene = dam_break_16m(args.gpus)  # carried as a sub-record of the weak-scaling lines
    slab.bench_weak_scaling(args, scene_fn, METRIC, UNIT, ClockSampler, peaks, cpu_sample_for(args))


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
    ap.add_argument("--reference-budget-s", type=float, default=150.0)
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-weak-base", action="store_true", help="skip the single-GPU run of the weak-scaling per-GPU workload")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the config-0 / config-4 sub-records of the N=1 line")
    ap.add_argument("--no-slab-parity", action="store_true", help="skip the cross-GPU bit-parity check of the N>1 lines")
    ap.add_argument("--no-strong-subrecord", action="store_true", help="skip the config-2 (16 M, strong scaling) sub-record of the N>1 lines")
    ap.add_argument("--multi-cpu-budget-s", type=float, default=25.0, help="CPU seconds for the cpu_baseline of an N>1 line")
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
