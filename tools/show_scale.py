"""Print the essentials of the multi-GPU bench lines in gpurun_out/ (or the files given)."""
import glob, json, sys
files = sys.argv[1:] or sorted(glob.glob("gpurun_out/bench_weak_n*_r02.json"))
for f in files:
    for l in open(f):
        if not l.startswith("{"):
            continue
        d = json.loads(l); c = d["config"]
        b = c.get("single_gpu_same_workload") or {}
        print(f"N={d['n_gpus']} value {d['value']/1e9:.2f} G  ms/step {d['ms_per_step']:.4f}  base {b.get('ms_per_step', 0):.4f} ms  "
              f"eff {c.get('scaling_efficiency_vs_same_workload') or 0:.3f}  launches/step {d['gpu_launches']/d['steps']:.0f}  general {c['general_steps_in_timed_region']}")
        print("   parity:", c.get("slab_parity"))
        ph = c.get("phase_ms_per_step") or {}
        print("   phases max:", ph.get("max_over_ranks"), "limiting", ph.get("limiting_phase"))
        s = c.get("strong_scaling_16M")
        if s:
            print(f"   strong 16M: {s['value']/1e9:.2f} G  {s['ms_per_step']:.4f} ms  per gpu {s['particles_per_gpu']}")
        print(f"   e2e {d['e2e']['value']/1e9:.3f} G  cpu {d['cpu_baseline']['value']/1e6 if d.get('cpu_baseline') else 0:.2f} M ({(d.get('cpu_baseline') or {}).get('cores')} cores)  clocks {d['clocks']}")
