cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python tools/ab_env.py SPH_B200_HEAVY_BLOCKS 4 8 2 > gpurun_out/ab_heavy_blocks.jsonl 2> gpurun_out/ab_heavy_blocks.err; cut -c1-420 gpurun_out/ab_heavy_blocks.jsonl; tail -3 gpurun_out/ab_heavy_blocks.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 5405 -c 80 --csv --log-file gpurun_out/launches_fluid.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-weak-base --no-other-configs --e2e-steps 3 > /dev/null 2>&1
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_fluid.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ki].split("(")[0][:50]].append(float(r[vi].replace(",", "")))
    except ValueError: pass
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])): print(f"{k:52s} n={len(v):3d} mean={sum(v)/len(v)/1000:9.2f} us")
P
