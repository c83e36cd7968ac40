"""Hot SASS regions of one kernel from an .ncu-rep source page: instructions executed, average
active threads and stall samples, grouped in address order (so loops show up as blocks).

    python tools/ncu_hot.py <rep> <kernel-regex> [top]
"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
lines = raw.splitlines()
# first line is the kernel name row
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[start:end]))) if r.get("Instructions Executed") not in (None, "")]
tot_inst = sum(int(r["Instructions Executed"]) for r in rows)
tot_samp = sum(int(r["# Samples"]) for r in rows)
print(f"total warp insts {tot_inst}, samples {tot_samp}, SASS lines {len(rows)}")
ranked = sorted(range(len(rows)), key=lambda i: -int(rows[i]["Instructions Executed"]))[:top]
keep = sorted(ranked)
for i in keep:
    r = rows[i]
    print(f"{i:5d} {int(r['Instructions Executed']):>10d} {100*int(r['Instructions Executed'])/tot_inst:5.1f}% "
          f"thr {float(r['Avg. Threads Executed']):5.1f} samp {100*int(r['# Samples'])/max(tot_samp,1):5.1f}%  {r['Source'].strip()[:90]}")
