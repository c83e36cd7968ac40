"""Basic-block view of one kernel from an .ncu-rep source page: consecutive SASS lines with the same
execution count are merged; prints warp-instructions per block and its share.

    python tools/ncu_blocks.py <rep> <kernel-regex> [min-share-%]
"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[start:end]))) if r.get("Instructions Executed") not in (None, "")]
tot = sum(int(r["Instructions Executed"]) for r in rows)
samp = sum(int(r["# Samples"]) for r in rows)
print(f"total warp insts {tot}, SASS lines {len(rows)}")
i = 0
while i < len(rows):
    j = i
    c = int(rows[i]["Instructions Executed"])
    while j + 1 < len(rows) and int(rows[j + 1]["Instructions Executed"]) == c:
        j += 1
    n = j - i + 1
    share = 100.0 * c * n / tot
    if share >= minshare:
        s = sum(int(rows[k]["# Samples"]) for k in range(i, j + 1))
        thr = float(rows[i]["Avg. Threads Executed"])
        ops = " ".join(rows[k]["Source"].strip().split()[0 if not rows[k]["Source"].strip().startswith("@") else 1] for k in range(i, min(j + 1, i + 14)))
        print(f"[{i:4d}-{j:4d}] n={n:3d} exec={c:>9d} share={share:5.1f}% thr={thr:4.1f} samp={100*s/max(samp,1):4.1f}%  {ops[:150]}")
    i = j + 1
