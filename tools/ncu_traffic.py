"""Write profiles/r01_traffic.json: DRAM bytes (read + write) per launch of each kernel in an
ncu --set full report, averaged over the captured launches.

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep
"""
import csv, io, json, os, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = collections.defaultdict(list)
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].split("<")[0].replace("sphb::", "").replace("void ", "").strip()
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[col[k]].replace(",", "")) * scale[units[col[k]]]
    acc[name].append(tot)
out = {k: sum(v) / len(v) for k, v in acc.items()}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r01_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
