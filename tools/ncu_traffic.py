"""Write profiles/traffic.json: DRAM bytes (read + write) per launch of each kernel in an ncu --set full
report, averaged over the captured launches, together with the hash of the CUDA sources the capture was
taken on (bench.py only quotes a traffic figure whose hash matches the sources it runs).

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep [source-hash-file]

The hash is read from gpurun_out/source_hash.txt when the capture script left it there (the sources may have
moved on since), else computed from the tree.
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
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
hash_file = sys.argv[2] if len(sys.argv) > 2 else os.path.join(root, "gpurun_out", "source_hash.txt")
if os.path.exists(hash_file):
    src = open(hash_file).read().strip()
else:
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    src = b.source_hash()
out = {"source_hash": src, "report": os.path.basename(rep), "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
       "kernels": {k: sum(v) / len(v) for k, v in acc.items()}}
path = os.path.join(root, "profiles", "traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
