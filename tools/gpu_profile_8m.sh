cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/profile_step.py --h 0.02 --dims 61 256 512 --settle 600 --steps 50
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 4214 -c 14 --csv --log-file gpurun_out/launches_8m.csv python tools/profile_step.py --h 0.02 --dims 61 256 512 --settle 600 --steps 4 > /dev/null 2>&1
python - <<'PY'
import csv, io, collections
lines=[l for l in open('gpurun_out/launches_8m.csv') if not l.startswith('==')]
r=csv.DictReader(io.StringIO(''.join(lines)))
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for row in r:
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']; m=row['Metric Name']
    if m=='gpu__time_duration.sum':
        v = v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
    else:
        v = v*{'byte':1e-6,'Kbyte':1e-3,'Mbyte':1,'Gbyte':1e3}[u]
    agg[row['Kernel Name'].split('(')[0]][m].append(v)
for k,d in agg.items():
    t=d['gpu__time_duration.sum']; print(f"{k:36s} n={len(t)} {sum(t)/len(t):8.1f} us  dram R {sum(d['dram__bytes_read.sum'])/len(t):8.1f} MB  W {sum(d['dram__bytes_write.sum'])/len(t):8.1f} MB")
PY
