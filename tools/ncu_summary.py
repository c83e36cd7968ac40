"""Summarise an .ncu-rep (read here, no GPU needed): per kernel the duration, DRAM bytes, throughput
percentages, occupancy, registers, and the top warp-stall reasons.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xxx.txt]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %peak"),
    ("sm__inst_executed.sum", "warp insts"),
    ("smsp__inst_executed.avg.per_cycle_active", "IPC/SMSP"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active thr/inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed_pipe_fma.sum", "fma-pipe insts"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_fp64.sum", "fp64 insts"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_xu.sum", "xu insts"),
]
stall_cols = [n for n in hdr if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
if not stall_cols:
    stall_cols = [n for n in hdr if "warp_issue_stalled" in n and n.endswith("_per_warp_active.pct")]

for r in data:
    name = r[col["Kernel Name"]].split("(")[0]
    print(f"=== {name}  (id {r[col['ID']]})")
    for key, label in WANT:
        if key in col:
            print(f"  {label:18s} {r[col[key]]:>16s} {units[col[key]]}")
    st = []
    for n in stall_cols:
        try:
            st.append((float(r[col[n]].replace(",", "")), n))
        except ValueError:
            pass
    st.sort(reverse=True)
    print("  top stalls:")
    for v, n in st[:6]:
        short = n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
        short = short.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", "")
        print(f"    {short:28s} {v:8.2f}")
