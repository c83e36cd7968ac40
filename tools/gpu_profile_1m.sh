# ncu evidence for the bench workload (1 M dam break): launch list of the bench command itself, then a
# full capture of one launch of every step kernel. gpurun --timeout 1500 -- 'bash tools/gpu_profile_1m.sh <tag>'
cd $GRAFT_REPO_ROOT
TAG=${1:-r02}
mkdir -p gpurun_out
python -c "import importlib.util as u; s=u.spec_from_file_location('b','bench.py'); m=u.module_from_spec(s); s.loader.exec_module(m); print(m.source_hash())" > gpurun_out/source_hash.txt
# 9 kernels per step; skip the 600 settle steps (+ the upload's few launches), list the warm-up and timed steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 5405 -c 160 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-weak-base --no-other-configs --e2e-steps 3 > gpurun_out/ncu_bench_$TAG.log 2>&1
tail -2 gpurun_out/ncu_bench_$TAG.log | cut -c1-300
ncu --set full --clock-control none --import-source on \
    -k regex:'k_forces_integrate|k_forces_heavy|k_density|k_scan_exclusive|k_order_gather|k_cell_hist|k_place|k_plan_zero' \
    -s 5409 -c 9 -f -o gpurun_out/prof_$TAG python tools/profile_step.py --settle 601 --steps 3 > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log
