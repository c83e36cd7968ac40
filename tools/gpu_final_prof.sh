# Final round-1 evidence for the 1 M dam break (bench workload): launch list + full capture of the hot kernels.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# launch list of the bench command itself (2 timed steps), skipping the 600 settle steps (7 launches each + 3 upload launches)
ncu --metrics gpu__time_duration.sum --clock-control none -s 4210 -c 120 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_bench_final.log 2>&1
tail -2 gpurun_out/ncu_bench_final.log | cut -c1-300
# full capture, one launch of each step kernel at the bench state
ncu --set full --clock-control none --import-source on -k regex:'k_forces_integrate|k_density|k_scan_exclusive|k_order_gather|k_cell_hist|k_place|k_plan_zero' -s 4207 -c 7 -f -o gpurun_out/prof_r01_final python tools/profile_step.py --settle 601 --steps 3 > gpurun_out/prof_final.log 2>&1
tail -2 gpurun_out/prof_final.log
