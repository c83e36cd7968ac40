set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# each step launches 10 kernels; after upload: reset_bbox + bbox + import = 3 launches. settle 600 steps = 6000 launches.
ncu --set full --clock-control none --import-source on -k regex:'k_forces|k_density|k_integrate|k_cell_hist|k_gather_sorted|k_stable_order|k_scan_exclusive' -s 4213 -c 7 -f -o gpurun_out/prof_r01_v1 python tools/profile_step.py --settle 600 --steps 3 > gpurun_out/prof1.log 2>&1
tail -5 gpurun_out/prof1.log
ls -la gpurun_out
