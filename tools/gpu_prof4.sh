cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_forces_integrate|k_density|k_scan_exclusive|k_order_gather|k_cell_hist|k_place' -s 3606 -c 6 -f -o gpurun_out/prof_r01_8m python tools/profile_step.py --h 0.02 --dims 61 256 512 --settle 600 --steps 3 > gpurun_out/prof4.log 2>&1
tail -2 gpurun_out/prof4.log
