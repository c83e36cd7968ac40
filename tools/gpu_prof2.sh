set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_forces|k_density' -s 1604 -c 2 -f -o gpurun_out/prof_r01_v2 python tools/profile_step.py --settle 800 --steps 4 > gpurun_out/prof2.log 2>&1
tail -3 gpurun_out/prof2.log
