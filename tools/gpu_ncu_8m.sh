# ncu --set full of the neighbour passes on the settled 8 M block (config 3 on one GPU).  gpurun --timeout 900 -- 'bash tools/gpu_ncu_8m.sh <tag>'
cd $GRAFT_REPO_ROOT
TAG=${1:-r02_8m}
mkdir -p gpurun_out
# settle 601 steps = 601 launches of each kernel; capture the first timed step's density / forces kernels and one of each grid kernel
ncu --set full --clock-control none --import-source on -k regex:'k_density_staged|k_forces_integrate|k_order_gather|k_cell_hist|k_place|k_scan' \
    -s 3606 -c 6 -f -o gpurun_out/prof_$TAG python tools/profile_step.py --h 0.02 --dims 61 256 512 --settle 601 --steps 3 > gpurun_out/prof_$TAG.log 2>&1
tail -3 gpurun_out/prof_$TAG.log
