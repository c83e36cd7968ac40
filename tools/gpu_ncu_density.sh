# ncu --set full of the density kernels on the settled 1 M dam break.  gpurun --timeout 900 -- 'bash tools/gpu_ncu_density.sh <tag> [cfg]'
cd $GRAFT_REPO_ROOT
TAG=${1:-r02_pair}
export SPH_B200_DENSITY_CFG=${2:-0}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_density' \
    -s ${3:-1800} -c 3 -f -o gpurun_out/prof_$TAG python tools/profile_step.py --settle 601 --steps 3 > gpurun_out/prof_$TAG.log 2>&1
tail -3 gpurun_out/prof_$TAG.log
