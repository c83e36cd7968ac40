# Full ncu capture of the two neighbour passes at a later, denser state of the 1 M dam break (step ~1000).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
S=${1:-1000}
ncu --set full --clock-control none --import-source on -k regex:'k_forces_integrate|k_density' \
    -s $(( (S + 1) * 3 )) -c 3 -f -o gpurun_out/prof_late python tools/profile_step.py --settle $((S + 1)) --steps 3 > gpurun_out/prof_late.log 2>&1
tail -2 gpurun_out/prof_late.log
