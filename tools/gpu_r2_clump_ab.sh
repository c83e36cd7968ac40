# Collapsed regime A/B on one saved state + ncu of the two heavy kernels.  gpurun --timeout 900 -- 'bash tools/gpu_r2_clump_ab.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python tools/clump_ab.py > gpurun_out/clump_ab.jsonl 2> gpurun_out/clump_ab.err; cat gpurun_out/clump_ab.jsonl; tail -3 gpurun_out/clump_ab.err
timeout 300 ncu --set full --import-source on --clock-control none -k regex:heavy -s 6 -c 2 -o gpurun_out/prof_r02_clump -f python tools/clump_ab.py --cells 64 --steps 2 > gpurun_out/prof_r02_clump.log 2>&1; tail -3 gpurun_out/prof_r02_clump.log
