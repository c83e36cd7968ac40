# Collapsed regime: suites, A/B on one saved state, determinism of two long runs; fluid-regime step times (the heavy
# kernels are launched every step).   gpurun --timeout 900 -- 'bash tools/gpu_r2_clump_ab.sh [ncu]'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -s -C oracle oracle
timeout 400 python -m pytest tests/test_gpu_edge.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_clump.txt 2>&1; tail -15 gpurun_out/pytest_clump.txt
timeout 200 python tools/clump_ab.py > gpurun_out/clump_ab.jsonl 2> gpurun_out/clump_ab.err; cat gpurun_out/clump_ab.jsonl; tail -3 gpurun_out/clump_ab.err
timeout 120 python tools/long_run.py 500 12 > gpurun_out/long_run_det1.txt 2>&1
timeout 120 python tools/long_run.py 500 12 > gpurun_out/long_run_det2.txt 2>&1
tail -3 gpurun_out/long_run_det1.txt
cut -c70- gpurun_out/long_run_det1.txt > /tmp/d1; cut -c70- gpurun_out/long_run_det2.txt > /tmp/d2; cmp /tmp/d1 /tmp/d2 && echo "DETERMINISTIC: statistics of the two runs are identical"
timeout 200 python tools/ab_env.py SPH_B200_CLUMP_CELL 64 > gpurun_out/ab_fluid.jsonl 2> gpurun_out/ab_fluid.err; cut -c1-300 gpurun_out/ab_fluid.jsonl
if [ "$1" = "ncu" ]; then
timeout 300 ncu --set full --import-source on --clock-control none -k regex:heavy -s 6 -c 2 -o gpurun_out/prof_r02_clump -f python tools/clump_ab.py --cells 64 --steps 2 > gpurun_out/prof_r02_clump.log 2>&1; tail -2 gpurun_out/prof_r02_clump.log
fi
