# Clump regime: edge / parity suites, then the long run with and without the tiled clump phase.
#   gpurun --timeout 900 -- 'bash tools/gpu_r2_clump.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -s -C oracle oracle
timeout 400 python -m pytest tests/test_gpu_edge.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_clump.txt 2>&1; tail -15 gpurun_out/pytest_clump.txt
timeout 120 python tools/long_run.py 500 10 > gpurun_out/long_run_clump.txt 2>&1; tail -4 gpurun_out/long_run_clump.txt
SPH_B200_CLUMP_CELL=0 timeout 120 python tools/long_run.py 500 10 > gpurun_out/long_run_noclump.txt 2>&1; tail -4 gpurun_out/long_run_noclump.txt
