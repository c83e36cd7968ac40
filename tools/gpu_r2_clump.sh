# Clump regime: edge / parity suites, the long run with (and, with an argument, without) the tiled clump phase, and the
# density launch shapes at 1 M / 8 M.   gpurun --timeout 900 -- 'bash tools/gpu_r2_clump.sh [ab]'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -s -C oracle oracle
timeout 400 python -m pytest tests/test_gpu_edge.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_clump.txt 2>&1; tail -15 gpurun_out/pytest_clump.txt
timeout 120 python tools/long_run.py 500 10 > gpurun_out/long_run_clump.txt 2>&1; tail -4 gpurun_out/long_run_clump.txt
if [ "$1" = "ab" ]; then
SPH_B200_CLUMP_CELL=0 timeout 120 python tools/long_run.py 500 10 > gpurun_out/long_run_noclump.txt 2>&1; tail -4 gpurun_out/long_run_noclump.txt
fi
timeout 300 python tools/sweep_density.py --cfgs 10 4 3 0 --big > gpurun_out/sweep_density_shapes.jsonl 2> gpurun_out/sweep_density_shapes.err; cut -c1-330 gpurun_out/sweep_density_shapes.jsonl; tail -3 gpurun_out/sweep_density_shapes.err
