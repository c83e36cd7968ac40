# Round 2, first GPU call: packed-arithmetic microbenchmark, parity + edge tests with the pair-walk density kernel,
# A/B sweep of the density variants at 1 M and 8 M.   gpurun --timeout 1200 -- 'bash tools/gpu_r2_first.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -s -C oracle oracle
./tools/ubench/packed_f32x2 > gpurun_out/ubench_packed.jsonl 2>&1; cat gpurun_out/ubench_packed.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -x -q -m gpu 2>&1 | tail -15
timeout 400 python tools/sweep_density.py --big --steps 20 > gpurun_out/sweep_density_a.jsonl 2> gpurun_out/sweep_density_a.err
cat gpurun_out/sweep_density_a.jsonl; tail -3 gpurun_out/sweep_density_a.err
