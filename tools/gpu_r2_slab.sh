# slab tests on one GPU (ranks share the device, gloo) + N=1 bench short.  gpurun --timeout 1200 -- 'bash tools/gpu_r2_slab.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -s -C oracle oracle
timeout 900 python -m pytest tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1_r02a.json 2> gpurun_out/bench_n1_r02a.err; tail -3 gpurun_out/bench_n1_r02a.err; cat gpurun_out/bench_n1_r02a.json
