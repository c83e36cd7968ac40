# The two arms the driver runs at N = 1, with its arguments.   gpurun --timeout 1500 -- 'bash tools/gpu_bench.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -s -C oracle oracle
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; tail -2 gpurun_out/bench_ref_n1.err
cat gpurun_out/bench_ref_n1.json | cut -c1-900
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
