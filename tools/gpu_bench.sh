cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; tail -2 gpurun_out/bench_ref_n1.err
cat gpurun_out/bench_ref_n1.json | cut -c1-700
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
