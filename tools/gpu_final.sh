# Final single-GPU evidence of a round: whole GPU suite, smoke, both bench arms with the driver's arguments, ncu launch
# list + full capture.   gpurun --timeout 2400 -- 'bash tools/gpu_final.sh r02'
cd $GRAFT_REPO_ROOT
TAG=${1:-r02}
mkdir -p gpurun_out
make -s -C oracle oracle
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_$TAG.txt
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; tail -2 gpurun_out/bench_ref_n1.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
cut -c1-400 gpurun_out/bench_n1.json
bash tools/gpu_profile_1m.sh $TAG
