set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -s -C oracle oracle
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 300 --warmup 5 --cpu-steps 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -5 gpurun_out/bench1.err
cat gpurun_out/bench1.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 20 --warmup 3 --settle 300 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
