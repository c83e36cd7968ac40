set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
python bench.py --gpus 1 --workload weak --steps 100 --warmup 3 > gpurun_out/bench_weak_1.json 2> gpurun_out/bench_weak_1.err
tail -3 gpurun_out/bench_weak_1.err; cat gpurun_out/bench_weak_1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/bench_multi_$N.json 2> gpurun_out/bench_multi_$N.err
tail -5 gpurun_out/bench_multi_$N.err
cat gpurun_out/bench_multi_$N.json
