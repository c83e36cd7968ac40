cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 3 > gpurun_out/bench_multi_$N.json 2> gpurun_out/bench_multi_$N.err
tail -3 gpurun_out/bench_multi_$N.err
cat gpurun_out/bench_multi_$N.json | head -c 3000
