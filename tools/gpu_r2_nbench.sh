# The multi-GPU bench line only (no tests): gpurun --gpus N --timeout 600 -- 'bash tools/gpu_r2_nbench.sh N'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_weak_n${N}_r02.json 2> gpurun_out/bench_weak_n${N}_r02.err
tail -3 gpurun_out/bench_weak_n${N}_r02.err; cut -c1-300 gpurun_out/bench_weak_n${N}_r02.json
