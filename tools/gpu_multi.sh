set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
make -s -C oracle oracle
timeout 600 python -m pytest tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -5
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 3 --settle 100 > gpurun_out/bench_multi_$N.json 2> gpurun_out/bench_multi_$N.err
tail -5 gpurun_out/bench_multi_$N.err
cat gpurun_out/bench_multi_$N.json
