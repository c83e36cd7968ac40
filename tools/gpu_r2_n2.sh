# Two GPUs: the NCCL / peer-mailbox slab tests between distinct devices, then the weak-scaling bench at N = 2.
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_r2_n2.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -s -C oracle oracle
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_slab.py -x -q -m gpu -k "nccl" > gpurun_out/pytest_nccl.txt 2>&1; tail -5 gpurun_out/pytest_nccl.txt
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_weak_n${N}_r02.json 2> gpurun_out/bench_weak_n${N}_r02.err
tail -5 gpurun_out/bench_weak_n${N}_r02.err; cat gpurun_out/bench_weak_n${N}_r02.json
