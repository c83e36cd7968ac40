cd $GRAFT_REPO_ROOT
N=${1:-4}
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 300 --warmup 3 --workload dam-break-16M --e2e-steps 0 2> gpurun_out/config2_n$N.err | tee gpurun_out/config2_n$N.json | cut -c1-1500
tail -5 gpurun_out/config2_n$N.err
