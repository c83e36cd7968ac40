# The driver's round-end sequence in miniature: N = 1 then N = 2 (twice) back to back on one box.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline 2> gpurun_out/b2b_1.err | python tools/show_bench.py
for N in ${NS:-2 2}; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 200 --warmup 5 2> gpurun_out/b2b_$N.err | python tools/show_bench.py
done
