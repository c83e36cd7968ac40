# diagnostic: weak scaling at N GPUs, short: edge scans off / on
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
for edge in 0 1; do
for steps in 20 200; do
SPH_B200_EDGE_SCAN=$edge timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps $steps --warmup 5 --no-weak-base --no-slab-parity --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('edge $edge steps $steps ms', round(d['ms_per_step'],4), d['config']['phase_ms_per_step']['max_over_ranks'])
"
done; done
