set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
SLAB_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/bench_multi_prof_$N.json 2> gpurun_out/bench_multi_prof_$N.err
tail -3 gpurun_out/bench_multi_prof_$N.err
python -c "
import json; d=json.load(open('gpurun_out/bench_multi_prof_$N.json')); print(d['ms_per_step']); print(d['config']['phase_ms_per_step_rank0'])"
