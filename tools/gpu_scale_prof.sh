cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-4}
SLAB_PROFILE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 3 > gpurun_out/bench_multi_prof_$N.json 2> gpurun_out/bench_multi_prof_$N.err
grep "phase ms" gpurun_out/bench_multi_prof_$N.err
grep '^{' gpurun_out/bench_multi_prof_$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g ms/step %.4f' % (d['value'], d['ms_per_step']))"
