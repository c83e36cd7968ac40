set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -3
make -s -C oracle oracle
timeout 900 python -m pytest tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -15
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/bench_multi_$N.json 2> gpurun_out/bench_multi_$N.err
tail -5 gpurun_out/bench_multi_$N.err
grep '^{' gpurun_out/bench_multi_$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g ms/step %.4f' % (d['value'], d['ms_per_step'])); print(d['config']['particles_per_gpu'], d['config'].get('halo_message_rows'))"
