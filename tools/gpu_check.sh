cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -s -C oracle oracle
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -x -q -m gpu 2>&1 | tail -3
python tools/profile_step.py --h 0.02 --dims 61 256 512 --settle 600 --steps 50 2>&1 | grep pass
python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('1M: ms/step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['roofline']['pass_ms'].items()})"
