cd $GRAFT_REPO_ROOT
make -s -C oracle oracle
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -x -q -m gpu 2>&1 | tail -5
echo "--- graphs"; timeout 300 python tools/small_n.py
echo "--- host launches"; SPH_B200_GRAPH=0 timeout 300 python tools/small_n.py
python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('1M: ms/step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['roofline']['pass_ms'].items()}, d['value'], d['gpu_launches'])"
