# quick iteration: parity tests + short bench (no cpu baseline)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -s -C oracle oracle
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 5 > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_iter.json'))
print('value %.4g  ms/step %.4f  warm %.4g  e2e %.4g' % (d['value'], d['ms_per_step'], d['config']['value_l2_warm'], d['e2e']['value']))
print('pass_ms', d['roofline']['pass_ms'])
print('nbr', d['config']['neighbours_mean'], d['config']['neighbours_max'], 'grid', d['config']['grid_dim'])
PY
