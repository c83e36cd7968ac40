cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in 0 1 2 3 4 5; do
  echo cfg $cfg
  SPH_B200_DENSITY_CFG=$cfg python tools/profile_step.py --h 0.02 --dims 61 256 512 --settle 600 --steps 50 2>&1 | grep pass
  SPH_B200_DENSITY_CFG=$cfg python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('   1M: ms/step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['roofline']['pass_ms'].items()})"
done
