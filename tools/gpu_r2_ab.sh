# A/B of a library switch plus the parity / edge suites.  gpurun --timeout 900 -- 'bash tools/gpu_r2_ab.sh SPH_B200_PDL 0 1 --big'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -s -C oracle oracle
timeout 300 python tools/ab_env.py "$@" > gpurun_out/ab_$1.jsonl 2> gpurun_out/ab_$1.err
cat gpurun_out/ab_$1.jsonl; tail -3 gpurun_out/ab_$1.err
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -6
