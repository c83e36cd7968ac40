# Quick regression + A/B while working on a kernel.  gpurun --timeout 900 -- 'bash tools/gpu_r2_quick.sh "<sweep args>"'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -s -C oracle oracle
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -x -q -m gpu 2>&1 | tail -6
timeout 400 python tools/sweep_density.py $1 > gpurun_out/sweep_density.jsonl 2> gpurun_out/sweep_density.err
cat gpurun_out/sweep_density.jsonl; tail -3 gpurun_out/sweep_density.err
