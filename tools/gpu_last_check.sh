# One test on the GPU box (what is left of a round's budget).  gpurun --timeout 120 -- 'bash tools/gpu_last_check.sh <pytest -k expression>'
cd $GRAFT_REPO_ROOT
timeout 90 python -m pytest tests/test_gpu_edge.py -x -q -m gpu -k "${1:-sums_are}" 2>&1 | tail -12 | cut -c1-400
