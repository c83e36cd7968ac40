# compute-sanitizer over the tests that exercise the heavy / clump kernels.  gpurun --timeout 900 -- 'bash tools/gpu_sanitize.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -s -C oracle oracle
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --target-processes all --print-limit 5 python -m pytest tests/test_gpu_edge.py -x -q -m gpu -k "clump or sums_are_taken" > gpurun_out/sanitize_$tool.txt 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Hazard|Invalid" gpurun_out/sanitize_$tool.txt | head -12
done
