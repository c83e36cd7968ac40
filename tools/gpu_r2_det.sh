# Determinism of the collapsed regime (two runs, same build) + edge / parity suites + density launch shapes.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -s -C oracle oracle
timeout 400 python -m pytest tests/test_gpu_edge.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_clump.txt 2>&1; tail -15 gpurun_out/pytest_clump.txt
timeout 120 python tools/long_run.py 500 10 > gpurun_out/long_run_det1.txt 2>&1
timeout 120 python tools/long_run.py 500 10 > gpurun_out/long_run_det2.txt 2>&1
tail -3 gpurun_out/long_run_det1.txt; cut -c1-30,75-220 gpurun_out/long_run_det2.txt | tail -3
cut -c70- gpurun_out/long_run_det1.txt > /tmp/d1; cut -c70- gpurun_out/long_run_det2.txt > /tmp/d2; cmp /tmp/d1 /tmp/d2 && echo "DETERMINISTIC: statistics of the two runs are identical"
timeout 300 python tools/sweep_density.py --cfgs 10 4 3 0 --big > gpurun_out/sweep_density_shapes.jsonl 2> gpurun_out/sweep_density_shapes.err; cut -c1-330 gpurun_out/sweep_density_shapes.jsonl; tail -3 gpurun_out/sweep_density_shapes.err
