# slab tests on one GPU (ranks share the device, gloo) + quick A/B of the force variants.  gpurun --timeout 1200 -- 'bash tools/gpu_r2_slab.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -s -C oracle oracle
timeout 900 python -m pytest tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/sweep_density.py --big --cfgs 0 --forces 7 2 2>&1 | cut -c1-260
