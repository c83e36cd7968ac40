# Whole GPU suite on one B200 (the slab tests run their ranks on the same GPU over gloo, and over NCCL
# when two GPUs are visible): gpurun --timeout 1500 -- 'bash tools/gpu_tests.sh'
cd $GRAFT_REPO_ROOT
make -s -C oracle oracle
timeout 1400 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
