set -x
cd $GRAFT_REPO_ROOT
make -s -C oracle oracle
timeout 900 python -m pytest tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -30
