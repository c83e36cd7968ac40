cd $GRAFT_REPO_ROOT
make -s -C oracle oracle
timeout 1400 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_edge.py -x -q -m gpu 2>&1 | tail -15
