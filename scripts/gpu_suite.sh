# usage: gpu_suite.sh <tag>: the whole GPU suite on the in-tree library
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${1:-r02}_pytest_gpu.log
