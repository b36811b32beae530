# usage: bash scripts/gpu_test.sh [debug codec names...]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python tests/gpu_debug.py "$@" ) > gpurun_out/debug.log 2>&1
tail -60 gpurun_out/debug.log
