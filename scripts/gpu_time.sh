# per-codec kernel timings only: bash scripts/gpu_time.sh codec...
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in "$@"; do timeout 300 python scripts/prof_one.py $c 5 both 2>&1 | tail -2; done | tee gpurun_out/time.log
