# quick check: debug parity sweep over all codecs + timing of a few codecs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python tests/gpu_debug.py ) > gpurun_out/debug.log 2>&1
grep -E "BAD|TOTAL|Error|error" gpurun_out/debug.log | head -20
for c in "$@"; do timeout 300 python scripts/prof_one.py $c 5 both 2>&1 | tail -3; done
