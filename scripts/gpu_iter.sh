# one iteration on the GPU box: debug parity sweep -> (if clean) GPU test suite -> per-codec kernel timing -> bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python tests/gpu_debug.py ) > gpurun_out/debug.log 2>&1
grep -E "BAD|TOTAL|Error|error" gpurun_out/debug.log | head -20
if grep -q "TOTAL BAD 0" gpurun_out/debug.log; then
  ( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
  tail -5 gpurun_out/pytest_gpu.log
fi
for c in "$@"; do timeout 300 python scripts/prof_one.py $c 5 both 2>&1 | tail -2; done
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("value", d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print("kernel_ms", d["roofline"]["kernel_ms"])
except Exception as e:
    print("bench parse failed", e)
PY
