# full GPU parity suite + bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("value", d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print("kernel_ms", d["roofline"]["kernel_ms"])
    for k, v in d["per_codec"].items():
        print(k, v)
except Exception as e:
    print("bench parse failed", e)
PY
