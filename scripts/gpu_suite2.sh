# GPU parity suite + bench (single-stream and multi-stream), r01 session 5
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --streams 1 --threads 1 > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err
tail -3 gpurun_out/bench_s1.err
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_s1.json", "gpurun_out/bench.json"):
    try:
        d = json.load(open(f))
        print(f, "value", d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"])
        print("kernel_ms", d["roofline"]["kernel_ms"])
    except Exception as e:
        print("bench parse failed", f, e)
PY
