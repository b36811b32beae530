# bench at several stream counts: bash scripts/gpu_bench2.sh 4 8 ...
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for s in "$@"; do
  timeout 600 python bench.py --steps 3 --warmup 3 --streams $s > gpurun_out/bench_s$s.json 2> gpurun_out/bench_s$s.err
  tail -2 gpurun_out/bench_s$s.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_s$s.json"))
print("streams $s value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
print("kernel_ms", d["roofline"]["kernel_ms"])
PY
done
