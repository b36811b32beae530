# e2e vs host thread count: bash scripts/gpu_bench3.sh 4 8 ...
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for s in "$@"; do
  timeout 600 python bench.py --steps 3 --warmup 3 --threads $s > gpurun_out/bench_t$s.json 2> gpurun_out/bench_t$s.err
  tail -2 gpurun_out/bench_t$s.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_t$s.json"))
print("threads $s value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"])
PY
done
nproc; lscpu | grep -E "Model name|Socket|Core|Thread" 
