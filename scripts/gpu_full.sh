# usage: gpu_full.sh <tag>: whole GPU suite, N=1 bench line (configs[1]), configs[3] at N=1 on 4 frames, frame / codec timings
cd $GRAFT_REPO_ROOT
tag=${1:-r02}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_gpu.log
python bench.py --gpus 1 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['kernel_ms'])
PY
python bench.py --gpus 1 --workload configs3 --frames 4 --no-slices > gpurun_out/${tag}_c3.json 2> gpurun_out/${tag}_c3.err; python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_c3.json").read().strip().splitlines()[-1])
print("configs3 N=1 4 frames:", {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
timeout 600 python scripts/prof_frame.py rle8_multi,rle64_byte,rle64_byte_packed 0 3 both 2>&1 | tail -3 | cut -c1-330
