# full GPU suite (single GPU) + short bench.  usage: gpu_suite.sh <tag>
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -6 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --quick > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench.json"))
    print("value",d["value"],"ms/step",d["ms_per_step"],"e2e",d["e2e"]["value"],"launches",d["gpu_launches"])
    print(d["roofline"]["kernel_ms"])
except Exception as e:
    print("bench failed",e); print(open("gpurun_out/${TAG}_bench.err").read()[-2000:])
PY
