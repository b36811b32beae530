# quick iteration: a parity subset, per-call timings of a few codecs, short bench.  usage: gpu_iter.sh <tag> [pytest -k expr]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; K=${2:-"golden or fuzz or structured or edge"}
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sliced.py -m gpu -x -q -k "$K" ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.log
for c in rle8_multi rle8_packed_multi rle8_3symlut rle8_7symlut rle16_7symlut_byte rle32_3symlut_byte rle64_byte_packed; do timeout 120 python scripts/prof_one.py $c 3 both 2>&1 | tail -1 | cut -c1-400; done | tee gpurun_out/${TAG}_prof.log
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
