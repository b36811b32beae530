# usage: gpu_bench_all.sh <tag>: N=1 bench line + per-codec kernel timings of all 44 codecs (88 MB DCT)
cd $GRAFT_REPO_ROOT
tag=${1:-r02}
python bench.py --gpus 1 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 3000 gpurun_out/${tag}_bench.json
python - <<'PY' > gpurun_out/all_codecs.list
import sys
sys.path.insert(0, "tests")
from common import CODECS
print(" ".join(c.name for c in CODECS))
PY
for c in $(cat gpurun_out/all_codecs.list); do timeout 120 python scripts/prof_one.py $c 3 both 2>&1 | tail -1; done > gpurun_out/${tag}_all_codecs.log
python - <<PY
import re
for l in open("gpurun_out/${tag}_all_codecs.log"):
    m = re.match(r"(\S+) n \d+ clen (\d+) enc us ([\d.]+) dec us ([\d.]+) \| kernel us: (.*?) \|", l)
    if m: print(f"{m.group(1):26s} clen {int(m.group(2))>>20:4d}M enc {float(m.group(3)):7.0f} dec {float(m.group(4)):7.0f} | {m.group(5)}")
    else: print(l[:200])
PY
