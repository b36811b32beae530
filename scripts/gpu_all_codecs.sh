# per-call timings of all 44 codecs on the 88 MB DCT stream -> gpurun_out/all_codecs.log
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/all_codecs.list
import sys, os
sys.path.insert(0, "tests")
from common import CODECS
print(" ".join(c.name for c in CODECS))
PY
for c in $(cat gpurun_out/all_codecs.list); do timeout 120 python scripts/prof_one.py $c 3 both 2>&1 | tail -1; done | tee gpurun_out/all_codecs.log | cut -c1-200
