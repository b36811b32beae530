cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frames.py -x -q -k "gib or huge or long_literals or edge_inputs or corrupt or frames" 2>&1 | tail -3
timeout 300 python scripts/bench_edge_1gib.py > gpurun_out/r02bi_edge.jsonl 2> gpurun_out/r02bi_edge.err; python - <<PY
import json
for l in open("gpurun_out/r02bi_edge.jsonl"):
    if l.startswith("{"):
        d=json.loads(l); print(d["input"], d["codec"], "enc", d["enc_ms"], "dec", d["dec_ms"], d["dec_roofline"], d["roundtrip_ok"])
PY
