# usage: gpu_round.sh <tag>   -- timing of the hot codecs, 1 GiB edge inputs, then the whole GPU suite
cd $GRAFT_REPO_ROOT
tag=${1:-r02}
for c in rle8_3symlut rle8_7symlut rle8_multi rle64_byte_packed rle16_3symlut_byte; do timeout 120 python scripts/prof_one.py $c 3 both 2>&1 | tail -1 | cut -c1-420; done
timeout 600 python scripts/bench_edge_1gib.py > gpurun_out/${tag}_edge.jsonl 2> gpurun_out/${tag}_edge.err; python - <<PY
import json
for l in open("gpurun_out/${tag}_edge.jsonl"):
    try: d=json.loads(l)
    except Exception: continue
    print(d.get("input"), d.get("codec"), "enc", d.get("enc_ms"), "dec", d.get("dec_ms"), d.get("roundtrip_ok"), d.get("kernel_us"))
PY
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/${tag}_pytest_gpu.log
