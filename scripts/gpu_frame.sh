cd $GRAFT_REPO_ROOT
timeout 600 python scripts/prof_frame.py rle8_multi,rle64_byte,rle64_byte_packed 0 3 both 2>&1 | tail -4 | cut -c1-600
timeout 300 python scripts/bench_edge_1gib.py rle8_multi 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(d['input'], d['codec'], 'enc', d['enc_ms'], 'dec', d['dec_ms'], d['roundtrip_ok'], d['kernel_us'])"
timeout 120 python scripts/prof_one.py rle8_multi 3 both 2>&1 | tail -1 | cut -c1-420
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
