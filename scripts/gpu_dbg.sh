cd $GRAFT_REPO_ROOT
timeout 600 python scripts/prof_frame.py rle8_multi,rle64_byte_packed 0 3 both 2>&1 | tail -2 | cut -c1-600
timeout 120 python scripts/prof_one.py rle8_multi 3 both 2>&1 | tail -1 | cut -c1-420
timeout 120 python scripts/prof_one.py rle64_byte_packed 3 both 2>&1 | tail -1 | cut -c1-420
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
export HSRLE_LIB=libhsrle_b200_dbg.so HSRLE_DEBUG=1
for c in rle8_multi rle64_byte_packed; do timeout 300 python scripts/prof_frame.py $c 0 1 dec 2>&1 | grep -E "k_dec_emit done" | tail -1 | cut -c1-700; done
timeout 120 python scripts/prof_one.py rle8_multi 1 dec 2>&1 | grep -E "k_dec_emit done" | tail -1 | cut -c1-600
timeout 120 python scripts/prof_one.py rle64_byte_packed 1 dec 2>&1 | grep -E "k_dec_emit done" | tail -1 | cut -c1-600
