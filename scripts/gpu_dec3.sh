cd $GRAFT_REPO_ROOT
timeout 600 python scripts/prof_frame.py rle8_multi,rle64_byte_packed 0 3 dec 2>&1 | tail -2 | cut -c1-330
for c in rle8_multi rle32_byte_packed rle64_byte_packed; do timeout 120 python scripts/prof_one.py $c 3 dec 2>&1 | tail -1 | cut -c1-330; done
export HSRLE_LIB=libhsrle_b200_dbg.so HSRLE_DEBUG=1
timeout 300 python scripts/prof_frame.py rle8_multi 0 1 dec 2>&1 | grep -E "k_dec_emit done" | tail -1 | cut -c1-700
timeout 120 python scripts/prof_one.py rle8_multi 1 dec 2>&1 | grep -E "k_dec_emit done" | tail -1 | cut -c1-600
timeout 120 python scripts/prof_one.py rle64_byte_packed 1 dec 2>&1 | grep -E "k_dec_emit done" | tail -1 | cut -c1-600
