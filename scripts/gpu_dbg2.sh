cd $GRAFT_REPO_ROOT
export HSRLE_LIB=libhsrle_b200_dbg.so HSRLE_DEBUG=1
timeout 120 python scripts/prof_one.py rle64_byte_packed 1 dec 2>&1 | grep -E "k_dec_emit done" | tail -1 | cut -c1-700
timeout 300 python scripts/prof_frame.py rle8_multi 0 1 dec 2>&1 | grep -E "k_dec_emit done" | tail -1 | cut -c1-700
timeout 300 python scripts/prof_frame.py rle64_byte_packed 0 1 dec 2>&1 | grep -E "k_dec_emit done" | tail -1 | cut -c1-700
