cd $GRAFT_REPO_ROOT
timeout 600 python scripts/prof_frame.py rle8_multi,rle64_byte_packed 0 3 both 2>&1 | tail -2 | cut -c1-600
for c in rle8_multi rle32_byte_packed rle48_byte_packed rle64_byte_packed rle64_3symlut_byte; do timeout 120 python scripts/prof_one.py $c 3 dec 2>&1 | tail -1 | cut -c1-330; done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
