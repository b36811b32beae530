cd $GRAFT_REPO_ROOT
for c in rle8_3symlut rle8_7symlut; do timeout 120 python scripts/prof_one.py $c 3 both 2>&1 | tail -1 | cut -c1-420; done
HSRLE_LUTWALK=0 timeout 120 python scripts/prof_one.py rle8_3symlut 3 enc 2>&1 | tail -1 | cut -c1-420
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "rle8_3symlut or rle8_7symlut or rle8" 2>&1 | tail -5
