cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "either_way or segment_table or lut_walk" 2>&1 | tail -3
bash scripts/gpu_sanitize.sh r02ba
