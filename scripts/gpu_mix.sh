cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1
HSRLE_DEBUG=1 timeout 120 python scripts/prof_one.py rle8_multi 1 dec 2>&1 | grep -E "phase|kernel us" | tail -2 | cut -c1-300
HSRLE_DEBUG=1 timeout 120 python scripts/prof_one.py rle16_7symlut_byte 1 dec 2>&1 | grep -E "phase|kernel us" | tail -2 | cut -c1-300
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sliced.py -m gpu -x -q -k "golden or fuzz or structured or edge or sliced_shared_gpu_world2" ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
for c in rle8_3symlut rle8_7symlut rle16_7symlut_byte rle8_packed_multi; do timeout 120 python scripts/prof_one.py $c 3 enc 2>&1 | tail -1 | cut -c1-330; done
