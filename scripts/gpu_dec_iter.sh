# decoder iteration: debug cases, then per-call timings.  usage: gpu_dec_iter.sh <tag>
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1
HSRLE_DEBUG=1 timeout 200 python scripts/dec_debug.py > gpurun_out/${TAG}_dec_debug.log 2>&1
grep -v "hsrle\] k_dec" gpurun_out/${TAG}_dec_debug.log | grep -v "^ok" | tail -20
for c in rle8_multi rle8_packed_multi rle8_3symlut rle16_7symlut_byte rle32_3symlut_byte rle48_byte_packed rle64_byte_packed rle64_7symlut_byte; do timeout 120 python scripts/prof_one.py $c 3 dec 2>&1 | tail -1 | cut -c1-300; done | tee gpurun_out/${TAG}_prof.log
