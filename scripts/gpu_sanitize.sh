# usage: gpu_sanitize.sh <tag>
cd $GRAFT_REPO_ROOT
tag=${1:-r02}
for tool in memcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool python scripts/sanitize_small.py 400000 > gpurun_out/${tag}_san_$tool.log 2>&1
  HSRLE_DEC_MODE=segtab timeout 1500 compute-sanitizer --tool $tool python scripts/sanitize_small.py 300000 rle8_multi,rle16_7symlut_byte,rle64_byte_packed > gpurun_out/${tag}_san_${tool}_segtab.log 2>&1
  echo "$tool (segment tables forced): $(grep -E "ERROR SUMMARY" gpurun_out/${tag}_san_${tool}_segtab.log | tail -1) | round trips: $(grep -c True gpurun_out/${tag}_san_${tool}_segtab.log) ok"
  echo "$tool: $(grep -E "ERROR SUMMARY" gpurun_out/${tag}_san_$tool.log | tail -1) | round trips: $(grep -c True gpurun_out/${tag}_san_$tool.log) ok, $(grep -c False gpurun_out/${tag}_san_$tool.log) bad"
done
timeout 1500 compute-sanitizer --tool racecheck python scripts/sanitize_small.py 200000 rle8_multi,rle8_7symlut,rle16_7symlut_byte,rle64_byte_packed > gpurun_out/${tag}_san_racecheck.log 2>&1
echo "racecheck: $(grep -E "RACECHECK SUMMARY" gpurun_out/${tag}_san_racecheck.log | tail -1) | round trips: $(grep -c True gpurun_out/${tag}_san_racecheck.log) ok"
timeout 1500 compute-sanitizer --tool initcheck python scripts/sanitize_small.py 300000 rle8_multi,rle8_3symlut,rle16_7symlut_byte,rle32_byte_packed > gpurun_out/${tag}_san_initcheck.log 2>&1
echo "initcheck: $(grep -E "ERROR SUMMARY" gpurun_out/${tag}_san_initcheck.log | tail -1) | round trips: $(grep -c True gpurun_out/${tag}_san_initcheck.log) ok"
