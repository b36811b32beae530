cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests/test_gpu_sliced.py -x -q -m gpu ) > gpurun_out/pytest_sliced.log 2>&1
tail -30 gpurun_out/pytest_sliced.log
for cfg in "1 1" "2 2" "4 4" "6 6"; do
  set -- $cfg
  timeout 300 python bench.py --steps 3 --warmup 3 --quick --streams $1 --threads $2 > gpurun_out/sweep_$1.json 2> gpurun_out/sweep_$1.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/sweep_$1.json'))
print('streams/threads $1', 'value', d['value'], 'e2e', d['e2e']['value'])
" || tail -3 gpurun_out/sweep_$1.err
done
