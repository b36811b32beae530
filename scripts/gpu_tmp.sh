cd $GRAFT_REPO_ROOT
python bench.py --gpus 1 --workload configs3 --frames 4 --no-slices --no-e2e 2> gpurun_out/tmp_c3.err | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('configs3 4 frames N=1: value', d['value'], 'ms', d['ms_per_step'], 'streams', d['config']['streams'])"; tail -3 gpurun_out/tmp_c3.err
python bench.py --gpus 1 --workload configs3 --no-slices --no-e2e 2> gpurun_out/tmp_c3b.err | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('configs3 16 frames N=1: value', d['value'], 'ms', d['ms_per_step'], 'streams', d['config']['streams'])"; tail -3 gpurun_out/tmp_c3b.err; nvidia-smi --query-gpu=memory.used --format=csv
