# round record: GPU suite, bench (both arms), ncu launch list of the bench command, full captures of the top kernels
# usage: bash scripts/gpu_profile_round.sh <tag>
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
tail -c 400 gpurun_out/${TAG}_bench_reference.json
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 300 gpurun_out/${TAG}_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --quick > gpurun_out/${TAG}_ncu_l.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_l.log
# (what comes back in gpurun_out/ must stay below 64 MiB: source import only for one of the two captures)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ -s 8 -c 8 -f -o gpurun_out/${TAG}_full_rle8_multi python scripts/prof_one.py rle8_multi 0 both > gpurun_out/${TAG}_full_rle8_multi.log 2>&1
tail -1 gpurun_out/${TAG}_full_rle8_multi.log
# (the second capture, rle32_3symlut_byte, is its own call: scripts/gpu_ncu.sh -- two reports do not fit the 64 MiB that come back)
ls -la gpurun_out | tail -12
timeout 600 python scripts/bench_edge_1gib.py > gpurun_out/${TAG}_edge.jsonl 2> gpurun_out/${TAG}_edge.err; grep -c roundtrip_ok gpurun_out/${TAG}_edge.jsonl
