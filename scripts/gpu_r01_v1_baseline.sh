set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r01_pytest_gpu_v1.log 2>&1
tail -3 gpurun_out/r01_pytest_gpu_v1.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r01_bench_v1.json 2> gpurun_out/r01_bench_v1.err
tail -c 600 gpurun_out/r01_bench_v1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 400 --csv --log-file gpurun_out/r01_launches_v1.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_enc_scan_check -s 10 -c 2 -o gpurun_out/r01_prof_v1_scan_check python bench.py --steps 1 --warmup 1 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
