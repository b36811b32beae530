# r02 first check: new at-size parity tests + bench N=1 (both arms)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
nproc; free -g | head -2
( time timeout 2400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frames.py -m gpu -x -q -k "full_size or one_gib or four_gib or frame_size or two_devices or corrupt or frame_codec" ) > gpurun_out/r02a_pytest.log 2>&1
tail -15 gpurun_out/r02a_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -c 600 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err
tail -c 900 gpurun_out/r02a_bench_ref.json
