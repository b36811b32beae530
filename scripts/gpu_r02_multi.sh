# multi-GPU record: sliced NCCL parity tests + bench.py --gpus N (configs[3] sharding), both arms.  usage: gpu_r02_multi.sh <N> <tag>
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1; TAG=$2
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv
( time timeout 1500 python -m pytest tests/test_gpu_sliced.py tests/test_gpu_parity.py -m gpu -x -v -k "sliced_nccl or two_devices" ) > gpurun_out/${TAG}_pytest_multi.log 2>&1
grep -E "PASSED|FAILED|passed|failed" gpurun_out/${TAG}_pytest_multi.log | tail -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
tail -c 2500 gpurun_out/${TAG}_bench_n$N.json; tail -5 gpurun_out/${TAG}_bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_n$N.json 2> gpurun_out/${TAG}_bench_ref_n$N.err
tail -c 1200 gpurun_out/${TAG}_bench_ref_n$N.json; tail -3 gpurun_out/${TAG}_bench_ref_n$N.err
