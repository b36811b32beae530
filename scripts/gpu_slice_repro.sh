cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29655 scripts/slice_frame_repro.py rle8_multi,rle64_byte_packed 0 2>&1 | grep -E "^rank|Error" | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 3 --master-addr 127.0.0.1 --master-port 29656 scripts/slice_frame_repro.py rle8_multi,rle64_byte_packed 1 2>&1 | grep -E "^rank|Error|Assert" | cut -c1-300
timeout 1200 python -m pytest tests/test_gpu_sliced.py -x -q 2>&1 | tail -3
