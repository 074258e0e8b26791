timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_check.py > gpurun_out/mgpu.log 2>&1
grep -E "MGPU|Error|error|assert" gpurun_out/mgpu.log | head
