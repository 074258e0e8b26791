set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_check.py > gpurun_out/mgpu.log 2>&1
grep -E "MGPU|Error|error|assert" gpurun_out/mgpu.log | head -5
for mode in "" "--no-p2p"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $mode > gpurun_out/bench_2gpu_p2p$mode.json 2> gpurun_out/bench_2gpu.err
python -c "
import json,sys;d=json.loads([l for l in open('gpurun_out/bench_2gpu_p2p$mode.json') if l.startswith('{')][0]);print(d['ms_per_step'],d['lm_iters_per_sec'],d['config'].get('pcg_vector_allreduce'),d['pcg_iters'],{k:v['avg_ms'] for k,v in d['kernels'].items()})"; tail -2 gpurun_out/bench_2gpu.err
done
