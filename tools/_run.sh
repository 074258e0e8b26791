N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-p2p > gpurun_out/bench_venice_${N}gpu_r01u.json 2> gpurun_out/bench_${N}gpu_u.err
echo rc=$?
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_venice_${N}gpu_r01u.json') if l.startswith('{')][0]);print(d['n_gpus'],d['value'],d['ms_per_step'],d['lm_iters_per_sec'],0,d['config'].get('pcg_vector_allreduce'),d['pcg_iters'],{k:v['avg_ms'] for k,v in d['kernels'].items()})"; tail -3 gpurun_out/bench_${N}gpu_u.err
