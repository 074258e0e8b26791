N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_venice_${N}gpu_r01E.json 2> gpurun_out/bench_${N}gpu_E.err
echo rc=$?
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_venice_${N}gpu_r01E.json') if l.startswith('{')][0]);print(d['n_gpus'],d['value'],d['ms_per_step'],d['lm_iters_per_sec'],d['e2e']['value'],d['config'].get('pcg_vector_allreduce'),d['pcg_iters'],{k:(v['avg_ms'],v['count']) for k,v in d['kernels'].items()})"; tail -3 gpurun_out/bench_${N}gpu_E.err
