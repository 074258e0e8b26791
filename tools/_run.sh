set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 300 python bench.py --config venice --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_venice_r01D.json 2> gpurun_out/bench_D.err
python -c "
import json;d=json.load(open('gpurun_out/bench_venice_r01D.json'));print(d['ms_per_step'],d['lm_iters_per_sec'],d['e2e']['ms_per_step'],d['cost_first_last'],d['pcg_iters'],{k:v['avg_ms'] for k,v in d['kernels'].items()})"; tail -3 gpurun_out/bench_D.err
