set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
python bench.py --config trafalgar --solver pcg --steps 10 --warmup 3 --no-cpu-baseline --no-e2e | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['ms_per_step'],d['lm_iters_per_sec'],d['cost_first_last'],d['pcg_iters'],{k:v['avg_ms'] for k,v in d['kernels'].items()})"
timeout 1200 python bench.py --config final --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_final_r01v.json 2> gpurun_out/bench_final_v.err; python -c "
import json;d=json.load(open('gpurun_out/bench_final_r01v.json'));print(d['ms_per_step'],d['lm_iters_per_sec'],d['cost_first_last'],d['pcg_iters'],{k:v['avg_ms'] for k,v in d['kernels'].items()})"; tail -3 gpurun_out/bench_final_v.err
