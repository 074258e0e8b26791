set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
for c in ladybug venice; do
python bench.py --config $c --model projective --steps 10 --warmup 3 > gpurun_out/bench_${c}_projective_r01w.json 2> gpurun_out/bench_w.err
python -c "
import json;d=json.load(open('gpurun_out/bench_${c}_projective_r01w.json'));print(d['ms_per_step'],d['lm_iters_per_sec'],d['e2e'] and d['e2e']['ms_per_step'],d['config']['solver'],d['cost_first_last'],d['pcg_iters'],d['accepted_steps'],{k:v['avg_ms'] for k,v in d['kernels'].items()})"; tail -3 gpurun_out/bench_w.err
done
