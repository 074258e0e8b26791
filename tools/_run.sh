set -x
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
for c in ladybug trafalgar; do
timeout 300 python bench.py --config $c --solver chol --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_${c}_chol3.json 2>gpurun_out/b_${c}_chol3.err
python -c "
import json;d=json.load(open('gpurun_out/b_${c}_chol3.json'));print(d['ms_per_step'],d['lm_iters_per_sec'],d['cost_first_last'],{k:v['avg_ms'] for k,v in d['kernels'].items()})"; tail -3 gpurun_out/b_${c}_chol3.err
done
