set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_default_r01F.json 2> gpurun_out/bench_F.err
python -c "
import json;d=json.load(open('gpurun_out/bench_default_r01F.json'));print(d['value'],d['ms_per_step'],d['lm_iters_per_sec'],d['e2e'],d['roofline'],d['cpu_baseline']['value'],d['gpu_launches'],d['clocks'])"; tail -3 gpurun_out/bench_F.err
