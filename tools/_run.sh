#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open(f"gpurun_out/bench_s.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"], d["pcg_iters"], {k:v["avg_ms"] for k,v in d["kernels"].items()})
PY
python tools/step_timeline.py > gpurun_out/timeline.txt 2>&1; tail -3 gpurun_out/timeline.txt
