#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
VLG_BA_PERSIST_PROF=1 python bench.py --steps 5 --warmup 4 --no-cpu-baseline > gpurun_out/prof_z.json 2> gpurun_out/prof_z.err
grep k_pcg_persistent gpurun_out/prof_z.err | tail -1 | cut -c1-700
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open(f"gpurun_out/bench_z.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["pcg_iters"], {k:v["avg_ms"] for k,v in d["kernels"].items()})
PY
