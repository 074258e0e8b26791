#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
VLG_BA_PERSIST_PROF=1 python bench.py --steps 5 --warmup 4 --no-cpu-baseline > gpurun_out/prof_frag.json 2> gpurun_out/prof_frag.err
grep -A1 k_pcg_persistent gpurun_out/prof_frag.err | tail -2 | cut -c1-1700
for at in 3 0; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --autotune $at > gpurun_out/bench_at$at.json 2> gpurun_out/bench_at$at.err; echo "bench rc=$?"
done
python - <<'PY'
import json
for at in (3,0):
    d=json.loads(open(f"gpurun_out/bench_at{at}.json").read().strip().splitlines()[-1])
    print(at, d["ms_per_step"], d["value"], d["pcg_iters"], {k:v["avg_ms"] for k,v in d["kernels"].items()})
PY
