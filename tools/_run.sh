#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log; grep "autotuned cut" gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final_default.json 2> gpurun_out/bench_final_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_final_default.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],4), round(d["value"]/1e6,1), "e2e", round(d["e2e"]["value"]/1e6,1), d["pcg_iters"][:6], d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["gpu_launches"], d["clocks"])
PY
