#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final_default.json 2> gpurun_out/bench_final_default.err; echo "bench rc=$?"
for cfg in final ladybug trafalgar; do
python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err; echo "$cfg rc=$?"
done
python bench.py --model projective --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_proj.json 2> gpurun_out/bench_proj.err; echo "proj rc=$?"
python - <<'PY'
import json
for f in ("bench_final_default","bench_final","bench_ladybug","bench_trafalgar","bench_proj"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), round(d["value"]/1e6,1), "e2e", round(d["e2e"]["value"]/1e6,1) if d.get("e2e") else None, d["pcg_iters"][:6], d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["config"]["solver"][:20])
    except Exception as e: print(f, "ERR", e)
PY
