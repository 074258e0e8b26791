#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_4gpu_J.json 2> gpurun_out/bench_4gpu_J.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_4gpu_J.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["pcg_iters"], d.get("accepted_steps"), {k:v["avg_ms"] for k,v in d["kernels"].items()})
PY
