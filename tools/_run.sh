set -x
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
python bench.py --config venice --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_venice_r01k.json 2> gpurun_out/bench_k.err
tail -c 600 gpurun_out/bench_venice_r01k.json; tail -5 gpurun_out/bench_k.err
