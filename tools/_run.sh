set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default_r01o.json 2> gpurun_out/bench_o.err; tail -c 600 gpurun_out/bench_default_r01o.json; tail -3 gpurun_out/bench_o.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_r01o.json 2> gpurun_out/bench_ref_o.err; cat gpurun_out/bench_reference_r01o.json; tail -3 gpurun_out/bench_ref_o.err
timeout 1200 python bench.py --config final --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_final_r01o.json 2> gpurun_out/bench_final_o.err; tail -c 1500 gpurun_out/bench_final_r01o.json; tail -3 gpurun_out/bench_final_o.err
