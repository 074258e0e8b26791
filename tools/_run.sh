set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
VLG_BA_RING=1 python bench.py --config venice --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_venice_r01h_ring.json 2> gpurun_out/bench_h.err
VLG_BA_RING=0 python bench.py --config venice --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_venice_r01h_noring.json 2> gpurun_out/bench_h0.err
tail -c 1500 gpurun_out/bench_venice_r01h_ring.json
tail -c 1500 gpurun_out/bench_venice_r01h_noring.json
