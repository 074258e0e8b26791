#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
