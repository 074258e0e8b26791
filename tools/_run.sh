#!/bin/bash
mkdir -p gpurun_out
VLG_BA_PERSIST=0 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "teacher or venice or cluster or autotuned or deflation or medium or large" 2>&1 | tail -2
VLG_BA_OVERLAP=0 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "teacher or venice or cluster or autotuned" 2>&1 | tail -2
