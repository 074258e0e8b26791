"""Runs tests/mgpu_check.py under torchrun at every world size the box offers (gpurun --gpus 2 / 4 / 8): one whole
problem sharded by point over G ranks against the single-GPU result, teacher-forced, all three solvers, NCCL and
peer-memory exchange.  The driver's 1-GPU test box skips these; bench.py --gpus N repeats the check on its own
workload (`parity` in its JSON line), so the SCALE runs carry correctness as well."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_ranks_match_one_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(29611 + world), os.path.join(ROOT, "tests", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
