"""Multi-GPU parity check, launched by torchrun (see tests/test_multi_gpu.py):
G ranks, one whole problem sharded by point, against the single-GPU result of rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bundleadjustmentmatlab_b200 import capi, shard, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = synth.make_problem(60, 20000, 90000, seed=2)
    a = np.ascontiguousarray(np.vstack([P.w, P.Te]).T); b = np.ascontiguousarray(P.Xe[:3].T)
    def fresh_uid():
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.Context.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        return bytes(uid.cpu().numpy().tobytes())

    results = {}
    for solver in (capi.SOLVER_PCG, capi.SOLVER_CHOL, capi.SOLVER_PCG_EXPLICIT):
        ctx = capi.Context(num_variableK=0, solver=solver, pcg_rtol=1e-12, device=local)
        ctx.set_comm(rank, world, fresh_uid())
        xy, pt, cam, bl, _ = shard.shard_points(P.obs_xy, P.obs_pt, P.obs_cam, b, rank, world)
        ctx.set_problem_sparse(P.K.T, a, bl, xy, pt, cam)
        infos = [ctx.trial_step() for _ in range(3)]
        results[solver] = infos
        ctx.close()
    if rank == 0:
        ref = {}
        for solver in (capi.SOLVER_PCG, capi.SOLVER_CHOL, capi.SOLVER_PCG_EXPLICIT):
            ctx = capi.Context(num_variableK=0, solver=solver, pcg_rtol=1e-12, device=local)
            ctx.set_problem_sparse(P.K.T, a, b, P.obs_xy, P.obs_pt, P.obs_cam)
            ref[solver] = [ctx.trial_step() for _ in range(3)]
            ctx.close()
        worst = 0.0
        for solver in ref:
            for k in range(3):
                for key in ("old_cost", "new_cost"):
                    d = abs(results[solver][k][key] - ref[solver][k][key]) / ref[solver][k][key]
                    worst = max(worst, d)
                assert results[solver][k]["accepted"] == ref[solver][k]["accepted"]
        print(f"MGPU_OK world={world} worst relative cost deviation vs 1 GPU: {worst:.3e}")
        assert worst <= 1e-9
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
