"""Debug: free-running LM steps at Venice shape on N ranks, every step printed (rank 0)."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bundleadjustmentmatlab_b200 import capi, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P = synth.make_config("venice", seed=0, point_seed=rank)
a0 = np.ascontiguousarray(np.vstack([P.w, P.Te]).T); b0 = np.ascontiguousarray(P.Xe[:3].T)
for at in [int(x) for x in sys.argv[1:]] or [0, 3]:
    ctx = capi.Context(num_variableK=0, device=local, pcg_autotune=at)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.Context.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.set_comm(rank, world, bytes(uid.cpu().numpy().tobytes()))
    ctx.set_problem_sparse(P.K.T, a0, b0, P.obs_xy, P.obs_pt, P.obs_cam)
    if world > 1 and not os.environ.get('DBG_NOP2P'):
        mine = torch.frombuffer(bytearray(ctx.p2p_export()), dtype=torch.uint8).cuda()
        allh = [torch.zeros(capi.P2P_HANDLE_BYTES, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(allh, mine)
        ctx.p2p_import(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
    for k in range(int(os.environ.get('DBG_STEPS', '14'))):
        cont = ctx.lm_continue()
        if not cont:
            ctx.lm_reset(a0, b0)
        i = ctx.trial_step()
        if rank == 0:
            print(f"at={at} step {k:2d} {'RESET ' if not cont else '      '}lam {i['lambda_used']:.3e} old {i['old_cost']:.10e} new {i['new_cost']:.10e} "
                  f"acc {i['accepted']} its {i['pcg_iters']} relres {i['pcg_relres']:.2e} denom {i['denom']:.8e}", flush=True)
    ctx.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
