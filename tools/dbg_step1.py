"""Debug: after one free-running accepted step, the second step stage by stage -- 2 ranks (torchrun) or the union on 1 GPU."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bundleadjustmentmatlab_b200 import capi, synth
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
NS = 2
Ps = [synth.make_config("venice", seed=0, point_seed=r) for r in range(NS)]
a0 = np.ascontiguousarray(np.vstack([Ps[0].w, Ps[0].Te]).T)
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = Ps[rank]
    b0 = np.ascontiguousarray(P.Xe[:3].T)
    ctx = capi.Context(num_variableK=0, device=local, solver=capi.SOLVER_PCG_EXPLICIT)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.Context.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ctx.set_comm(rank, world, bytes(uid.cpu().numpy().tobytes()))
    ctx.set_problem_sparse(P.K.T, a0, b0, P.obs_xy, P.obs_pt, P.obs_cam)
    if not os.environ.get("DBG_NOP2P"):
        mine = torch.frombuffer(bytearray(ctx.p2p_export()), dtype=torch.uint8).cuda()
        allh = [torch.zeros(capi.P2P_HANDLE_BYTES, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(allh, mine)
        ctx.p2p_import(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
    shards = [(rank, slice(0, P.n))]
else:
    off = np.cumsum([0] + [p.n for p in Ps])
    b0 = np.ascontiguousarray(np.vstack([p.Xe[:3].T for p in Ps]))
    pt = np.concatenate([p.obs_pt.astype(np.int64) + off[r] for r, p in enumerate(Ps)])
    cam = np.concatenate([p.obs_cam for p in Ps]).astype(np.int64)
    xy = np.concatenate([p.obs_xy for p in Ps])
    order = np.lexsort((pt, cam))
    ctx = capi.Context(num_variableK=0, solver=capi.SOLVER_PCG_EXPLICIT)
    ctx.set_problem_sparse(Ps[0].K.T, a0, b0, np.ascontiguousarray(xy[order]), pt[order].astype(np.int32), cam[order].astype(np.int32))
    shards = [(r, slice(off[r], off[r + 1])) for r in range(NS)]
tag = f"[w{world} r{rank}]"
i = ctx.trial_step()
print(tag, "step0", f"old {i['old_cost']:.10e} new {i['new_cost']:.10e} acc {i['accepted']} lam_next {i['lambda_next']:.4e}", flush=True)
st = ctx.get_state()
print(tag, "state a sum %.12e" % st["a"].sum(), *["b[%d] sum %.12e" % (r, st["b"][sl].sum()) for r, sl in shards], flush=True)
lam = st["lam"]
c1 = ctx.stage1()
blk = ctx.get_blocks(want_W=False)
print(tag, "stage1 cost %.10e U %.10e eA %.10e" % (c1, np.abs(blk["U"]).sum(), np.abs(blk["eA"]).sum()),
      *["V[%d] %.10e eB[%d] %.10e" % (r, np.abs(blk["V"][sl]).sum(), r, np.abs(blk["eB"][sl]).sum()) for r, sl in shards], flush=True)
ctx.stage2(lam)
red = ctx.get_reduced()
print(tag, "stage2 e_ %.12e da %.12e da[6:9] %s" % (np.abs(red["e_"]).sum(), np.abs(red["da"]).sum(), red["da"][6:9]),
      *["Vinv[%d] %.10e" % (r, np.abs(red["Vinv"][sl]).sum()) for r, sl in shards], flush=True)
nc, dn = ctx.stage3(lam)
up = ctx.get_update()
print(tag, "stage3 new %.10e denom %.8e a_new %.12e" % (nc, dn, up["a_new"].sum()),
      *["db[%d] %.10e b_new[%d] %.12e" % (r, np.abs(up["db"][sl]).sum(), r, up["b_new"][sl].sum()) for r, sl in shards], flush=True)
ctx.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
