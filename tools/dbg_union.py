"""Debug: the 2-rank weak-scaling problem (two Venice shards sharing the cameras) solved on ONE GPU."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bundleadjustmentmatlab_b200 import capi, synth

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
Ps = [synth.make_config("venice", seed=0, point_seed=r) for r in range(world)]
P = Ps[0]
a0 = np.ascontiguousarray(np.vstack([P.w, P.Te]).T)
b0 = np.ascontiguousarray(np.vstack([p.Xe[:3].T for p in Ps]))
off = np.cumsum([0] + [p.n for p in Ps])
pt = np.concatenate([p.obs_pt.astype(np.int64) + off[r] for r, p in enumerate(Ps)])
cam = np.concatenate([p.obs_cam for p in Ps]).astype(np.int64)
xy = np.concatenate([p.obs_xy for p in Ps])
order = np.lexsort((pt, cam))
ctx = capi.Context(num_variableK=0, solver=capi.SOLVER_PCG_EXPLICIT)
ctx.set_problem_sparse(P.K.T, a0, b0, np.ascontiguousarray(xy[order]), pt[order].astype(np.int32), cam[order].astype(np.int32))
for k in range(3):
    i = ctx.trial_step()
    print(f"union step {k:2d} lam {i['lambda_used']:.3e} old {i['old_cost']:.10e} new {i['new_cost']:.10e} acc {i['accepted']} its {i['pcg_iters']} "
          f"relres {i['pcg_relres']:.2e} denom {i['denom']:.8e}", flush=True)
ctx.close()
