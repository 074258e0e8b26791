import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bundleadjustmentmatlab_b200 import capi, synth
P = synth.make_config("venice", seed=0)
a = np.ascontiguousarray(np.vstack([P.w, P.Te]).T); b = np.ascontiguousarray(P.Xe[:3].T)
for defl in (1, 0):
    ctx = capi.Context(num_variableK=0, pcg_deflate=defl)
    ctx.set_problem_sparse(P.K.T, a, b, P.obs_xy, P.obs_pt, P.obs_cam)
    rows = []
    for _ in range(8):
        if not ctx.lm_continue():
            ctx.lm_reset(a, b)
        t0 = time.perf_counter(); i = ctx.trial_step(); w = (time.perf_counter() - t0) * 1e3
        rows.append((round(w, 2), i["pcg_iters"], float("%.3g" % i["lambda_used"]), i["accepted"], i["new_cost"]))
    print("deflate", defl, rows)
    ctx.close()
