import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from bundleadjustmentmatlab_b200 import capi
from common import load_golden
from test_gpu_parity import ctx_from_golden
g = load_golden("euclid_fullK")
print("ref error_:", g["error_"])
print("golden t_old/t_new/t_accept/t_lam:")
for k in range(len(g["t_lam"])):
    print(k, float(g["t_old"][k]), float(g["t_new"][k]), int(g["t_accept"][k]), float(g["t_lam"][k]))
for solver in (capi.SOLVER_CHOL, capi.SOLVER_PCG):
    ctx = ctx_from_golden(g, solver=solver, pcg_rtol=1e-12)
    print("solver", solver)
    for it in range(12):
        if not ctx.lm_continue():
            break
        i = ctx.trial_step()
        print(it, i["old_cost"], i["new_cost"], i["accepted"], i["lambda_used"], i["rho"], i["pcg_iters"])
    ctx.close()
