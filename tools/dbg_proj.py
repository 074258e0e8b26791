import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from bundleadjustmentmatlab_b200 import capi
from common import load_golden
from test_projective import ctx_from_proj_golden
g = load_golden("proj_fixstructure")
print("ref error_:", g["error_"])
for k in range(len(g["t_lam"])):
    print("ref", k, float(g["t_old"][k]), float(g["t_new"][k]), int(g["t_accept"][k]), float(g["t_lam"][k]))
ctx = ctx_from_proj_golden(g, solver=capi.SOLVER_CHOL)
for it in range(12):
    if not ctx.lm_continue():
        break
    i = ctx.trial_step()
    print("gpu", it, i["old_cost"], i["new_cost"], i["accepted"], i["lambda_used"])
st = ctx.get_state()
print("db-check: b unchanged:", np.array_equal(st["b"], g["t_b"][0].T))
ctx.close()
