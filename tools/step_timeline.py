"""Where does an LM trial step go?  Event-timed stage spans (vlg_ba_trial_info.ms_*) next to the wall time of the
call, Venice shape.  The difference between the spans and the kernel sums of bench.py is host latency."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bundleadjustmentmatlab_b200 import capi, synth  # noqa: E402

P = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "venice", seed=0)
a = np.ascontiguousarray(np.vstack([P.w, P.Te]).T); b = np.ascontiguousarray(P.Xe[:3].T)
ctx = capi.Context(num_variableK=0)
ctx.set_problem_sparse(P.K.T, a, b, P.obs_xy, P.obs_pt, P.obs_cam)
for _ in range(3):
    if not ctx.lm_continue():
        ctx.lm_reset(a, b)
    ctx.trial_step()
rows = []
for _ in range(10):
    if not ctx.lm_continue():
        ctx.lm_reset(a, b)
    t0 = time.perf_counter()
    i = ctx.trial_step()
    wall = (time.perf_counter() - t0) * 1e3
    rows.append((wall, i["ms_stage1"], i["ms_schur"], i["ms_solve"], i["ms_stage3"], i["pcg_iters"], i["accepted"]))
print("wall    stage1  schur   solve   stage3  (sum)   iters acc")
for r in rows:
    print(" ".join(f"{v:7.3f}" for v in r[:5]), f"{sum(r[1:5]):7.3f}", f"{r[5]:5d} {r[6]}")
m = np.mean(np.array([r[:5] for r in rows]), axis=0)
print("mean:", " ".join(f"{v:7.3f}" for v in m), f"{m[1:].sum():7.3f}")
