"""Cholesky vs PCG crossover (BASELINE.json config 3): LM trial-step time of the three solvers on synthetic problems of
growing camera count at a fixed shape (about 250 points per camera, 3.5 observations per point -- Trafalgar's ratios).
Usage: python tools/crossover.py [m ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bundleadjustmentmatlab_b200 import capi, synth  # noqa: E402


def main():
    ms = [int(v) for v in sys.argv[1:]] or [257, 400, 600, 900, 1300]
    print(f"{'m':>6} {'nobs':>9} | {'chol ms':>9} {'pcg-S ms':>9} {'it':>4} {'pcg-impl ms':>11} {'it':>4}")
    for m in ms:
        n = 253 * m
        nobs = int(3.47 * n)
        P = synth.make_problem(m, n, nobs, seed=5)
        a = np.ascontiguousarray(np.vstack([P.w, P.Te]).T); b = np.ascontiguousarray(P.Xe[:3].T)
        row = []
        for solver in (capi.SOLVER_CHOL, capi.SOLVER_PCG_EXPLICIT, capi.SOLVER_PCG):
            ctx = capi.Context(num_variableK=0, solver=solver)
            ctx.set_problem_sparse(P.K.T, a, b, P.obs_xy, P.obs_pt, P.obs_cam)
            ctx.trial_step()
            ctx.lm_reset(a, b)
            ts, its = [], []
            for _ in range(3):
                ctx.lm_reset(a, b)
                t0 = time.perf_counter()
                info = ctx.trial_step()
                ts.append((time.perf_counter() - t0) * 1e3)
                its.append(info["pcg_iters"])
            row.append((min(ts), its[-1]))
            ctx.close()
        print(f"{m:6d} {P.nobs:9d} | {row[0][0]:9.2f} {row[1][0]:9.2f} {row[1][1]:4d} {row[2][0]:11.2f} {row[2][1]:4d}", flush=True)


if __name__ == "__main__":
    main()
