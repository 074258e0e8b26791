"""Parity on the shapes BASELINE.json names, against the CPU oracle, through the C ABI.

  C1  4-view UCLA fixture            -> goldens euclid_ucla4_* (reference's own C), covered by every parametrised
                                        test of test_gpu_parity.py / test_oracle.py / test_mex_dropin.py
  C2  Ladybug-shaped   (49 cameras)  -> teacher-forced LM trial steps vs lm.lm_trial (SVD pinv of the dense S,
  C3  Trafalgar-shaped (257 cameras)    bundle_euclid.m:193), all three solvers
  C4  Venice-shaped (1 778 cameras)  -> one LM trial step vs the sparse port's PCG at rtol 1e-12 (the dense reference
                                        cannot hold this size: W alone would be 255 GB), AUTO and implicit solvers

Bars (north_star): old cost <= 1e-12, new cost <= 1e-9 relative, accept decision equal.
"""
import numpy as np
import pytest

from bundleadjustmentmatlab_b200 import capi, synth
from oracle import lm

from common import rel

pytestmark = pytest.mark.gpu

COST_RTOL = 1e-9


def _problem(name):
    P = synth.make_config(name, seed=0)
    a = np.vstack([P.w, P.Te]); b = P.Xe[:3].copy()
    obs = lm.ObsList(P.m, P.n, P.obs_xy, P.obs_pt, P.obs_cam)
    o = lm.parse_options(P.m, P.n, np.ones((2, 1, 1)), ["fix_calibration", "visibility", np.zeros(0)])
    return P, a, b, obs, o


@pytest.mark.parametrize("name", ["ladybug", "trafalgar"])
def test_c2_c3_teacher_forced_vs_pinv_oracle(name):
    P, a, b, obs, o = _problem(name)
    # the oracle's own trajectory: three trips of the loop with the reference's lambda schedule
    traj = []
    aa, bb, lam, nu = a, b, 1e-3, 2.0
    for _ in range(3):
        t = lm.lm_trial(P.K, aa, bb, obs, lam, o, backend="sparse")
        traj.append((aa, bb, lam, nu, t))
        if t["old"] - t["new"] > 0:
            rho = (t["old"] - t["new"]) / t["denom"]
            aa, bb, lam, nu = t["a_new"], t["b_new"], lam * max(1 / 3, 1 - (2 * rho - 1) ** 3), 2.0
        else:
            lam, nu = lam * nu, 2 * nu
    for solver in (capi.SOLVER_CHOL, capi.SOLVER_PCG, capi.SOLVER_PCG_EXPLICIT):
        ctx = capi.Context(num_variableK=0, solver=solver, pcg_rtol=1e-12)
        ctx.set_problem_sparse(P.K.T, a.T, b.T, P.obs_xy, P.obs_pt, P.obs_cam)
        if solver == capi.SOLVER_CHOL:
            J = ctx.get_jacobians()
            s1 = traj[0][4]["blocks"]["s1"]
            for k in ("X_hat", "A", "B", "e"):
                assert np.array_equal(J[k], s1[k]), k
        worst = 0.0
        for k, (aa, bb, lam, nu, t) in enumerate(traj):
            ctx.set_state(a=aa.T, b=bb.T, lam=lam, nu=nu)
            info = ctx.trial_step()
            assert info["solver_used"] == solver
            assert rel(info["old_cost"], t["old"]) <= 1e-12, (name, solver, k)
            r = rel(info["new_cost"], t["new"])
            worst = max(worst, r)
            assert r <= COST_RTOL, (name, solver, k, r)
            assert bool(info["accepted"]) == bool(t["old"] - t["new"] > 0)
            assert rel(info["denom"], t["denom"]) <= 1e-6
        print(f"{name} solver {solver}: worst teacher-forced cost deviation over {len(traj)} steps {worst:.2e}")
        ctx.close()


def test_c4_venice_step_vs_sparse_port():
    P, a, b, obs, _ = _problem("venice")
    r = lm.trial_step_pcg(P.K, a, b, obs, 1e-3, pcg_rtol=1e-12, pcg_max_iter=3000)
    assert r["old"] - r["new"] > 0
    for solver in (capi.SOLVER_AUTO, capi.SOLVER_PCG):
        ctx = capi.Context(num_variableK=0, solver=solver, pcg_rtol=1e-12)
        ctx.set_problem_sparse(P.K.T, np.ascontiguousarray(a.T), np.ascontiguousarray(b.T), P.obs_xy, P.obs_pt, P.obs_cam)
        info = ctx.trial_step()
        up = ctx.get_state()
        d_old, d_new = rel(info["old_cost"], r["old"]), rel(info["new_cost"], r["new"])
        print(f"venice solver {info['solver_used']}: old {d_old:.2e} new {d_new:.2e}, PCG iterations {info['pcg_iters']} vs {r['pcg_iters']} (CPU)")
        assert d_old <= 1e-12 and d_new <= COST_RTOL
        assert info["accepted"]
        assert np.array_equal(np.isfinite(up["a"]), np.ones_like(up["a"], dtype=bool))
        # the accepted state is the oracle's candidate to the solver tolerance
        assert np.abs(up["a"] - r["a_new"].T).max() <= 1e-7 * max(np.abs(r["a_new"]).max(), 1.0)
        assert np.abs(up["b"] - r["b_new"].T).max() <= 1e-7 * max(np.abs(r["b_new"]).max(), 1.0)
        ctx.close()
    # library defaults (pcg_rtol 1e-8) stay inside the bar as well
    ctx = capi.Context(num_variableK=0)
    ctx.set_problem_sparse(P.K.T, np.ascontiguousarray(a.T), np.ascontiguousarray(b.T), P.obs_xy, P.obs_pt, P.obs_cam)
    info = ctx.trial_step()
    print(f"venice defaults: new {rel(info['new_cost'], r['new']):.2e}, {info['pcg_iters']} iterations")
    assert rel(info["new_cost"], r["new"]) <= COST_RTOL
    ctx.close()
