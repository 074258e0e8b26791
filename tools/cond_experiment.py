"""Why the free-intrinsics cases (num_variableK = 1 or 4, bundle_euclid's DEFAULT: bundle_euclid.m:49) are held to 1e-6
instead of 1e-9 on the one-step cost: how far apart are two backward-stable solves of the SAME
reduced system, and how far is each from the exact solution?

For every golden with free intrinsics: S, e_ from the reference-built golden; da by (a) the oracle's SVD pinv (MATLAB's
tolerance), (b) LU with partial pivoting, (c) Cholesky with zero-row elimination (what the GPU does), (d) (c) + two steps
of iterative refinement with the residual in extended precision (numpy longdouble, 64-bit significand), taken as the
reference answer.  Prints cond(S), the relative distance of each da from (d), and the resulting one-step cost of each
(teacher-forced, same stage 3) relative to the golden's.  CPU only:  python tools/cond_experiment.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import golden_names, load_golden, oracle_options  # noqa: E402
from oracle import lm  # noqa: E402


def chol_elim(S, e):
    keep = np.flatnonzero(np.abs(S).sum(axis=0) > 0)
    L = np.linalg.cholesky(S[np.ix_(keep, keep)])
    x = np.zeros_like(e)
    x[keep] = np.linalg.solve(L.T, np.linalg.solve(L, e[keep]))
    return x, keep


def refine(S, e, x, keep, steps=3):
    Sl = S[np.ix_(keep, keep)].astype(np.longdouble); el = e[keep].astype(np.longdouble)
    xl = x[keep].astype(np.longdouble)
    L = np.linalg.cholesky(S[np.ix_(keep, keep)])
    for _ in range(steps):
        r = (el - Sl @ xl).astype(np.float64)
        xl = xl + np.linalg.solve(L.T, np.linalg.solve(L, r)).astype(np.longdouble)
    out = np.zeros_like(e)
    out[keep] = xl.astype(np.float64)
    return out


def main():
    for name in golden_names():
        g = load_golden(name)
        if int(g["num_variableK"]) == 0 and "ucla" not in name:
            continue
        o = oracle_options(g)
        obs = lm.ObsList(int(g["m"]), int(g["n"]), g["obs_xy"], g["obs_pt"], g["obs_cam"])
        a, b, lam = g["t_a"][0], g["t_b"][0], float(g["t_lam"][0])
        t = lm.lm_trial(g["K"], a, b, obs, lam, o, backend="sparse")
        S, e_ = t["blocks"]["S"], t["blocks"]["e_"]
        sv = np.linalg.svd(S, compute_uv=False)
        nz = sv[sv > sv[0] * 1e-300]
        x_pinv = t["da"]
        x_lu = np.zeros_like(e_)
        x_ch, keep = chol_elim(S, e_)
        x_lu[keep] = np.linalg.solve(S[np.ix_(keep, keep)], e_[keep])
        x_ref = refine(S, e_, x_ch, keep)

        def cost(da):
            W, eB, Vinv = t["blocks"]["W"], t["blocks"]["eB"], t["blocks"]["Vinv"]
            _, _, _, X_hat_new = lm.stage3_sparse(W, da, eB, Vinv, g["K"], a, b, obs)
            en = obs.xy - X_hat_new
            return float(np.dot(en.reshape(-1), en.reshape(-1)))
        c_ref = cost(x_ref)
        rel = lambda x: np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)
        crel = lambda x: abs(cost(x) - c_ref) / c_ref
        print(f"{name}: num_variableK {int(g['num_variableK'])}, cond(S) over the non-zero part {nz[0] / nz[-1]:.2e} ({S.shape[0] - len(keep)} zero rows)")
        print(f"   |da - da_ref| / |da_ref|:  SVD pinv (oracle) {rel(x_pinv):.1e}   LU {rel(x_lu):.1e}   Cholesky+elimination (GPU's method) {rel(x_ch):.1e}")
        print(f"   one-step cost vs the refined solve:  SVD pinv {crel(x_pinv):.1e}   LU {crel(x_lu):.1e}   Cholesky {crel(x_ch):.1e}")


if __name__ == "__main__":
    main()


def projective():
    """The same question for the projective model (row N3: full-problem teacher-forced cost 4.8e-7 from the oracle): the
    15-dimensional projective gauge leaves S with 15 eigenvalues ~ lambda * diag(U), cond(S) ~ 1 / lambda larger."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import proj_golden_names
    for name in proj_golden_names():
        g = load_golden(name)
        fs = "fix_structure" in g["options"]
        a, b, lam = g["t_a"][0], g["t_b"][0], float(g["t_lam"][0])
        X = np.asfortranarray(g["x"][:2]); vis = np.asfortranarray(g["visible"])
        t = lm.lm_trial_proj(a, b, X, vis, lam, fix_structure=fs)
        S, e_ = t["blocks"]["S"], t["blocks"]["e_"]
        sv = np.linalg.svd(S, compute_uv=False)
        nz = sv[sv > sv[0] * 1e-300]
        x_ch, keep = chol_elim(S, e_)
        x_lu = np.zeros_like(e_); x_lu[keep] = np.linalg.solve(S[np.ix_(keep, keep)], e_[keep])
        x_ref = refine(S, e_, x_ch, keep)
        W, eB, Vinv = t["blocks"]["W_dense"], np.asfortranarray(t["blocks"]["eB"].T), np.asfortranarray(np.transpose(t["blocks"]["Vinv"], (2, 1, 0)))

        def cost(da):
            _, _, _, X_hat_new = lm.stage3_ref_proj(W, da, eB, Vinv, a, b, X, vis)
            en = X - X_hat_new
            return float(np.dot(en.reshape(-1, order="F"), en.reshape(-1, order="F")))
        c_ref = cost(x_ref)
        rel = lambda x: np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)
        crel = lambda x: abs(cost(x) - c_ref) / c_ref
        print(f"{name}: cond(S) over the non-zero part {nz[0] / nz[-1]:.2e} ({S.shape[0] - len(keep)} zero rows), lambda {lam:g}")
        print(f"   |da - da_ref| / |da_ref|:  SVD pinv (oracle) {rel(t['da']):.1e}   LU {rel(x_lu):.1e}   Cholesky+elimination (GPU's method) {rel(x_ch):.1e}")
        print(f"   one-step cost vs the refined solve:  SVD pinv {crel(t['da']):.1e}   LU {crel(x_lu):.1e}   Cholesky {crel(x_ch):.1e}")


if __name__ == "__main__":
    projective()
