"""Generates tests/golden/proj_*.npz by running THE REFERENCE ITSELF on the projective path: the reference's
mex_bundle_proj_{1_XABeUVWeAeB,2_Se_,3_db_new}.c compiled unmodified (oracle/_ref/libvlgref.so, oracle/Makefile)
driven by the restated bundle_projective.m loop (oracle/lm.py: bundle_projective).

Run in the dev container (needs /root/reference):  python tests/golden/make_golden_proj.py
The .npz files are committed; /root/reference does not exist on the GPU box.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bundleadjustmentmatlab_b200 import synth  # noqa: E402
from oracle import lm  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (m, n, nobs, seed, options)
    "proj_full": (5, 60, 260, 31, ()),
    "proj_fixstructure": (4, 50, 180, 32, ("fix_structure",)),      # multi_view.m:190
}


def rodrigues(w):
    th = np.linalg.norm(w)
    if th < 1e-6:
        return np.eye(3)
    k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def projection_matrices(P):
    Pp = np.zeros((3, 4, P.m))
    for j in range(P.m):
        Kj = np.array([[P.K[0, j], 0, P.K[2, j]], [0, P.K[1, j], P.K[3, j]], [0, 0, 1]])
        Pp[:, :, j] = Kj @ np.hstack([rodrigues(P.w[:, j]), P.Te[:, j:j + 1]])
    return Pp


def main():
    for name, (m, n, nobs, seed, opts) in CASES.items():
        P = synth.make_problem(m, n, nobs, seed=seed)
        Pp = projection_matrices(P)
        x, vis = P.dense()
        res = lm.bundle_projective(Pp, P.Xe, x, *opts, "visibility", vis)
        X = np.asfortranarray(x[:2])
        obs = lm.ObsList.from_dense(X, vis)
        pt, cam = obs.pt, obs.cam
        a0, b0 = res.trials[0].a, res.trials[0].b
        first = lm.lm_trial_proj(a0, b0, X, np.asfortranarray(vis), 1e-3, "fix_structure" in opts, "fix_motion" in opts)
        s1 = first["blocks"]["s1"]
        out = dict(
            m=m, n=n, options=np.array([str(t) for t in opts]),
            Pp=Pp, Xp=P.Xe, x=x, visible=vis, obs_xy=obs.xy, obs_pt=pt, obs_cam=cam,
            X_hat=np.ascontiguousarray(s1["X_hat"][:, pt, cam].T),
            A=np.ascontiguousarray(np.transpose(s1["A"][:, :, pt, cam], (2, 1, 0))),
            B=np.ascontiguousarray(np.transpose(s1["B"][:, :, pt, cam], (2, 1, 0))),
            e=np.ascontiguousarray(s1["e"][:, pt, cam].T),
            W=np.ascontiguousarray(np.transpose(first["blocks"]["W_dense"][:, :, pt, cam], (2, 1, 0))),
            U=first["blocks"]["U"], V=first["blocks"]["V"], eA=first["blocks"]["eA"], eB=first["blocks"]["eB"],
            error_=res.error_, Pp_=res.Pp_, Xp_=res.Xp_,
            t_a=np.stack([t.a for t in res.trials]), t_b=np.stack([t.b for t in res.trials]),
            t_lam=np.array([t.lam for t in res.trials]),
            t_old=np.array([t.old_cost for t in res.trials]), t_new=np.array([t.new_cost for t in res.trials]),
            t_accept=np.array([t.accept for t in res.trials]),
            t_da=np.stack([t.da for t in res.trials]), t_db=np.stack([t.db for t in res.trials]),
            t_a_new=np.stack([t.a_new for t in res.trials]), t_b_new=np.stack([t.b_new for t in res.trials]),
        )
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "trials", len(res.trials), "error_", res.error_, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
