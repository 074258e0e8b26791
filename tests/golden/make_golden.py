"""Generates tests/golden/*.npz by running THE REFERENCE ITSELF: the three mex C files of
/root/reference/toolbox/bundle compiled unmodified (oracle/_ref/libvlgref.so, see
oracle/Makefile) driven by the restated bundle_euclid.m loop (oracle/lm.py, backend="ref").

Run in the dev container (needs /root/reference):  python tests/golden/make_golden.py
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
    "euclid_fixcal": (6, 80, 400, 11, ("fix_calibration",)),
    "euclid_fullK": (5, 60, 260, 12, ()),
    "euclid_fixprincipal": (5, 60, 260, 13, ("fix_principal",)),
    "euclid_fixstructure": (4, 50, 180, 14, ("fix_calibration", "fix_structure")),
    "euclid_fixpivot": (6, 70, 330, 15, ("fix_calibration", "fix_pivot", [1, 1, 0, 0, 0, 0])),
    # structure-only BA (bundle_euclid.m:145-149): U, W, eA zeroed -> S = 0, da = pinv(0) e_ = 0, only the points move
    "euclid_fixmotion": (5, 60, 270, 16, ("fix_calibration", "fix_motion")),
}


# BASELINE.json config 1: the 4-view UCLA fixture (tests/golden/ucla4_tracks.npz, made by make_ucla4.py from the
# reference's data/UCLA_0[1-4].jpg), with the options of mview_reconstruction.m:196 ('fix_calibration') and with
# bundle_euclid's default (free intrinsics, bundle_euclid.m:49)
UCLA_CASES = {
    "euclid_ucla4_fixcal": ("fix_calibration",),
    "euclid_ucla4_fullK": (),
}


class _Fixture:
    def __init__(self, path):
        f = np.load(path)
        self.K, self.Te, self.w, self.Xe = f["K"], f["Te"], f["w"], f["Xe"]
        self._x, self._vis = f["x"], f["visible"]
        self.m, self.n = self.w.shape[1], self.Xe.shape[1]

    def dense(self):
        return np.asfortranarray(self._x), np.asfortranarray(self._vis)


def main():
    cases = [(name, synth.make_problem(m, n, nobs, seed=seed), opts) for name, (m, n, nobs, seed, opts) in CASES.items()]
    fx = os.path.join(HERE, "ucla4_tracks.npz")
    if os.path.exists(fx):
        cases += [(name, _Fixture(fx), opts) for name, opts in UCLA_CASES.items()]
    for name, P, opts in cases:
        m, n = P.m, P.n
        x, vis = P.dense()
        res = lm.bundle_euclid(P.K, P.Te, P.w, P.Xe, x, *opts, "visibility", vis, backend="ref")
        o = lm.parse_options(m, n, x, list(opts) + ["visibility", vis])
        nk = o["num_variableK"]
        a0 = res.trials[0].a
        b0 = res.trials[0].b
        X = np.asfortranarray(x[:2])
        obs = lm.ObsList.from_dense(X, vis)
        first = lm.lm_trial(P.K, a0, b0, obs, 1e-3, o, backend="ref", dense=(X, np.asfortranarray(vis)))
        s1 = first["blocks"]["s1"]
        pt, cam = obs.pt, obs.cam
        out = dict(
            m=m, n=n, num_variableK=nk, options=np.array([str(t) for t in opts if isinstance(t, str)]),
            pivot=np.asarray(o["pivot"], dtype=np.float64),
            K=P.K, Te=P.Te, w=P.w, Xe=P.Xe, x=x, visible=vis,
            obs_xy=obs.xy, obs_pt=pt, obs_cam=cam,
            # stage-1 outputs of the reference on the visible cells, list order
            X_hat=np.ascontiguousarray(s1["X_hat"][:, pt, cam].T),
            A=np.ascontiguousarray(np.transpose(s1["A"][:, :, pt, cam], (2, 1, 0))),
            B=np.ascontiguousarray(np.transpose(s1["B"][:, :, pt, cam], (2, 1, 0))),
            e=np.ascontiguousarray(s1["e"][:, pt, cam].T),
            W=np.ascontiguousarray(np.transpose(first["blocks"]["W_dense"][:, :, pt, cam], (2, 1, 0))),
            U=first["blocks"]["U"], V=first["blocks"]["V"], eA=first["blocks"]["eA"], eB=first["blocks"]["eB"],
            Vinv=first["blocks"]["Vinv"], S=first["blocks"]["S"], e_=first["blocks"]["e_"],
            error_=res.error_,
            K_=res.K_, Te_=res.Te_, w_=res.w_, Xe_=res.Xe_,
            t_a=np.stack([t.a for t in res.trials]), t_b=np.stack([t.b for t in res.trials]),
            t_lam=np.array([t.lam for t in res.trials]), t_nu=np.array([t.nu for t in res.trials]),
            t_old=np.array([t.old_cost for t in res.trials]), t_new=np.array([t.new_cost for t in res.trials]),
            t_rho=np.array([t.rho for t in res.trials]), t_accept=np.array([t.accept for t in res.trials]),
            t_denom=np.array([t.denom for t in res.trials]),
            t_da=np.stack([t.da for t in res.trials]), t_db=np.stack([t.db for t in res.trials]),
            t_a_new=np.stack([t.a_new for t in res.trials]), t_b_new=np.stack([t.b_new for t in res.trials]),
        )
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "trials", len(res.trials), "error_", res.error_, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
