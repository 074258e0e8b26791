"""Host-side mirror of the reference's operator interface for the BA hot path.

``bundle_euclid(K, Te, w, Xe, x, *options)`` has the same name, argument meaning, option
strings and outputs as ``toolbox/bundle/bundle_euclid.m:1-26`` of the reference; it only
packs the arguments (bundle_euclid.m:81-102) and calls the C ABI (libvlgba.so), where the
whole LM loop (bundle_euclid.m:111-267) runs on the GPU.  The three mex stages are exposed
under their reference names for stage-level drop-in and parity tests.  No CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import capi


def parse_options(m, n, x, opts):
    """bundle_euclid.m:43-79 (same option strings, same defaults)."""
    o = dict(fix_structure=False, fix_motion=False, pivot=None, num_variableK=4, visible=None, verbose=False)
    opts = list(opts)
    k = 0
    while k < len(opts):
        key = str(opts[k]).lower()
        if key == "fix_structure":
            o["fix_structure"] = True
        elif key == "fix_motion":
            o["fix_motion"] = True
        elif key == "fix_pivot":
            o["pivot"] = np.asarray(opts[k + 1]).astype(bool).reshape(-1)
            k += 1
        elif key == "fix_calibration":
            o["num_variableK"] = 0
        elif key == "fix_principal":
            o["num_variableK"] = 1
        elif key == "visibility":
            o["visible"] = np.asarray(opts[k + 1])
            k += 1
        elif key == "verbose":
            o["verbose"] = True
        k += 1
    if o["visible"] is None:
        o["visible"] = ((x[0] != 0) | (x[1] != 0)).reshape(n, m)      # bundle_euclid.m:50
    return o


def pack(K, Te, w, Xe, num_variableK):
    """a = [w; Te; K-part], b = Xe(1:3,:)  (bundle_euclid.m:89-99), returned in wire layout."""
    m = w.shape[1]
    a = np.zeros((m, 6 + num_variableK))
    a[:, 0:3] = w.T
    a[:, 3:6] = Te.T
    if num_variableK == 1:
        a[:, 6] = K[0]
    elif num_variableK == 4:
        a[:, 6:10] = K.T
    b = np.ascontiguousarray(Xe[0:3].T)
    return a, b


def make_context(K, Te, w, Xe, x, *options, **ctx_opts) -> capi.Context:
    K = np.asarray(K, dtype=np.float64); Te = np.asarray(Te, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64); Xe = np.asarray(Xe, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    m, n = w.shape[1], x.shape[1]
    o = parse_options(m, n, x, options)
    ctx = capi.Context(num_variableK=o["num_variableK"], fix_structure=int(o["fix_structure"]),
                       fix_motion=int(o["fix_motion"]), verbose=int(o["verbose"]), **ctx_opts)
    a, b = pack(K, Te, w, Xe, o["num_variableK"])
    vis = np.asfortranarray(np.asarray(o["visible"], dtype=np.float64).reshape(n, m))
    ctx.set_problem_dense(np.ascontiguousarray(K.T), a, b, np.asfortranarray(x[0:2]), vis, pivot=o["pivot"])
    return ctx


def bundle_euclid(K, Te, w, Xe, x, *options, **ctx_opts):
    """[K_ Te_ w_ Xe_ error_] = bundle_euclid(K, Te, w, Xe, x, ...)  -- bundle_euclid.m:1-26."""
    Xe = np.asarray(Xe, dtype=np.float64)
    ctx = make_context(K, Te, w, Xe, x, *options, **ctx_opts)
    try:
        K_, Te_, w_, Xe_, err = ctx.solve(Xe4=Xe[3])
    finally:
        ctx.close()
    return K_.T.copy(), Te_.T.copy(), w_.T.copy(), Xe_.T.copy(), err


def bundle_euclid_sparse(K, Te, w, Xe, obs_xy, obs_pt, obs_cam, *options, **ctx_opts):
    """Same solve on an observation list (no dense n x m arrays): SURVEY.md row N1."""
    K = np.asarray(K, dtype=np.float64); Xe = np.asarray(Xe, dtype=np.float64)
    m, n = np.asarray(w).shape[1], Xe.shape[1]
    o = parse_options(m, n, np.ones((2, 1, 1)), list(options) + ["visibility", np.zeros((0,))])
    ctx = capi.Context(num_variableK=o["num_variableK"], fix_structure=int(o["fix_structure"]),
                       fix_motion=int(o["fix_motion"]), verbose=int(o["verbose"]), **ctx_opts)
    try:
        a, b = pack(K, np.asarray(Te, dtype=np.float64), np.asarray(w, dtype=np.float64), Xe, o["num_variableK"])
        ctx.set_problem_sparse(np.ascontiguousarray(K.T), a, b, obs_xy, obs_pt, obs_cam, pivot=o["pivot"])
        K_, Te_, w_, Xe_, err = ctx.solve(Xe4=Xe[3])
    finally:
        ctx.close()
    return K_.T.copy(), Te_.T.copy(), w_.T.copy(), Xe_.T.copy(), err


# ------------------------------------------------------------------------------------------
# The three mex stages under their reference names, dense Fortran-ordered arrays in and out
# exactly as bundle_euclid.m:139,192,204 passes them (drop-in / parity use).
# ------------------------------------------------------------------------------------------
import ctypes as _C


def _F(x):
    return np.asfortranarray(x, dtype=np.float64)


def _check(r):
    if r != 0:
        raise capi.VlgBaError(f"libvlgba error {r}: {capi.lib().vlg_ba_last_error(None).decode()}")


def mex_bundle_1_XABeUVWeAeB(K, a, b, X, visible):
    """[X_hat A B e U V W eA eB] = mex_bundle_1_XABeUVWeAeB(K, a, b, X, visible)."""
    K, a, b, X, visible = _F(K), _F(a), _F(b), _F(X), _F(visible)
    na, m = a.shape
    n = b.shape[1]
    o = [np.zeros(s, order="F") for s in ((2, n, m), (2, na, n, m), (2, 3, n, m), (2, n, m), (na, na, m), (3, 3, n),
                                          (na, 3, n, m), (na, m), (3, n))]
    _check(capi.lib().vlg_ba_mex1_dense(_C.c_int(m), _C.c_int(n), _C.c_int(na), capi._d(K), capi._d(a), capi._d(b),
                                        capi._d(X), capi._d(visible), *[capi._d(t) for t in o]))
    return tuple(o)


def mex_bundle_2_Se_(Y, W, U_, eA, eB):
    """[S e_] = mex_bundle_2_Se_(Y, W, U_, eA, eB)."""
    Y, W, U_, eA, eB = _F(Y), _F(W), _F(U_), _F(eA), _F(eB)
    na, m = eA.shape
    n = eB.shape[1]
    S = np.zeros((na * m, na * m), order="F"); e_ = np.zeros(na * m)
    _check(capi.lib().vlg_ba_mex2_dense(_C.c_int(m), _C.c_int(n), _C.c_int(na), capi._d(Y), capi._d(W), capi._d(U_),
                                        capi._d(eA), capi._d(eB), capi._d(S), capi._d(e_)))
    return S, e_


def mex_bundle_3_db_new(W, da, eB, V_inv, K, a, b, X, visible):
    """[db a_new b_new X_hat] = mex_bundle_3_db_new(W, da, eB, V_inv, K, a, b, X, visible)."""
    W, eB, V_inv, K, a, b, X, visible = _F(W), _F(eB), _F(V_inv), _F(K), _F(a), _F(b), _F(X), _F(visible)
    da = np.ascontiguousarray(da, dtype=np.float64).reshape(-1)
    na, m = a.shape
    n = b.shape[1]
    db = np.zeros((3, n), order="F"); a_new = np.zeros((na, m), order="F"); b_new = np.zeros((3, n), order="F")
    X_hat = np.zeros((2, n, m), order="F")
    _check(capi.lib().vlg_ba_mex3_dense(_C.c_int(m), _C.c_int(n), _C.c_int(na), capi._d(W), capi._d(da), capi._d(eB),
                                        capi._d(V_inv), capi._d(K), capi._d(a), capi._d(b), capi._d(X), capi._d(visible),
                                        capi._d(db), capi._d(a_new), capi._d(b_new), capi._d(X_hat)))
    return db, a_new, b_new, X_hat


def mex_bundle_proj_1_XABeUVWeAeB(a, b, X, visible):
    """[X_hat A B e U V W eA eB] = mex_bundle_proj_1_XABeUVWeAeB(a, b, X, visible)  (a is 12 x m)."""
    a, b, X, visible = _F(a), _F(b), _F(X), _F(visible)
    na, m = a.shape
    if na != 12:
        raise ValueError("projective cameras have 12 parameters (a must be 12 x m)")
    n = b.shape[1]
    o = [np.zeros(s, order="F") for s in ((2, n, m), (2, na, n, m), (2, 3, n, m), (2, n, m), (na, na, m), (3, 3, n),
                                          (na, 3, n, m), (na, m), (3, n))]
    _check(capi.lib().vlg_ba_mex1_dense(_C.c_int(m), _C.c_int(n), _C.c_int(na), None, capi._d(a), capi._d(b),
                                        capi._d(X), capi._d(visible), *[capi._d(t) for t in o]))
    return tuple(o)


def mex_bundle_proj_2_Se_(Y, W, U_, eA, eB):
    """[S e_] = mex_bundle_proj_2_Se_(Y, W, U_, eA, eB)  (12-row blocks)."""
    return mex_bundle_2_Se_(Y, W, U_, eA, eB)


def mex_bundle_proj_3_db_new(W, da, eB, V_inv, a, b, X, visible):
    """[db a_new b_new X_hat] = mex_bundle_proj_3_db_new(W, da, eB, V_inv, a, b, X, visible)."""
    W, eB, V_inv, a, b, X, visible = _F(W), _F(eB), _F(V_inv), _F(a), _F(b), _F(X), _F(visible)
    da = np.ascontiguousarray(da, dtype=np.float64).reshape(-1)
    na, m = a.shape
    n = b.shape[1]
    db = np.zeros((3, n), order="F"); a_new = np.zeros((na, m), order="F"); b_new = np.zeros((3, n), order="F")
    X_hat = np.zeros((2, n, m), order="F")
    _check(capi.lib().vlg_ba_mex3_dense(_C.c_int(m), _C.c_int(n), _C.c_int(na), capi._d(W), capi._d(da), capi._d(eB),
                                        capi._d(V_inv), None, capi._d(a), capi._d(b), capi._d(X), capi._d(visible),
                                        capi._d(db), capi._d(a_new), capi._d(b_new), capi._d(X_hat)))
    return db, a_new, b_new, X_hat


def bundle_projective(Pp, Xp, x, *options, **ctx_opts):
    """[Pp_ Xp_ error_] = bundle_projective(Pp, Xp, x, ...)  -- toolbox/bundle/bundle_projective.m:1-229.
    Pp (3,4,m), Xp (4,n), x (3,n,m); options 'fix_structure', 'fix_motion', 'visibility', vis, 'verbose'."""
    Pp = np.asarray(Pp, dtype=np.float64); Xp = np.asarray(Xp, dtype=np.float64); x = np.asarray(x, dtype=np.float64)
    m, n = Pp.shape[2], x.shape[1]
    o = parse_options(m, n, x, options)
    opts = capi.default_opts(model=capi.MODEL_PROJECTIVE, fix_structure=int(o["fix_structure"]),
                             fix_motion=int(o["fix_motion"]), verbose=int(o["verbose"]), **ctx_opts)
    Ppf = np.asfortranarray(Pp); Xpf = np.asfortranarray(Xp); xf = np.asfortranarray(x)
    vis = np.asfortranarray(np.asarray(o["visible"], dtype=np.float64).reshape(n, m))
    Pp_ = np.zeros((3, 4, m), order="F"); Xp_ = np.zeros((4, n), order="F")
    err = np.zeros(opts.max_iter + 2); ne = _C.c_int(0)
    _check(capi.lib().vlg_ba_bundle_projective(_C.byref(opts), _C.c_int(m), _C.c_int(n), capi._d(Ppf), capi._d(Xpf), capi._d(xf),
                                               capi._d(vis), capi._d(Pp_), capi._d(Xp_), capi._d(err), _C.byref(ne)))
    return Pp_, Xp_, err[:ne.value].copy()
