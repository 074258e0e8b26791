"""oracle/lm.py -- TEST INFRASTRUCTURE ONLY.

CPU oracle for the Levenberg-Marquardt Euclidean bundle-adjustment hot path of
caomw/BundleAdjustmentMatlab (toolbox/bundle).  Two interchangeable back ends:

* ``ref``    -- the reference's own three mex C files compiled UNMODIFIED into
                ``oracle/_ref/libvlgref.so`` (oracle/Makefile), called on the dense
                ``n x m`` arrays exactly as ``bundle_euclid.m:139,192,204`` calls them;
* ``sparse`` -- ``oracle/oracle_sparse.c`` (this repo's restatement on an observation
                list, ``liboracle.so``), asserted bit-identical to ``ref`` by
                tests/test_oracle.py.

The MATLAB driver ``bundle_euclid.m:81-267`` (not runnable here: no MATLAB/Octave) is
restated line by line in :func:`bundle_euclid`; MATLAB's closed-source ``pinv`` is
restated as an SVD pseudo-inverse with ``tol = max(size) * eps(sigma_max)``.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  Parity status: pinned against the
reference's own C run in this container (the reference ships no golden vectors).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(quiet: bool = True) -> None:
    """Compile liboracle.so and (when /root/reference is present) _ref/libvlgref.so."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _load(path):
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


_sparse = None
_ref = None


def sparse_lib():
    global _sparse
    if _sparse is None:
        lib = _load(os.path.join(_HERE, "liboracle.so"))
        lib.orc_cost.restype = C.c_double
        lib.orc_trial_step_pcg.restype = C.c_int
        lib.orc_num_threads.restype = C.c_int
        _sparse = lib
    return _sparse


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libvlgref.so")) or os.path.isdir("/root/reference")


def ref_lib():
    global _ref
    if _ref is None:
        _ref = _load(os.path.join(_HERE, "_ref", "libvlgref.so"))
    return _ref


def use_ref_library(path=None):
    """Point the ``ref`` back end at another library exporting vlgref_stage1/2/3 (the flat
    wrappers of ref_glue.c).  tests/test_mex_dropin.py uses this to run the restated
    bundle_euclid.m driver over the GPU mex drop-ins (mex/_build/libvlgmex_shim.so) instead of
    the reference's own mex files; ``None`` restores the reference build."""
    global _ref
    _ref = None if path is None else C.CDLL(path)


def _d(x):
    return x.ctypes.data_as(_dp)


def _i(x):
    return x.ctypes.data_as(_ip)


def _f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


# ----------------------------------------------------------------------------------
# pinv (MATLAB semantics): tol = max(size(A)) * eps(norm(A))
# ----------------------------------------------------------------------------------
def pinv_matlab(A: np.ndarray) -> np.ndarray:
    A = np.asarray(A, dtype=np.float64)
    if A.size == 0:
        return A.T.copy()
    if not np.any(A):
        return np.zeros(A.T.shape)
    U, s, Vt = np.linalg.svd(A, full_matrices=False)
    tol = max(A.shape) * np.spacing(s[0])
    r = int(np.sum(s > tol))
    if r == 0:
        return np.zeros(A.T.shape)
    return (Vt[:r].T / s[:r]) @ U[:, :r].T


def pinv3_batch(V: np.ndarray) -> np.ndarray:
    """pinv of each 3x3 block of V (n,3,3), MATLAB tolerance (bundle_euclid.m:180)."""
    n = V.shape[0]
    out = np.zeros_like(V)
    if n == 0:
        return out
    U, s, Vt = np.linalg.svd(V)
    tol = 3 * np.spacing(s[:, 0])
    keep = s > tol[:, None]
    sinv = np.where(keep, 1.0 / np.where(keep, s, 1.0), 0.0)
    out = np.einsum("nji,nj,nkj->nik", Vt, sinv, U)
    return out


# ----------------------------------------------------------------------------------
# problem containers
# ----------------------------------------------------------------------------------
@dataclass
class ObsList:
    """Observation list in the reference's traversal order (ascending i + n*j)."""
    m: int
    n: int
    xy: np.ndarray      # (nobs, 2) float64
    pt: np.ndarray      # (nobs,) int32
    cam: np.ndarray     # (nobs,) int32

    @property
    def nobs(self):
        return int(self.pt.shape[0])

    @staticmethod
    def from_dense(X: np.ndarray, visible: np.ndarray) -> "ObsList":
        """X (2,n,m), visible (n,m).  Order: ascending i + n*j (mex1:192-196)."""
        n, m = visible.shape
        flat = np.flatnonzero(np.asarray(visible, dtype=np.float64).reshape(-1, order="F") != 0)
        cam = (flat // n).astype(np.int32)
        pt = (flat % n).astype(np.int32)
        xy = np.stack([X[0, pt, cam], X[1, pt, cam]], axis=1).astype(np.float64)
        return ObsList(m, n, np.ascontiguousarray(xy), pt, cam)

    def to_dense(self):
        X = np.zeros((2, self.n, self.m), order="F")
        vis = np.zeros((self.n, self.m), order="F")
        X[0, self.pt, self.cam] = self.xy[:, 0]
        X[1, self.pt, self.cam] = self.xy[:, 1]
        vis[self.pt, self.cam] = 1.0
        return X, vis


# ----------------------------------------------------------------------------------
# stage wrappers -- sparse back end (oracle_sparse.c)
# ----------------------------------------------------------------------------------
def stage1_sparse(K, a, b, obs: ObsList):
    """K (4,m), a (na,m), b (3,n) -> dict of per-observation and per-block outputs."""
    lib = sparse_lib()
    na, m = a.shape
    n = b.shape[1]
    no = obs.nobs
    Kf, af, bf = _f64(K.T), _f64(a.T), _f64(b.T)          # column-major == row-major of transpose
    out = dict(
        X_hat=np.zeros((no, 2)), A=np.zeros((no, na, 2)), B=np.zeros((no, 3, 2)), e=np.zeros((no, 2)),
        U=np.zeros((m, na, na)), V=np.zeros((n, 3, 3)), W=np.zeros((no, 3, na)),
        eA=np.zeros((m, na)), eB=np.zeros((n, 3)))
    lib.orc_stage1(C.c_int(m), C.c_int(n), C.c_int(na), _d(Kf), _d(af), _d(bf), C.c_long(no),
                   _d(obs.xy), _i(obs.pt), _i(obs.cam),
                   _d(out["X_hat"]), _d(out["A"]), _d(out["B"]), _d(out["e"]),
                   _d(out["U"]), _d(out["V"]), _d(out["W"]), _d(out["eA"]), _d(out["eB"]))
    return out


def make_Y_sparse(W, Vinv, obs: ObsList):
    lib = sparse_lib()
    na = W.shape[2]
    Y = np.zeros_like(W)
    lib.orc_make_Y(C.c_int(na), C.c_long(obs.nobs), _i(obs.pt), _d(_f64(W)), _d(_f64(Vinv)), _d(Y))
    return Y


def stage2_sparse(Y, W, U_, eA, eB, obs: ObsList):
    lib = sparse_lib()
    m, na = eA.shape
    n = eB.shape[0]
    N = na * m
    S = np.zeros((N, N))
    e_ = np.zeros(N)
    lib.orc_stage2_dense(C.c_int(m), C.c_int(n), C.c_int(na), C.c_long(obs.nobs), _i(obs.pt), _i(obs.cam),
                         _d(_f64(Y)), _d(_f64(W)), _d(_f64(U_)), _d(_f64(eA)), _d(_f64(eB)), _d(S), _d(e_))
    return S.T.copy(), e_       # S was written column-major


def stage3_sparse(W, da, eB, Vinv, K, a, b, obs: ObsList, all_rows=False):
    lib = sparse_lib()
    na, m = a.shape
    n = b.shape[1]
    db = np.zeros((n, 3)); a_new = np.zeros((m, na)); b_new = np.zeros((n, 3))
    X_hat = np.zeros((obs.nobs, 2))
    lib.orc_stage3(C.c_int(m), C.c_int(n), C.c_int(na), _d(_f64(K.T)), _d(_f64(a.T)), _d(_f64(b.T)),
                   C.c_long(obs.nobs), _i(obs.pt), _i(obs.cam), _d(_f64(W)), _d(_f64(da)), _d(_f64(eB)),
                   _d(_f64(Vinv)), C.c_int(1 if all_rows else 0), _d(db), _d(a_new), _d(b_new), _d(X_hat))
    return db, a_new.T.copy(), b_new.T.copy(), X_hat


# ----------------------------------------------------------------------------------
# stage wrappers -- dense reference back end (the reference's own mex C)
# ----------------------------------------------------------------------------------
def _F(x):
    return np.asfortranarray(x, dtype=np.float64)


def stage1_ref(K, a, b, X, visible):
    """mex_bundle_1_XABeUVWeAeB(K, a, b, X, visible) -> 9 dense outputs (Fortran order)."""
    lib = ref_lib()
    na, m = a.shape
    n = b.shape[1]
    o = dict(X_hat=np.zeros((2, n, m), order="F"), A=np.zeros((2, na, n, m), order="F"),
             B=np.zeros((2, 3, n, m), order="F"), e=np.zeros((2, n, m), order="F"),
             U=np.zeros((na, na, m), order="F"), V=np.zeros((3, 3, n), order="F"),
             W=np.zeros((na, 3, n, m), order="F"), eA=np.zeros((na, m), order="F"),
             eB=np.zeros((3, n), order="F"))
    K, a, b, X, visible = _F(K), _F(a), _F(b), _F(X), _F(visible)
    lib.vlgref_stage1(C.c_int(m), C.c_int(n), C.c_int(na), _d(K), _d(a), _d(b), _d(X), _d(visible),
                      _d(o["X_hat"]), _d(o["A"]), _d(o["B"]), _d(o["e"]), _d(o["U"]), _d(o["V"]),
                      _d(o["W"]), _d(o["eA"]), _d(o["eB"]))
    return o


def stage2_ref(Y, W, U_, eA, eB):
    lib = ref_lib()
    na, m = eA.shape
    n = eB.shape[1]
    S = np.zeros((na * m, na * m), order="F")
    e_ = np.zeros(na * m)
    lib.vlgref_stage2(C.c_int(m), C.c_int(n), C.c_int(na), _d(_F(Y)), _d(_F(W)), _d(_F(U_)), _d(_F(eA)),
                      _d(_F(eB)), _d(S), _d(e_))
    return S, e_


def stage3_ref(W, da, eB, Vinv, K, a, b, X, visible):
    lib = ref_lib()
    na, m = a.shape
    n = b.shape[1]
    db = np.zeros((3, n), order="F"); a_new = np.zeros((na, m), order="F")
    b_new = np.zeros((3, n), order="F"); X_hat = np.zeros((2, n, m), order="F")
    lib.vlgref_stage3(C.c_int(m), C.c_int(n), C.c_int(na), _d(_F(W)), _d(_f64(da)), _d(_F(eB)), _d(_F(Vinv)),
                      _d(_F(K)), _d(_F(a)), _d(_F(b)), _d(_F(X)), _d(_F(visible)),
                      _d(db), _d(a_new), _d(b_new), _d(X_hat))
    return db, a_new, b_new, X_hat


# ----------------------------------------------------------------------------------
# the LM driver: bundle_euclid.m:81-267
# ----------------------------------------------------------------------------------
@dataclass
class Trial:
    """Everything one trip of the while loop (bundle_euclid.m:120-249) saw and produced."""
    a: np.ndarray
    b: np.ndarray
    lam: float
    nu: float
    old_cost: float
    new_cost: float
    rho: float
    accept: bool
    da: np.ndarray
    db: np.ndarray
    a_new: np.ndarray
    b_new: np.ndarray
    denom: float


@dataclass
class Result:
    K_: np.ndarray
    Te_: np.ndarray
    w_: np.ndarray
    Xe_: np.ndarray
    error_: np.ndarray
    trials: list = field(default_factory=list)


def parse_options(m, n, x, opts):
    """bundle_euclid.m:43-79."""
    o = dict(fix_structure=False, fix_motion=False, fix_pivot=False, pivot=np.zeros(m, dtype=bool),
             num_variableK=4, visible=None, verbose=False)
    k = 0
    opts = list(opts)
    while k < len(opts):
        key = str(opts[k]).lower()
        if key == "fix_structure":
            o["fix_structure"] = True
        elif key == "fix_motion":
            o["fix_motion"] = True
        elif key == "fix_pivot":
            o["fix_pivot"] = True
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
        o["visible"] = ((x[0] != 0) | (x[1] != 0)).reshape(n, m)
    return o


def lm_trial(K, a, b, obs: ObsList, lam, o, backend="sparse", dense=None, all_rows=False):
    """One trip of the loop body, bundle_euclid.m:139-217.  Returns the pieces the
    accept/reject logic needs plus the intermediate blocks (for stage-level parity)."""
    na, m = a.shape
    n = b.shape[1]
    if backend == "ref":
        X, vis = dense
        s1 = stage1_ref(K, a, b, X, vis)
        U, V, W, eA, eB, e = s1["U"], s1["V"], s1["W"], s1["eA"], s1["eB"], s1["e"]
        if o["fix_structure"]:
            V[...] = 0; W[...] = 0; eB[...] = 0
        if o["fix_motion"]:
            U[...] = 0; W[...] = 0; eA[...] = 0
        if o["fix_pivot"]:
            U[:, :, o["pivot"]] = 0; W[:, :, :, o["pivot"]] = 0; eA[:, o["pivot"]] = 0
        U_ = U.copy(order="F")
        V_ = V.copy(order="F")
        for k in range(na):
            U_[k, k, :] = (1 + lam) * U[k, k, :]
        for k in range(3):
            V_[k, k, :] = (1 + lam) * V[k, k, :]
        Vinv = np.asfortranarray(np.transpose(pinv3_batch(np.ascontiguousarray(np.transpose(V_, (2, 0, 1)))), (1, 2, 0)))
        # Y_ij = W_ij * V_inv_i as the plain left-to-right triple sum (same statement as
        # orc_make_Y; MATLAB's own BLAS order is unknown)
        Y = np.zeros_like(W)
        for c in range(3):
            Y[:, c] = (W[:, 0] * Vinv[0, c][None, :, None] + W[:, 1] * Vinv[1, c][None, :, None]
                       + W[:, 2] * Vinv[2, c][None, :, None])
        S, e_ = stage2_ref(Y, W, U_, eA, eB)
        da = pinv_matlab(S) @ e_
        db, a_new, b_new, X_hat_new = stage3_ref(W, da, eB, Vinv, K, a, b, X, vis)
        e_new = X - X_hat_new
        old = float(np.dot(e.reshape(-1, order="F"), e.reshape(-1, order="F")))
        new = float(np.dot(e_new.reshape(-1, order="F"), e_new.reshape(-1, order="F")))
        g = np.concatenate([eA.reshape(-1, order="F"), eB.reshape(-1, order="F")])
        dp = np.concatenate([da, db.reshape(-1, order="F")])
        blocks = dict(U=np.transpose(U, (2, 1, 0)).copy(), V=np.transpose(V, (2, 1, 0)).copy(),
                      eA=eA.T.copy(), eB=eB.T.copy(), Vinv=np.transpose(Vinv, (2, 1, 0)).copy(),
                      S=np.array(S), e_=e_, W_dense=W, s1=s1)
    else:
        s1 = stage1_sparse(K, a, b, obs)
        U, V, W, eA, eB, e = s1["U"], s1["V"], s1["W"], s1["eA"], s1["eB"], s1["e"]
        if o["fix_structure"]:
            V[...] = 0; W[...] = 0; eB[...] = 0
        if o["fix_motion"]:
            U[...] = 0; W[...] = 0; eA[...] = 0
        if o["fix_pivot"]:
            piv = o["pivot"]
            U[piv] = 0; W[piv[obs.cam]] = 0; eA[piv] = 0
        U_ = U.copy(); V_ = V.copy()
        for k in range(na):
            U_[:, k, k] = (1 + lam) * U[:, k, k]
        for k in range(3):
            V_[:, k, k] = (1 + lam) * V[:, k, k]
        # blocks are stored as (idx, col, row) == column-major per block
        Vinv = np.ascontiguousarray(np.transpose(pinv3_batch(np.ascontiguousarray(np.transpose(V_, (0, 2, 1)))), (0, 2, 1)))
        Y = make_Y_sparse(W, Vinv, obs)
        S, e_ = stage2_sparse(Y, W, U_, eA, eB, obs)
        da = pinv_matlab(S) @ e_
        db, a_new, b_new, X_hat_new = stage3_sparse(W, da, eB, Vinv, K, a, b, obs, all_rows=all_rows)
        e_new = obs.xy - X_hat_new
        old = float(np.dot(e.reshape(-1), e.reshape(-1)))
        new = float(np.dot(e_new.reshape(-1), e_new.reshape(-1)))
        g = np.concatenate([eA.reshape(-1), eB.reshape(-1)])
        dp = np.concatenate([da, db.reshape(-1)])
        db = db.T.copy()
        blocks = dict(U=U, V=V, eA=eA, eB=eB, Vinv=Vinv, S=S, e_=e_, W=W, s1=s1)
    denom = float(dp @ (lam * dp + g))
    return dict(old=old, new=new, denom=denom, da=np.array(da), db=np.array(db), a_new=np.array(a_new),
                b_new=np.array(b_new), blocks=blocks)


def bundle_euclid(K, Te, w, Xe, x, *opts, backend="sparse", record=True, max_iter=20, max_iter2=10,
                  all_rows=False):
    """[K_ Te_ w_ Xe_ error_] = bundle_euclid(K, Te, w, Xe, x, ...)  (bundle_euclid.m:1-269).

    K (4,m), Te (3,m), w (3,m), Xe (4,n), x (3,n,m); options as the MATLAB strings."""
    K = np.asarray(K, dtype=np.float64); Te = np.asarray(Te, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64); Xe = np.asarray(Xe, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    m = w.shape[1]; n = x.shape[1]
    o = parse_options(m, n, x, opts)
    visible = np.asfortranarray(np.asarray(o["visible"], dtype=np.float64).reshape(n, m))
    num_vis = float(visible.sum())
    nk = o["num_variableK"]
    na = 6 + nk
    a = np.zeros((na, m))
    a[0:3] = w; a[3:6] = Te
    if nk == 1:
        a[6] = K[0]
    elif nk == 4:
        a[6:10] = K
    b = Xe[0:3].copy()
    X = np.asfortranarray(x[0:2])
    obs = ObsList.from_dense(X, visible)
    lam, nu = 1e-3, 2.0
    it, it2 = 1, 0
    error_ = []
    trials = []

    def cont():
        if not (it < max_iter and it2 < max_iter2):
            return False
        if it < 3:
            return True
        return error_[it - 1] > 1e-20 and error_[it - 2] - error_[it - 1] > 1e-3 * error_[it - 2]

    while cont():
        t = lm_trial(K, a, b, obs, lam, o, backend=backend, dense=(X, visible), all_rows=all_rows)
        old, new = t["old"], t["new"]
        rho = (old - new) / t["denom"] if t["denom"] != 0 else np.inf
        accept = (old - new) > 0
        if record:
            trials.append(Trial(a.copy(), b.copy(), lam, nu, old, new, rho, bool(accept), t["da"], t["db"],
                                t["a_new"], t["b_new"], t["denom"]))
        if accept:
            a = np.array(t["a_new"]); b = np.array(t["b_new"])
            lam = lam * max(1.0 / 3.0, 1 - (2 * rho - 1) ** 3)
            nu = 2.0
            while len(error_) < it + 1:
                error_.append(0.0)
            error_[it - 1] = old / num_vis
            it += 1
            error_[it - 1] = new / num_vis
            it2 = 0
        else:
            lam = lam * nu
            nu = 2 * nu
            it2 += 1
    K_ = K.copy()
    if nk == 1:
        K_[0] = a[6]; K_[1] = a[6]
    elif nk == 4:
        K_[:] = a[6:10]
    return Result(K_, a[3:6].copy(), a[0:3].copy(), np.vstack([b, Xe[3:4]]), np.array(error_), trials)


# ----------------------------------------------------------------------------------
# projective BA: bundle_projective.m:1-229 over mex_bundle_proj_{1,2,3} (reference build only)
# ----------------------------------------------------------------------------------
def stage1_ref_proj(a, b, X, visible):
    """mex_bundle_proj_1_XABeUVWeAeB(a, b, X, visible) -> 9 dense outputs (Fortran order)."""
    lib = ref_lib()
    na, m = a.shape
    assert na == 12
    n = b.shape[1]
    o = dict(X_hat=np.zeros((2, n, m), order="F"), A=np.zeros((2, na, n, m), order="F"),
             B=np.zeros((2, 3, n, m), order="F"), e=np.zeros((2, n, m), order="F"),
             U=np.zeros((na, na, m), order="F"), V=np.zeros((3, 3, n), order="F"),
             W=np.zeros((na, 3, n, m), order="F"), eA=np.zeros((na, m), order="F"),
             eB=np.zeros((3, n), order="F"))
    a, b, X, visible = _F(a), _F(b), _F(X), _F(visible)
    lib.vlgref_pstage1(C.c_int(m), C.c_int(n), _d(a), _d(b), _d(X), _d(visible),
                       _d(o["X_hat"]), _d(o["A"]), _d(o["B"]), _d(o["e"]), _d(o["U"]), _d(o["V"]),
                       _d(o["W"]), _d(o["eA"]), _d(o["eB"]))
    return o


def stage2_ref_proj(Y, W, U_, eA, eB):
    lib = ref_lib()
    na, m = eA.shape
    n = eB.shape[1]
    S = np.zeros((na * m, na * m), order="F")
    e_ = np.zeros(na * m)
    lib.vlgref_pstage2(C.c_int(m), C.c_int(n), _d(_F(Y)), _d(_F(W)), _d(_F(U_)), _d(_F(eA)), _d(_F(eB)), _d(S), _d(e_))
    return S, e_


def stage3_ref_proj(W, da, eB, Vinv, a, b, X, visible):
    lib = ref_lib()
    na, m = a.shape
    n = b.shape[1]
    db = np.zeros((3, n), order="F"); a_new = np.zeros((na, m), order="F")
    b_new = np.zeros((3, n), order="F"); X_hat = np.zeros((2, n, m), order="F")
    lib.vlgref_pstage3(C.c_int(m), C.c_int(n), _d(_F(W)), _d(_f64(da)), _d(_F(eB)), _d(_F(Vinv)),
                       _d(_F(a)), _d(_F(b)), _d(_F(X)), _d(_F(visible)), _d(db), _d(a_new), _d(b_new), _d(X_hat))
    return db, a_new, b_new, X_hat


def lm_trial_proj(a, b, X, vis, lam, fix_structure=False, fix_motion=False):
    """One trip of the loop body of bundle_projective.m:116-178."""
    na, m = a.shape
    s1 = stage1_ref_proj(a, b, X, vis)
    U, V, W, eA, eB, e = s1["U"], s1["V"], s1["W"], s1["eA"], s1["eB"], s1["e"]
    if fix_structure:
        V[...] = 0; W[...] = 0; eB[...] = 0
    if fix_motion:
        U[...] = 0; W[...] = 0; eA[...] = 0
    U_ = U.copy(order="F"); V_ = V.copy(order="F")
    for k in range(na):
        U_[k, k, :] = (1 + lam) * U[k, k, :]
    for k in range(3):
        V_[k, k, :] = (1 + lam) * V[k, k, :]
    Vinv = np.asfortranarray(np.transpose(pinv3_batch(np.ascontiguousarray(np.transpose(V_, (2, 0, 1)))), (1, 2, 0)))
    Y = np.zeros_like(W)
    for c in range(3):
        Y[:, c] = (W[:, 0] * Vinv[0, c][None, :, None] + W[:, 1] * Vinv[1, c][None, :, None]
                   + W[:, 2] * Vinv[2, c][None, :, None])
    S, e_ = stage2_ref_proj(Y, W, U_, eA, eB)
    da = pinv_matlab(S) @ e_
    db, a_new, b_new, X_hat_new = stage3_ref_proj(W, da, eB, Vinv, a, b, X, vis)
    e_new = X - X_hat_new
    old = float(np.dot(e.reshape(-1, order="F"), e.reshape(-1, order="F")))
    new = float(np.dot(e_new.reshape(-1, order="F"), e_new.reshape(-1, order="F")))
    blocks = dict(U=np.transpose(U, (2, 1, 0)).copy(), V=np.transpose(V, (2, 1, 0)).copy(),
                  eA=eA.T.copy(), eB=eB.T.copy(), Vinv=np.transpose(Vinv, (2, 1, 0)).copy(),
                  S=np.array(S), e_=e_, W_dense=W, s1=s1)
    return dict(old=old, new=new, da=np.array(da), db=np.array(db), a_new=np.array(a_new), b_new=np.array(b_new),
                blocks=blocks)


@dataclass
class ProjResult:
    Pp_: np.ndarray
    Xp_: np.ndarray
    error_: np.ndarray
    trials: list


def bundle_projective(Pp, Xp, x, *opts, record=True, max_iter=20, max_iter2=10):
    """[Pp_ Xp_ error_] = bundle_projective(Pp, Xp, x, ...)  (bundle_projective.m:1-229).
    Pp (3,4,m), Xp (4,n), x (3,n,m); options 'fix_structure', 'fix_motion', 'visibility', vis, 'verbose'."""
    Pp = np.asarray(Pp, dtype=np.float64); Xp = np.asarray(Xp, dtype=np.float64); x = np.asarray(x, dtype=np.float64)
    m = Pp.shape[2]; n = x.shape[1]
    fix_structure = fix_motion = False
    visible = ((x[0] != 0) | (x[1] != 0)).reshape(n, m)                     # :38
    opts = list(opts); k = 0
    while k < len(opts):
        key = str(opts[k]).lower()
        if key == "fix_structure": fix_structure = True
        elif key == "fix_motion": fix_motion = True
        elif key == "visibility": visible = np.asarray(opts[k + 1]); k += 1
        k += 1
    visible = np.asfortranarray(np.asarray(visible, dtype=np.float64).reshape(n, m))
    num_vis = float(visible.sum())
    a = np.asfortranarray(Pp.reshape(12, m, order="F"))                      # :69-72
    b = Xp[0:3].copy()
    X = np.asfortranarray(x[0:2])
    lam = 0.001                                                              # :87
    it, it2 = 1, 0
    error_ = []
    trials = []

    def cont():
        if not (it < max_iter and it2 < max_iter2):
            return False
        if it < 3:
            return True
        return error_[it - 1] > 1e-20 and error_[it - 2] - error_[it - 1] > 1e-3 * error_[it - 2]

    while cont():
        t = lm_trial_proj(a, b, X, visible, lam, fix_structure, fix_motion)
        old_error, new_error = t["old"] / num_vis, t["new"] / num_vis      # :181-182
        accept = new_error < old_error                                       # :187
        if record:
            trials.append(Trial(a.copy(), b.copy(), lam, 0.0, t["old"], t["new"], 0.0, bool(accept), t["da"], t["db"],
                                t["a_new"], t["b_new"], 0.0))
        if accept:
            a = np.array(t["a_new"]); b = np.array(t["b_new"])
            lam = lam / 10                                                   # :194
            while len(error_) < it + 1:
                error_.append(0.0)
            error_[it - 1] = old_error
            it += 1
            error_[it - 1] = new_error
            it2 = 0
        else:
            lam = lam * 10                                                   # :204
            it2 += 1
    Pp_ = np.asarray(a).reshape(3, 4, m, order="F").copy()
    return ProjResult(Pp_, np.vstack([b, Xp[3:4]]), np.array(error_), trials)


# ----------------------------------------------------------------------------------
# reprojection-error map: toolbox/test/error_reproj.m:1-88 and remove_outlier, incr_reconstruction.m:354-414
# ----------------------------------------------------------------------------------
def _rodr(w):
    """vl_rodr, the formula of oracle/shim/vl/rodrigues.h."""
    th = float(np.sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]))
    if th < 1e-6:
        return np.eye(3)
    x, y, z = w[0] / th, w[1] / th, w[2] / th
    s, c = np.sin(th), np.cos(th)
    mc = 1.0 - c
    return np.array([[1 - mc * (y * y + z * z), -s * z + mc * x * y, s * y + mc * x * z],
                     [s * z + mc * x * y, 1 - mc * (z * z + x * x), -s * x + mc * y * z],
                     [-s * y + mc * x * z, s * x + mc * y * z, 1 - mc * (x * x + y * y)]])


def error_reproj(x, K, T, w, X, vis=None):
    """[err error] = error_reproj(x, K, T, w, X, 'visibility', vis)  (error_reproj.m:1-88), K as 4 x m parameters.
    Restated literally, including its test `vis(n,m)` (the LAST cell, error_reproj.m:76) where `vis(i,j)` was meant:
    with the last cell visible every cell is evaluated, also the invisible ones."""
    m, n = T.shape[1], X.shape[1]
    if vis is None:
        vis = np.ones((n, m))
    error = np.zeros((n, m))
    for j in range(m):
        Kj = np.array([[K[0, j], 0, K[2, j]], [0, K[1, j], K[3, j]], [0, 0, 1]])      # calibration_matrix.m:9-11
        Pj = Kj @ np.hstack([_rodr(w[:, j]), T[:, j:j + 1]])
        for i in range(n):
            if X[3, i] == 1 and vis[n - 1, m - 1]:
                xr = Pj @ X[:, i]
                xr = xr / xr[2]
                error[i, j] = np.linalg.norm(x[0:2, i, j] - xr[0:2])
    return float(error.sum() / vis.sum()), error


def remove_outlier_stats(K, T, Omega, X, x, vis, depth_max=10.0):
    """The quantities remove_outlier() computes before it edits vis / X (incr_reconstruction.m:363-390):
    number of cells failing the depth test, largest squared reprojection error and its (point, frame), 0-based."""
    m, n = K.shape[1], X.shape[1]
    max_error, arg, negative = 0.0, (-1, -1), 0
    for j in range(m):
        Kj = np.array([[K[0, j], 0, K[2, j]], [0, K[1, j], K[3, j]], [0, 0, 1]])
        Pj = Kj @ np.hstack([_rodr(Omega[:, j]), T[:, j:j + 1]])
        for i in range(n):
            if X[3, i] == 1 and vis[i, j]:
                xr = Pj @ X[:, i]
                if xr[2] < 0 or xr[2] > depth_max:
                    negative += 1
                else:
                    xr = xr / xr[2]
                    e = float(np.linalg.norm(x[0:2, i, j] - xr[0:2]) ** 2)
                    if e > max_error:
                        max_error, arg = e, (i, j)
    return dict(n_bad_depth=negative, max_sq_err=max_error, argmax=arg)


# ----------------------------------------------------------------------------------
# CPU arm for the large configurations (port, PCG solve) -- see oracle_sparse.c
# ----------------------------------------------------------------------------------
def trial_step_pcg(K, a, b, obs: ObsList, lam, pcg_rtol=1e-8, pcg_max_iter=500):
    lib = sparse_lib()
    na, m = a.shape
    n = b.shape[1]
    a_new = np.zeros((m, na)); b_new = np.zeros((n, 3))
    costs = np.zeros(3); secs = np.zeros(4)
    it = lib.orc_trial_step_pcg(C.c_int(m), C.c_int(n), C.c_int(na), _d(_f64(K.T)), _d(_f64(a.T)), _d(_f64(b.T)),
                                C.c_long(obs.nobs), _d(obs.xy), _i(obs.pt), _i(obs.cam),
                                C.c_double(lam), C.c_double(pcg_rtol), C.c_int(pcg_max_iter),
                                _d(a_new), _d(b_new), _d(costs), _d(secs))
    return dict(a_new=a_new.T.copy(), b_new=b_new.T.copy(), old=costs[0], new=costs[1], denom=costs[2],
                pcg_iters=int(it), stage_seconds=secs)
