"""ctypes binding of libvlgba.so (include/vlg_ba.h).

This is plumbing only: every computation happens in the CUDA library.  There is no CPU
fallback -- if the shared library is missing or no CUDA device is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VLG_BA_LIB") or os.path.join(_HERE, "libvlgba.so")      # VLG_BA_LIB: an experimental build of the same library

SOLVER_AUTO, SOLVER_CHOL, SOLVER_PCG, SOLVER_PCG_EXPLICIT = 0, 1, 2, 3
MODEL_EUCLID, MODEL_PROJECTIVE = 0, 1
P2P_HANDLE_BYTES = 128      # VLG_BA_P2P_HANDLE_BYTES
RTABLE_HOST_LIBM, RTABLE_DEVICE = 0, 1
ORDER_CHUNKED, ORDER_REFERENCE = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class Opts(C.Structure):
    _fields_ = [
        ("num_variableK", C.c_int), ("fix_structure", C.c_int), ("fix_motion", C.c_int),
        ("lambda0", C.c_double), ("nu0", C.c_double), ("max_iter", C.c_int), ("max_iter2", C.c_int),
        ("rel_tol", C.c_double), ("abs_tol", C.c_double), ("backsub_all_rows", C.c_int),
        ("solver", C.c_int), ("chol_max_cams", C.c_int), ("pcg_rtol", C.c_double), ("pcg_max_iter", C.c_int),
        ("rtable", C.c_int), ("order", C.c_int), ("device", C.c_int), ("verbose", C.c_int),
        ("pcg_deflate", C.c_int), ("pcg_cluster", C.c_int), ("model", C.c_int), ("pcg_autotune", C.c_int),
    ]


class TrialInfo(C.Structure):
    _fields_ = [
        ("old_cost", C.c_double), ("new_cost", C.c_double), ("denom", C.c_double), ("rho", C.c_double),
        ("lambda_used", C.c_double), ("lambda_next", C.c_double), ("nu_next", C.c_double),
        ("accepted", C.c_int), ("solver_used", C.c_int), ("pcg_iters", C.c_int), ("pcg_relres", C.c_double),
        ("ms_stage1", C.c_float), ("ms_schur", C.c_float), ("ms_solve", C.c_float), ("ms_stage3", C.c_float),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class VlgBaError(RuntimeError):
    pass


_lib = None

# every symbol include/vlg_ba.h declares
SYMBOLS = [
    "vlg_ba_opts_default", "vlg_ba_version", "vlg_ba_create", "vlg_ba_destroy", "vlg_ba_last_error",
    "vlg_ba_nccl_unique_id", "vlg_ba_set_comm", "vlg_ba_p2p_export", "vlg_ba_p2p_import", "vlg_ba_set_problem_dense", "vlg_ba_set_problem_sparse",
    "vlg_ba_set_num_vis", "vlg_ba_nobs", "vlg_ba_get_obs", "vlg_ba_set_state", "vlg_ba_get_state",
    "vlg_ba_stage1", "vlg_ba_get_blocks", "vlg_ba_get_jacobians", "vlg_ba_stage2", "vlg_ba_get_reduced",
    "vlg_ba_set_da", "vlg_ba_stage3", "vlg_ba_get_update", "vlg_ba_trial_step", "vlg_ba_solve",
    "vlg_ba_trial_step_host", "vlg_ba_get_schur_structure", "vlg_ba_kernel_launches", "vlg_ba_kernel_time",
    "vlg_ba_reset_timers", "vlg_ba_timer_start", "vlg_ba_timer_stop", "vlg_ba_lm_reset", "vlg_ba_lm_continue",
    "vlg_ba_mex1_dense", "vlg_ba_mex2_dense", "vlg_ba_mex3_dense", "vlg_ba_bundle_euclid", "vlg_ba_bundle_euclid_sparse", "vlg_ba_bundle_projective", "vlg_ba_reproj_errors", "vlg_ba_symv_plan", "vlg_ba_selftest_quotients", "vlg_ba_symv_plan_occ", "vlg_ba_symv_bytes", "vlg_ba_solve_cameras_independent", "vlg_ba_dense_release", "vlg_ba_dense_cache_stats",
]


def lib():
    """Load libvlgba.so; fails loudly if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VlgBaError(f"{LIB_PATH} is missing: build it with `make -C {_HERE}/csrc` "
                             "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.vlg_ba_version.restype = C.c_char_p
        L.vlg_ba_last_error.restype = C.c_char_p
        L.vlg_ba_last_error.argtypes = [C.c_void_p]
        L.vlg_ba_nobs.restype = C.c_int64
        L.vlg_ba_nobs.argtypes = [C.c_void_p]
        L.vlg_ba_kernel_launches.restype = C.c_int64
        L.vlg_ba_kernel_launches.argtypes = [C.c_void_p]
        L.vlg_ba_destroy.argtypes = [C.c_void_p]
        L.vlg_ba_destroy.restype = None
        _lib = L
    return _lib


def _d(x):
    return None if x is None else x.ctypes.data_as(_dp)


def _i(x):
    return None if x is None else x.ctypes.data_as(_ip)


def _c(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def default_opts(**kw) -> Opts:
    o = Opts()
    lib().vlg_ba_opts_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown option {k}")
        setattr(o, k, v)
    return o


class Context:
    """One GPU context (vlg_ba_ctx).  Arrays cross this boundary in the reference's
    column-major layouts, passed as C-contiguous numpy arrays of the transposed shape:
    K (m,4), a (m,na), b (n,3), obs_xy (nobs,2), U (m,na,na), V (n,3,3), W (nobs,3,na)."""

    def __init__(self, **opts):
        self._L = lib()
        self.opts = default_opts(**opts)
        self._h = C.c_void_p()
        r = self._L.vlg_ba_create(C.byref(self.opts), C.byref(self._h))
        if r != 0:
            raise VlgBaError(f"vlg_ba_create failed ({r}): {self._L.vlg_ba_last_error(None).decode()}")
        self.m = self.n = 0
        self.na = 12 if self.opts.model == MODEL_PROJECTIVE else 6 + self.opts.num_variableK

    def close(self):
        if self._h:
            self._L.vlg_ba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, r):
        if r != 0:
            raise VlgBaError(f"libvlgba error {r}: {self._L.vlg_ba_last_error(self._h).decode()}")

    # ---- multi-GPU
    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        r = lib().vlg_ba_nccl_unique_id(buf)
        if r != 0:
            raise VlgBaError(f"vlg_ba_nccl_unique_id failed: {lib().vlg_ba_last_error(None).decode()}")
        return buf.raw

    def set_comm(self, rank: int, nranks: int, unique_id: bytes):
        self._ck(self._L.vlg_ba_set_comm(self._h, C.c_int(rank), C.c_int(nranks), C.c_char_p(unique_id)))

    def p2p_export(self) -> bytes:
        """P2P_HANDLE_BYTES of CUDA IPC handles (PCG mailbox, share of S) of this rank (after set_comm and set_problem_*)."""
        buf = C.create_string_buffer(P2P_HANDLE_BYTES)
        self._ck(self._L.vlg_ba_p2p_export(self._h, buf))
        return buf.raw

    def p2p_import(self, handles: bytes):
        """`handles`: the ranks' P2P_HANDLE_BYTES blobs concatenated in rank order."""
        self._ck(self._L.vlg_ba_p2p_import(self._h, C.c_char_p(handles)))

    # ---- problem
    def set_problem_sparse(self, K, a, b, obs_xy, obs_pt, obs_cam, pivot=None):
        a, b = _c(a), _c(b)
        K = None if K is None else _c(K)          # the projective model has no K
        m, n = a.shape[0], b.shape[0]
        assert a.shape == (m, self.na), (a.shape, m, self.na)
        obs_xy = _c(obs_xy)
        obs_pt = np.ascontiguousarray(obs_pt, dtype=np.int32)
        obs_cam = np.ascontiguousarray(obs_cam, dtype=np.int32)
        piv = None if pivot is None else _c(np.asarray(pivot, dtype=np.float64))
        self._ck(self._L.vlg_ba_set_problem_sparse(self._h, C.c_int(m), C.c_int(n), _d(K), _d(a), _d(b),
                                                   C.c_int64(obs_pt.shape[0]), _d(obs_xy), _i(obs_pt), _i(obs_cam),
                                                   _d(piv)))
        self.m, self.n = m, n

    def set_problem_dense(self, K, a, b, X, visible, pivot=None):
        """X (2,n,m) and visible (n,m) Fortran-ordered, as the reference's mex1 takes them."""
        a, b = _c(a), _c(b)
        K = None if K is None else _c(K)
        m, n = a.shape[0], b.shape[0]
        X = np.asfortranarray(X, dtype=np.float64)
        visible = np.asfortranarray(visible, dtype=np.float64)
        assert X.shape == (2, n, m) and visible.shape == (n, m)
        piv = None if pivot is None else _c(np.asarray(pivot, dtype=np.float64))
        self._ck(self._L.vlg_ba_set_problem_dense(self._h, C.c_int(m), C.c_int(n), _d(K), _d(a), _d(b), _d(X),
                                                  _d(visible), _d(piv)))
        self.m, self.n = m, n

    @property
    def nobs(self) -> int:
        return int(self._L.vlg_ba_nobs(self._h))

    def get_obs(self):
        no = self.nobs
        xy = np.zeros((no, 2)); pt = np.zeros(no, dtype=np.int32); cam = np.zeros(no, dtype=np.int32)
        self._ck(self._L.vlg_ba_get_obs(self._h, _d(xy), _i(pt), _i(cam)))
        return xy, pt, cam

    def set_state(self, a=None, b=None, lam=-1.0, nu=-1.0):
        a = None if a is None else _c(a)
        b = None if b is None else _c(b)
        self._ck(self._L.vlg_ba_set_state(self._h, _d(a), _d(b), C.c_double(lam), C.c_double(nu)))

    def get_state(self):
        a = np.zeros((self.m, self.na)); b = np.zeros((self.n, 3))
        lam, nu = C.c_double(), C.c_double()
        it, it2 = C.c_int(), C.c_int()
        self._ck(self._L.vlg_ba_get_state(self._h, _d(a), _d(b), C.byref(lam), C.byref(nu), C.byref(it), C.byref(it2)))
        return dict(a=a, b=b, lam=lam.value, nu=nu.value, iter=it.value, iter2=it2.value)

    def reproj_errors(self, depth_max: float = 10.0) -> dict:
        """error_reproj.m / remove_outlier statistics of the current state (vlg_ba_reproj_errors)."""
        no = self.nobs
        err = np.zeros(no); depth = np.zeros(no)
        mean, mx = C.c_double(), C.c_double()
        arg, nbad = C.c_int64(), C.c_int64()
        self._ck(self._L.vlg_ba_reproj_errors(self._h, C.c_double(depth_max), _d(err), _d(depth), C.byref(mean), C.byref(mx),
                                              C.byref(arg), C.byref(nbad)))
        return dict(err=err, depth=depth, mean_err=mean.value, max_sq_err=mx.value, argmax=arg.value, n_bad_depth=nbad.value)

    # ---- stages
    def stage1(self) -> float:
        cost = C.c_double()
        self._ck(self._L.vlg_ba_stage1(self._h, C.byref(cost)))
        return cost.value

    def get_blocks(self, want_W=True):
        na, m, n, no = self.na, self.m, self.n, self.nobs
        out = dict(U=np.zeros((m, na, na)), V=np.zeros((n, 3, 3)), eA=np.zeros((m, na)), eB=np.zeros((n, 3)))
        W = np.zeros((no, 3, na)) if want_W else None
        self._ck(self._L.vlg_ba_get_blocks(self._h, _d(out["U"]), _d(out["V"]), _d(W), _d(out["eA"]), _d(out["eB"])))
        out["W"] = W
        return out

    def get_jacobians(self):
        na, no = self.na, self.nobs
        out = dict(X_hat=np.zeros((no, 2)), A=np.zeros((no, na, 2)), B=np.zeros((no, 3, 2)), e=np.zeros((no, 2)))
        self._ck(self._L.vlg_ba_get_jacobians(self._h, _d(out["X_hat"]), _d(out["A"]), _d(out["B"]), _d(out["e"])))
        return out

    def stage2(self, lam: float):
        self._ck(self._L.vlg_ba_stage2(self._h, C.c_double(lam)))

    def get_reduced(self, want_S=False):
        na, m, n = self.na, self.m, self.n
        N = na * m
        out = dict(Vinv=np.zeros((n, 3, 3)), e_=np.zeros(N), da=np.zeros(N))
        S = np.zeros((N, N)) if want_S else None
        self._ck(self._L.vlg_ba_get_reduced(self._h, _d(out["Vinv"]), _d(S), _d(out["e_"]), _d(out["da"])))
        out["S"] = None if S is None else S.T.copy()      # column-major on the wire
        return out

    def set_da(self, da):
        da = _c(da)
        self._ck(self._L.vlg_ba_set_da(self._h, _d(da)))

    def stage3(self, lam: float):
        nc, dn = C.c_double(), C.c_double()
        self._ck(self._L.vlg_ba_stage3(self._h, C.c_double(lam), C.byref(nc), C.byref(dn)))
        return nc.value, dn.value

    def get_update(self):
        out = dict(db=np.zeros((self.n, 3)), a_new=np.zeros((self.m, self.na)), b_new=np.zeros((self.n, 3)))
        self._ck(self._L.vlg_ba_get_update(self._h, _d(out["db"]), _d(out["a_new"]), _d(out["b_new"])))
        return out

    def trial_step(self) -> dict:
        info = TrialInfo()
        self._ck(self._L.vlg_ba_trial_step(self._h, C.byref(info)))
        return info.as_dict()

    def trial_step_host(self, a, b, obs_xy, lam, a_new, b_new) -> dict:
        info = TrialInfo()
        self._ck(self._L.vlg_ba_trial_step_host(self._h, _d(a), _d(b), _d(obs_xy), C.c_double(lam), _d(a_new), _d(b_new),
                                                C.byref(info)))
        return info.as_dict()

    def lm_reset(self, a=None, b=None):
        a = None if a is None else _c(a)
        b = None if b is None else _c(b)
        self._ck(self._L.vlg_ba_lm_reset(self._h, _d(a), _d(b)))

    def lm_continue(self) -> bool:
        return bool(self._L.vlg_ba_lm_continue(self._h))

    def solve(self, Xe4=None):
        m, n = self.m, self.n
        K_ = np.zeros((m, 4)); Te_ = np.zeros((m, 3)); w_ = np.zeros((m, 3)); Xe_ = np.zeros((n, 4))
        err = np.zeros(max(self.opts.max_iter, 2) + 2)
        ne = C.c_int()
        x4 = None if Xe4 is None else _c(Xe4)
        self._ck(self._L.vlg_ba_solve(self._h, _d(K_), _d(Te_), _d(w_), _d(Xe_), _d(x4), _d(err), C.byref(ne)))
        return K_, Te_, w_, Xe_, err[:ne.value].copy()

    def solve_cameras_independent(self):
        """Per-camera LM loops side by side (fix_structure contexts): -> a (m, na), list of per-camera error_ arrays, rounds."""
        m, mi = self.m, self.opts.max_iter
        a = np.zeros((m, self.na)); err = np.zeros((m, mi)); ne = np.zeros(m, dtype=np.int32)
        rounds = C.c_int()
        self._ck(self._L.vlg_ba_solve_cameras_independent(self._h, _d(a), _d(err), _i(ne), C.byref(rounds)))
        return a, [err[j, :ne[j]].copy() for j in range(m)], rounds.value

    # ---- introspection
    def schur_structure(self):
        nb = C.c_int64()
        self._ck(self._L.vlg_ba_get_schur_structure(self._h, C.byref(nb), None, None))
        bj = np.zeros(nb.value, dtype=np.int32); bk = np.zeros(nb.value, dtype=np.int32)
        self._ck(self._L.vlg_ba_get_schur_structure(self._h, C.byref(nb), _i(bj), _i(bk)))
        return bj, bk

    @property
    def kernel_launches(self) -> int:
        return int(self._L.vlg_ba_kernel_launches(self._h))

    def reset_timers(self, enable=True):
        self._ck(self._L.vlg_ba_reset_timers(self._h, C.c_int(1 if enable else 0)))

    def timer_start(self):
        self._ck(self._L.vlg_ba_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._ck(self._L.vlg_ba_timer_stop(self._h, C.byref(ms)))
        return ms.value

    @property
    def symv_bytes(self) -> int:
        """Bytes of S one assembled-S matvec streams on this rank (kept tiles of the lower triangle)."""
        self._L.vlg_ba_symv_bytes.restype = C.c_int64
        self._L.vlg_ba_symv_bytes.argtypes = [C.c_void_p]
        return int(self._L.vlg_ba_symv_bytes(self._h))

    def kernel_time(self, name: str):
        ms, cnt = C.c_double(), C.c_int64()
        self._ck(self._L.vlg_ba_kernel_time(self._h, name.encode(), C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value


def dense_cache_stats():
    """(re-uses, rebuilds) of the cached contexts behind vlg_ba_mex{1,2,3}_dense on this thread."""
    h, b = C.c_int64(), C.c_int64()
    lib().vlg_ba_dense_cache_stats(C.byref(h), C.byref(b))
    return int(h.value), int(b.value)


def dense_release():
    lib().vlg_ba_dense_release()


def selftest_quotients(nsamples: int = 10**9, seed: int = 1, device: int = -1) -> int:
    """Bitwise mismatches between stage 1's shared-reciprocal quotients and __ddiv_rn over random operands (must be 0)."""
    L = lib()
    bad = C.c_int64()
    r = L.vlg_ba_selftest_quotients(C.c_int(device), C.c_int64(nsamples), C.c_uint64(seed), C.byref(bad))
    if r != 0:
        raise VlgBaError(f"vlg_ba_selftest_quotients: {L.vlg_ba_last_error(None).decode()}")
    return int(bad.value)


def symv_plan(Np: int, G: int, J0: int = 0, J1: int | None = None, speed=None, occ=None) -> dict:
    """Host-side plan of the assembled-S matvec (vlg_ba_symv_plan): tiles, per-CTA pieces and fold lists.  No GPU needed."""
    L = lib()
    if J1 is None:
        J1 = Np // 32
    sizes = (C.c_int64 * 8)()
    sp = None if speed is None else np.ascontiguousarray(speed, dtype=np.float64)
    spp = None if sp is None else sp.ctypes.data_as(C.POINTER(C.c_double))

    oc = None if occ is None else np.ascontiguousarray(occ, dtype=np.uint8)
    ocp = None if oc is None else oc.ctypes.data_as(C.POINTER(C.c_uint8))

    def call(*arrs):
        r = L.vlg_ba_symv_plan_occ(C.c_int(Np), C.c_int(G), C.c_int(J0), C.c_int(J1), spp, ocp, *arrs, sizes)
        if r != 0:
            raise VlgBaError(f"vlg_ba_symv_plan: {L.vlg_ba_last_error(None).decode()}")

    call(None, None, None, None, None, None)
    nt, nfrag, nrl, ncl, nrb = (int(sizes[k]) for k in range(5))
    out = dict(tiles=np.zeros((nt, 4), np.int32), tile_ptr=np.zeros(G + 1, np.int32), row_ptr=np.zeros(nrb + 1, np.int32),
               row_list=np.zeros(max(nrl, 1), np.int32), col_ptr=np.zeros(Np // 32 + 1, np.int32), col_list=np.zeros(max(ncl, 1), np.int32))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    call(ip(out["tiles"]), ip(out["tile_ptr"]), ip(out["row_ptr"]), ip(out["row_list"]), ip(out["col_ptr"]), ip(out["col_list"]))
    out.update(nfrag=nfrag, blk_rows=int(sizes[5]), slab=int(sizes[6]), ncell=int(sizes[7]))
    out["row_list"] = out["row_list"][:nrl]; out["col_list"] = out["col_list"][:ncl]
    return out

