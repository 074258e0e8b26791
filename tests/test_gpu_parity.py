"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden
vectors of the reference's own C.  Tolerances follow BASELINE.json's north_star:
  * per-observation quantities (X_hat, A, B, e, W), V, eB, visibility indexing and the block
    structure of S: bit-exact;
  * U, eA: bit-exact in VLG_BA_ORDER_REFERENCE, a few ulp in the default chunked order;
  * per-iteration cost, teacher-forced: 1e-9 relative (observed ~1e-13);
  * final reprojection RMS, free-running: 1e-6 relative.
"""
import numpy as np
import pytest

from bundleadjustmentmatlab_b200 import bundle, capi, synth
from oracle import lm

from common import golden_names, golden_opts, load_golden, oracle_options, rel, ulp_diff

pytestmark = pytest.mark.gpu

COST_RTOL = 1e-9       # north_star: per-iteration cost within 1e-9 relative (FP64)
RMS_RTOL = 1e-6        # north_star: final reprojection RMS within 1e-6 relative
# With free intrinsics (num_variableK = 1 or 4; NOT the reference's test path, which passes
# 'fix_calibration': mview_reconstruction.m:196) the reduced system has cond(S) ~ 1e10, so da itself
# is only defined to ~cond*eps ~ 1e-6 relative: the oracle's SVD pinv and an exact LU solve of the
# SAME S already differ by 3e-7 in da and 9e-8 in the one-step cost (measured on the goldens on
# the CPU).  Those cases are held to the conditioning limit instead of 1e-9.
COST_RTOL_FREE_K = 1e-6
RMS_RTOL_FREE_K = 1e-3


def cost_rtol(g):
    return COST_RTOL if int(g["num_variableK"]) == 0 else COST_RTOL_FREE_K


def ctx_from_golden(g, **kw):
    o = oracle_options(g)
    ctx = capi.Context(num_variableK=int(g["num_variableK"]), fix_structure=int(o["fix_structure"]),
                       fix_motion=int(o["fix_motion"]), **kw)
    a0 = np.ascontiguousarray(g["t_a"][0].T)
    b0 = np.ascontiguousarray(g["t_b"][0].T)
    piv = g["pivot"] if o["fix_pivot"] else None
    ctx.set_problem_dense(np.ascontiguousarray(g["K"].T), a0, b0, np.asfortranarray(g["x"][:2]),
                          np.asfortranarray(g["visible"]), pivot=piv)
    return ctx


@pytest.mark.parametrize("name", golden_names())
def test_visibility_indexing_bit_exact(name):
    g = load_golden(name)
    ctx = ctx_from_golden(g)
    xy, pt, cam = ctx.get_obs()
    assert np.array_equal(pt, g["obs_pt"]) and np.array_equal(cam, g["obs_cam"]) and np.array_equal(xy, g["obs_xy"])
    ctx.close()


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("rtable", [capi.RTABLE_HOST_LIBM])
def test_stage1_bit_exact_vs_reference_golden(name, rtable):
    g = load_golden(name)
    ctx = ctx_from_golden(g, order=capi.ORDER_REFERENCE, rtable=rtable)
    J = ctx.get_jacobians()
    for k in ("X_hat", "A", "B", "e"):
        assert np.array_equal(J[k], g[k]), f"{k}: {ulp_diff(J[k], g[k])} ulp"
    cost = ctx.stage1()
    blk = ctx.get_blocks()
    for k in ("W", "V", "eB", "U", "eA"):
        assert np.array_equal(blk[k], g[k]), f"{k}: {ulp_diff(blk[k], g[k])} ulp"
    assert rel(cost, float(g["t_old"][0])) <= 1e-13
    ctx.close()


@pytest.mark.parametrize("name", golden_names())
def test_stage1_chunked_order_within_ulps(name):
    g = load_golden(name)
    ctx = ctx_from_golden(g)
    ctx.stage1()
    blk = ctx.get_blocks()
    for k in ("W", "V", "eB"):
        assert np.array_equal(blk[k], g[k]), k
    scale = np.abs(g["U"]).max()
    assert np.abs(blk["U"] - g["U"]).max() <= 1e-13 * scale
    assert np.abs(blk["eA"] - g["eA"]).max() <= 1e-13 * max(np.abs(g["eA"]).max(), 1.0)
    ctx.close()


@pytest.mark.parametrize("name", golden_names())
def test_stage2_stage3_vs_golden(name):
    g = load_golden(name)
    ctx = ctx_from_golden(g, solver=capi.SOLVER_CHOL)
    ctx.stage1()
    lam = float(g["t_lam"][0])
    ctx.stage2(lam)
    red = ctx.get_reduced(want_S=True)
    # V*^-1: Cholesky-with-elimination vs the oracle's SVD pinv
    assert np.abs(red["Vinv"] - g["Vinv"]).max() <= 1e-9 * np.abs(g["Vinv"]).max()
    sS = np.abs(g["S"]).max()
    assert np.abs(red["S"] - g["S"]).max() <= 1e-11 * sS
    assert np.abs(red["e_"] - g["e_"]).max() <= 1e-11 * max(np.abs(g["e_"]).max(), 1.0)
    # block-sparsity structure of S, bit-exact: cameras sharing a point (plus the diagonal)
    na = ctx.na
    bj, bk = ctx.schur_structure()
    m = int(g["m"])
    nz = np.zeros((m, m), dtype=bool)
    Sg = g["S"]
    for j in range(m):
        for k in range(m):
            nz[j, k] = np.any(Sg[na * j:na * j + na, na * k:na * k + na] != 0)
    vis = g["visible"] != 0
    share = (vis.T.astype(np.int64) @ vis.astype(np.int64)) > 0
    vis_blocks = {(j, k) for j in range(m) for k in range(j, m) if share[j, k] or j == k}
    mine = set(zip(bj.tolist(), bk.tolist()))
    assert mine == vis_blocks
    # the reference's dense S is exactly zero outside that structure ...
    ref_nz = {(min(j, k), max(j, k)) for j in range(m) for k in range(m) if nz[j, k]}
    assert ref_nz <= mine
    # ... and non-zero everywhere inside it unless a fix_* option zeroed W (bundle_euclid.m:140-154)
    if not g["options"] or g["options"] == ["fix_calibration"] or g["options"] == ["fix_principal"]:
        assert ref_nz == mine
    # teacher-forced stage 3: feed the reference's own da
    ctx.set_da(g["t_da"][0])
    new_cost, denom = ctx.stage3(lam)
    up = ctx.get_update()
    assert np.abs(up["db"] - g["t_db"][0].T).max() <= 1e-9 * max(np.abs(g["t_db"][0]).max(), 1e-30)
    assert np.array_equal(up["a_new"], g["t_a_new"][0].T)
    assert rel(new_cost, float(g["t_new"][0])) <= COST_RTOL      # the reference's own da: no conditioning excuse
    assert rel(denom, float(g["t_denom"][0])) <= 1e-7
    ctx.close()


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("solver", [capi.SOLVER_CHOL, capi.SOLVER_PCG, capi.SOLVER_PCG_EXPLICIT])
def test_teacher_forced_cost_trajectory(name, solver):
    """Every iteration of the reference's trajectory, restarted from the reference's exact
    (a, b, lambda): old cost, new cost, accept decision."""
    g = load_golden(name)
    ctx = ctx_from_golden(g, solver=solver, pcg_rtol=1e-12)
    worst = 0.0
    for k in range(len(g["t_lam"])):
        ctx.set_state(a=g["t_a"][k].T, b=g["t_b"][k].T, lam=float(g["t_lam"][k]), nu=float(g["t_nu"][k]))
        info = ctx.trial_step()
        assert rel(info["old_cost"], float(g["t_old"][k])) <= 1e-12
        r = rel(info["new_cost"], float(g["t_new"][k]))
        worst = max(worst, r)
        assert r <= cost_rtol(g), (k, r, info)
        margin = abs(float(g["t_old"][k]) - float(g["t_new"][k])) / float(g["t_old"][k])
        if margin > 10 * cost_rtol(g):          # a decision closer than the tolerance may flip
            assert bool(info["accepted"]) == bool(g["t_accept"][k])
        if info["accepted"]:
            assert abs(info["rho"] - float(g["t_rho"][k])) <= 1e3 * cost_rtol(g) * max(abs(float(g["t_rho"][k])), 1.0)
    print(f"{name}: worst teacher-forced relative cost deviation {worst:.2e}")
    ctx.close()


@pytest.mark.parametrize("name", golden_names())
def test_free_running_solve_vs_golden(name):
    g = load_golden(name)
    K_, Te_, w_, Xe_, err = bundle.bundle_euclid(g["K"], g["Te"], g["w"], g["Xe"], g["x"], *golden_opts(g))
    ref = g["error_"]
    assert len(err) >= 2 and rel(err[0], ref[0]) <= 1e-12
    rms, rms_ref = np.sqrt(err[-1]), np.sqrt(ref[-1])
    print(f"{name}: iterations {len(err)} vs {len(ref)}, final RMS rel diff {rel(rms, rms_ref):.2e}")
    if len(err) == len(ref):
        assert rel(rms, rms_ref) <= (RMS_RTOL if int(g["num_variableK"]) == 0 else RMS_RTOL_FREE_K)
    elif int(g["num_variableK"]) == 0:
        # the stop rule (bundle_euclid.m:120-123) sits on a 1e-3 relative-decrease threshold; a
        # flipped iteration count is reported, and the RMS must still agree to the decrease scale
        assert rel(rms, rms_ref) <= 2e-3
    else:
        # free intrinsics: cond(S) ~ 1e10 makes the free-running trajectory chaotic after the first
        # accepted step (two exact solvers of the same S already part ways at the 1 % level in the
        # second step's cost; tools/dbg_fullk.py prints it), so an accept decision sitting on the
        # stop rule's 1e-3 threshold can end the solve early.  What holds: the first step agrees to
        # the conditioning limit and the accepted costs descend.
        assert rel(err[1], ref[1]) <= COST_RTOL_FREE_K and np.all(np.diff(err) < 0)
    assert Xe_.shape == g["Xe_"].shape and K_.shape == g["K_"].shape


def test_fast_quotients_are_ieee_quotients():
    """Stage 1 replaces the reference's 34 divisions per observation by 8 correctly rounded reciprocals and one exact
    remainder step per quotient (ba_math.cuh).  That is only legitimate if every quotient is still the IEEE quotient:
    1e9 random operand pairs (generic, reprojection-like, forward-difference-like) against __ddiv_rn, bitwise."""
    for seed in (1, 2):
        assert capi.selftest_quotients(10**9, seed=seed) == 0


def test_repeated_steps_are_bit_identical_on_every_solver():
    """The hand-rolled synchronisation (mbarrier rings, grid barriers, cp.async pipelines, cooperative kernels) has no
    race-detector record -- compute-sanitizer is closed on the GPU pool (profiles/compute_sanitizer_r02_refused.txt).  What a
    race in them would break first is run-to-run reproducibility: every reduction here has a fixed shape, so forty repeats
    of the same LM trial step must give the same bits -- costs, iteration counts and the whole step da."""
    P = synth.make_problem(130, 14000, 66000, seed=41)
    a = np.ascontiguousarray(np.vstack([P.w, P.Te]).T); b = np.ascontiguousarray(P.Xe[:3].T)
    for solver in (capi.SOLVER_CHOL, capi.SOLVER_PCG, capi.SOLVER_PCG_EXPLICIT):
        ctx = capi.Context(num_variableK=0, solver=solver)
        ctx.set_problem_sparse(P.K.T, a, b, P.obs_xy, P.obs_pt, P.obs_cam)
        first = None
        for rep in range(40):
            ctx.set_state(a=a, b=b, lam=1e-3, nu=2.0)
            ctx.stage1(); ctx.stage2(1e-3)
            da = ctx.get_reduced()["da"].copy()
            new_cost, denom = ctx.stage3(1e-3)
            cur = (new_cost, denom, da.tobytes())
            if first is None:
                first = cur
            assert cur == first, (solver, rep)
        ctx.close()


def test_banded_scene_streams_only_the_band_of_S():
    """A scene whose tracks are contiguous camera windows (no loop closures) has a banded reduced system: the assembled-S
    matvec must skip the 256 x 32 tiles without a non-zero block and still give the direct
    solver's step."""
    P = synth.make_problem(420, 30000, 150000, seed=19, banded=True)
    a = np.ascontiguousarray(np.vstack([P.w, P.Te]).T); b = np.ascontiguousarray(P.Xe[:3].T)
    out = {}
    for solver in (capi.SOLVER_CHOL, capi.SOLVER_PCG_EXPLICIT):
        ctx = capi.Context(num_variableK=0, solver=solver, pcg_rtol=1e-11)
        ctx.set_problem_sparse(P.K.T, a, b, P.obs_xy, P.obs_pt, P.obs_cam)
        out[solver] = ctx.trial_step()
        if solver == capi.SOLVER_PCG_EXPLICIT:
            Np = (6 * P.m + 31) // 32 * 32
            dense = 4 * Np * (Np + 32)
            print(f"banded scene: the matvec streams {ctx.symv_bytes / 1e6:.2f} MB of S per product, the whole lower triangle is {dense / 1e6:.2f} MB")
            assert 0 < ctx.symv_bytes < 0.6 * dense
        ctx.close()
    assert out[capi.SOLVER_CHOL]["old_cost"] == out[capi.SOLVER_PCG_EXPLICIT]["old_cost"]
    assert rel(out[capi.SOLVER_PCG_EXPLICIT]["new_cost"], out[capi.SOLVER_CHOL]["new_cost"]) <= COST_RTOL


def test_medium_problem_vs_sparse_oracle():
    P = synth.make_problem(30, 3000, 13000, seed=4)
    a = np.vstack([P.w, P.Te]); b = P.Xe[:3].copy()
    obs = lm.ObsList(P.m, P.n, P.obs_xy, P.obs_pt, P.obs_cam)
    o = lm.parse_options(P.m, P.n, np.ones((2, 1, 1)), ["fix_calibration", "visibility", np.zeros(0)])
    t = lm.lm_trial(P.K, a, b, obs, 1e-3, o, backend="sparse")
    for solver in (capi.SOLVER_CHOL, capi.SOLVER_PCG, capi.SOLVER_PCG_EXPLICIT):
        ctx = capi.Context(num_variableK=0, solver=solver, pcg_rtol=1e-12)
        ctx.set_problem_sparse(P.K.T, a.T, b.T, P.obs_xy, P.obs_pt, P.obs_cam)
        J = ctx.get_jacobians()
        s1 = t["blocks"]["s1"]
        for k in ("X_hat", "A", "B", "e"):
            assert np.array_equal(J[k], s1[k]), k
        info = ctx.trial_step()
        blk_ok = rel(info["old_cost"], t["old"]) <= 1e-12 and rel(info["new_cost"], t["new"]) <= COST_RTOL
        assert blk_ok, (solver, info, t["old"], t["new"])
        ctx.close()


def test_default_pcg_tolerance_meets_the_cost_bar():
    """Library defaults (pcg_rtol = 1e-8): the one-step cost stays within 1e-9 of the oracle."""
    P = synth.make_problem(40, 4000, 18000, seed=12)
    a = np.vstack([P.w, P.Te]); b = P.Xe[:3].copy()
    obs = lm.ObsList(P.m, P.n, P.obs_xy, P.obs_pt, P.obs_cam)
    o = lm.parse_options(P.m, P.n, np.ones((2, 1, 1)), ["fix_calibration", "visibility", np.zeros(0)])
    ctx = capi.Context(num_variableK=0, solver=capi.SOLVER_PCG)
    ctx.set_problem_sparse(P.K.T, a.T, b.T, P.obs_xy, P.obs_pt, P.obs_cam)
    aa, bb, lam = a, b, 1e-3
    for k in range(3):
        t = lm.lm_trial(P.K, aa, bb, obs, lam, o, backend="sparse")
        ctx.set_state(a=aa.T, b=bb.T, lam=lam, nu=2.0)
        info = ctx.trial_step()
        r = rel(info["new_cost"], t["new"])
        print(f"trial {k}: lambda {lam:.2e}, PCG iterations {info['pcg_iters']}, cost deviation {r:.2e}")
        assert r <= COST_RTOL and rel(info["old_cost"], t["old"]) <= 1e-12
        rho = (t["old"] - t["new"]) / t["denom"]
        aa, bb, lam = t["a_new"], t["b_new"], lam * max(1 / 3, 1 - (2 * rho - 1) ** 3)
    ctx.close()


def test_cluster_preconditioner_same_answer_fewer_iterations():
    """Explicit-S PCG: the cluster-Jacobi preconditioner (inverses of 21-camera diagonal blocks of S)
    must give the same step as per-camera blocks and the direct solve, in fewer iterations."""
    P = synth.make_problem(90, 9000, 42000, seed=15)
    a = np.vstack([P.w, P.Te]); b = P.Xe[:3].copy()
    out = {}
    for cl in (0, 1):
        ctx = capi.Context(num_variableK=0, solver=capi.SOLVER_PCG_EXPLICIT, pcg_cluster=cl, pcg_rtol=1e-10)
        ctx.set_problem_sparse(P.K.T, a.T, b.T, P.obs_xy, P.obs_pt, P.obs_cam)
        out[cl] = ctx.trial_step()
        ctx.close()
    ctx = capi.Context(num_variableK=0, solver=capi.SOLVER_CHOL)
    ctx.set_problem_sparse(P.K.T, a.T, b.T, P.obs_xy, P.obs_pt, P.obs_cam)
    exact = ctx.trial_step()
    ctx.close()
    print("PCG iterations per-camera/cluster blocks:", out[0]["pcg_iters"], out[1]["pcg_iters"])
    for cl in (0, 1):
        assert rel(out[cl]["new_cost"], exact["new_cost"]) <= COST_RTOL
    assert out[1]["pcg_iters"] < out[0]["pcg_iters"]


def test_gauge_deflation_same_answer_fewer_iterations():
    P = synth.make_problem(60, 6000, 28000, seed=14)
    a = np.vstack([P.w, P.Te]); b = P.Xe[:3].copy()
    out = {}
    for defl in (0, 1):
        ctx = capi.Context(num_variableK=0, solver=capi.SOLVER_PCG, pcg_deflate=defl, lambda0=1e-5)
        ctx.set_problem_sparse(P.K.T, a.T, b.T, P.obs_xy, P.obs_pt, P.obs_cam)
        out[defl] = ctx.trial_step()
        ctx.close()
    ctx = capi.Context(num_variableK=0, solver=capi.SOLVER_CHOL, lambda0=1e-5)
    ctx.set_problem_sparse(P.K.T, a.T, b.T, P.obs_xy, P.obs_pt, P.obs_cam)
    exact = ctx.trial_step()
    ctx.close()
    print("PCG iterations plain/deflated:", out[0]["pcg_iters"], out[1]["pcg_iters"],
          "cost deviation from the direct solve:", rel(out[0]["new_cost"], exact["new_cost"]), rel(out[1]["new_cost"], exact["new_cost"]))
    # at lambda = 1e-5 plain PCG leaves the gauge component of da under-resolved (second-order effect
    # on the cost); the deflated solve resolves it exactly
    assert rel(out[1]["new_cost"], exact["new_cost"]) <= COST_RTOL
    assert rel(out[0]["new_cost"], exact["new_cost"]) <= 1e-7
    assert out[1]["pcg_iters"] < out[0]["pcg_iters"]
    # fixed cameras must keep da = 0 with deflation on (pinv semantics of the zero rows of S)
    piv = np.zeros(P.m); piv[:2] = 1
    ctx = capi.Context(num_variableK=0, solver=capi.SOLVER_PCG, pcg_deflate=1)
    ctx.set_problem_sparse(P.K.T, a.T, b.T, P.obs_xy, P.obs_pt, P.obs_cam, pivot=piv)
    ctx.stage1(); ctx.stage2(1e-3)
    da = ctx.get_reduced()["da"].reshape(P.m, 6)
    assert np.all(da[:2] == 0) and np.any(da[2:] != 0)
    ctx.close()


def test_device_rtable_mode_reports_mismatch():
    """VLG_BA_RTABLE_DEVICE uses CUDA's sin/cos: count rotation-table entries that differ from
    the host libm table and the resulting cost deviation (reported, loose bound)."""
    P = synth.make_problem(49, 2000, 9000, seed=6)
    a = np.vstack([P.w, P.Te]); b = P.Xe[:3].copy()
    costs = []
    for rt in (capi.RTABLE_HOST_LIBM, capi.RTABLE_DEVICE):
        ctx = capi.Context(num_variableK=0, rtable=rt, solver=capi.SOLVER_CHOL)
        ctx.set_problem_sparse(P.K.T, a.T, b.T, P.obs_xy, P.obs_pt, P.obs_cam)
        info = ctx.trial_step()
        costs.append((info["old_cost"], info["new_cost"]))
        ctx.close()
    print("device-rtable vs host-libm: old", rel(costs[1][0], costs[0][0]), "new", rel(costs[1][1], costs[0][1]))
    assert rel(costs[1][0], costs[0][0]) <= 1e-12
    assert rel(costs[1][1], costs[0][1]) <= 1e-3


def test_rejected_step_path_and_empty_error():
    """Bad initialisation: the first trial steps are rejected (iter2 path, bundle_euclid.m:233-241);
    lambda grows by nu, nu doubles."""
    P = synth.make_problem(6, 60, 300, seed=8)
    a = np.vstack([P.w, P.Te]); b = P.Xe[:3].copy()
    ctx = capi.Context(num_variableK=0, lambda0=1e-12)
    ctx.set_problem_sparse(P.K.T, a.T, b.T, P.obs_xy, P.obs_pt, P.obs_cam)
    lam, nu = 1e-12, 2.0
    for _ in range(3):
        info = ctx.trial_step()
        st = ctx.get_state()
        if info["accepted"]:
            assert st["nu"] == 2.0 and st["iter2"] == 0
        else:
            assert st["lam"] == lam * nu and st["nu"] == 2 * nu
        lam, nu = st["lam"], st["nu"]
    ctx.close()


def test_empty_and_ragged_inputs():
    # a point with no observation, a camera with no observation
    P = synth.make_problem(5, 40, 160, seed=2)
    keep = (P.obs_pt != 7) & (P.obs_cam != 3)
    a = np.vstack([P.w, P.Te]); b = P.Xe[:3].copy()
    obs = lm.ObsList(P.m, P.n, np.ascontiguousarray(P.obs_xy[keep]), P.obs_pt[keep], P.obs_cam[keep])
    o = lm.parse_options(P.m, P.n, np.ones((2, 1, 1)), ["fix_calibration", "visibility", np.zeros(0)])
    t = lm.lm_trial(P.K, a, b, obs, 1e-3, o, backend="sparse")
    for solver in (capi.SOLVER_CHOL, capi.SOLVER_PCG, capi.SOLVER_PCG_EXPLICIT):
        ctx = capi.Context(num_variableK=0, solver=solver, pcg_rtol=1e-12)
        ctx.set_problem_sparse(P.K.T, a.T, b.T, obs.xy, obs.pt, obs.cam)
        info = ctx.trial_step()
        assert rel(info["old_cost"], t["old"]) <= 1e-12 and rel(info["new_cost"], t["new"]) <= COST_RTOL, (solver, info)
        ctx.close()
    # out-of-order list is refused
    ctx = capi.Context(num_variableK=0)
    with pytest.raises(capi.VlgBaError):
        ctx.set_problem_sparse(P.K.T, a.T, b.T, obs.xy[::-1].copy(), obs.pt[::-1].copy(), obs.cam[::-1].copy())
    ctx.close()


def test_incremental_cadence_context_reuse():
    """The reference's heaviest caller adds one camera at a time and runs BA on the growing problem
    (incr_reconstruction.m:223-348, about 3 bundle_euclid calls per camera; estimate_camera.m:247-253 is a
    single-camera motion-only BA).  One context is re-used for every call -- set_problem_* on a live context
    must rebuild every structure (solver path, tile lists, optional buffers) -- and each call must match the
    CPU oracle run on the same sub-problem."""
    P = synth.make_problem(12, 400, 2600, seed=21)
    a_all = np.vstack([P.w, P.Te]); b_all = P.Xe[:3].copy()
    o = lm.parse_options(P.m, P.n, np.ones((2, 1, 1)), ["fix_calibration", "visibility", np.zeros(0)])
    ctx = capi.Context(num_variableK=0)
    ctx_fs = capi.Context(num_variableK=0, fix_structure=1)
    for mc in (2, 3, 5, 8, 12):
        keep = P.obs_cam < mc
        obs = lm.ObsList(mc, P.n, np.ascontiguousarray(P.obs_xy[keep]), P.obs_pt[keep], P.obs_cam[keep])
        a = a_all[:, :mc].copy()
        # (i) motion-only BA of the newest camera alone (fix_structure, one camera)
        k1 = P.obs_cam[keep] == mc - 1
        obs1 = lm.ObsList(1, P.n, np.ascontiguousarray(obs.xy[k1]), obs.pt[k1], np.zeros(int(k1.sum()), dtype=np.int32))
        o1 = dict(o); o1["fix_structure"] = True
        t1 = lm.lm_trial(P.K[:, mc - 1:mc], a[:, mc - 1:mc], b_all, obs1, 1e-3, o1, backend="sparse")
        ctx_fs.set_problem_sparse(P.K.T[mc - 1:mc], a.T[mc - 1:mc], b_all.T, obs1.xy, obs1.pt, obs1.cam)
        i1 = ctx_fs.trial_step()
        assert rel(i1["old_cost"], t1["old"]) <= 1e-12 and rel(i1["new_cost"], t1["new"]) <= COST_RTOL, (mc, i1, t1["new"])
        # (ii) full BA over the cameras so far, same context as the previous (smaller) problem
        t = lm.lm_trial(P.K[:, :mc], a, b_all, obs, 1e-3, o, backend="sparse")
        ctx.set_problem_sparse(P.K.T[:mc], a.T, b_all.T, obs.xy, obs.pt, obs.cam)
        info = ctx.trial_step()
        assert rel(info["old_cost"], t["old"]) <= 1e-12 and rel(info["new_cost"], t["new"]) <= COST_RTOL, (mc, info, t["new"])
    ctx.close(); ctx_fs.close()


def test_large_problem_invariants():
    """Trafalgar-shaped problem (BASELINE.json config 3) -- size-independent properties: the
    first LM step is accepted, the cost decreases monotonically over accepted steps, Cholesky and
    PCG agree teacher-forced, and two runs are bit-identical (deterministic reductions)."""
    P = synth.make_config("trafalgar", seed=0)
    a = np.vstack([P.w, P.Te]); b = P.Xe[:3].copy()
    news = {}
    for solver in (capi.SOLVER_CHOL, capi.SOLVER_PCG, capi.SOLVER_PCG_EXPLICIT):
        ctx = capi.Context(num_variableK=0, solver=solver, pcg_rtol=1e-10)
        ctx.set_problem_sparse(P.K.T, a.T, b.T, P.obs_xy, P.obs_pt, P.obs_cam)
        i1 = ctx.trial_step()
        assert i1["accepted"] and i1["new_cost"] < i1["old_cost"]
        assert i1["solver_used"] == solver
        news[solver] = i1["new_cost"]
        if solver != capi.SOLVER_CHOL:
            ctx.set_state(a=a.T, b=b.T, lam=1e-3, nu=2.0)
            i2 = ctx.trial_step()
            assert i2["new_cost"] == i1["new_cost"] and i2["old_cost"] == i1["old_cost"]
        ctx.close()
    assert rel(news[capi.SOLVER_PCG], news[capi.SOLVER_CHOL]) <= COST_RTOL
    assert rel(news[capi.SOLVER_PCG_EXPLICIT], news[capi.SOLVER_CHOL]) <= COST_RTOL


def test_venice_shape_invariants():
    """Venice-shaped problem (BASELINE.json config 4, the bench workload: 1 778 cameras, 994 k points,
    5.0 M observations) -- size-independent properties at full size: AUTO picks the assembled-S PCG, the
    first LM step is accepted and lowers the cost, the assembled and the implicit Schur operators give
    the same step (teacher-forced, 1e-9), and a repeated step is bit-identical (deterministic kernels)."""
    P = synth.make_config("venice", seed=0)
    a = np.ascontiguousarray(np.vstack([P.w, P.Te]).T); b = np.ascontiguousarray(P.Xe[:3].T)
    news = {}
    for solver in (capi.SOLVER_AUTO, capi.SOLVER_PCG):
        ctx = capi.Context(num_variableK=0, solver=solver, pcg_rtol=1e-10)
        ctx.set_problem_sparse(P.K.T, a, b, P.obs_xy, P.obs_pt, P.obs_cam)
        i1 = ctx.trial_step()
        assert i1["accepted"] and i1["new_cost"] < i1["old_cost"]
        if solver == capi.SOLVER_AUTO:
            assert i1["solver_used"] == capi.SOLVER_PCG_EXPLICIT
            ctx.set_state(a=a, b=b, lam=1e-3, nu=2.0)
            i2 = ctx.trial_step()
            assert i2["new_cost"] == i1["new_cost"] and i2["old_cost"] == i1["old_cost"] and i2["pcg_iters"] == i1["pcg_iters"]
        else:
            assert i1["solver_used"] == capi.SOLVER_PCG
        news[solver] = (i1["old_cost"], i1["new_cost"], i1["pcg_iters"])
        ctx.close()
    print("venice: (old, new, PCG iterations) assembled / implicit:", news[capi.SOLVER_AUTO], news[capi.SOLVER_PCG])
    assert news[capi.SOLVER_AUTO][0] == news[capi.SOLVER_PCG][0]
    assert rel(news[capi.SOLVER_AUTO][1], news[capi.SOLVER_PCG][1]) <= COST_RTOL


def test_many_clusters_implicit_path():
    """4 000 cameras = 667 preconditioner clusters: more update CTAs than can be co-resident with the
    register-hungry kernel variants (cooperative launches must fall back to the light variants, as at Final
    shape with 13 682 cameras).  Cluster-Jacobi on/off must give the same step."""
    P = synth.make_problem(4000, 40000, 200000, seed=23)
    a = np.ascontiguousarray(np.vstack([P.w, P.Te]).T); b = np.ascontiguousarray(P.Xe[:3].T)
    out = {}
    for cl in (0, 1):
        ctx = capi.Context(num_variableK=0, pcg_cluster=cl, pcg_rtol=1e-10)
        ctx.set_problem_sparse(P.K.T, a, b, P.obs_xy, P.obs_pt, P.obs_cam)
        out[cl] = ctx.trial_step()
        assert out[cl]["solver_used"] == capi.SOLVER_PCG
        ctx.close()
    print("many clusters: PCG iterations per-camera/cluster blocks:", out[0]["pcg_iters"], out[1]["pcg_iters"])
    assert out[0]["old_cost"] == out[1]["old_cost"]
    assert rel(out[0]["new_cost"], out[1]["new_cost"]) <= COST_RTOL


def test_autotuned_matvec_cut_same_steps():
    """opts.pcg_autotune re-cuts the assembled-S matvec by the measured per-SM rate after each of the first solves: the
    steps must not change beyond the PCG tolerance (teacher-forced from the equal-cut trajectory), and the overlapping
    preconditioner switched off (VLG_BA_OVERLAP=0 semantics are covered by the solver-variant tests) is not needed here."""
    P = synth.make_problem(700, 60000, 330000, seed=31)
    a = np.ascontiguousarray(np.vstack([P.w, P.Te]).T); b = np.ascontiguousarray(P.Xe[:3].T)
    ref = capi.Context(num_variableK=0, solver=capi.SOLVER_PCG_EXPLICIT, pcg_rtol=1e-11)
    ref.set_problem_sparse(P.K.T, a, b, P.obs_xy, P.obs_pt, P.obs_cam)
    tun = capi.Context(num_variableK=0, solver=capi.SOLVER_PCG_EXPLICIT, pcg_rtol=1e-11, pcg_autotune=3)
    tun.set_problem_sparse(P.K.T, a, b, P.obs_xy, P.obs_pt, P.obs_cam)
    worst = 0.0
    for _ in range(5):
        st = ref.get_state()
        i0 = ref.trial_step()
        tun.set_state(a=st["a"], b=st["b"], lam=st["lam"], nu=st["nu"])
        i1 = tun.trial_step()
        assert i1["solver_used"] == capi.SOLVER_PCG_EXPLICIT
        assert i0["old_cost"] == i1["old_cost"]
        worst = max(worst, rel(i0["new_cost"], i1["new_cost"]))
    print(f"autotuned cut: worst relative cost deviation over 5 steps {worst:.2e}")
    assert worst <= COST_RTOL
    ref.close(); tun.close()


def test_long_tracks_fall_back_to_untiled_kernels():
    """Points seen by more than 512 cameras do not fit a point tile: the stage-1 point pass, the back-substitution
    and the implicit PCG point sweep fall back to their thread-per-point forms.  All three solvers must still agree
    with each other and with the tiled result structure (first step accepted, same cost)."""
    rng = np.random.default_rng(5)
    P = synth.make_problem(640, 300, 5000, seed=9)
    # rebuild the observation list so that every point is seen by every camera (tracks of 640)
    a = np.ascontiguousarray(np.vstack([P.w, P.Te]).T); b = np.ascontiguousarray(P.Xe[:3].T)
    m, n = P.m, P.n
    cam = np.repeat(np.arange(m, dtype=np.int32), n)
    pt = np.tile(np.arange(n, dtype=np.int32), m)
    # observations = exact projections of the initial estimate + noise (so the problem is well posed)
    ctx = capi.Context(num_variableK=0, solver=capi.SOLVER_CHOL)
    xy0 = np.zeros((m * n, 2))
    ctx.set_problem_sparse(P.K.T, a, b, xy0, pt, cam)
    xhat = ctx.get_jacobians()["X_hat"]
    ctx.close()
    xy = np.ascontiguousarray(xhat + rng.normal(0.0, 0.5, xhat.shape))
    res = {}
    for solver in (capi.SOLVER_CHOL, capi.SOLVER_PCG, capi.SOLVER_PCG_EXPLICIT):
        ctx = capi.Context(num_variableK=0, solver=solver, pcg_rtol=1e-11)
        ctx.set_problem_sparse(P.K.T, a, b, xy, pt, cam)
        info = ctx.trial_step()
        assert info["solver_used"] == solver and info["accepted"], (solver, info)
        res[solver] = info
        ctx.close()
    for solver in (capi.SOLVER_PCG, capi.SOLVER_PCG_EXPLICIT):
        assert res[solver]["old_cost"] == res[capi.SOLVER_CHOL]["old_cost"]
        assert rel(res[solver]["new_cost"], res[capi.SOLVER_CHOL]["new_cost"]) <= COST_RTOL, (solver, res)


def test_reprojection_error_map_vs_error_reproj():
    """vlg_ba_reproj_errors against the restated error_reproj.m / remove_outlier
    (oracle/lm.py).  Fully visible problem, so that error_reproj.m's `vis(n,m)` slip (it tests the last cell instead of
    cell (i,j)) does not matter; a few points are pushed behind their cameras to exercise the depth test."""
    P = synth.make_problem(5, 40, 200, seed=17)
    m, n = P.m, P.n
    a = np.ascontiguousarray(np.vstack([P.w, P.Te]).T); b = np.ascontiguousarray(P.Xe[:3].T).copy()
    b[3] *= -1.0; b[11] *= -1.0                                   # behind the cameras
    cam = np.repeat(np.arange(m, dtype=np.int32), n); pt = np.tile(np.arange(n, dtype=np.int32), m)
    rng = np.random.default_rng(3)
    xy = rng.uniform(0.0, 500.0, (m * n, 2))
    x = np.zeros((3, n, m)); x[0, pt, cam] = xy[:, 0]; x[1, pt, cam] = xy[:, 1]; x[2] = 1.0
    X4 = np.vstack([b.T, np.ones((1, n))])
    vis = np.ones((n, m))
    ctx = capi.Context(num_variableK=0)
    ctx.set_problem_sparse(P.K.T, a, b, xy, pt, cam)
    depth_max = 1e4
    g = ctx.reproj_errors(depth_max=depth_max)
    ctx.close()
    err_ref, emap = lm.error_reproj(x, P.K, P.Te, P.w, X4, vis)
    st = lm.remove_outlier_stats(P.K, P.Te, P.w, X4, x, vis, depth_max=depth_max)
    got = np.zeros((n, m)); got[pt, cam] = g["err"]
    assert np.abs(got - emap).max() <= 1e-10 * emap.max()
    assert rel(g["mean_err"], err_ref) <= 1e-12
    assert g["n_bad_depth"] == st["n_bad_depth"] and g["n_bad_depth"] >= 2
    assert rel(g["max_sq_err"], st["max_sq_err"]) <= 1e-10
    assert (int(pt[g["argmax"]]), int(cam[g["argmax"]])) == st["argmax"]



@pytest.mark.parametrize("name", ["euclid_fixcal", "euclid_ucla4_fixcal"])
def test_reprojection_error_map_vs_reference_built_reprojections(name):
    """Row N4 pinned on reference-run data: the goldens' X_hat is the output of the reference's own mex1 (oracle/_ref), so
    ||x - X_hat|| per observation is what error_reproj.m / remove_outlier would see; vlg_ba_reproj_errors must give it."""
    g = load_golden(name)
    ctx = ctx_from_golden(g)
    out = ctx.reproj_errors(depth_max=1e9)
    ctx.close()
    ref = np.sqrt(((g["obs_xy"] - g["X_hat"]) ** 2).sum(axis=1))
    assert np.abs(out["err"] - ref).max() <= 1e-12 * max(ref.max(), 1.0)
    assert rel(out["mean_err"], ref.mean()) <= 1e-12
    ok = out["depth"] >= 0
    assert rel(out["max_sq_err"], float((ref[ok] ** 2).max())) <= 1e-12 and out["argmax"] == int(np.argmax(np.where(ok, ref, -1.0) ** 2 * np.where(ok, 1, -1)))
