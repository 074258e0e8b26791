"""the incremental caller's single-camera motion-only bundle adjustments
(toolbox/geometry/estimate_camera.m:247-253, one per camera added in incr_reconstruction.m:223-348) as ONE batched
solve -- vlg_ba_solve_cameras_independent -- against the CPU oracle run on every camera alone."""
import time

import numpy as np
import pytest

from bundleadjustmentmatlab_b200 import capi, synth
from oracle import lm

from common import rel

pytestmark = pytest.mark.gpu


def _single_camera_oracle(P, j):
    """bundle_euclid(K_j, T_j, w_j, X, x_j, 'fix_structure', 'fix_calibration', 'visibility', vis_j) on the CPU."""
    sel = P.obs_cam == j
    n = P.n
    x = np.zeros((3, n, 1), order="F"); vis = np.zeros((n, 1), order="F")
    x[0, P.obs_pt[sel], 0] = P.obs_xy[sel, 0]; x[1, P.obs_pt[sel], 0] = P.obs_xy[sel, 1]; x[2] = 1.0
    vis[P.obs_pt[sel], 0] = 1.0
    return lm.bundle_euclid(P.K[:, j:j + 1], P.Te[:, j:j + 1], P.w[:, j:j + 1], P.Xe, x, "fix_structure", "fix_calibration",
                            "visibility", vis, backend="sparse", record=False)


def test_batched_single_camera_problems_match_the_oracle_camera_by_camera():
    P = synth.make_problem(24, 600, 4200, seed=33)
    # perturb the cameras more than the generator does: resection starts from a RANSAC-DLT estimate (estimate_camera.m:230-246)
    rng = np.random.default_rng(7)
    P.w[:, 1:] += rng.normal(0, 5e-3, P.w[:, 1:].shape); P.Te[:, 1:] += rng.normal(0, 5e-2, P.Te[:, 1:].shape)
    a = np.ascontiguousarray(np.vstack([P.w, P.Te]).T); b = np.ascontiguousarray(P.Xe[:3].T)
    ctx = capi.Context(num_variableK=0, fix_structure=1)
    ctx.set_problem_sparse(P.K.T, a, b, P.obs_xy, P.obs_pt, P.obs_cam)
    t0 = time.perf_counter()
    a_out, errs, rounds = ctx.solve_cameras_independent()
    t_batch = time.perf_counter() - t0
    ctx.close()
    worst_first, worst_rms, same_len = 0.0, 0.0, 0
    for j in range(P.m):
        ref = _single_camera_oracle(P, j)
        e = errs[j]
        assert len(e) >= 2 and len(ref.error_) >= 2, (j, e, ref.error_)
        assert rel(e[0], ref.error_[0]) <= 1e-12, (j, e[0], ref.error_[0])
        worst_first = max(worst_first, rel(e[1], ref.error_[1]))
        assert rel(e[1], ref.error_[1]) <= 1e-9, (j, e[1], ref.error_[1])
        assert np.all(np.diff(e) < 0)
        if len(e) == len(ref.error_):
            same_len += 1
            r = rel(np.sqrt(e[-1]), np.sqrt(ref.error_[-1]))
            worst_rms = max(worst_rms, r)
            assert r <= 1e-6, (j, r)
            # the parameters themselves: the free-running trajectory is chaotic at ~1e-6 (SURVEY.md finding 2), the bars are on costs
            assert np.abs(a_out[j, :3] - ref.w_[:, 0]).max() <= 1e-4 and np.abs(a_out[j, 3:] - ref.Te_[:, 0]).max() <= 1e-4 * max(1.0, np.abs(ref.Te_).max())
        else:
            # the stop rule sits on a 1e-3 relative-decrease threshold (bundle_euclid.m:123): a flipped count is reported
            assert rel(np.sqrt(e[-1]), np.sqrt(ref.error_[-1])) <= 2e-3
    # the same problems one bundle_euclid call at a time through the library (what the caller would do without the batch entry)
    t0 = time.perf_counter()
    for j in range(P.m):
        sel = P.obs_cam == j
        c1 = capi.Context(num_variableK=0, fix_structure=1)
        c1.set_problem_sparse(P.K.T[j:j + 1], a[j:j + 1], b, P.obs_xy[sel], P.obs_pt[sel], np.zeros(int(sel.sum()), dtype=np.int32))
        _, _, _, _, e1 = c1.solve()
        c1.close()
        assert rel(e1[0], errs[j][0]) <= 1e-12 and rel(e1[1], errs[j][1]) <= 1e-9
    t_loop = time.perf_counter() - t0
    print(f"batched: {P.m} single-camera problems in {rounds} rounds, {1e3 * t_batch:.2f} ms; one call per camera: {1e3 * t_loop:.2f} ms; "
          f"first-step cost deviation vs the oracle {worst_first:.1e}, final RMS {worst_rms:.1e} ({same_len}/{P.m} with equal iteration counts)")
    assert same_len >= P.m - 2


def test_batch_entry_refuses_coupled_problems():
    P = synth.make_problem(4, 60, 220, seed=3)
    a = np.ascontiguousarray(np.vstack([P.w, P.Te]).T); b = np.ascontiguousarray(P.Xe[:3].T)
    ctx = capi.Context(num_variableK=0)
    ctx.set_problem_sparse(P.K.T, a, b, P.obs_xy, P.obs_pt, P.obs_cam)
    with pytest.raises(capi.VlgBaError):
        ctx.solve_cameras_independent()
    ctx.close()
