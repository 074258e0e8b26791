"""CPU: the host-side logic of bench.py and of the generator that the GPU runs rely on -- strong-scaling shards of ONE
problem, the byte accounting behind `roofline` / `jacobian_schur_roofline_frac`, the banded scenes."""
import argparse
import importlib.util
import os

import numpy as np

from bundleadjustmentmatlab_b200 import shard, synth

from conftest import ROOT

_spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(bench)


def _args(**kw):
    d = dict(config="ladybug", seed=0, scale=1.0, model="euclid", banded=False)
    d.update(kw)
    return argparse.Namespace(**d)


def test_strong_scaling_shards_partition_one_problem():
    P, a, b = bench.make_global(_args())
    for world in (1, 2, 4, 8):
        parts = [bench.make_local(P, b, r, world) for r in range(world)]
        assert sum(L.nobs for L in parts) == P.nobs and sum(L.n for L in parts) == P.n
        assert [L.lo for L in parts] == [0] + [L.hi for L in parts[:-1]] and parts[-1].hi == P.n
        # every rank keeps all cameras, its own contiguous point range re-indexed from 0, the list order of the reference
        for L in parts:
            assert L.m == P.m and L.b.shape == (L.n, 3)
            if L.nobs:
                assert L.obs_pt.min() >= 0 and L.obs_pt.max() < L.n
                key = L.obs_pt.astype(np.int64) + L.n * L.obs_cam.astype(np.int64)
                assert np.all(np.diff(key) > 0)
            sel = (P.obs_pt >= L.lo) & (P.obs_pt < L.hi)
            assert np.array_equal(L.obs_xy, P.obs_xy[sel]) and np.array_equal(L.obs_cam, P.obs_cam[sel])
        # balanced by observation count
        if world > 1:
            cnt = np.array([L.nobs for L in parts], dtype=float)
            assert cnt.max() / cnt.mean() < 1.05


def test_byte_accounting_is_per_rank_and_matches_the_survey_figures():
    P, a, b = bench.make_global(_args(config="trafalgar"))
    L1 = bench.make_local(P, b, 0, 1)
    Np = (6 * P.m + 31) // 32 * 32
    # Jacobian+Schur = 320 B/obs + 216 B/pt (+ the dense S when it is assembled)
    assert bench.jacobian_schur_bytes(L1, 6, True) == 320.0 * P.nobs + 216.0 * P.n + 8.0 * Np * Np
    assert bench.jacobian_schur_bytes(L1, 6, False) == 320.0 * P.nobs + 216.0 * P.n
    # the assembled-S matvec: the lower triangle on one rank, 1/world of it per rank (frac > 1 at N > 1)
    one = bench.algorithmic_bytes("pcg_symv", L1, 6, 0.0, 1)
    four = bench.algorithmic_bytes("pcg_symv", bench.make_local(P, b, 0, 4), 6, 0.0, 4)
    assert abs(one - (4 * Np * (Np + 32) + 32 * Np)) < 1 and abs(four - (Np * (Np + 32) + 32 * Np)) < 1
    its = 57.0
    assert abs(bench.algorithmic_bytes("pcg_persistent", L1, 6, its, 1) - its * (4 * Np * (Np + 32) + 64 * Np)) < 1
    # ... and the library's own count of kept tiles takes precedence (banded scenes)
    L1.symv_bytes = 12345
    assert bench.algorithmic_bytes("pcg_symv", L1, 6, 0.0, 1) == 12345 + 32 * Np


def test_banded_scenes_have_contiguous_tracks():
    for banded in (False, True):
        P = synth.make_problem(300, 4000, 22000, seed=5, banded=banded)
        order = np.lexsort((P.obs_cam, P.obs_pt))
        pt, cam = P.obs_pt[order], P.obs_cam[order].astype(np.int64)
        ptr = np.concatenate([[0], np.cumsum(np.bincount(pt, minlength=P.n))])
        spans = []
        for i in range(P.n):
            c = np.sort(cam[ptr[i]:ptr[i + 1]])
            gaps = np.diff(np.concatenate([c, [c[0] + P.m]]))      # cyclic: the generator wraps windows around camera 0
            spans.append(P.m - gaps.max() + 1)                      # shortest cyclic window that holds the track
        spans = np.array(spans); tlen = np.diff(ptr)
        if banded:
            assert np.all(spans == tlen)                            # every track is one contiguous camera window
        else:
            assert np.mean(spans > tlen) > 0.05                     # the default generator adds loop-closure cameras
    # a point still has at least two views and the list is in the reference's traversal order
    key = P.obs_pt.astype(np.int64) + P.n * P.obs_cam.astype(np.int64)
    assert np.all(np.diff(key) > 0) and np.bincount(P.obs_pt, minlength=P.n).min() >= 2


def test_strip_bounds_cover_the_triangle_with_equal_areas():
    for nstrips, nranks in ((334, 2), (334, 8), (12, 4), (5, 2)):
        J = shard.strip_bounds(nstrips, nranks)
        assert J[0] == 0 and J[-1] == nstrips and np.all(np.diff(J) >= 0)
        area = np.array([sum(nstrips - s for s in range(J[r], J[r + 1])) for r in range(nranks)], dtype=float)
        if nstrips >= 100:
            assert area.max() / area.mean() < 1.05
