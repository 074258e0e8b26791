"""Deterministic synthetic BAL-shaped bundle-adjustment problems.

Mirrors the conventions of the reference's scene generator
(toolbox/test/generate_scene_and_motion.m:32-39: 500x500 image, f=500, cx=cy=250, the same
intrinsics for every camera; outputs K(4xm) T(3xm) w(3xm) X(4xn) and a visibility map) and
of its BA demo (toolbox/test/demo_bundle_euclid.m:29-31: initial estimate = truth + N(0,1e-3)
on w, N(0,1e-4)*scale on T, N(0,1e-3)*scale on X; 0.5 px image noise as test_mview.m:45),
but with a BAL-like sparse visibility pattern: per-point track length 2 + Geom clipped to
[2, min(m,64)], ~90 % of a track a contiguous camera window, the rest random cameras.

Camera 0 sits exactly at w = 0, T = 0 (as in the reference's calibrated pipelines,
multi_view.m:82-84), which exercises the theta < 1e-6 identity branch of vl_rodrigues and the
pinv-of-zero-rows behaviour of the reduced system.

The observation list is produced in the reference's traversal order: ascending i + n*j
(camera j outer, point i inner; mex_bundle_1_XABeUVWeAeB.c:192-196).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# nominal shapes of BASELINE.json's configs (m cameras, n points, nobs observations)
CONFIGS = {
    "ucla4": (4, 1000, 3200),
    "ladybug": (49, 7776, 31843),
    "trafalgar": (257, 65132, 225911),
    "venice": (1778, 993923, 5001946),
    "final": (13682, 4456117, 28987644),
}


@dataclass
class Problem:
    m: int
    n: int
    K: np.ndarray        # (4, m)  [fx fy cx cy]'
    w: np.ndarray        # (3, m)  initial rotation vectors
    Te: np.ndarray       # (3, m)  initial translations
    Xe: np.ndarray       # (4, n)  initial points, homogeneous
    obs_xy: np.ndarray   # (nobs, 2) measured image points
    obs_pt: np.ndarray   # (nobs,) int32 point index i
    obs_cam: np.ndarray  # (nobs,) int32 camera index j
    w_true: np.ndarray
    Te_true: np.ndarray
    X_true: np.ndarray

    @property
    def nobs(self) -> int:
        return int(self.obs_pt.shape[0])

    def dense(self):
        """x (3,n,m) homogeneous image points and visible (n,m), as bundle_euclid.m:9,18 takes them."""
        x = np.zeros((3, self.n, self.m), order="F")
        vis = np.zeros((self.n, self.m), order="F")
        x[0, self.obs_pt, self.obs_cam] = self.obs_xy[:, 0]
        x[1, self.obs_pt, self.obs_cam] = self.obs_xy[:, 1]
        x[2, self.obs_pt, self.obs_cam] = 1.0
        vis[self.obs_pt, self.obs_cam] = 1.0
        return x, vis


def rodrigues(w: np.ndarray) -> np.ndarray:
    """Rotation matrices (m,3,3) of rotation vectors w (3,m); identity below 1e-6 like vl_rodrigues."""
    th = np.sqrt((w * w).sum(axis=0))
    m = w.shape[1]
    R = np.tile(np.eye(3), (m, 1, 1))
    big = th >= 1e-6
    if np.any(big):
        u = w[:, big] / th[big]
        s, c = np.sin(th[big]), np.cos(th[big])
        x, y, z = u
        mc = 1 - c
        Rb = np.empty((u.shape[1], 3, 3))
        Rb[:, 0, 0] = 1 - mc * (y * y + z * z); Rb[:, 0, 1] = -s * z + mc * x * y; Rb[:, 0, 2] = s * y + mc * x * z
        Rb[:, 1, 0] = s * z + mc * x * y; Rb[:, 1, 1] = 1 - mc * (z * z + x * x); Rb[:, 1, 2] = -s * x + mc * y * z
        Rb[:, 2, 0] = -s * y + mc * x * z; Rb[:, 2, 1] = s * x + mc * y * z; Rb[:, 2, 2] = 1 - mc * (x * x + y * y)
        R[big] = Rb
    return R


def project(K, w, Te, X, obs_pt, obs_cam):
    """Pinhole projection of X[:, obs_pt] into cameras obs_cam -> (nobs, 2)."""
    R = rodrigues(w)
    Xc = np.einsum("oij,jo->io", R[obs_cam], X[:, obs_pt]) + Te[:, obs_cam]
    u = K[0, obs_cam] * Xc[0] / Xc[2] + K[2, obs_cam]
    v = K[1, obs_cam] * Xc[1] / Xc[2] + K[3, obs_cam]
    return np.stack([u, v], axis=1)


def _tracks(m: int, n: int, nobs: int, rng, window_frac: float = 0.9) -> tuple[np.ndarray, np.ndarray]:
    tmax = min(m, 64)
    mean = max(nobs / n, 2.0)
    if tmax <= 2 or mean <= 2.0:
        t = np.full(n, 2, dtype=np.int64)
    else:
        p = 1.0 / (mean - 1.0)
        t = np.minimum(2 + rng.geometric(p, size=n) - 1, tmax).astype(np.int64)
    # nudge the total towards the requested count
    for _ in range(8):
        diff = int(nobs - t.sum())
        if diff == 0:
            break
        idx = rng.integers(0, n, size=min(abs(diff), n))
        if diff > 0:
            np.add.at(t, idx, 1)
        else:
            np.add.at(t, idx, -1)
        t = np.clip(t, 2, tmax)
    pt = np.repeat(np.arange(n, dtype=np.int64), t)
    start = np.repeat(np.cumsum(t) - t, t)
    r = np.arange(pt.shape[0], dtype=np.int64) - start
    tt = np.repeat(t, t)
    nwin = np.maximum(2, np.ceil(window_frac * tt).astype(np.int64))
    nwin = np.minimum(nwin, tt)
    centre = np.repeat(rng.integers(0, m, size=n), t)
    first = centre - nwin // 2
    in_win = r < nwin
    extra = rng.integers(0, np.maximum(m - nwin, 1))
    cam = np.where(in_win, first + r, first + nwin + extra) % m
    key = np.unique(pt * m + cam)
    pt = key // m
    cam = key % m
    # a point must keep at least two views: add the next free camera where dedup removed one
    cnt = np.bincount(pt, minlength=n)
    short = np.flatnonzero(cnt < 2)
    if short.size:
        have = cam[np.searchsorted(pt, short)]
        key = np.unique(np.concatenate([key, short * m + (have + 1) % m]))
        pt = key // m
        cam = key % m
    return pt.astype(np.int32), cam.astype(np.int32)


def make_problem(m: int, n: int, nobs: int, seed: int = 0, noise_px: float = 0.5,
                 point_seed: int | None = None, banded: bool = False) -> Problem:
    """`seed` fixes the cameras; `point_seed` (default: continue the same stream) fixes the points,
    tracks and image noise -- shards of one scene share `seed` and differ in `point_seed`.
    `banded`: every track is one contiguous camera window (no random "loop closure" cameras), so cameras share points
    only with their <= 63 neighbours and S is banded -- what a sequence without revisits looks like."""
    rng = np.random.default_rng(seed)
    cam_noise = np.random.default_rng([seed, 7])
    depth = 100.0
    K = np.tile(np.array([[500.0], [500.0], [250.0], [250.0]]), (1, m))
    # cameras on an arc of radius `depth` around the scene centre (0,0,depth), looking inwards
    span = 1.5 * np.pi if m > 8 else 0.25 * np.pi * max(m - 1, 1) / 3.0
    theta = np.linspace(0.0, span, m) if m > 1 else np.zeros(1)
    centre = np.array([0.0, 0.0, depth])
    cpos = centre[:, None] + depth * np.stack([-np.sin(theta), np.zeros(m), -np.cos(theta)])
    cpos += rng.normal(0.0, 1.0, size=(3, m))
    w_true = np.stack([rng.normal(0, 0.02, m), -theta + rng.normal(0, 0.02, m), rng.normal(0, 0.02, m)])
    w_true[:, 0] = 0.0
    cpos[:, 0] = 0.0
    R = rodrigues(w_true)
    Te_true = -np.einsum("mij,jm->im", R, cpos)
    Te_true[:, 0] = 0.0
    w0 = w_true + cam_noise.normal(0, 1e-3, size=w_true.shape)
    T0 = Te_true + cam_noise.normal(0, 1e-4 * depth, size=Te_true.shape)
    if point_seed is not None:
        rng = np.random.default_rng([seed, 1000 + point_seed])
    # points in a ball of radius 25 around the centre: inside every camera's frustum
    d = rng.normal(size=(3, n))
    d /= np.linalg.norm(d, axis=0)
    X_true = centre[:, None] + d * (25.0 * rng.random(n) ** (1.0 / 3.0))

    pt, cam = _tracks(m, n, nobs, rng, 1.0 if banded else 0.9)
    order = np.lexsort((pt, cam))            # ascending i + n*j
    pt, cam = pt[order], cam[order]
    xy = project(K, w_true, Te_true, X_true, pt, cam)
    xy += rng.normal(0.0, noise_px, size=xy.shape)

    X0 = X_true + rng.normal(0, 1e-3 * depth, size=X_true.shape)
    w0[:, 0] = 0.0
    T0[:, 0] = 0.0
    Xe = np.vstack([X0, np.ones((1, n))])
    return Problem(m, n, K, w0, T0, Xe, np.ascontiguousarray(xy), pt, cam, w_true, Te_true, X_true)


def make_config(name: str, seed: int = 0, scale: float = 1.0, point_seed: int | None = None, banded: bool = False) -> Problem:
    m, n, nobs = CONFIGS[name]
    if scale != 1.0:
        n = max(int(n * scale), 8)
        nobs = max(int(nobs * scale), 2 * n)
    return make_problem(m, n, nobs, seed, point_seed=point_seed, banded=banded)
