"""BASELINE.json config 1: the 4-view UCLA fixture.

The reference's UCLA path is toolbox/test/demo_vlmvg.m:13-22 -> generate_feature_track.m (VLFeat vl_sift +
vl_ubcmatch + RANSAC on the epipolar constraint) -> VLmvg.m -> mview_reconstruction.m:196
(bundle_euclid(Kparam, Te, w, Xe, x, 'fix_calibration', 'visibility', vis, 'verbose')).  MATLAB and VLFeat are
not in the image, so the FRONT END (everything before the bundle_euclid call -- out of scope)
is substituted here by OpenCV: SIFT on data/UCLA_0[1-4].jpg, tracks seeded by frame 1 (generate_feature_track.m:52-56),
Lowe ratio test + mutual check (match_sift_unique.m), RANSAC fundamental matrix per pair
(ransac_epipolar_constraint.m), points seen by views 1 and 2 kept (demo_vlmvg.m:19), K guessed from the feature
extents exactly as VLmvg.m:146-150, first camera at w = 0, T = 0 (multi_view.m:82-84), the other cameras from
recoverPose / solvePnP and the points from triangulation.  What the fixture pins is the INPUT of the hot path;
the golden trajectory on it is produced by the reference's own C (make_golden.py, oracle/_ref).

Run in the dev container (needs /root/reference and cv2):  python tests/golden/make_ucla4.py
Writes tests/golden/ucla4_tracks.npz (committed; the GPU box has no /root/reference).
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = "/root/reference/data"
MAX_POINTS = 400        # keeps the dense golden (x is 3 x n x m, A is 2 x 6 x n x m) small


def main():
    files = [os.path.join(DATA, f"UCLA_0{k}.jpg") for k in (1, 2, 3, 4)]
    sift = cv2.SIFT_create()
    kps, descs = [], []
    for f in files:
        img = cv2.imread(f, cv2.IMREAD_GRAYSCALE)
        kp, d = sift.detectAndCompute(img, None)
        kps.append(np.array([k.pt for k in kp], dtype=np.float64))
        descs.append(d)
        print(os.path.basename(f), img.shape, len(kp), "keypoints")
    m = len(files)
    n1 = len(kps[0])
    featx = np.zeros((n1, m)); featy = np.zeros((n1, m)); vis = np.zeros((n1, m))
    featx[:, 0] = kps[0][:, 0]; featy[:, 0] = kps[0][:, 1]; vis[:, 0] = 1
    bf = cv2.BFMatcher(cv2.NORM_L2)
    for i in range(1, m):
        fwd = bf.knnMatch(descs[i], descs[0], k=2)
        bwd = bf.knnMatch(descs[0], descs[i], k=2)
        best_b = {q.queryIdx: q.trainIdx for q, r in bwd if q.distance < 0.8 * r.distance}
        pairs = [(q.queryIdx, q.trainIdx) for q, r in fwd if q.distance < 0.8 * r.distance and best_b.get(q.trainIdx) == q.queryIdx]
        pi = np.array([p[0] for p in pairs]); p0 = np.array([p[1] for p in pairs])
        F, mask = cv2.findFundamentalMat(kps[i][pi], kps[0][p0], cv2.FM_RANSAC, 1.0, 0.999)
        inl = mask.ravel() != 0
        print(f"frame {i + 1}: {len(pairs)} unique matches, {int(inl.sum())} epipolar inliers")
        featx[p0[inl], i] = kps[i][pi[inl], 0]; featy[p0[inl], i] = kps[i][pi[inl], 1]; vis[p0[inl], i] = 1
    keep = (vis[:, 0] != 0) & (vis[:, 1] != 0)                        # demo_vlmvg.m:19
    featx, featy, vis = featx[keep], featy[keep], vis[keep]
    # K guess, VLmvg.m:146-150 (over the features that are passed in)
    cx = (featx.max() - featx.min()) / 2
    cy = (featy.max() - featy.min()) / 2
    fx = 2 * cx - featx.min()
    K3 = np.array([[fx, 0, cx], [0, fx, cy], [0, 0, 1.0]])
    # two-view initialisation (views 1, 2), then resection of views 3, 4 and triangulation
    x1 = np.stack([featx[:, 0], featy[:, 0]], 1); x2 = np.stack([featx[:, 1], featy[:, 1]], 1)
    E, _ = cv2.findEssentialMat(x1, x2, K3, cv2.RANSAC, 0.999, 1.0)
    _, R2, t2, pose_mask = cv2.recoverPose(E, x1, x2, K3)
    good = pose_mask.ravel() != 0
    featx, featy, vis, x1, x2 = featx[good], featy[good], vis[good], x1[good], x2[good]
    P1 = K3 @ np.hstack([np.eye(3), np.zeros((3, 1))]); P2 = K3 @ np.hstack([R2, t2])
    Xh = cv2.triangulatePoints(P1, P2, x1.T, x2.T)
    X = (Xh[:3] / Xh[3]).T
    front = (X[:, 2] > 0) & ((X @ R2.T + t2.ravel())[:, 2] > 0)
    featx, featy, vis, X = featx[front], featy[front], vis[front], X[front]
    if len(X) > MAX_POINTS:       # strongest coverage first: points seen in more views, then list order
        order = np.argsort(-vis.sum(1), kind="stable")[:MAX_POINTS]
        order.sort()
        featx, featy, vis, X = featx[order], featy[order], vis[order], X[order]
    n = len(X)
    R = [np.eye(3), R2]; T = [np.zeros(3), t2.ravel()]
    for i in (2, 3):
        s = vis[:, i] != 0
        ok, rvec, tvec, inl = cv2.solvePnPRansac(X[s], np.stack([featx[s, i], featy[s, i]], 1), K3, None, reprojectionError=3.0)
        assert ok
        # observations that disagree with the resected pose are dropped from the track (the reference prunes with
        # remove_outliers.m before BA)
        bad = np.setdiff1d(np.arange(int(s.sum())), inl.ravel())
        idx = np.flatnonzero(s)[bad]
        vis[idx, i] = 0; featx[idx, i] = 0; featy[idx, i] = 0
        R.append(cv2.Rodrigues(rvec)[0]); T.append(tvec.ravel())
        print(f"view {i + 1}: {int(s.sum())} tracked, {len(inl)} PnP inliers")
    w = np.stack([cv2.Rodrigues(Ri)[0].ravel() for Ri in R], 1)
    w[:, 0] = 0.0                                                     # vl_irodr(I) = 0, multi_view.m:82-84
    Te = np.stack(T, 1)
    K = np.tile(np.array([[fx], [fx], [cx], [cy]]), (1, m))           # calibration_parameter(K), repmat (mview_reconstruction.m:191)
    Xe = np.vstack([X.T, np.ones((1, n))])
    x = np.zeros((3, n, m))
    x[0] = featx * (vis != 0); x[1] = featy * (vis != 0); x[2] = 1.0    # x = ones(3,n,m) with the features filled in (VLmvg.m:121-123)
    out = os.path.join(HERE, "ucla4_tracks.npz")
    np.savez_compressed(out, K=K, Te=Te, w=w, Xe=Xe, x=x, visible=(vis != 0).astype(np.float64))
    print(f"ucla4: m = {m}, n = {n}, nobs = {int((vis != 0).sum())}, K guess fx = {fx:.2f} cx = {cx:.2f} cy = {cy:.2f}; wrote {out} ({os.path.getsize(out)} bytes)")


if __name__ == "__main__":
    main()
