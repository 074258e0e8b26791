"""N>1 host logic on CPU: two gloo ranks each run the oracle on their point shard; the all-reduced
per-camera sums, cost and PCG-style matvec equal the oracle on the whole problem."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bundleadjustmentmatlab_b200 import shard, synth
from test_symv_plan import replay
from oracle import lm


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = synth.make_problem(9, 400, 1900, seed=3)
    a = np.vstack([P.w, P.Te]); b = np.ascontiguousarray(P.Xe[:3].T)
    xy, pt, cam, bl, (lo, hi) = shard.shard_points(P.obs_xy, P.obs_pt, P.obs_cam, b, rank, world)
    obs = lm.ObsList(P.m, hi - lo, xy, pt, cam)
    s1 = lm.stage1_sparse(P.K, a, bl.T.copy(), obs)
    cost = float((s1["e"] ** 2).sum())
    # a matvec-shaped per-camera sum: sum_i W_ij (W_ij' p_j summed over the point's cameras)
    rng = np.random.default_rng(0)
    p = rng.normal(size=(P.m, 6))
    t = np.zeros((hi - lo, 3))
    np.add.at(t, pt, np.einsum("ocr,or->oc", s1["W"], p[cam]))
    q = np.zeros((P.m, 6))
    np.add.at(q, cam, np.einsum("ocr,oc->or", s1["W"], t[pt]))
    buf = torch.from_numpy(np.concatenate([s1["U"].ravel(), s1["eA"].ravel(), q.ravel(), [cost, float(obs.nobs)]]))
    dist.all_reduce(buf)
    if rank == 0:
        np.save(out, buf.numpy())
    dist.destroy_process_group()


def test_point_sharding_sums_over_two_gloo_ranks(tmp_path):
    world = 2
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    red = np.load(out)
    P = synth.make_problem(9, 400, 1900, seed=3)
    a = np.vstack([P.w, P.Te]); b = P.Xe[:3].copy()
    obs = lm.ObsList(P.m, P.n, P.obs_xy, P.obs_pt, P.obs_cam)
    s1 = lm.stage1_sparse(P.K, a, b, obs)
    rng = np.random.default_rng(0)
    p = rng.normal(size=(P.m, 6))
    t = np.zeros((P.n, 3))
    np.add.at(t, obs.pt, np.einsum("ocr,or->oc", s1["W"], p[obs.cam]))
    q = np.zeros((P.m, 6))
    np.add.at(q, obs.cam, np.einsum("ocr,oc->or", s1["W"], t[obs.pt]))
    full = np.concatenate([s1["U"].ravel(), s1["eA"].ravel(), q.ravel(), [float((s1["e"] ** 2).sum()), float(obs.nobs)]])
    assert red[-1] == full[-1]                                   # every observation lives on exactly one rank
    scale = np.maximum(np.abs(full), 1e-300)
    assert np.max(np.abs(red - full) / np.maximum(scale, np.abs(full).max() * 1e-12)) <= 1e-12


def test_point_ranges_balance_and_cover():
    P = synth.make_problem(12, 3000, 14000, seed=5)
    for world in (1, 2, 4, 8):
        r = shard.point_ranges(P.obs_pt, P.n, world)
        assert r[0] == 0 and r[-1] == P.n and np.all(np.diff(r) >= 0)
        cnt = np.bincount(P.obs_pt, minlength=P.n)
        per = [cnt[r[k]:r[k + 1]].sum() for k in range(world)]
        assert sum(per) == P.nobs and max(per) - min(per) <= 64 + 1
        b = np.ascontiguousarray(P.Xe[:3].T)
        seen = 0
        for k in range(world):
            xy, pt, cam, bl, (lo, hi) = shard.shard_points(P.obs_xy, P.obs_pt, P.obs_cam, b, k, world)
            key = pt.astype(np.int64) + (hi - lo) * cam.astype(np.int64)
            assert np.all(np.diff(key) > 0)                      # still the reference's traversal order
            seen += pt.shape[0]
        assert seen == P.nobs


def _matvec_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Np = 3200
    J = shard.strip_bounds(Np // 32, world)
    # this rank's column block through the library's own plan (vlg_ba_symv_plan), replayed with the kernel's semantics
    y, ref, full, _ = replay(Np, 19, int(J[rank]), int(J[rank + 1]))
    buf = torch.from_numpy(y.copy())
    dist.all_reduce(buf)                                         # the per-iteration vector exchange
    if rank == 0:
        np.save(out, np.stack([buf.numpy(), full]))
    dist.destroy_process_group()


def test_symmetric_matvec_split_over_two_gloo_ranks(tmp_path):
    """Multi-GPU assembled-S path on the host side: the ranks multiply equal-area column blocks of the lower triangle and
    their partial products meet in one all-reduce."""
    world = 2
    out = str(tmp_path / "matvec.npy")
    mp.spawn(_matvec_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got, full = np.load(out)
    assert np.abs(got - full).max() <= 1e-10 * np.abs(full).max()
    J = shard.strip_bounds(100, 4)
    area = [sum(100 - j for j in range(J[r], J[r + 1])) for r in range(4)]
    assert J[0] == 0 and J[-1] == 100 and max(area) - min(area) <= 2 * 100      # equal areas up to a strip

