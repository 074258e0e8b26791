"""Point sharding for multi-GPU runs.

Points are independent given the cameras: V_i, eB_i, V_i^-1, db_i and every W_ij of point i live
with the point, so a global observation list is split into contiguous POINT ranges balanced by
observation count; each rank keeps all cameras and re-indexes its points from 0.  The order inside
a shard stays the reference's traversal order (ascending i + n*j), so every per-shard sum over
observations is a sub-sequence of the global one and the per-camera sums U_j, eA_j, S_jj, e_j and the
PCG matvec add up over ranks (NCCL all-reduce inside libvlgba).
"""
from __future__ import annotations

import numpy as np


def point_ranges(obs_pt: np.ndarray, n: int, world: int) -> np.ndarray:
    """Boundaries lo[0..world] of contiguous point ranges with ~equal observation counts."""
    cnt = np.bincount(obs_pt, minlength=n).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(cnt)])
    total = csum[-1]
    targets = (np.arange(1, world) * total) // world
    cuts = np.searchsorted(csum, targets, side="left")
    return np.concatenate([[0], cuts, [n]]).astype(np.int64)


def shard_points(obs_xy, obs_pt, obs_cam, b, rank: int, world: int):
    """-> (obs_xy, obs_pt_local, obs_cam, b_local, (lo, hi)) for this rank; b is (n, 3)."""
    n = b.shape[0]
    lo_hi = point_ranges(np.asarray(obs_pt), n, world)
    lo, hi = int(lo_hi[rank]), int(lo_hi[rank + 1])
    sel = (obs_pt >= lo) & (obs_pt < hi)
    return (np.ascontiguousarray(obs_xy[sel]), (obs_pt[sel] - lo).astype(np.int32), np.ascontiguousarray(obs_cam[sel]),
            np.ascontiguousarray(b[lo:hi]), (lo, hi))


def strip_bounds(nstrips: int, nranks: int) -> np.ndarray:
    """Column blocks of the assembled S for `nranks` ranks: strips [J_r, J_r+1) of 32 columns with equal lower-triangle
    area, J_r = round(nstrips (1 - sqrt(1 - r / nranks))) -- the rule build_problem applies when set_comm preceded
    set_problem (csrc/vlg_ba.cu, s_bounds); rank r multiplies only its block."""
    J = [0]
    for r in range(1, nranks):
        J.append(min(nstrips, max(J[-1], int(round(nstrips * (1.0 - np.sqrt(1.0 - r / nranks)))))))
    J.append(nstrips)
    return np.asarray(J, dtype=np.int64)

