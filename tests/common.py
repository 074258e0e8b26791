"""Helpers shared by the parity tests."""
import glob
import os

import numpy as np

from oracle import lm

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    """Euclidean goldens (bundle_euclid.m over mex_bundle_{1,2,3}; tests/golden/make_golden.py)."""
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "euclid_*.npz")))


def proj_golden_names():
    """Projective goldens (bundle_projective.m over mex_bundle_proj_{1,2,3}; tests/golden/make_golden_proj.py)."""
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "proj_*.npz")))


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    g["options"] = [str(s) for s in g["options"]]
    return g


def golden_opts(g):
    """Option list as the reference's varargin (strings + values)."""
    opts = []
    for s in g["options"]:
        opts.append(s)
        if s == "fix_pivot":
            opts.append(g["pivot"] != 0)
    return opts + ["visibility", g["visible"]]


def oracle_options(g):
    return lm.parse_options(int(g["m"]), int(g["n"]), g["x"], golden_opts(g))


def ulp_diff(a, b):
    """max |a-b| in units of spacing(b) (0 where both are exactly equal)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    sp = np.spacing(np.maximum(np.abs(a), np.abs(b)))
    with np.errstate(invalid="ignore", divide="ignore"):
        u = np.where(d == 0, 0.0, d / sp)
    return float(u.max()) if u.size else 0.0


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)
