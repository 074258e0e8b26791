"""CPU-side checks of the drop-in boundary: libvlgba.so loads, exports every symbol that
include/vlg_ba.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bundleadjustmentmatlab_b200 import bundle, capi

from conftest import ROOT, has_gpu


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "vlg_ba.h")).read()
    return sorted(set(re.findall(r"\b(vlg_ba_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"libvlgba.so does not export {n}"
    assert set(names) == set(capi.SYMBOLS)


def test_defaults_are_the_reference_constants():
    o = capi.default_opts()
    # bundle_euclid.m:49,111-123
    assert o.num_variableK == 4 and o.lambda0 == 1e-3 and o.nu0 == 2.0
    assert o.max_iter == 20 and o.max_iter2 == 10 and o.rel_tol == 1e-3 and o.abs_tol == 1e-20
    assert o.backsub_all_rows == 0 and o.rtable == capi.RTABLE_HOST_LIBM
    # solver-side defaults: everything that changes bits from run to run is opt-in
    assert o.solver == capi.SOLVER_AUTO and o.pcg_deflate == 1 and o.pcg_cluster == 1 and o.model == capi.MODEL_EUCLID
    assert o.pcg_autotune == 0 and o.pcg_rtol == 1e-8


def test_opts_struct_matches_the_header():
    """The ctypes mirror lists the fields of vlg_ba_opts in the header's order (the last one is written by
    vlg_ba_opts_default through the C struct, so a layout slip shows up as a wrong default above too)."""
    hdr = open(os.path.join(ROOT, "include", "vlg_ba.h")).read()
    body = hdr[hdr.index("typedef struct vlg_ba_opts {"):hdr.index("} vlg_ba_opts;")]
    fields = re.findall(r"^\s*(int|double)\s+([a-zA-Z0-9_]+);", body, flags=re.M)
    assert [(n, {"int": C.c_int, "double": C.c_double}[t]) for t, n in fields] == list(capi.Opts._fields_)


def test_option_parsing_mirrors_reference():
    x = np.zeros((3, 4, 2)); x[0, 1, 0] = 5.0; x[1, 3, 1] = 2.0
    o = bundle.parse_options(2, 4, x, ["FIX_CALIBRATION", "fix_pivot", [1, 0], "verbose"])
    assert o["num_variableK"] == 0 and o["verbose"] and list(o["pivot"]) == [True, False]
    assert o["visible"].shape == (4, 2) and o["visible"][1, 0] and o["visible"][3, 1] and o["visible"].sum() == 2
    assert bundle.parse_options(2, 4, x, ["fix_principal"])["num_variableK"] == 1
    assert bundle.parse_options(2, 4, x, [])["num_variableK"] == 4


def test_pack_layout():
    K = np.arange(8.0).reshape(4, 2); Te = np.ones((3, 2)); w = 2 * np.ones((3, 2)); Xe = np.ones((4, 5))
    a, b = bundle.pack(K, Te, w, Xe, 4)
    assert a.shape == (2, 10) and b.shape == (5, 3)
    assert np.array_equal(a[1], [2, 2, 2, 1, 1, 1, 1, 3, 5, 7])
    a1, _ = bundle.pack(K, Te, w, Xe, 1)
    assert a1.shape == (2, 7) and a1[1, 6] == 1.0


@pytest.mark.skipif(has_gpu(), reason="only meaningful without a device")
def test_no_cpu_fallback():
    with pytest.raises(capi.VlgBaError):
        capi.Context(num_variableK=0)
