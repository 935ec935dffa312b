"""Binning parity against the bins of the unmodified reference (tests/golden/binning_golden.json, produced by
LGBM_DatasetDumpText of oracle/_ref): CPU part = the host boundary search (value -> bin restated with numpy.searchsorted in the
test); GPU part = the device binning kernel through LGBM_DatasetCreateFromMat / LGBM_DatasetDumpText of the product library.
Integer outputs: bit-exact."""
import base64
import ctypes as C
import json
import os
import tempfile
import zlib

import numpy as np
import pytest

import bindata

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "binning_golden.json")) as f:
    GOLD = {c["name"]: c for c in json.load(f)["cases"]}


def gold_bins(name):
    g = GOLD[name]
    return np.frombuffer(zlib.decompress(base64.b64decode(g["bins_z"])), dtype="<i2").reshape(g["shape"]).astype(np.int64)


def make_dataset(lib, X, params, row_major=True, dtype=np.float64):
    from gpboost_b200.basic import c_str
    from gpboost_b200.booster import param_dict_to_str
    Xc = np.ascontiguousarray(X, dtype=dtype) if row_major else np.asfortranarray(X, dtype=dtype)
    h = C.c_void_p()
    rc = lib.LGBM_DatasetCreateFromMat(Xc.ctypes.data_as(C.c_void_p), C.c_int(1 if dtype == np.float64 else 0), C.c_int32(X.shape[0]),
                                       C.c_int32(X.shape[1]), C.c_int(1 if row_major else 0), c_str(param_dict_to_str(params)), None, C.byref(h))
    assert rc == 0, lib.LGBM_GetLastError().decode()
    return h


@pytest.mark.parametrize("case", bindata.CASES, ids=[c["name"] for c in bindata.CASES])
def test_host_bin_boundaries_reproduce_reference_bins(product_lib, case):
    """No device needed: the Dataset keeps its boundaries when binning cannot run; value -> bin is restated here."""
    lib = product_lib
    X = bindata.make_matrix(case["n"], case["seed"])
    h = make_dataset(lib, X, dict(case["params"], verbose=-1))
    want = gold_bins(case["name"])
    for j in range(X.shape[1]):
        nb, triv = C.c_int(0), C.c_int(0)
        ub = np.zeros(256)
        assert lib.GPB200_DatasetGetFeatureBins(h, j, C.byref(nb), C.byref(triv), ub.ctypes.data_as(C.POINTER(C.c_double))) == 0
        if np.all(want[:, j] < 0):
            assert triv.value == 1, (case["name"], j)
            continue
        assert triv.value == 0, (case["name"], j)
        b = ub[:nb.value]
        assert np.all(np.diff(b) > 0) and np.isinf(b[-1])
        got = np.searchsorted(b, X[:, j], side="left")  # smallest l with v <= b[l]  (bin.h:465-488)
        assert np.array_equal(got, want[:, j]), (case["name"], j, int(np.sum(got != want[:, j])))
    lib.LGBM_DatasetFree(h)


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["row_f64", "col_f64", "row_f32"])
@pytest.mark.parametrize("case", bindata.CASES, ids=[c["name"] for c in bindata.CASES])
def test_device_bins_equal_reference_bins(product_lib, case, layout):
    from gpboost_b200.basic import c_str
    lib = product_lib
    X = bindata.make_matrix(case["n"], case["seed"])
    want = gold_bins(case["name"])
    if layout == "row_f32":
        # float32 input: the reference bins the widened values; regenerate the expectation from the boundaries of the f32 matrix
        X = X.astype(np.float32)
    h = make_dataset(lib, X, dict(case["params"], verbose=-1), row_major=not layout.startswith("col"),
                     dtype=np.float32 if layout == "row_f32" else np.float64)
    path = os.path.join(tempfile.mkdtemp(), "dump.txt")
    assert lib.LGBM_DatasetDumpText(h, c_str(path)) == 0, lib.LGBM_GetLastError().decode()
    rows = open(path).read().split("\n")[-case["n"]:]
    got = np.array([[-1 if t.strip() == "NA" else int(t) for t in ln.split(",") if t.strip()] for ln in rows], dtype=np.int64)
    if layout != "row_f32":
        assert np.array_equal(got, want), (case["name"], layout, int(np.sum(got != want)))
    else:
        Xd = X.astype(np.float64)
        for j in range(X.shape[1]):
            nb, triv = C.c_int(0), C.c_int(0)
            ub = np.zeros(256)
            assert lib.GPB200_DatasetGetFeatureBins(h, j, C.byref(nb), C.byref(triv), ub.ctypes.data_as(C.POINTER(C.c_double))) == 0
            if triv.value:
                assert np.all(got[:, j] == -1)
            else:
                assert np.array_equal(got[:, j], np.searchsorted(ub[:nb.value], Xd[:, j], side="left")), (case["name"], j)
    lib.LGBM_DatasetFree(h)
