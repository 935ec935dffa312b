"""Bins of the UNMODIFIED reference (oracle/_ref/lib_gpboost.so) for tests/bindata.py's matrices: LGBM_DatasetCreateFromMat +
LGBM_DatasetDumpText (Dataset::DumpTextFile, src/LightGBM/io/dataset.cpp:1070) -> tests/golden/binning_golden.json.
Run in the build container: python tests/golden/make_binning_golden.py"""
import base64
import ctypes
import json
import os
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bindata  # noqa: E402
from gpboost_b200.booster import Dataset  # noqa: E402
from gpboost_b200.basic import c_str  # noqa: E402
from gpboost_b200.libpath import load_lib  # noqa: E402
from oracle import ref_lib_path  # noqa: E402


def dump_rows(lib, ds, n):
    """per-row bins of the dump: int array n x num_total_features, -1 for a filtered feature"""
    path = os.path.join(tempfile.mkdtemp(), "dump.txt")
    assert lib.LGBM_DatasetDumpText(ds.handle, c_str(path)) == 0, lib.LGBM_GetLastError()
    lines = open(path).read().split("\n")
    rows = [ln for ln in lines[-n:]]
    out = np.array([[-1 if t.strip() == "NA" else int(t) for t in ln.split(",") if t.strip()] for ln in rows], dtype=np.int16)
    assert out.shape[0] == n
    return out


if __name__ == "__main__":
    lib = load_lib(ref_lib_path())
    gold = []
    for c in bindata.CASES:
        X = bindata.make_matrix(c["n"], c["seed"])
        ds = Dataset(X, np.zeros(c["n"]), params=dict(c["params"], verbose=-1), _lib=lib)
        rows = dump_rows(lib, ds, c["n"])
        gold.append(dict(name=c["name"], shape=list(rows.shape), bins_z=base64.b64encode(zlib.compress(rows.astype("<i2").tobytes(), 9)).decode()))
        print(c["name"], rows.shape, "bins per feature:", rows.max(axis=0) + 1)
    with open(os.path.join(ROOT, "tests", "golden", "binning_golden.json"), "w") as f:
        json.dump(dict(generator="tests/golden/make_binning_golden.py", cases=gold), f)
