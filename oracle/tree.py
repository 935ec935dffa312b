"""TEST INFRASTRUCTURE ONLY — numpy front-end of oracle/tree_oracle.c plus a restatement of the boosting loop around it
(GBDT::TrainOneIter, src/LightGBM/boosting/gbdt.cpp:411-567; RegressionL2loss::GetGradients
objective/regression_objective.hpp:153-201 without a GP model: grad = score - label, hess = 1; BoostFromScore :259-265)."""
import ctypes as C

import numpy as np

from .build import build_oracle


class TreeConfig(C.Structure):
    _fields_ = [("num_leaves", C.c_int), ("min_data_in_leaf", C.c_int), ("min_sum_hessian_in_leaf", C.c_double),
                ("lambda_l2", C.c_double), ("min_gain_to_split", C.c_double), ("max_depth", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        _lib.orc_tree_train.restype = C.c_int
    return _lib


def make_config(num_leaves=31, min_data_in_leaf=20, min_sum_hessian_in_leaf=1e-3, lambda_l2=0., min_gain_to_split=0., max_depth=-1):
    return TreeConfig(num_leaves, min_data_in_leaf, min_sum_hessian_in_leaf, lambda_l2, min_gain_to_split, max_depth)


def train_tree(bins_fm, num_bin, grad, cfg, hess_const=1.0):
    """bins_fm: uint8 (F, n) feature-major. Returns dict with the tree arrays (unshrunk leaf values)."""
    F, n = bins_fm.shape
    L = cfg.num_leaves
    bins_fm = np.ascontiguousarray(bins_fm, dtype=np.uint8)
    num_bin = np.ascontiguousarray(num_bin, dtype=np.int32)
    grad = np.ascontiguousarray(grad, dtype=np.float64)
    sf = np.zeros(L, np.int32); tb = np.zeros(L, np.int32); lc = np.zeros(L, np.int32); rc = np.zeros(L, np.int32)
    sg = np.zeros(L, np.float32); lv = np.zeros(L, np.float64); cnt = np.zeros(L, np.int32); lor = np.zeros(n, np.int32)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    nl = lib().orc_tree_train(P(bins_fm, C.c_uint8), C.c_int(n), C.c_int(F), P(num_bin, C.c_int), P(grad, C.c_double),
                              C.c_double(hess_const), C.byref(cfg), P(sf, C.c_int), P(tb, C.c_int), P(lc, C.c_int), P(rc, C.c_int),
                              P(sg, C.c_float), P(lv, C.c_double), P(cnt, C.c_int), P(lor, C.c_int32))
    return {"num_leaves": nl, "split_feature": sf[:nl - 1], "threshold_bin": tb[:nl - 1], "left_child": lc[:nl - 1],
            "right_child": rc[:nl - 1], "split_gain": sg[:nl - 1], "leaf_value": lv[:nl], "leaf_count": cnt[:nl], "leaf_of_row": lor}


def boost_l2(bins_fm, num_bin, label, cfg, learning_rate, num_iter):
    """Plain L2 boosting on pre-binned data. Returns (trees, scores): leaf values are shrunk; the first tree does NOT carry the
    init-score bias (the reference adds it to the stored tree only, gbdt.cpp:498-500)."""
    label32 = np.asarray(label, dtype=np.float32)
    n = label32.shape[0]
    init = float(np.sum(label32.astype(np.float64)) / n)
    score = np.full(n, init if abs(init) > 1e-15 else 0.0)
    trees = []
    for _ in range(num_iter):
        grad = score - label32.astype(np.float64)
        t = train_tree(bins_fm, num_bin, grad, cfg)
        if t["num_leaves"] <= 1:
            break
        t["leaf_value"] = t["leaf_value"] * learning_rate
        score = score + t["leaf_value"][t["leaf_of_row"]]
        trees.append(t)
    return trees, score, init
