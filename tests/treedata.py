"""Seeded data and parameters of the tree / GPBoost parity cases (shared by the golden generator and the tests)."""
import numpy as np

CASES = [
    {"name": "int_features", "n": 5000, "F": 6, "kind": "int", "levels": 40, "num_leaves": 8, "min_data_in_leaf": 20, "num_iter": 3, "seed": 3},
    {"name": "int_features_deep", "n": 20000, "F": 12, "kind": "int", "levels": 200, "num_leaves": 63, "min_data_in_leaf": 5, "num_iter": 4, "seed": 4,
     "lambda_l2": 1.5, "min_gain_to_split": 0.01, "max_depth": 7},
    {"name": "real_features", "n": 20000, "F": 10, "kind": "real", "num_leaves": 31, "min_data_in_leaf": 20, "num_iter": 5, "seed": 5},
    {"name": "real_features_sampled_bins", "n": 250000, "F": 5, "kind": "real", "num_leaves": 15, "min_data_in_leaf": 50, "num_iter": 2, "seed": 6},
    {"name": "mixed_signs_and_zeros", "n": 8000, "F": 7, "kind": "mixed", "num_leaves": 16, "min_data_in_leaf": 10, "num_iter": 3, "seed": 7},
    {"name": "gpboost_vecchia", "n": 3000, "F": 5, "kind": "real", "num_leaves": 8, "min_data_in_leaf": 20, "num_iter": 3, "seed": 8,
     "gp": True, "num_neighbors": 15},
    # covariance parameters held fixed (train_gp_model_cov_pars = false): every iteration is a deterministic function of the data,
    # so ALL trees are compared (structure bit-exact, values 1e-8) — and the same with Newton leaf updates
    {"name": "gpboost_vecchia_fixed_pars", "n": 4000, "F": 6, "kind": "real", "num_leaves": 12, "min_data_in_leaf": 20, "num_iter": 4, "seed": 9,
     "gp": True, "num_neighbors": 20, "train_cov": False, "init_cov_pars": [0.12, 0.3, 0.15]},
    {"name": "gpboost_vecchia_newton", "n": 4000, "F": 6, "kind": "real", "num_leaves": 12, "min_data_in_leaf": 20, "num_iter": 4, "seed": 9,
     "gp": True, "num_neighbors": 20, "train_cov": False, "init_cov_pars": [0.12, 0.3, 0.15], "newton": True},
    {"name": "gpboost_vecchia_newton_many_leaves", "n": 6000, "F": 4, "kind": "real", "num_leaves": 80, "min_data_in_leaf": 10, "num_iter": 2, "seed": 10,
     "gp": True, "num_neighbors": 10, "train_cov": False, "init_cov_pars": [0.1, 0.4, 0.1], "newton": True},
    # step length of every tree from the closed-form line search -(F - y)' Psi^-1 f / f' Psi^-1 f (line_search_step_length),
    # alone and behind the Newton leaf update. The covariance parameters are trained: the reference takes F - y from its last
    # OptimCovPar call (re_model_template.h:1211-1214) and reads an empty vector when train_gp_model_cov_pars = false.
    {"name": "gpboost_vecchia_line_search", "n": 4000, "F": 6, "kind": "real", "num_leaves": 12, "min_data_in_leaf": 20, "num_iter": 3, "seed": 11,
     "gp": True, "num_neighbors": 20, "line_search": True},
    {"name": "gpboost_vecchia_newton_line_search", "n": 4000, "F": 6, "kind": "real", "num_leaves": 12, "min_data_in_leaf": 20, "num_iter": 3, "seed": 12,
     "gp": True, "num_neighbors": 20, "newton": True, "line_search": True},
]


def make_case(spec):
    rng = np.random.default_rng(spec["seed"])
    n, F = spec["n"], spec["F"]
    if spec["kind"] == "int":
        X = rng.integers(0, spec["levels"], size=(n, F)).astype(np.float64)
        f = np.sin(X[:, 0] / (spec["levels"] / 6.0)) + 0.3 * (X[:, 1] > spec["levels"] / 2) + X[:, 2] / spec["levels"]
    elif spec["kind"] == "real":
        X = rng.random((n, F))
        f = 2 * np.sin(3 * X[:, 0]) + X[:, 1] ** 2 + 0.5 * (X[:, 2] > 0.6)
    else:  # mixed: negative values, exact zeros, few distinct values, a constant column
        X = rng.standard_normal((n, F))
        X[:, 1] = np.where(rng.random(n) < 0.4, 0.0, X[:, 1])
        X[:, 2] = np.round(X[:, 2] * 2) / 2
        X[:, 3] = 1.25
        X[:, 4] = rng.integers(-3, 4, size=n)
        f = X[:, 0] + np.abs(X[:, 1]) + 0.7 * X[:, 2] - 0.2 * X[:, 4]
    y = f + 0.3 * rng.standard_normal(n)
    coords = None
    if spec.get("gp"):
        coords = rng.random((n, 2))
        y = y + np.sin(5 * coords[:, 0]) * np.cos(4 * coords[:, 1])
    return X, y, coords


def booster_params(spec, reference):
    p = {"objective": "regression", "num_leaves": spec["num_leaves"], "min_data_in_leaf": spec["min_data_in_leaf"], "learning_rate": 0.1,
         "max_bin": 255, "verbose": -1}
    for k in ("lambda_l2", "min_gain_to_split", "max_depth"):
        if k in spec:
            p[k] = spec[k]
    if spec.get("train_cov") is False:
        p["train_gp_model_cov_pars"] = False
    if spec.get("newton"):
        p["leaves_newton_update"] = True
    if spec.get("line_search"):
        p["line_search_step_length"] = True
    if reference:  # make the reference's summation order deterministic (column-wise histograms, fixed threads)
        p.update({"force_col_wise": True, "deterministic": True, "num_threads": 4})
    return p
