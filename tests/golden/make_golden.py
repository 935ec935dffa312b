"""Generates tests/golden/vecchia_golden.json by driving the UNMODIFIED reference library
(oracle/_ref/lib_gpboost.so, built from /root/reference by oracle/Makefile.ref) through its own C API with the
same frontend the product uses (gpboost_b200.GPModel(..., _lib=reference)). Run in the build container:
    python tests/golden/make_golden.py
The GPU box has no /root/reference; the JSON produced here is what travels."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datagen  # noqa: E402
from gpboost_b200 import GPModel  # noqa: E402
from gpboost_b200.libpath import load_lib  # noqa: E402
from oracle import ref_lib_path  # noqa: E402

ref = load_lib(ref_lib_path())
out = {"generator": "tests/golden/make_golden.py", "reference": "fabsig/GPBoost c93fa49 (v1.7.3), CPU build", "nll": [], "fit": []}


def data(spec):
    if spec["data"] == "r_test":
        return datagen.r_test_data()
    if spec["data"] == "lattice":
        c = datagen.lattice(spec["k"])
        rng = np.random.default_rng(spec["seed"])
        return c, rng.standard_normal(c.shape[0])
    return datagen.synth(spec["n"], spec.get("d", 2), spec["seed"])


nll_cases = []
for cov, shape in (("exponential", 0.5), ("matern", 1.5), ("matern", 2.5), ("gaussian", 0.)):
    for m, ordering, seed in ((30, "none", 0), (20, "random", 0), (10, "random", 7)):
        nll_cases.append({"data": "r_test", "cov_function": cov, "cov_fct_shape": shape, "num_neighbors": m,
                          "vecchia_ordering": ordering, "seed": seed, "cov_pars": [0.1, 1.6, 0.2]})
for n, d, m in ((2000, 2, 30), (5000, 2, 15), (3000, 3, 20), (3000, 1, 10), (1500, 4, 12)):
    nll_cases.append({"data": "synth", "n": n, "d": d, "seed": 3, "cov_function": "matern", "cov_fct_shape": 1.5,
                      "num_neighbors": m, "vecchia_ordering": "random", "seed_model": 1, "cov_pars": [0.5, 1.0, 0.1]})
nll_cases.append({"data": "lattice", "k": 40, "seed": 5, "cov_function": "matern", "cov_fct_shape": 1.5, "num_neighbors": 12,
                  "vecchia_ordering": "random", "seed_model": 2, "cov_pars": [0.3, 1.2, 0.15]})
nll_cases.append({"data": "lattice", "k": 30, "seed": 5, "cov_function": "exponential", "cov_fct_shape": 0.5, "num_neighbors": 30,
                  "vecchia_ordering": "none", "seed_model": 0, "cov_pars": [0.3, 1.2, 0.15]})
for spec in nll_cases:
    coords, y = data(spec)
    mdl = GPModel(gp_coords=coords, cov_function=spec["cov_function"], cov_fct_shape=spec["cov_fct_shape"], gp_approx="vecchia",
                  num_neighbors=spec["num_neighbors"], vecchia_ordering=spec["vecchia_ordering"],
                  seed=spec.get("seed_model", spec["seed"]), _lib=ref)
    spec = dict(spec)
    spec["negll"] = mdl.neg_log_likelihood(np.array(spec["cov_pars"]), y)
    out["nll"].append(spec)
    print(spec)

fit_cases = [
    {"data": "r_test", "cov_function": "exponential", "cov_fct_shape": 0.5, "num_neighbors": 30, "vecchia_ordering": "none", "seed": 0},
    {"data": "r_test", "cov_function": "matern", "cov_fct_shape": 1.5, "num_neighbors": 20, "vecchia_ordering": "random", "seed": 0},
    {"data": "synth", "n": 3000, "d": 2, "seed": 3, "cov_function": "matern", "cov_fct_shape": 1.5, "num_neighbors": 20,
     "vecchia_ordering": "random", "seed_model": 1},
    {"data": "synth", "n": 3000, "d": 2, "seed": 4, "cov_function": "matern", "cov_fct_shape": 2.5, "num_neighbors": 15,
     "vecchia_ordering": "random", "seed_model": 1},
    {"data": "synth", "n": 4000, "d": 2, "seed": 5, "cov_function": "gaussian", "cov_fct_shape": 0., "num_neighbors": 15,
     "vecchia_ordering": "random", "seed_model": 1},
]
for spec in fit_cases:
    coords, y = data(spec)
    mdl = GPModel(gp_coords=coords, cov_function=spec["cov_function"], cov_fct_shape=spec["cov_fct_shape"], gp_approx="vecchia",
                  num_neighbors=spec["num_neighbors"], vecchia_ordering=spec["vecchia_ordering"],
                  seed=spec.get("seed_model", spec["seed"]), _lib=ref)
    mdl.fit(y)
    spec = dict(spec)
    spec["cov_pars"] = mdl.get_cov_pars().tolist()
    spec["negll"] = mdl.get_current_neg_log_likelihood()
    spec["num_it"] = mdl._get_num_optim_iter()
    out["fit"].append(spec)
    print(spec)

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vecchia_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
