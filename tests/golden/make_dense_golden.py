"""Generates tests/golden/dense_golden.json: exact (dense) GP likelihood values from the UNMODIFIED reference library
(gp_approx="none"), incl. BASELINE config 1 (n=2000, 2-D, Matern-1.5). Run: python tests/golden/make_dense_golden.py"""
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
out = {"generator": "tests/golden/make_dense_golden.py", "nll": []}
cases = [{"data": "r_test", "cov_function": c, "cov_fct_shape": s, "cov_pars": [0.1, 1.6, 0.2]}
         for c, s in (("exponential", 0.5), ("matern", 1.5), ("matern", 2.5), ("gaussian", 0.))]
cases += [{"data": "synth", "n": 2000, "d": 2, "seed": 1, "cov_function": "matern", "cov_fct_shape": 1.5, "cov_pars": [0.25, 1.0, 0.1]},
          {"data": "synth", "n": 777, "d": 3, "seed": 2, "cov_function": "matern", "cov_fct_shape": 2.5, "cov_pars": [0.4, 0.7, 0.3]},
          {"data": "synth", "n": 63, "d": 1, "seed": 3, "cov_function": "exponential", "cov_fct_shape": 0.5, "cov_pars": [0.4, 0.7, 0.3]},
          {"data": "synth", "n": 64, "d": 2, "seed": 4, "cov_function": "gaussian", "cov_fct_shape": 0., "cov_pars": [0.4, 0.7, 0.3]},
          {"data": "synth", "n": 3001, "d": 2, "seed": 5, "cov_function": "matern", "cov_fct_shape": 1.5, "cov_pars": [0.5, 1.0, 0.05]}]
for spec in cases:
    coords, y = datagen.r_test_data() if spec["data"] == "r_test" else datagen.synth(spec["n"], spec["d"], spec["seed"])
    m = GPModel(gp_coords=coords, cov_function=spec["cov_function"], cov_fct_shape=spec["cov_fct_shape"], gp_approx="none", _lib=ref)
    spec = dict(spec)
    spec["negll"] = m.neg_log_likelihood(np.array(spec["cov_pars"]), y)
    out["nll"].append(spec)
    print(spec)
# fits (GPB_OptimCovPar, default optimiser lbfgs): BASELINE configs[0] and two smaller cases
out["fit"] = []
for spec in ({"data": "r_test", "cov_function": "exponential", "cov_fct_shape": 0.5},
             {"data": "synth", "n": 2000, "d": 2, "seed": 1, "cov_function": "matern", "cov_fct_shape": 1.5},
             {"data": "synth", "n": 777, "d": 3, "seed": 2, "cov_function": "matern", "cov_fct_shape": 2.5},
             {"data": "synth", "n": 500, "d": 2, "seed": 8, "cov_function": "gaussian", "cov_fct_shape": 0.}):
    coords, y = datagen.r_test_data() if spec["data"] == "r_test" else datagen.synth(spec["n"], spec["d"], spec["seed"])
    m = GPModel(gp_coords=coords, cov_function=spec["cov_function"], cov_fct_shape=spec["cov_fct_shape"], gp_approx="none", _lib=ref)
    m.fit(y)
    spec = dict(spec)
    spec.update({"cov_pars": m.get_cov_pars().tolist(), "negll": m.get_current_neg_log_likelihood(), "num_it": m._get_num_optim_iter()})
    out["fit"].append(spec)
    print(spec)
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dense_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
