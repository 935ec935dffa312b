"""Generates tests/golden/predict_golden.json with the UNMODIFIED reference library (oracle/_ref) through the shared frontend:
Vecchia GP, Gaussian likelihood, GPB_SetPredictionData + GPB_PredictREModel (mean and variance, response and latent), default
vecchia_pred_type. Run in the build container:  python tests/golden/make_predict_golden.py"""
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
CASES = [
    dict(n=1500, dseed=3, npred=200, pseed=5, cov_function="matern", shape=1.5, m=15, seed=1, cov_pars=[0.3, 1.2, 0.1]),
    dict(n=3000, dseed=4, npred=300, pseed=6, cov_function="exponential", shape=0.5, m=30, seed=2, cov_pars=[0.5, 0.8, 0.2]),
    dict(n=2000, dseed=5, npred=250, pseed=7, cov_function="matern", shape=2.5, m=20, seed=3, cov_pars=[0.1, 2.0, 0.05]),
    dict(n=1000, dseed=6, npred=100, pseed=8, cov_function="gaussian", shape=0., m=10, seed=4, cov_pars=[0.4, 1.0, 0.08]),
]


def pred_points(c):
    return np.random.default_rng(c["pseed"]).random((c["npred"], 2))


if __name__ == "__main__":
    out = {"generator": "tests/golden/make_predict_golden.py", "cases": []}
    for c in CASES:
        X, y = datagen.synth(c["n"], 2, c["dseed"])
        Xp = pred_points(c)
        m = GPModel(gp_coords=X, cov_function=c["cov_function"], cov_fct_shape=c["shape"], gp_approx="vecchia", num_neighbors=c["m"],
                    vecchia_ordering="random", seed=c["seed"], _lib=ref)
        rec = dict(c)
        r = m.predict(y, Xp, np.array(c["cov_pars"]), predict_var=True, predict_response=True)
        rl = m.predict(y, Xp, np.array(c["cov_pars"]), predict_var=True, predict_response=False)
        rec["mu_head"] = r["mu"][:32].tolist(); rec["mu_sum"] = float(r["mu"].sum())
        rec["var_response_head"] = r["var"][:32].tolist(); rec["var_response_sum"] = float(r["var"].sum())
        rec["var_latent_head"] = rl["var"][:32].tolist(); rec["var_latent_sum"] = float(rl["var"].sum())
        print(c["cov_function"], rec["mu_sum"], rec["var_response_sum"], rec["var_latent_sum"])
        out["cases"].append(rec)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "predict_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
