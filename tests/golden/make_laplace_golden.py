"""Generates tests/golden/laplace_golden.json with the UNMODIFIED reference library (oracle/_ref) through the shared
frontend: latent Vecchia GP + bernoulli_logit likelihood, Laplace-approximated negative log-likelihood
with matrix_inversion_method "cholesky" and "iterative" (VADU preconditioner, 50 SLQ probes, seed 1), and the gradient of
both w.r.t. the log covariance parameters (recovered from one gradient-descent step of the reference's optimiser).
Run in the build container:  python tests/golden/make_laplace_golden.py"""
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
    dict(name="r_binary", cov_function="exponential", shape=0.5, m=20, ordering="none", seed=0, cov_pars=[0.9, 0.2]),
    dict(name="synth", n=2000, dseed=3, offset=False, cov_function="matern", shape=1.5, m=10, ordering="random", seed=1, cov_pars=[1.0, 0.1]),
    dict(name="synth", n=3000, dseed=4, offset=True, cov_function="exponential", shape=0.5, m=20, ordering="random", seed=2, cov_pars=[1.7, 0.15]),
    dict(name="synth", n=2500, dseed=5, offset=False, cov_function="matern", shape=2.5, m=30, ordering="random", seed=3, cov_pars=[0.6, 0.08]),
    dict(name="synth", n=1500, dseed=6, offset=True, cov_function="gaussian", shape=0., m=15, ordering="random", seed=4, cov_pars=[2.5, 0.05]),
    dict(name="synth", n=6000, dseed=7, offset=False, cov_function="matern", shape=1.5, m=30, ordering="random", seed=5, cov_pars=[1.2, 0.05]),
]


def case_data(c):
    if c["name"] == "r_binary":
        X, y = datagen.r_binary_test_data()
        return X, y, None
    return datagen.binary_synth(c["n"], c["dseed"], c["offset"])


def reference_gradient(c, X, y, off, method):
    """The reference does not export its gradient, but ONE step of its plain gradient descent on log(cov_pars) does
    (REModelTemplate::UpdateCovAuxPars, re_model_template.h:8741-8744: cov_pars_new = exp(log(cov_pars) - lr * gradient)):
    gradient = -log(cov_pars_new / cov_pars) / lr, w.r.t. (log variance, log range) on the original scale. Two learning rates
    must agree (no step halving, no learning-rate cap involved)."""
    th0 = np.array(c["cov_pars"], dtype=np.float64)
    got = []
    for lr in (1e-4, 5e-5):
        g = GPModel(likelihood="bernoulli_logit", gp_coords=X, cov_function=c["cov_function"], cov_fct_shape=c["shape"],
                    gp_approx="vecchia", num_neighbors=c["m"], vecchia_ordering=c["ordering"], seed=c["seed"],
                    matrix_inversion_method=method, _lib=ref)
        g.params["use_nesterov_acc"] = False
        g.fit(y, params=dict(optimizer_cov="gradient_descent", lr_cov=lr, maxit=1, init_cov_pars=th0, delta_rel_conv=1e-30), offset=off)
        got.append(-np.log(g.get_cov_pars() / th0) / lr)
    assert np.all(np.abs(got[0] - got[1]) <= 1e-7 * np.abs(got[0])), got
    return got[0].tolist()


if __name__ == "__main__":
    out = {"generator": "tests/golden/make_laplace_golden.py", "cases": []}
    for c in CASES:
        X, y, off = case_data(c)
        rec = dict(c)
        for method in ("cholesky", "iterative"):
            m = GPModel(likelihood="bernoulli_logit", gp_coords=X, cov_function=c["cov_function"], cov_fct_shape=c["shape"],
                        gp_approx="vecchia", num_neighbors=c["m"], vecchia_ordering=c["ordering"], seed=c["seed"],
                        matrix_inversion_method=method, _lib=ref)
            rec["negll_" + method] = m.neg_log_likelihood(np.array(c["cov_pars"]), y, fixed_effects=off)
            rec["grad_" + method] = reference_gradient(c, X, y, off, method)
        if c.get("n", 100) <= 2500:  # GPB_OptimCovPar with the reference's defaults (L-BFGS, iterative method): target of the device fit
            g = GPModel(likelihood="bernoulli_logit", gp_coords=X, cov_function=c["cov_function"], cov_fct_shape=c["shape"],
                        gp_approx="vecchia", num_neighbors=c["m"], vecchia_ordering=c["ordering"], seed=c["seed"],
                        matrix_inversion_method="iterative", _lib=ref)
            g.fit(y, offset=off)
            init = np.zeros(2)
            g._safe_call(g._LIB.GPB_GetInitCovPar(g.handle, init.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double))))
            rec["fit_iterative"] = dict(init_cov_pars=init.tolist(), cov_pars=g.get_cov_pars().tolist(), num_it=int(g._get_num_optim_iter()),
                                        negll=float(g.get_current_neg_log_likelihood()))
        print(rec)
        out["cases"].append(rec)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "laplace_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
