"""Exact (dense) GP likelihood — SURVEY §8 a6 / BASELINE config 1: oracle pinned on CPU, device path on the GPU."""
import json
import os

import numpy as np
import pytest

import datagen
from oracle import vecchia as ov

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dg():
    with open(os.path.join(ROOT, "tests", "golden", "dense_golden.json")) as f:
        return json.load(f)


def _data(spec):
    return datagen.r_test_data() if spec["data"] == "r_test" else datagen.synth(spec["n"], spec["d"], spec["seed"])


def test_dense_oracle_pinned(dg):
    for spec in dg["nll"]:
        if spec.get("n", 0) > 2500:
            continue
        coords, y = _data(spec)
        v = ov.dense_neg_log_likelihood(coords, spec["cov_pars"], y, spec["cov_function"], spec["cov_fct_shape"])
        assert abs(v - spec["negll"]) <= 1e-10 * abs(spec["negll"]), spec


@pytest.mark.gpu
def test_dense_device_negll_matches_reference_golden(dg, product_lib):
    from gpboost_b200 import GPModel
    assert product_lib.gpbdev_device_count() > 0
    for spec in dg["nll"]:
        coords, y = _data(spec)
        m = GPModel(gp_coords=coords, cov_function=spec["cov_function"], cov_fct_shape=spec["cov_fct_shape"], gp_approx="none")
        v = m.neg_log_likelihood(np.array(spec["cov_pars"]), y)
        assert abs(v - spec["negll"]) <= 1e-8 * abs(spec["negll"]), (spec, v)
    # R known answers (test_GPModel_gaussian_process.R:86-120)
    coords, y = datagen.r_test_data()
    for cov, shape, want in (("exponential", 0.5, 124.2549533), ("matern", 1.5, 141.3502172), ("matern", 2.5, 158.1111626)):
        m = GPModel(gp_coords=coords, cov_function=cov, cov_fct_shape=shape, gp_approx="none")
        assert abs(m.neg_log_likelihood(np.array([0.1, 1.6, 0.2]), y) - want) < 1e-6


@pytest.mark.gpu
def test_dense_device_psi_inv_y(product_lib):
    from gpboost_b200 import GPModel
    coords, y = datagen.synth(900, 2, 6)
    cp = np.array([0.3, 0.9, 0.15])
    m = GPModel(gp_coords=coords, cov_function="matern", cov_fct_shape=1.5, gp_approx="none")
    m.set_optim_params({"init_cov_pars": cp, "maxit": 0})
    m.fit(y)
    g = m.response_gradient(y)
    D = np.sqrt(((coords[:, None, :] - coords[None, :, :]) ** 2).sum(-1))
    r = np.sqrt(3.) / cp[2]
    Psi = cp[0] * np.eye(900) + cp[1] * (1 + r * D) * np.exp(-r * D)
    want = np.linalg.solve(Psi, y)
    assert np.abs(g - want).max() <= 1e-9 * np.abs(want).max()


@pytest.mark.gpu
def test_dense_device_fit_matches_reference_golden(dg, product_lib):
    """GPB_OptimCovPar with gp_approx="none" (BASELINE configs[0]): same optimum as the reference's L-BFGS fit."""
    from gpboost_b200 import GPModel
    for spec in dg["fit"]:
        coords, y = _data(spec)
        m = GPModel(gp_coords=coords, cov_function=spec["cov_function"], cov_fct_shape=spec["cov_fct_shape"], gp_approx="none")
        m.fit(y)
        # both optimisers stop on a 1e-6 relative change of the likelihood: compare the optimum, not the last digits of the path
        assert abs(m.get_current_neg_log_likelihood() - spec["negll"]) <= 2e-6 * abs(spec["negll"]), (spec, m.get_current_neg_log_likelihood())
        cp = m.get_cov_pars()
        assert np.all(np.abs(np.log(cp) - np.log(spec["cov_pars"])) < 5e-2), (spec, cp)
        assert abs(m._get_num_optim_iter() - spec["num_it"]) <= 3, (spec, m._get_num_optim_iter())


@pytest.mark.gpu
def test_dense_device_gradient_sums(product_lib):
    """tr(Psi^-1 dPsi_k) and alpha^T dPsi_k alpha from the device (L^-1, Psi^-1 = L^-T L^-1 tiles) against numpy."""
    import ctypes as C
    n = 333
    coords, y = datagen.synth(n, 2, 12)
    var, rho = 1.7, 0.21
    r = np.sqrt(3.) / rho
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    h = C.c_void_p()
    co = np.ascontiguousarray(coords)
    assert product_lib.gpbdev_dense_create(C.byref(h), 0, n, 2, P(co)) == 0
    assert product_lib.gpbdev_dense_set_y(h, P(np.ascontiguousarray(y))) == 0
    o3 = np.zeros(3)
    assert product_lib.gpbdev_dense_eval(h, 1, C.c_double(var), C.c_double(r), P(o3)) == 0
    g = np.zeros(4)
    assert product_lib.gpbdev_dense_grad(h, P(g)) == 0, product_lib.gpbdev_dense_last_error()
    D = np.sqrt(((coords[:, None, :] - coords[None, :, :]) ** 2).sum(-1))
    Sig = var * (1 + r * D) * np.exp(-r * D)
    G = -var * r * r * D * D * np.exp(-r * D)  # d Sigma / d log(range_transformed), cov_fcts.h:1155-1160
    Pinv = np.linalg.inv(np.eye(n) + Sig)
    a = Pinv @ y
    want = np.array([np.sum(Pinv * Sig), np.sum(Pinv * G), a @ Sig @ a, a @ G @ a])
    assert np.all(np.abs(g - want) <= 1e-9 * (1 + np.abs(want))), (g, want)
    product_lib.gpbdev_dense_free(h)
