"""Vecchia prediction on the device through GPB_SetPredictionData / GPB_PredictREModel of the product library, against
golden vectors of the unmodified reference (tests/golden/predict_golden.json: mean, response and latent variances, four kernels,
default prediction type, 2 x num_neighbors neighbours = up to 60) and against the pinned oracle. Tolerance 1e-8 relative (the
Gaussian kernel's neighbour blocks are ill-conditioned: 1e-6 there, like the oracle's own pinning)."""
import json
import os

import numpy as np
import pytest

import datagen
from oracle import predict as op

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "predict_golden.json")) as f:
    GOLD = json.load(f)["cases"]


def model_of(c, X):
    from gpboost_b200 import GPModel
    return GPModel(gp_coords=X, cov_function=c["cov_function"], cov_fct_shape=c["shape"], gp_approx="vecchia", num_neighbors=c["m"],
                   vecchia_ordering="random", seed=c["seed"])


@pytest.mark.parametrize("idx", range(len(GOLD)))
def test_prediction_matches_reference_golden(idx):
    c = GOLD[idx]
    X, y = datagen.synth(c["n"], 2, c["dseed"])
    Xp = np.random.default_rng(c["pseed"]).random((c["npred"], 2))
    mdl = model_of(c, X)
    r = mdl.predict(y, Xp, np.array(c["cov_pars"]), predict_var=True, predict_response=True)
    rl = mdl.predict(y, Xp, np.array(c["cov_pars"]), predict_var=True, predict_response=False)
    tol = 1e-6 if c["cov_function"] == "gaussian" else 1e-8
    assert np.abs(r["mu"][:32] - np.array(c["mu_head"])).max() <= tol * np.abs(c["mu_head"]).max()
    assert abs(r["mu"].sum() - c["mu_sum"]) <= tol * np.abs(r["mu"]).sum()
    assert np.abs(r["var"][:32] - np.array(c["var_response_head"])).max() <= tol * np.abs(c["var_response_head"]).max()
    assert abs(r["var"].sum() - c["var_response_sum"]) <= tol * abs(c["var_response_sum"])
    assert np.abs(rl["var"][:32] - np.array(c["var_latent_head"])).max() <= tol * np.abs(c["var_response_head"]).max()
    assert abs(rl["var"].sum() - c["var_latent_sum"]) <= tol * abs(c["var_response_sum"])
    assert np.array_equal(rl["mu"], r["mu"])


def test_prediction_matches_oracle_every_point_and_explicit_neighbour_count():
    """all prediction points against the numpy restatement, with num_neighbors_pred set explicitly (GPB_SetPredictionData) and
    prediction points that coincide with observed ones (zero distance to the first neighbour)"""
    X, y = datagen.synth(2500, 2, 21)
    Xp = np.concatenate([np.random.default_rng(3).random((150, 2)), X[:25]])
    mdl = model_of(dict(cov_function="matern", shape=1.5, m=20, seed=2), X)
    cp = np.array([0.2, 1.1, 0.12])
    for nnp in (7, 33, 60):
        r = mdl.predict(y, Xp, cp, predict_var=True, predict_response=False, num_neighbors_pred=nnp)
        mu, var = op.predict_gaussian(X, y, Xp, cp, "matern", 1.5, 20, predict_response=False, num_neighbors_pred=nnp)
        assert np.abs(r["mu"] - mu).max() <= 1e-8 * np.abs(mu).max(), nnp
        assert np.abs(r["var"] - var).max() <= 1e-8 * np.abs(var).max(), nnp


def test_prediction_after_fit_uses_the_fitted_parameters_and_errors_are_clean():
    from gpboost_b200.basic import GPBoostError
    import ctypes as C
    X, y = datagen.synth(1200, 2, 4)
    mdl = model_of(dict(cov_function="exponential", shape=0.5, m=10, seed=1), X)
    mdl.fit(y)
    cp = mdl.get_cov_pars()
    Xp = np.random.default_rng(9).random((40, 2))
    r = mdl.predict(y, Xp, cp, predict_var=True)
    mu, var = op.predict_gaussian(X, y, Xp, cp, "exponential", 0.5, 10, predict_response=True)
    assert np.abs(r["mu"] - mu).max() <= 1e-8 * np.abs(mu).max() and np.abs(r["var"] - var).max() <= 1e-8 * np.abs(var).max()
    with pytest.raises(GPBoostError):
        mdl.predict(y, Xp, cp, vecchia_pred_type="order_pred_first")
    with pytest.raises(GPBoostError):
        mdl.predict(y, Xp, cp, num_neighbors_pred=61)
