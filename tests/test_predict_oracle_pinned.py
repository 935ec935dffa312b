"""Pins oracle/predict.py (Vecchia prediction, Gaussian likelihood, default vecchia_pred_type) against golden
vectors produced by the unmodified reference library (tests/golden/make_predict_golden.py)."""
import json
import os

import numpy as np
import pytest

import datagen
from oracle import predict as op

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "predict_golden.json")) as f:
    GOLD = json.load(f)["cases"]


@pytest.mark.parametrize("idx", range(len(GOLD)))
def test_prediction_matches_reference_golden(idx):
    c = GOLD[idx]
    X, y = datagen.synth(c["n"], 2, c["dseed"])
    Xp = np.random.default_rng(c["pseed"]).random((c["npred"], 2))
    mu, var = op.predict_gaussian(X, y, Xp, c["cov_pars"], c["cov_function"], c["shape"], c["m"], predict_response=True)
    _, varl = op.predict_gaussian(X, y, Xp, c["cov_pars"], c["cov_function"], c["shape"], c["m"], predict_response=False)
    tol = 1e-7 if c["cov_function"] == "gaussian" else 1e-9  # the Gaussian kernel's neighbour blocks are ill-conditioned
    assert np.abs(mu[:32] - np.array(c["mu_head"])).max() <= tol * np.abs(c["mu_head"]).max()
    assert abs(mu.sum() - c["mu_sum"]) <= tol * np.abs(mu).sum()
    assert np.abs(var[:32] - np.array(c["var_response_head"])).max() <= tol * np.abs(c["var_response_head"]).max()
    assert abs(var.sum() - c["var_response_sum"]) <= tol * abs(c["var_response_sum"])
    assert np.abs(varl[:32] - np.array(c["var_latent_head"])).max() <= tol * np.abs(c["var_response_head"]).max()
    assert abs(varl.sum() - c["var_latent_sum"]) <= tol * abs(c["var_response_sum"])
