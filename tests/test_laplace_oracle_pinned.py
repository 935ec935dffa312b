"""Pins oracle/laplace.py (Laplace-Vecchia, bernoulli_logit) against
  * the R test's hard-coded value (test_GPModel_non_Gaussian_data.R:2541-2542, exact GP == Vecchia with all predecessors),
  * golden vectors produced by the unmodified reference library (tests/golden/make_laplace_golden.py), for both the
    sparse-Cholesky and the iterative (PCG + stochastic Lanczos quadrature, identical probe vectors) variants."""
import json
import os

import numpy as np
import pytest

import datagen
from oracle import laplace as ol
from oracle import vecchia as ov

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "laplace_golden.json")) as f:
    GOLD = json.load(f)["cases"]


def data_of(c):
    if c["name"] == "r_binary":
        X, y = datagen.r_binary_test_data()
        return X, y, None
    return datagen.binary_synth(c["n"], c["dseed"], c["offset"])


def oracle_negll(c, method):
    X, y, off = data_of(c)
    vo = ov.VecchiaOracle(X, c["m"], c["cov_function"], c["shape"], c["ordering"], c["seed"])
    _, pt = ov.transform_cov_pars([1.0] + list(c["cov_pars"]), c["cov_function"], c["shape"])
    fe = None if off is None else off[vo.perm]
    return ol.negll(vo.coords, vo.nn, vo.cid, c["cov_pars"][0], pt[1], y[vo.perm], fixed_effects=fe, method=method)


def test_r_known_answer():
    X, y = datagen.r_binary_test_data()
    vo = ov.VecchiaOracle(X, 99, "exponential", 0.5, "none", 0)
    r = ol.negll(vo.coords, vo.nn, vo.cid, 0.9, 1. / 0.2, y[vo.perm], method="cholesky")
    assert abs(r["negll"] - 66.299571) < 1e-6


@pytest.mark.parametrize("idx", range(len(GOLD)))
def test_golden_cholesky(idx):
    c = GOLD[idx]
    if c.get("n", 100) > 3000:
        pytest.skip("sparse LU of the larger case is slow in scipy")
    r = oracle_negll(c, "cholesky")
    assert abs(r["negll"] - c["negll_cholesky"]) <= 1e-9 * abs(c["negll_cholesky"])


@pytest.mark.parametrize("idx", range(len(GOLD)))
def test_golden_iterative(idx):
    c = GOLD[idx]
    r = oracle_negll(c, "iterative")
    # same probe vectors (std::mt19937 + seed_seq + normal_distribution) => deterministic agreement
    assert abs(r["negll"] - c["negll_iterative"]) <= 1e-9 * abs(c["negll_iterative"])
