"""Single-level grouped random effects: oracle pinned on CPU, device path on the GPU."""
import json
import os

import numpy as np
import pytest

import datagen
import treedata
from oracle import grouped as og

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gg():
    with open(os.path.join(ROOT, "tests", "golden", "grouped_golden.json")) as f:
        return json.load(f)


def _data(rec):
    return datagen.r_grouped_test_data() if rec["kind"] == "r_test" else datagen.grouped_synth(**rec["kw"])


def test_grouped_oracle_pinned(gg):
    for rec in gg["nll"]:
        group, y = _data(rec)
        v = og.negll_woodbury(group, y, rec["cov_pars"])
        assert abs(v - rec["negll"]) <= 1e-10 * abs(rec["negll"])
    group, y = datagen.r_grouped_test_data()
    assert abs(og.negll_dense(group, y, [0.5, 1.2]) - og.negll_woodbury(group, y, [0.5, 1.2])) <= 1e-9 * 1e3


def test_grouped_r_known_answer_is_the_likelihood_optimum():
    # test_GPModel_grouped_random_effects.R:62-69: MLE (0.49348532, 1.22299521)
    group, y = datagen.r_grouped_test_data()
    cp = np.array([0.49348532, 1.22299521])
    f0 = og.negll_woodbury(group, y, cp)
    for d in ([1e-3, 0], [-1e-3, 0], [0, 1e-3], [0, -1e-3]):
        assert og.negll_woodbury(group, y, cp + np.array(d)) >= f0 - 1e-9


@pytest.mark.gpu
def test_grouped_device_negll_fit_gradient(gg, product_lib):
    from gpboost_b200 import GPModel
    assert product_lib.gpbdev_device_count() > 0
    for rec in gg["nll"]:
        group, y = _data(rec)
        m = GPModel(group_data=group)
        v = m.neg_log_likelihood(np.array(rec["cov_pars"]), y)
        assert abs(v - rec["negll"]) <= 1e-8 * abs(rec["negll"]), rec
    for rec in gg["fit"]:
        group, y = _data(rec)
        m = GPModel(group_data=group)
        m.fit(y)
        cp = m.get_cov_pars()
        print(rec["kind"], rec["kw"], "iters", m._get_num_optim_iter(), "ref", rec["num_it"], cp, rec["cov_pars"])
        assert abs(m.get_current_neg_log_likelihood() - rec["negll"]) <= 1e-7 * abs(rec["negll"])
        assert np.all(np.abs(cp - np.array(rec["cov_pars"])) <= 2e-3 * np.abs(rec["cov_pars"]))
        assert abs(m._get_num_optim_iter() - rec["num_it"]) <= 2
        g = m.response_gradient(y)
        assert np.abs(g - og.grad_response(group, y, cp)).max() <= 1e-9 * np.abs(g).max()
    # R known answer (Fisher scoring in the reference's test; the optimum does not depend on the optimiser)
    group, y = datagen.r_grouped_test_data()
    m = GPModel(group_data=group)
    m.fit(y)
    assert np.abs(m.get_cov_pars() - np.array([0.49348532, 1.22299521])).sum() < 1e-4


@pytest.mark.gpu
def test_gpboost_grouped_iteration_matches_reference_golden(gg, product_lib):
    from gpboost_b200 import GPModel
    from gpboost_b200.booster import Booster, Dataset, parse_model_string
    rec = gg["boost"][0]
    spec = rec["spec"]
    X, y, _ = treedata.make_case(spec)
    rng = np.random.default_rng(99)
    group = rng.integers(0, 120, size=spec["n"])
    y = y + rng.standard_normal(120)[group]
    params = treedata.booster_params(spec, reference=False)
    gp = GPModel(group_data=group)
    b = Booster(params, Dataset(X, y, params=params), gp_model=gp)
    for _ in range(spec["num_iter"]):
        b.update()
    trees = parse_model_string(b.model_to_string())
    for t, g in zip(trees, rec["trees"]):
        assert np.array_equal(t["split_feature"], np.array(g["split_feature"]))
        assert np.array_equal(t["threshold"], np.array(g["threshold"]))
        assert np.max(np.abs(t["leaf_value"] - np.array(g["leaf_value"]))) <= 1e-3 * np.max(np.abs(g["leaf_value"]))
    assert np.all(np.abs(gp.get_cov_pars() - np.array(rec["cov_pars"])) <= 5e-3 * np.abs(rec["cov_pars"]))
    assert np.abs(b.inner_predict_train()[:64] - np.array(rec["score_head"])).max() <= 2e-3 * np.abs(rec["score_head"]).max()
