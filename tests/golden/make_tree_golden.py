"""Generates tests/golden/tree_golden.json by driving the UNMODIFIED reference library (oracle/_ref/lib_gpboost.so) through
its own C API (LGBM_DatasetCreateFromMat / LGBM_[GP]BoosterCreate / LGBM_BoosterUpdateOneIter / LGBM_BoosterSaveModelToString)
with the same frontend the product uses. Run in the build container:  python tests/golden/make_tree_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import treedata  # noqa: E402
from gpboost_b200 import GPModel  # noqa: E402
from gpboost_b200.booster import Booster, Dataset, parse_model_string  # noqa: E402
from gpboost_b200.libpath import load_lib  # noqa: E402
from oracle import ref_lib_path  # noqa: E402

ref = load_lib(ref_lib_path())
out = {"generator": "tests/golden/make_tree_golden.py", "reference": "fabsig/GPBoost c93fa49 (v1.7.3), CPU build", "cases": []}
for spec in treedata.CASES:
    X, y, coords = treedata.make_case(spec)
    params = treedata.booster_params(spec, reference=True)
    ds = Dataset(X, y, params=params, _lib=ref)
    gp = None
    if coords is not None:
        gp = GPModel(gp_coords=coords, cov_function="matern", cov_fct_shape=1.5, gp_approx="vecchia", num_neighbors=spec["num_neighbors"],
                     vecchia_ordering="random", seed=1, _lib=ref)
        if spec.get("init_cov_pars"):
            gp.set_optim_params({"init_cov_pars": np.array(spec["init_cov_pars"])})
    b = Booster(params, ds, gp_model=gp, _lib=ref)
    for _ in range(spec["num_iter"]):
        b.update()
    trees = parse_model_string(b.model_to_string())
    score = b.inner_predict_train()
    rec = {"spec": spec, "trees": [{k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in t.items()
                                    if k in ("num_leaves", "split_feature", "threshold", "left_child", "right_child", "leaf_value", "leaf_count", "shrinkage")}
                                   for t in trees],
           "score_head": score[:64].tolist(), "score_sum": float(score.sum()), "score_sq": float((score ** 2).sum())}
    if gp is not None:
        rec["cov_pars"] = gp.get_cov_pars().tolist()
    out["cases"].append(rec)
    print(spec["name"], "trees", len(trees), [t["num_leaves"] for t in trees], rec.get("cov_pars"))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tree_golden.json"), "w") as f:
    json.dump(out, f)
