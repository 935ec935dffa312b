"""Generates tests/golden/grouped_golden.json with the UNMODIFIED reference library (oracle/_ref) through the shared frontend:
single-level grouped random effects, Gaussian likelihood: likelihood values, fits and a GPBoost run.
Run in the build container:  python tests/golden/make_grouped_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datagen  # noqa: E402
import treedata  # noqa: E402
from gpboost_b200 import GPModel  # noqa: E402
from gpboost_b200.booster import Booster, Dataset, parse_model_string  # noqa: E402
from gpboost_b200.libpath import load_lib  # noqa: E402
from oracle import ref_lib_path  # noqa: E402

ref = load_lib(ref_lib_path())
out = {"generator": "tests/golden/make_grouped_golden.py", "nll": [], "fit": [], "boost": []}
cases = [("r_test", None), ("synth", dict(n=5000, G=50, seed=2, balanced=True)), ("synth", dict(n=20000, G=700, seed=3, balanced=False)),
         ("synth", dict(n=3000, G=2500, seed=4, balanced=False))]
for kind, kw in cases:
    group, y = datagen.r_grouped_test_data() if kind == "r_test" else datagen.grouped_synth(**kw)
    for cp in ([0.5, 1.2], [1.3, 0.2]):
        m = GPModel(group_data=group, _lib=ref)
        out["nll"].append({"kind": kind, "kw": kw, "cov_pars": cp, "negll": m.neg_log_likelihood(np.array(cp), y)})
    m = GPModel(group_data=group, _lib=ref)
    m.fit(y)
    out["fit"].append({"kind": kind, "kw": kw, "cov_pars": m.get_cov_pars().tolist(), "negll": m.get_current_neg_log_likelihood(),
                       "num_it": m._get_num_optim_iter()})
    print(out["fit"][-1])
# GPBoost with a grouped random effect (BASELINE config 3 in miniature)
spec = {"name": "gpboost_grouped", "n": 6000, "F": 5, "kind": "real", "num_leaves": 8, "min_data_in_leaf": 20, "num_iter": 4, "seed": 9}
X, y, _ = treedata.make_case(spec)
rng = np.random.default_rng(99)
group = rng.integers(0, 120, size=spec["n"])
y = y + rng.standard_normal(120)[group]
params = treedata.booster_params(spec, reference=True)
gp = GPModel(group_data=group, _lib=ref)
b = Booster(params, Dataset(X, y, params=params, _lib=ref), gp_model=gp, _lib=ref)
for _ in range(spec["num_iter"]):
    b.update()
trees = parse_model_string(b.model_to_string())
score = b.inner_predict_train()
out["boost"].append({"spec": spec, "cov_pars": gp.get_cov_pars().tolist(), "score_head": score[:64].tolist(),
                     "trees": [{k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in t.items()
                                if k in ("num_leaves", "split_feature", "threshold", "leaf_value", "leaf_count")} for t in trees]})
print("boost cov_pars", out["boost"][0]["cov_pars"])
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "grouped_golden.json"), "w") as f:
    json.dump(out, f)
