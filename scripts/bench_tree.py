"""Tree-side timing on one data set: plain L2 boosting and GPBoost with a grouped random effect, for the histogram kernel
versions / leaf-loop variants named on the command line (environment switches read at Booster creation).
Usage: python scripts/bench_tree.py [n] [variant ...]   variant = NAME or NAME:ENV=VAL[,ENV=VAL]  (default: v1, v2)
Prints ms/iter and a hash of the model text (identical trees <=> identical hash)."""
import hashlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpboost_b200 import GPModel
from gpboost_b200.booster import Booster, Dataset

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
variants = sys.argv[2:] or ["hist1:GPB200_HIST_KERNEL=1", "hist2:GPB200_HIST_KERNEL=2"]
F = 50
rng = np.random.default_rng(1)
X = rng.random((n, F)); y = 2 * np.sin(3 * X[:, 0]) + X[:, 1] ** 2 + 0.5 * rng.standard_normal(n)
group = rng.integers(0, 10000, size=n)
yg = y + rng.standard_normal(10000)[group]
params = dict(objective="regression", num_leaves=31, min_data_in_leaf=20, learning_rate=0.1, max_bin=255, verbose=-1)
t = time.perf_counter(); ds = Dataset(X, y, params=params); dsg = Dataset(X, yg, params=params)
print(f"datasets n={n} F={F}: {time.perf_counter() - t:.2f}s", flush=True)
managed = set()
for v in variants:
    name, _, envs = v.partition(":")
    for k in managed: os.environ.pop(k, None)
    for kv in filter(None, envs.split(",")):
        k, _, val = kv.partition("="); os.environ[k] = val; managed.add(k)
    b = Booster(params, ds)
    b.update()
    t = time.perf_counter()
    for _ in range(10): b.update()
    dt = (time.perf_counter() - t) / 10
    h = hashlib.sha1(b.model_to_string().encode()).hexdigest()[:12]
    print(f"[{name}] plain boosting: {dt*1e3:.3f} ms/iter ({1/dt:.1f} iters/s)  model {h}", flush=True)
    del b
    gp = GPModel(group_data=group)
    b = Booster(params, dsg, gp_model=gp)
    t = time.perf_counter(); b.update(); t_first = time.perf_counter() - t
    t = time.perf_counter()
    for _ in range(10): b.update()
    dt = (time.perf_counter() - t) / 10
    h = hashlib.sha1(b.model_to_string().encode()).hexdigest()[:12]
    print(f"[{name}] GPBoost grouped RE (1e4 groups): first {t_first*1e3:.1f} ms, then {dt*1e3:.3f} ms/iter ({1/dt:.1f} iters/s)  model {h} cov_pars {gp.get_cov_pars()}", flush=True)
    del b, gp
