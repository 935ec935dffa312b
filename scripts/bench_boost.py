"""Timing of boosting iterations: (a) plain L2 boosting n x 50 features (tree side only), (b) GPBoost with a Vecchia GP.
Usage: python scripts/bench_boost.py [n_tree] [n_gp] [--ref]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpboost_b200 import GPModel
from gpboost_b200.booster import Booster, Dataset
from gpboost_b200.libpath import load_lib
n_tree = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
n_gp = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
with_ref = "--ref" in sys.argv
libs = [("b200", None)]
if with_ref:
    from oracle import ref_lib_path
    libs.append(("reference", load_lib(ref_lib_path())))
rng = np.random.default_rng(1)
F = 50
X = rng.random((n_tree, F)); y = 2 * np.sin(3 * X[:, 0]) + X[:, 1] ** 2 + 0.5 * rng.standard_normal(n_tree)
params = dict(objective="regression", num_leaves=31, min_data_in_leaf=20, learning_rate=0.1, max_bin=255, verbose=-1)
for name, lib in libs:
    p = dict(params)
    t = time.perf_counter(); ds = Dataset(X, y, params=p, _lib=lib); t_ds = time.perf_counter() - t
    b = Booster(p, ds, _lib=lib)
    b.update()
    t = time.perf_counter()
    for _ in range(10): b.update()
    dt = (time.perf_counter() - t) / 10
    print(f"[{name}] plain boosting n={n_tree} F={F}: dataset {t_ds:.2f}s, {dt*1e3:.2f} ms/iter ({1/dt:.1f} iters/s)", flush=True)
    del b, ds
# (c) GPBoost with a single-level grouped random effect (BASELINE config 3): 1e4 groups
group = rng.integers(0, 10000, size=n_tree)
yg3 = y + rng.standard_normal(10000)[group]
for name, lib in libs:
    p = dict(params)
    gp = GPModel(group_data=group, _lib=lib)
    ds = Dataset(X, yg3, params=p, _lib=lib)
    b = Booster(p, ds, gp_model=gp, _lib=lib)
    t = time.perf_counter(); b.update(); t_first = time.perf_counter() - t
    t = time.perf_counter()
    for _ in range(10): b.update()
    dt = (time.perf_counter() - t) / 10
    print(f"[{name}] GPBoost grouped RE (1e4 groups) n={n_tree} F={F}: first iter {t_first:.3f}s, then {dt*1e3:.2f} ms/iter ({1/dt:.1f} iters/s) cov_pars {gp.get_cov_pars()}", flush=True)
    del b, ds, gp
coords = rng.random((n_gp, 2)); Xg = rng.random((n_gp, F))
yg = 2 * np.sin(3 * Xg[:, 0]) + Xg[:, 1] ** 2 + np.sin(5 * coords[:, 0]) * np.cos(4 * coords[:, 1]) + 0.5 * rng.standard_normal(n_gp)
for name, lib in libs:
    p = dict(params)
    t = time.perf_counter()
    gp = GPModel(gp_coords=coords, cov_function="matern", cov_fct_shape=1.5, gp_approx="vecchia", num_neighbors=30, vecchia_ordering="random", seed=1, _lib=lib)
    ds = Dataset(Xg, yg, params=p, _lib=lib)
    b = Booster(p, ds, gp_model=gp, _lib=lib)
    t_create = time.perf_counter() - t
    t = time.perf_counter(); b.update(); t_first = time.perf_counter() - t
    iters = 5 if lib is None else 2
    t = time.perf_counter()
    for _ in range(iters): b.update()
    dt = (time.perf_counter() - t) / iters
    print(f"[{name}] GPBoost Vecchia m=30 n={n_gp} F={F}: create {t_create:.2f}s first iter {t_first:.2f}s, then {dt*1e3:.1f} ms/iter ({1/dt:.3f} iters/s) cov_pars {gp.get_cov_pars()}", flush=True)
    del b, ds, gp
