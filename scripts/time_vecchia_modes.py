"""Wall-clock time of one factor pass per mode (likelihood / store / gradient) at the headline shape, through gpbdev_vecchia_eval
(synchronous: the nine sums come back to the host). Usage: python scripts/time_vecchia_modes.py [n] — GPB200_LIB picks a variant build."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpboost_b200 import GPModel

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
rng = np.random.default_rng(3)
coords = rng.random((n, 2)); y = rng.standard_normal(n)
gp = GPModel(gp_coords=coords, cov_function="matern", cov_fct_shape=1.5, gp_approx="vecchia", num_neighbors=30, vecchia_ordering="random", seed=1)
v0 = gp.neg_log_likelihood(np.array([0.5, 1.0, 0.1]), y)  # builds the engine, uploads y
L, eng = gp._LIB, gp.device_engine()
out = (C.c_double * 9)()
res = {}
for mode, name in ((0, "nll"), (1, "store"), (2, "grad")):
    for k in range(2):
        assert L.gpbdev_vecchia_eval(eng, 1, C.c_double(2.0 + 1e-7 * k), C.c_double(17.3), mode, out) == 0
    t = time.perf_counter()
    for k in range(8):  # (a store request at exactly the state of the last pass is answered without a launch: vary the variance)
        L.gpbdev_vecchia_eval(eng, 1, C.c_double(2.0 + 1e-7 * (k + 2)), C.c_double(17.3), mode, out)
    res[name] = ((time.perf_counter() - t) / 8 * 1e3, [out[k] for k in range(9)])
print("lib", os.environ.get("GPB200_LIB", "default"), {k: round(v[0], 3) for k, v in res.items()}, "ms per pass")
print("grad sums", ["%.12g" % x for x in res["grad"][1]])
