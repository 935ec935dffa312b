"""Time the three modes of the factor kernel for one or more library builds (tuning helper)."""
import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vecchia as ov
def P(a, t=C.c_double): return a.ctypes.data_as(C.POINTER(t))
n, m = 1000000, 30
rng = np.random.default_rng(1); coords = rng.random((n, 2)); y = rng.standard_normal(n)
perm = ov.random_order(n, 1); co = np.ascontiguousarray(coords[perm])
s2, pt = ov.transform_cov_pars([0.5, 1.0, 0.1], "matern", 1.5)
nn_keep = None
for name in sys.argv[1:]:
    L = C.CDLL(os.path.join(ROOT, "gpboost_b200", name)); L.gpbdev_last_error.restype = C.c_char_p
    def chk(rc):
        if rc: raise RuntimeError(L.gpbdev_last_error().decode())
    h = C.c_void_p()
    chk(L.gpbdev_vecchia_create(C.byref(h), 0, C.c_int64(n), 2, m, P(co), P(perm, C.c_int32), None if nn_keep is None else P(nn_keep, C.c_int32), C.c_int64(0), C.c_int64(n)))
    if nn_keep is None:
        nn_keep = np.empty((n, m), dtype=np.int32); chk(L.gpbdev_vecchia_get_nn(h, P(nn_keep, C.c_int32)))
    chk(L.gpbdev_vecchia_set_y(h, P(y)))
    out = np.zeros(9); res = []
    for mode in (0, 1, 2):
        chk(L.gpbdev_vecchia_eval(h, 1, C.c_double(pt[0]), C.c_double(pt[1]), mode, P(out)))
        for _ in range(3): chk(L.gpbdev_vecchia_eval_async(h, 1, C.c_double(pt[0]), C.c_double(pt[1]), mode))
        chk(L.gpbdev_vecchia_sync(h)); ms = C.c_float()
        chk(L.gpbdev_vecchia_timer_start(h))
        for _ in range(10): chk(L.gpbdev_vecchia_eval_async(h, 1, C.c_double(pt[0]), C.c_double(pt[1]), mode))
        chk(L.gpbdev_vecchia_timer_stop(h, C.byref(ms))); res.append(ms.value / 10)
    print(f"{name}: NLL {res[0]:.3f} ms  STORE {res[1]:.3f} ms  GRAD {res[2]:.3f} ms   quad {out[0]:.10e} logdet {out[1]:.10e} g {out[3]:.6e} {out[4]:.6e}")
    chk(L.gpbdev_vecchia_free(h))
