"""First GPU contact: device engine (C ABI) vs oracle; small parity + n=1e6 timing."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vecchia as ov
L = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpboost_b200", "lib_gpboost_b200.so"))
L.gpbdev_last_error.restype = C.c_char_p
def P(a, t=C.c_double): return a.ctypes.data_as(C.POINTER(t))
def chk(rc):
    if rc != 0: raise RuntimeError(L.gpbdev_last_error().decode())
print("devices", L.gpbdev_device_count())
def run(n, m, cov, shape, d=2, seed=1, check=True, reps=5):
    rng = np.random.default_rng(seed)
    coords = rng.random((n, d)); y = rng.standard_normal(n)
    perm = ov.random_order(n, seed)
    co = np.ascontiguousarray(coords[perm])
    h = C.c_void_p()
    t = time.time()
    chk(L.gpbdev_vecchia_create(C.byref(h), 0, C.c_int64(n), d, m, P(co), P(perm, C.c_int32), None, C.c_int64(0), C.c_int64(n)))
    print(f"n={n} m={m} create(+device kNN) {time.time()-t:.3f}s")
    cid = ov.cov_id(cov, shape)
    s2, pt = ov.transform_cov_pars([0.5, 1.0, 0.1], cov, shape)
    chk(L.gpbdev_vecchia_set_y(h, P(y)))
    out = np.zeros(9)
    if check:
        nn = np.empty((n, m), dtype=np.int32); chk(L.gpbdev_vecchia_get_nn(h, P(nn, C.c_int32)))
        t = time.time(); nn_o = ov.knn(co, m); print(f"  oracle knn {time.time()-t:.2f}s")
        print("  nn mismatches:", int((nn != nn_o).sum()))
        A, Dinv, Ag, Dg, bad = ov.factor(co, nn_o, cid, pt, calc_grad=True)
        yo = y[perm]
        ref = ov.nll_from_factor(nn_o, A, Dinv, yo, s2)
        for mode in (0, 1, 2):
            chk(L.gpbdev_vecchia_eval(h, cid, C.c_double(pt[0]), C.c_double(pt[1]), mode, P(out)))
            print(f"  mode {mode}: quad rel {abs(out[0]-ref[1])/abs(ref[1]):.2e} logdet rel {abs(out[1]-ref[2])/abs(ref[2]):.2e} nbad {out[2]}")
        g_ref = ov.grad_from_factor(nn_o, A, Dinv, Ag, Dg, yo, s2)
        g = np.array([(out[3+k] - 0.5*out[5+k])/s2 + 0.5*out[7+k] for k in range(2)])
        print("  grad", g, "ref", g_ref, "rel", np.abs(g-g_ref)/np.abs(g_ref))
        chk(L.gpbdev_vecchia_eval(h, cid, C.c_double(pt[0]), C.c_double(pt[1]), 1, P(out)))
        A_d = np.empty((n, m)); Di_d = np.empty(n)
        chk(L.gpbdev_vecchia_get_factor(h, P(A_d), P(Di_d)))
        print("  A max abs diff", np.abs(A_d - A).max(), "Dinv max rel", (np.abs(Di_d - Dinv)/Dinv).max())
        ya = np.empty(n); chk(L.gpbdev_vecchia_yaux(h, P(ya)))
        ya_o = np.empty(n); ya_o[perm] = ov.yaux(nn_o, A, Dinv, yo)
        print("  yaux max abs diff", np.abs(ya - ya_o).max(), "scale", np.abs(ya_o).max())
    for mode in (0, 1, 2):
        ms = C.c_float()
        for _ in range(3): chk(L.gpbdev_vecchia_eval_async(h, cid, C.c_double(pt[0]), C.c_double(pt[1]), mode))
        chk(L.gpbdev_vecchia_sync(h))
        chk(L.gpbdev_vecchia_timer_start(h))
        for _ in range(reps): chk(L.gpbdev_vecchia_eval_async(h, cid, C.c_double(pt[0]), C.c_double(pt[1]), mode))
        chk(L.gpbdev_vecchia_timer_stop(h, C.byref(ms)))
        print(f"  mode {mode}: {ms.value/reps:.3f} ms/eval")
    t = time.time()
    for _ in range(reps):
        chk(L.gpbdev_vecchia_set_y(h, P(y))); chk(L.gpbdev_vecchia_eval(h, cid, C.c_double(pt[0]), C.c_double(pt[1]), 0, P(out)))
    print(f"  e2e (H2D y + eval + D2H sums): {(time.time()-t)/reps*1e3:.3f} ms")
    chk(L.gpbdev_vecchia_free(h))
run(500, 10, "exponential", 0.5)
run(3000, 30, "matern", 1.5)
run(20000, 30, "matern", 2.5)
run(20000, 20, "gaussian", 0.)
run(100000, 30, "matern", 1.5, d=3)
run(100000, 15, "matern", 1.5, d=1)
run(1000000, 30, "matern", 1.5, check=True, reps=10)
