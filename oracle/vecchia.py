"""TEST INFRASTRUCTURE ONLY — numpy front-end of the C restatement (oracle/vecchia_oracle.c).

High-level entry points mirror what the reference's C API computes for the hot path:
  * `neg_log_likelihood(...)`  ==  GPB_EvalNegLogLikelihood on a Gaussian GP (Vecchia or exact)
  * `factor(...)`              ==  B = I - A, D^-1 (+ gradients) of CalcCovFactorGradientVecchia
  * `nll_and_grad(...)`        ==  one EvalLLforLBFGSpp objective + gradient call (optim_utils.h:244-340)
"""
import ctypes as C
import numpy as np

from .build import build_oracle

COV_IDS = {"exponential": 0, "matern0.5": 0, "matern1.5": 1, "matern2.5": 2, "gaussian": 3}

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        _lib.orc_vecchia_factor.restype = C.c_int
        _lib.orc_dense_nll.restype = C.c_int
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def cov_id(cov_function, shape=1.5):
    if cov_function == "matern":
        return COV_IDS["matern%.1f" % shape]
    return COV_IDS[cov_function]


def transform_cov_pars(cov_pars, cov_function, shape=1.5):
    """Original scale (sigma2, sigma1^2, rho) -> (sigma2, sigma1^2/sigma2, range_transformed):
    include/GPBoost/cov_fcts.h:485-552."""
    s2, s1, rho = [float(v) for v in cov_pars]
    cid = cov_id(cov_function, shape)
    rt = {0: 1. / rho, 1: np.sqrt(3.) / rho, 2: np.sqrt(5.) / rho, 3: 1. / (rho * rho)}[cid]
    return s2, np.array([s1 / s2, rt])


def random_order(n, seed):
    perm = np.empty(n, dtype=np.int32)
    lib().orc_vecchia_random_order(C.c_int(n), C.c_int(seed), _p(perm, C.c_int32))
    return perm


def knn(coords_ordered, m):
    """coords_ordered: (n, d) in Vecchia order. Returns int32 (n, m), -1 padded."""
    n, d = coords_ordered.shape
    cm = np.asfortranarray(coords_ordered, dtype=np.float64)
    nn = np.empty((n, m), dtype=np.int32)
    lib().orc_knn_vecchia(_p(cm), C.c_int(n), C.c_int(d), C.c_int(m), _p(nn, C.c_int32))
    return nn


def factor(coords_ordered, nn, cid, pars_trans, calc_grad=False):
    n, d = coords_ordered.shape
    m = nn.shape[1]
    cm = np.asfortranarray(coords_ordered, dtype=np.float64)
    A = np.empty((n, m)); Dinv = np.empty(n)
    Ag = np.empty((2, n, m)) if calc_grad else np.empty(1)
    Dg = np.empty((2, n)) if calc_grad else np.empty(1)
    pt = np.ascontiguousarray(pars_trans, dtype=np.float64)
    bad = lib().orc_vecchia_factor(_p(cm), C.c_int(n), C.c_int(d), C.c_int(m), _p(nn, C.c_int32), C.c_int(cid),
                                   _p(pt), _p(A), _p(Dinv), C.c_int(int(calc_grad)), _p(Ag), _p(Dg))
    return A, Dinv, (Ag if calc_grad else None), (Dg if calc_grad else None), bad


def nll_from_factor(nn, A, Dinv, y, sigma2):
    n, m = nn.shape
    out = np.empty(3)
    y = np.ascontiguousarray(y, dtype=np.float64)
    lib().orc_vecchia_nll(C.c_int(n), C.c_int(m), _p(nn, C.c_int32), _p(A), _p(Dinv), _p(y), C.c_double(sigma2), _p(out))
    return out  # negll, yTPsiInvy, log_det


def yaux(nn, A, Dinv, y):
    n, m = nn.shape
    out = np.empty(n)
    y = np.ascontiguousarray(y, dtype=np.float64)
    lib().orc_vecchia_yaux(C.c_int(n), C.c_int(m), _p(nn, C.c_int32), _p(A), _p(Dinv), _p(y), _p(out))
    return out


def grad_from_factor(nn, A, Dinv, Ag, Dg, y, sigma2):
    n, m = nn.shape
    g = np.empty(2)
    y = np.ascontiguousarray(y, dtype=np.float64)
    lib().orc_vecchia_grad(C.c_int(n), C.c_int(m), _p(nn, C.c_int32), _p(A), _p(Dinv), _p(Ag), _p(Dg), _p(y),
                           C.c_double(sigma2), _p(g))
    return g


class VecchiaOracle:
    """Ordering + neighbours fixed at construction (as GPB_CreateREModel does), evaluations afterwards."""

    def __init__(self, coords, num_neighbors=20, cov_function="matern", cov_fct_shape=1.5,
                 vecchia_ordering="random", seed=0):
        coords = np.asarray(coords, dtype=np.float64)
        n = coords.shape[0]
        self.n = n
        self.m = min(int(num_neighbors), n - 1)
        self.cov_function, self.shape = cov_function, cov_fct_shape
        self.cid = cov_id(cov_function, cov_fct_shape)
        if vecchia_ordering == "random":
            self.perm = random_order(n, seed)
        elif vecchia_ordering == "none":
            self.perm = np.arange(n, dtype=np.int32)
        else:
            raise ValueError(vecchia_ordering)
        self.coords = np.ascontiguousarray(coords[self.perm])
        self.nn = knn(self.coords, self.m)

    def neg_log_likelihood(self, cov_pars, y):
        s2, pt = transform_cov_pars(cov_pars, self.cov_function, self.shape)
        A, Dinv, _, _, _ = factor(self.coords, self.nn, self.cid, pt)
        return nll_from_factor(self.nn, A, Dinv, np.asarray(y, dtype=np.float64)[self.perm], s2)[0]

    def nll_and_grad_profiled(self, log_pars, y):
        """One L-BFGS objective call with the error variance profiled out (optim_utils.h:244-340):
        log_pars = log(sigma1^2/sigma2), log(range_transformed). Returns (negll, grad(2), sigma2)."""
        pt = np.exp(np.asarray(log_pars, dtype=np.float64))
        yo = np.asarray(y, dtype=np.float64)[self.perm]
        A, Dinv, Ag, Dg, _ = factor(self.coords, self.nn, self.cid, pt, calc_grad=True)
        _, ypy, ld = nll_from_factor(self.nn, A, Dinv, yo, 1.0)
        s2 = ypy / self.n  # ProfileOutSigma2, re_model_template.h:2640
        negll = ypy / 2. / s2 + ld / 2. + self.n / 2. * (np.log(s2) + np.log(2 * np.pi))
        g = grad_from_factor(self.nn, A, Dinv, Ag, Dg, yo, s2)
        return negll, g, s2

    def grad_response(self, cov_pars, y):
        """Psi^-1 y in the ORIGINAL data order: what REModel::CalcGradient writes back (re_model_template.h:3298)."""
        s2, pt = transform_cov_pars(cov_pars, self.cov_function, self.shape)
        A, Dinv, _, _, _ = factor(self.coords, self.nn, self.cid, pt)
        ya = yaux(self.nn, A, Dinv, np.asarray(y, dtype=np.float64)[self.perm])
        out = np.empty(self.n)
        out[self.perm] = ya
        return out, s2


def dense_neg_log_likelihood(coords, cov_pars, y, cov_function="matern", cov_fct_shape=1.5):
    coords = np.asarray(coords, dtype=np.float64)
    n, d = coords.shape
    s2, pt = transform_cov_pars(cov_pars, cov_function, cov_fct_shape)
    cm = np.asfortranarray(coords)
    out = np.empty(3)
    y = np.ascontiguousarray(y, dtype=np.float64)
    rc = lib().orc_dense_nll(_p(cm), C.c_int(n), C.c_int(d), C.c_int(cov_id(cov_function, cov_fct_shape)), _p(pt), _p(y),
                             C.c_double(s2), _p(out))
    if rc != 0:
        raise FloatingPointError("dense Cholesky failed")
    return out[0]
