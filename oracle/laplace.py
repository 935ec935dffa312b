"""TEST INFRASTRUCTURE ONLY (oracle). Laplace approximation for a latent Vecchia GP with a bernoulli_logit likelihood.

numpy/scipy restatement of the reference's
  FindModePostRandEffCalcMLLVecchia      include/GPBoost/likelihoods.h:3773-4059
  CheckConvergenceModeFinding            include/GPBoost/likelihoods.h:16079-16125
  Inv_SigmaI_plus_ZtWZ_Vecchia_iterative include/GPBoost/likelihoods.h:16264-16348  (VADU branch)
  CGVecchiaLaplaceVec                    src/GPBoost/CG_utils.cpp:21-108
  CalcLogDetStochVecchia                 include/GPBoost/likelihoods.h:16376-16521  (VADU branch)
  CGTridiagVecchiaLaplace                src/GPBoost/CG_utils.cpp:110-229
  LogDetStochTridiag                     src/GPBoost/CG_utils.cpp:1035-1052
  GenRandVecNormalParallel               src/GPBoost/CG_utils.cpp:978-994 (oracle/shuffle_oracle.cpp)
  LogLikBernoulliLogit / derivatives     include/GPBoost/likelihoods.h:11401, 12477, 13307; DF_utils.h:37-60
  CalcGradNegMargLikelihoodLaplaceApproxVecchia  include/GPBoost/likelihoods.h:6521-7044 (covariance-parameter gradient,
                                         Cholesky branch :6837-6905 and iterative / VADU branch :6567-6690)
  CalcLogDetStochDerivModeVecchia / CalcLogDetStochDerivCovParVecchia  include/GPBoost/likelihoods.h:16538-16692, 16706-16781
  CalcOptimalC / CalcOptimalCVectorized  src/GPBoost/CG_utils.cpp:1053-1088
Everything lives in the Vecchia ("ordered") index space. Pinned against the reference library by
tests/golden/make_laplace_golden.py -> tests/golden/laplace_golden.json (likelihood values and, recovered from one
gradient-descent step of the reference's own optimiser, its gradients for both matrix_inversion_methods).
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from . import vecchia as ov

DEFAULTS = dict(maxit_mode_newton=1000, delta_conv_mode_finding=1e-8, max_lr_shrink=20, c_armijo=1e-4,
                cg_max_num_it=1000, cg_max_num_it_tridiag=1000, cg_delta_conv=1e-2, num_rand_vec_trace=50,
                seed_rand_vec_trace=1)


def gen_rand_normal(seed, run_id, n, t):
    out = np.empty((t, n))
    ov.lib().orc_gen_rand_normal(C.c_int(seed), C.c_ulonglong(run_id), C.c_int(n), C.c_int(t),
                                 out.ctypes.data_as(C.POINTER(C.c_double)))
    return np.ascontiguousarray(out.T)  # n x t


def factor_latent(coords_ordered, nn, cid, var, range_trans):
    n, d = coords_ordered.shape
    m = nn.shape[1]
    cm = np.asfortranarray(coords_ordered, dtype=np.float64)
    A = np.empty((n, m)); Dinv = np.empty(n)
    pt = np.array([var, range_trans], dtype=np.float64)
    p = lambda a, t=C.c_double: a.ctypes.data_as(C.POINTER(t))
    nn = np.ascontiguousarray(nn, dtype=np.int32)
    bad = ov.lib().orc_vecchia_factor_latent(p(cm), C.c_int(n), C.c_int(d), C.c_int(m), p(nn, C.c_int32), C.c_int(cid),
                                             p(pt), p(A), p(Dinv))
    return A, Dinv, bad


def build_B(nn, A):
    n, m = nn.shape
    rows = np.repeat(np.arange(n), m)
    mask = nn.ravel() >= 0
    B = sp.csr_matrix((-A.ravel()[mask], (rows[mask], nn.ravel()[mask])), shape=(n, n)) + sp.identity(n, format="csr")
    return B.tocsr()


def sigmoid_stable(x):
    t = np.exp(-np.abs(x))
    return np.where(x >= 0, 1. / (1. + t), t / (1. + t))


def softplus(x):
    return np.log1p(np.exp(-np.abs(x))) + np.maximum(x, 0.)


def loglik(y, loc):
    return float(np.sum(y * loc - softplus(loc)))


def _vadu_solve(B, Bt, dw, R):
    Y = spl.spsolve_triangular(Bt, R, lower=False, unit_diagonal=True)
    return spl.spsolve_triangular(B, Y / (dw[:, None] if Y.ndim == 2 else dw), lower=True, unit_diagonal=True)


def cg_vadu(B, Bt, Dinv, W, rhs, u, p, delta_conv, initialize_to_zero):
    """CGVecchiaLaplaceVec, VADU preconditioner; returns (u, iterations)."""
    op = lambda h: Bt @ (Dinv * (B @ h)) + W * h
    if np.abs(rhs).sum() < 1e-100:
        return np.zeros_like(rhs), 0
    if initialize_to_zero or not u.any():
        u = np.zeros_like(rhs); r = rhs.copy()
    else:
        r = rhs - op(u)
    dw = Dinv + W
    z = _vadu_solve(B, Bt, dw, r)
    h = z.copy()
    for j in range(p):
        v = op(h)
        a = (r @ z) / (h @ v)
        u = u + a * h
        r_old, z_old = r, z
        r = r - a * v
        if np.linalg.norm(r) < delta_conv:
            return u, j + 1
        z = _vadu_solve(B, Bt, dw, r)
        b = (r @ z) / (r_old @ z_old)
        h = z + b * h
    return u, p


def cg_tridiag_vadu(B, Bt, Dinv, W, rhs, p, delta_conv):
    """CGTridiagVecchiaLaplace; returns the Lanczos tridiagonals (list of diag, subdiag) per column."""
    n, t = rhs.shape
    dw = Dinv + W
    R = rhs.copy()
    Z = _vadu_solve(B, Bt, dw, R)
    H = Z.copy()
    U = np.zeros_like(rhs)
    a = np.ones(t); b = np.zeros(t)
    Td = np.zeros((p, t)); Ts = np.zeros((max(p - 1, 0), t))
    its = p
    for j in range(p):
        V = Bt @ (Dinv[:, None] * (B @ H)) + W[:, None] * H
        a_old = a
        a = (R * Z).sum(0) / (H * V).sum(0)
        U = U + H * a
        R_old = R
        R = R - V * a
        early = np.linalg.norm(R, axis=0).mean() < delta_conv
        Z_old = Z
        Z = _vadu_solve(B, Bt, dw, R)
        b_old = b
        b = (R * Z).sum(0) / (R_old * Z_old).sum(0)
        H = Z + H * b
        Td[j] = 1. / a + b_old / a_old
        if j > 0:
            Ts[j - 1] = np.sqrt(b_old) / a_old
        if early:
            its = j + 1
            break
    return Td[:its], Ts[:max(its - 1, 0)], its, U


def logdet_tridiag(Td, Ts, n):
    from scipy.linalg import eigh_tridiagonal
    t = Td.shape[1]
    tot = 0.
    for i in range(t):
        if Td.shape[0] == 1:
            lam, vec = Td[:1, i], np.ones((1, 1))
        else:
            lam, vec = eigh_tridiagonal(Td[:, i], Ts[:, i])
        tot += float(np.sum(vec[0] ** 2 * np.log(lam)))
    return tot * n / t


def negll(coords_ordered, nn, cid, var, range_trans, y, fixed_effects=None, method="cholesky", probes=None, **kw):
    """Laplace-approximated negative marginal log-likelihood; returns dict(negll, mode, newton_it, cg_it, slq_it, logdet)."""
    cfg = dict(DEFAULTS); cfg.update(kw)
    n = y.shape[0]
    A, Dinv, bad = factor_latent(coords_ordered, nn, cid, var, range_trans)
    assert bad == 0
    B = build_B(np.asarray(nn), A)
    Bt = B.T.tocsr()
    F = np.zeros(n) if fixed_effects is None else fixed_effects
    mode = np.zeros(n)
    SigmaI = (Bt @ sp.diags(Dinv) @ B).tocsc() if method == "cholesky" else None
    quad = lambda v: float((B @ v) @ (Dinv * (B @ v)))
    mll = loglik(y, F + mode) - 0.5 * quad(mode)
    upd = np.zeros(n)
    cg_total = 0
    lu = None
    for it in range(cfg["maxit_mode_newton"]):
        p_ = sigmoid_stable(F + mode)
        grad = y - p_
        W = p_ * (1. - p_)
        rhs = W * mode + grad
        if method == "cholesky":
            lu = spl.splu((SigmaI + sp.diags(W)).tocsc())
            upd = lu.solve(rhs)
        else:
            upd, k = cg_vadu(B, Bt, Dinv, W, rhs, upd, cfg["cg_max_num_it"], cfg["cg_delta_conv"], it == 0)
            cg_total += k
        direction = upd - mode
        gdd = float(direction @ (Bt @ (Dinv * (B @ direction)) + W * direction))
        lr = 1.
        for ih in range(cfg["max_lr_shrink"]):
            mode_new = upd if ih == 0 else (1 - lr) * mode + lr * upd
            mll_new = loglik(y, F + mode_new) - 0.5 * quad(mode_new)
            if mll_new < mll + cfg["c_armijo"] * lr * gdd or not np.isfinite(mll_new):
                lr *= 0.5
            else:
                break
        mode = mode_new
        if it == 0:
            stop = abs(mll_new - mll) < cfg["delta_conv_mode_finding"] * abs(mll)
        else:
            stop = (mll_new - mll) < cfg["delta_conv_mode_finding"] * abs(mll)
        mll = mll_new
        if stop:
            break
    p_ = sigmoid_stable(F + mode)
    W = p_ * (1. - p_)
    out = dict(mode=mode, newton_it=it, cg_it=cg_total, mll_mode=mll)
    if method == "cholesky":
        lu = spl.splu((SigmaI + sp.diags(W)).tocsc())
        logdet_A = float(np.sum(np.log(np.abs(lu.U.diagonal()))) + np.sum(np.log(np.abs(lu.L.diagonal()))))
        ld = logdet_A - float(np.sum(np.log(Dinv)))
        out["slq_it"] = 0
    else:
        t = cfg["num_rand_vec_trace"]
        if probes is None:
            probes = gen_rand_normal(cfg["seed_rand_vec_trace"], 0, n, t)
        dw = Dinv + W
        Zp = Bt @ (np.sqrt(dw)[:, None] * probes)
        Td, Ts, its, AinvZ = cg_tridiag_vadu(B, Bt, Dinv, W, Zp, min(cfg["cg_max_num_it_tridiag"], n), cfg["cg_delta_conv"])
        out["_Zp"], out["_AinvZ"] = Zp, AinvZ
        ld = logdet_tridiag(Td, Ts, n) - float(np.sum(np.log(Dinv))) + float(np.sum(np.log(dw)))
        out["slq_it"] = its
    out["logdet"] = ld
    out["negll"] = -(mll - 0.5 * ld)
    out["_state"] = dict(A=A, Dinv=Dinv, B=B, Bt=Bt, W=W, p=p_, F=F, cfg=cfg)
    return out


def factor_latent_grad(coords_ordered, nn, cid, var, range_trans):
    """A, D^-1 and d/dlog(range) of A (= -B_grad) and of D for the latent factor (oracle C restatement)."""
    n, d = coords_ordered.shape
    m = nn.shape[1]
    cm = np.asfortranarray(coords_ordered, dtype=np.float64)
    A = np.empty((n, m)); Dinv = np.empty(n); Ag = np.empty((n, m)); Dg = np.empty(n)
    pt = np.array([var, range_trans], dtype=np.float64)
    p = lambda a, t=C.c_double: a.ctypes.data_as(C.POINTER(t))
    nn = np.ascontiguousarray(nn, dtype=np.int32)
    f = ov.lib().orc_vecchia_factor_latent_grad
    f.restype = C.c_int
    bad = f(p(cm), C.c_int(n), C.c_int(d), C.c_int(m), p(nn, C.c_int32), C.c_int(cid), p(pt), p(A), p(Dinv), p(Ag), p(Dg))
    return A, Dinv, Ag, Dg, bad


def _optimal_c(za, zb, tra, trb):  # CalcOptimalC
    cb = zb - trb
    den = float(np.mean(cb * cb))
    return 1. if den == 0 else float(np.mean((za - tra) * cb)) / den


def grad_negll(coords_ordered, nn, cid, var, range_trans, y, fixed_effects=None, method="cholesky", probes=None, **kw):
    """Gradient of the Laplace-approximated negative log-likelihood w.r.t. (log var, log range) at the mode
    (CalcGradNegMargLikelihoodLaplaceApproxVecchia with one GP: parameter 0 = marginal variance, 1 = range).
    Returns dict(grad, negll, ...). `method` selects the reference's Cholesky branch (exact traces) or its iterative branch
    (stochastic traces with the SLQ probe vectors and solutions, VADU variance reduction, implicit term by PCG)."""
    res = negll(coords_ordered, nn, cid, var, range_trans, y, fixed_effects=fixed_effects, method=method, probes=probes, **kw)
    st = res["_state"]
    B, Bt, Dinv, W, p_, cfg = st["B"], st["Bt"], st["Dinv"], st["W"], st["p"], st["cfg"]
    n = y.shape[0]
    mode = res["mode"]
    A_, Dinv2, Ag, Dg, bad = factor_latent_grad(coords_ordered, nn, cid, var, range_trans)
    assert bad == 0 and np.allclose(Dinv2, Dinv, rtol=0, atol=0)
    nn = np.asarray(nn)
    m = nn.shape[1]
    rows = np.repeat(np.arange(n), m)
    mask = nn.ravel() >= 0
    Bg = sp.csr_matrix((-Ag.ravel()[mask], (rows[mask], nn.ravel()[mask])), shape=(n, n))  # B_grad (zero diagonal)
    Dm = sp.diags(Dinv)
    SigmaI = (Bt @ Dm @ B).tocsr()
    X1 = (Bg.T @ Dm @ B)
    SigmaI_deriv = [(-SigmaI).tocsr(), (X1 + X1.T - Bt @ sp.diags(Dinv * Dg * Dinv) @ B).tocsr()]
    dW = p_ * (1. - p_) * (1. - 2. * p_)  # CalcFirstDerivInformationLocPar, bernoulli_logit
    grad = np.zeros(2)
    if method == "cholesky":
        Ainv = np.linalg.inv((SigmaI + sp.diags(W)).toarray())
        d_mll_d_mode = 0.5 * np.diag(Ainv) * dW
        Ainv_dmll = Ainv @ d_mll_d_mode
        for j in range(2):
            Sd = SigmaI_deriv[j]
            Sd_mode = Sd @ mode
            explicit = 0.5 * (mode @ Sd_mode + float(Sd.multiply(Ainv).sum()))
            explicit += 0.5 * n if j == 0 else 0.5 * float(np.sum(Dinv * Dg))
            grad[j] = explicit - Ainv_dmll @ Sd_mode
    else:
        Zp, AinvZ = res["_Zp"], res["_AinvZ"]
        dw = Dinv + W
        PI_Z = _vadu_solve(B, Bt, dw, Zp)
        # d log|Sigma W + I| / d mode (CalcLogDetStochDerivModeVecchia, VADU)
        ZA = AinvZ * dW[:, None] * PI_Z
        trA = ZA.mean(1)
        Dw_inv = 1. / dw
        trD = Dw_inv * dW
        BPZ = B @ PI_Z
        ZP = BPZ * dW[:, None] * BPZ
        trP = ZP.mean(1)
        cA = ZA - trA[:, None]; cP = ZP - trP[:, None]
        c_var = (cP * cP).mean(1)
        with np.errstate(divide="ignore", invalid="ignore"):
            c_opt = np.where(c_var == 0, 1., (cA * cP).mean(1) / c_var)
        d_logdet_d_mode = trA + c_opt * trD - c_opt * trP
        d_mll_d_mode = 0.5 * d_logdet_d_mode
        Ainv_dmll, _ = cg_vadu(B, Bt, Dinv, W, d_mll_d_mode, np.zeros(n), cfg["cg_max_num_it"], cfg["cg_delta_conv"], True)
        for j in range(2):
            Sd = SigmaI_deriv[j]
            zA = (AinvZ * (Sd @ PI_Z)).sum(0)
            tr1 = float(zA.mean())
            d = tr1 + (n if j == 0 else float(np.sum(Dinv * Dg)))
            if j == 0:
                trDd = -float(np.sum(Dw_inv * Dinv))
                zP = (PI_Z * (Sd @ PI_Z)).sum(0)
            else:
                trDd = -float(np.sum(Dw_inv * Dinv * Dg * Dinv))
                BtWBg = (Bt @ sp.diags(W) @ Bg)
                Pd = (Sd + BtWBg.T + BtWBg).tocsr()
                zP = (PI_Z * (Pd @ PI_Z)).sum(0)
            trPd = float(zP.mean())
            c = _optimal_c(zA, zP, tr1, trPd)
            d += c * trDd - c * trPd
            Sd_mode = Sd @ mode
            grad[j] = 0.5 * (mode @ Sd_mode + d) - Ainv_dmll @ Sd_mode
    # The optimiser's variable for non-Gaussian likelihoods is log(cov_pars) on the ORIGINAL scale. range_trans = c / range for
    # the Matern family and the exponential, 1 / range^2 for the Gaussian kernel (cov_fcts.h:485-552), so the chain rule gives
    # d/dlog(range) = -1 resp. -2 times d/dlog(range_trans). The reference's gradient for the Gaussian kernel is -1/2 times
    # (a quarter of the derivative of its own likelihood: central differences of GPB_EvalNegLogLikelihood give the -2 value,
    # tests/golden/make_laplace_golden.py records what its optimiser uses). "grad" reproduces the REFERENCE (parity is defined
    # against it); "grad_consistent" is the derivative of the likelihood this module evaluates.
    res["grad_trans"] = grad.copy()
    res["grad_consistent"] = grad * np.array([1., -2. if cid == 3 else -1.])
    grad[1] *= -0.5 if cid == 3 else -1.
    res["grad"] = grad
    return res
