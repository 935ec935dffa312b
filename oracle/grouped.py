"""TEST INFRASTRUCTURE ONLY — numpy restatement of the single-level grouped random-effects Gaussian likelihood.
`negll_woodbury` follows the reference's Woodbury path (re_model_template.h:9417-9420 diag(Sigma^-1 + Z^T Z), :9907-9918 y_tilde,
:3029-3031 log-det, :3132 formula); `negll_dense` evaluates the same likelihood from Psi = I + v Z Z^T directly (small n)."""
import numpy as np


def _index(group):
    _, gi = np.unique(np.asarray(group), return_inverse=True)
    return gi


def negll_woodbury(group, y, cov_pars):
    s2, s1 = float(cov_pars[0]), float(cov_pars[1])
    v = s1 / s2
    gi = _index(group)
    y = np.asarray(y, dtype=np.float64)
    n = y.shape[0]
    ng = np.bincount(gi).astype(np.float64)
    sg = np.bincount(gi, weights=y)
    quad = y @ y - np.sum(sg * sg / (1. / v + ng))
    logdet = np.sum(np.log(1. + v * ng))
    return quad / 2. / s2 + logdet / 2. + n / 2. * (np.log(s2) + np.log(2 * np.pi))


def negll_dense(group, y, cov_pars):
    s2, s1 = float(cov_pars[0]), float(cov_pars[1])
    gi = _index(group)
    y = np.asarray(y, dtype=np.float64)
    n = y.shape[0]
    Z = np.zeros((n, gi.max() + 1)); Z[np.arange(n), gi] = 1.
    Psi = np.eye(n) + (s1 / s2) * Z @ Z.T
    sign, logdet = np.linalg.slogdet(Psi)
    quad = y @ np.linalg.solve(Psi, y)
    return quad / 2. / s2 + logdet / 2. + n / 2. * (np.log(s2) + np.log(2 * np.pi))


def grad_response(group, y, cov_pars):
    """Psi^-1 y / sigma^2 (CalcYAux single-RE branch :9843-9891, CalcGradientF :3298)."""
    s2, s1 = float(cov_pars[0]), float(cov_pars[1])
    v = s1 / s2
    gi = _index(group)
    y = np.asarray(y, dtype=np.float64)
    ng = np.bincount(gi).astype(np.float64)
    sg = np.bincount(gi, weights=y)
    return (y - (sg / (1. / v + ng))[gi]) / s2
