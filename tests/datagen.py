"""Deterministic test data shared by the oracle-pinning tests, the GPU parity tests and the golden-vector script.

`sim_rand_unif` restates the pure-arithmetic LCG the reference's R tests use to build their data
(R-package/tests/testthat/test_GPModel_gaussian_process.R:37-43), in float64 like R."""
import numpy as np


def sim_rand_unif(n, init_c, mod=2.0 ** 32, a=22695477.0, c=1.0):
    s = np.empty(n)
    s[0] = np.floor(init_c * mod)
    for i in range(1, n):
        s[i] = (a * s[i - 1] + c) % mod
    return s / mod


def r_test_data():
    """n=100 coordinates and response of test_GPModel_gaussian_process.R:46-66."""
    from scipy.stats import norm
    n, d = 100, 2
    coords = sim_rand_unif(n * d, 0.1).reshape((n, d), order="F")
    D = np.sqrt(((coords[:, None, :] - coords[None, :, :]) ** 2).sum(-1))
    eps = np.linalg.cholesky(np.exp(-D / 0.1) + 1e-20 * np.eye(n)) @ norm.ppf(sim_rand_unif(n, 0.8))
    xi = norm.ppf(sim_rand_unif(n, 0.1)) / 5
    return coords, eps + xi


def synth(n, d=2, seed=1, smooth=True):
    """Synthetic coords U[0,1]^d and a response with spatial structure + noise (seeded, numpy PCG64)."""
    rng = np.random.default_rng(seed)
    coords = rng.random((n, d))
    y = rng.standard_normal(n) * 0.5
    if smooth:
        y = y + np.sin(4 * coords[:, 0]) + np.cos(3 * coords[:, -1]) * coords[:, 0]
    return coords, y


def lattice(k):
    """k x k regular lattice: ubiquitous distance ties — the neighbour tie-break rules matter here."""
    g = np.arange(k) / k
    xx, yy = np.meshgrid(g, g, indexing="ij")
    return np.stack([xx.ravel(), yy.ravel()], axis=1)
