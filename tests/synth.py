"""Seeded synthetic simulator of SURVEY.md section 8(d) (shared by tests, bench.py, golden script).

theta ~ U(0,1)^{n x d}; latent G = sin(1.5 theta W + phi), W ~ N(0,1)^{d x k}, phi ~ U(0, 2 pi)^k;
loadings Bm ~ N(0,1)^{k x m} scaled by 0.85^j per row; f = (G Bm + 1e-5 N(0,1))' is (m, n).
Yields exactly k principal components under the default epsilonPC = 0.001 for complete data.
"""
import numpy as np


def make_simulator(n, d, m, k, seed=0, missing=0.0, noise=1e-5):
    rng = np.random.default_rng(seed)
    theta = rng.uniform(size=(n, d))
    W = rng.normal(size=(d, k))
    ph = rng.uniform(0, 2 * np.pi, size=k)
    Bm = rng.normal(size=(k, m)) * (0.85 ** np.arange(k))[:, None]

    def latent(th):
        return np.sin(1.5 * th @ W + ph)

    f = (latent(theta) @ Bm + noise * rng.normal(size=(n, m))).T          # (m, n)
    if missing > 0:
        f = f.copy()
        f[rng.uniform(size=f.shape) < missing] = np.nan
    x = np.arange(m, dtype=float)[:, None]

    def truth(th):
        return (latent(np.atleast_2d(th)) @ Bm).T                          # (m, B)
    return x, theta, f, truth


def make_observation(truth, d, seed=1, sd=0.05):
    rng = np.random.default_rng(seed)
    theta_true = rng.uniform(0.3, 0.7, size=(1, d))
    y = truth(theta_true)[:, 0]
    y = y + sd * rng.normal(size=y.shape)
    yvar = np.full(y.shape, sd ** 2)
    return theta_true, y, yvar
