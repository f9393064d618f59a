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


# analytic log-posteriors for the sampler tests: every row is evaluated with elementwise arithmetic only, so a
# row's value does not depend on which other rows share the batch (needed to compare batched and one-at-a-time
# samplers draw for draw)
_MU = np.array([0.3, -0.2, 0.5])
_CI = np.linalg.inv(np.array([[0.04, 0.018, 0.0], [0.018, 0.02, 0.004], [0.0, 0.004, 0.01]]))


def gaussian_logpost(theta, return_grad=True):
    """Correlated 3-d Gaussian: (B, 1) values and (B, 3) gradients."""
    theta = np.atleast_2d(theta)
    r = theta - _MU
    q = np.zeros(theta.shape[0])
    g = np.zeros_like(r)
    for i in range(3):
        for j in range(3):
            q = q + r[:, i] * _CI[i, j] * r[:, j]
            g[:, i] = g[:, i] - _CI[i, j] * r[:, j]
    lp = (-0.5 * q).reshape(-1, 1)
    return (lp, g) if return_grad else lp


def boxed_logpost(theta, return_grad=False):
    """The same Gaussian restricted to [-1, 1]^3 (-inf outside), values only."""
    theta = np.atleast_2d(theta)
    lp = gaussian_logpost(theta, return_grad=False)
    lp[np.any(np.abs(theta) > 1, axis=1)] = -np.inf
    return lp


def box_draws(n):
    return np.random.uniform(-1, 1, size=(n, 3))
