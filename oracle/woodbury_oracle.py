"""CPU oracle for the directbayeswoodbury log-likelihood (TEST INFRASTRUCTURE).

Numpy restatement of src/surmise/calibrationmethods/directbayeswoodbury.py:261-383 (loglik,
loglik_grad).  Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import it.

Parity pinning: as for oracle/pcgpwm_oracle.py -- against live outputs of the reference,
committed under tests/golden/ (the reference's tests hold no known-answer vectors).

`loglik_mspace` walks the reference's m-space algebra per theta (what the CPU baseline times);
`loglik_pcspace` is the k-space restatement the CUDA kernel implements:
    v = predvars[b], pv = predvecs[b], D = diag(sqrt v), G = Phi' Sigma^-1 Phi,
    r = y - offset - Phi pv, q = Phi' Sigma^-1 r, A = I + D G D,
    l = -1/2 log|A| - 1/2 (r' Sigma^-1 r - (Dq)' A^-1 (Dq)).
"""
import numpy as np

from . import pcgpwm_oracle as po


def extravar_trigger(var, covxhalf):
    """Batch-wide test that inflates obsvar                  (directbayeswoodbury.py:277-279)."""
    return bool(np.any(np.abs(var / (1e-4 + (1 + 1e-4) * np.sum(np.square(covxhalf), 2))) > 1))


def _obsvar(yvar, fitinfo_emu, pred, xind):
    """obsvar = yvar (+ mean_theta |var_old - sum covxhalf_old^2| when triggered)  (277-284)."""
    obsvar = 1 * np.squeeze(yvar)
    fired = extravar_trigger(pred['var'], pred['covxhalf'])
    if fired:
        old = {}
        po.predict(old, fitinfo_emu, None, fitinfo_emu['theta'], xind=xind)
        obsvar = obsvar + np.mean(np.abs(old['var'] - np.sum(np.square(old['covxhalf']), 2)), 1)
    return obsvar, fired


def loglik_mspace(fitinfo_emu, yvar, theta, y, xind=None, return_grad=False):
    """Reference algebra, one theta at a time                 (directbayeswoodbury.py:261-383)."""
    pred = {}
    po.predict(pred, fitinfo_emu, None, theta, return_grad=return_grad, xind=xind)
    obsvar, _ = _obsvar(yvar, fitinfo_emu, pred, xind)
    mean, cxh = pred['mean'], pred['covxhalf']
    B = mean.shape[1]
    d = fitinfo_emu['theta'].shape[1]
    ll = np.zeros((B, 1))
    dll = np.zeros((B, d))
    isd = 1 / np.sqrt(obsvar)
    for b in range(B):
        S0 = cxh[:, b, :]
        resid = (np.squeeze(y) - mean[:, b]) * isd
        J = S0 * isd[:, None]
        J2 = J.T @ resid
        W, V = np.linalg.eigh(np.eye(J.shape[1]) + J.T @ J)
        Ainv = (V / W) @ V.T
        J3 = Ainv @ J2
        ll[b, 0] = -0.5 * (np.sum(resid ** 2) - np.sum(J3 * J2)) - 0.5 * np.sum(np.log(W))
        if return_grad:
            dresid = -pred['mean_gradtheta'][:, b, :] * isd[:, None]        # (m, d)
            dterm1 = 2 * resid @ dresid
            V2 = (Ainv @ S0.T) / obsvar                                     # (k, m)
            dterm2 = np.zeros(d)
            dterm3 = np.zeros(d)
            for i in range(d):
                dJ = pred['covxhalf_gradtheta'][:, b, :, i].T * isd         # (k, m)
                dJ2 = J.T @ dresid[:, i] + dJ @ resid
                ex = dJ @ J
                ex = ex + ex.T
                dJ3 = Ainv @ (dJ2 - ex @ J3)
                dterm2[i] = np.sum(J2 * dJ3) + np.sum(dJ2 * J3)
                dterm3[i] = 2 * np.sum(V2 * pred['covxhalf_gradtheta'][:, b, :, i].T)
            dll[b, :] = -0.5 * dterm3 - 0.5 * (dterm1 - dterm2)
    return (ll, dll) if return_grad else ll


def woodbury_constants(phi, offset, extravar, yvar, y):
    """theta-independent pieces for both obsvar variants (index 0: yvar, 1: yvar+extravar)."""
    out = []
    for ov in (np.squeeze(yvar) * np.ones(phi.shape[0]), np.squeeze(yvar) + np.abs(extravar)):
        out.append({'sinv': 1 / ov, 'G': phi.T @ (phi / ov[:, None]), 'resid0': np.squeeze(y) - offset})
    return out


def loglik_pcspace(pv, pvar, phi, offset, extravar, yvar, y, dpv=None, dpvar=None):
    """k-space restatement used by the CUDA kernel; returns (ll (B,1), dll (B,d) or None, fired)."""
    B, k = pv.shape
    s = pvar @ (phi ** 2).T                                     # (B, m) sum_k covxhalf^2
    fired = bool(np.any(np.abs((extravar[None, :] + s) / (1e-4 + (1 + 1e-4) * s)) > 1))
    C = woodbury_constants(phi, offset, extravar, yvar, y)[1 if fired else 0]
    sinv, G = C['sinv'], C['G']
    ll = np.zeros((B, 1))
    dll = None if dpv is None else np.zeros((B, dpv.shape[2]))
    for b in range(B):
        D = np.sqrt(pvar[b])
        r = C['resid0'] - phi @ pv[b]
        w = sinv * r
        q = phi.T @ w
        A = np.eye(k) + (D[:, None] * G) * D[None, :]
        Lc = np.linalg.cholesky(A)
        Ainv = np.linalg.inv(A)
        Dq = D * q
        J3 = Ainv @ Dq
        ll[b, 0] = -np.sum(np.log(np.diag(Lc))) - 0.5 * (r @ w - Dq @ J3)
        if dpv is not None:
            M = G * D[None, :]                                   # G D
            Nvec = np.sum(M * Ainv.T, 1)                         # (G D A^-1)_aa
            MJ3 = M @ J3
            for i in range(dpv.shape[2]):
                dD = 0.5 * dpvar[b, :, i] / D
                dJ2 = -D * (G @ dpv[b, :, i]) + dD * q
                dterm1 = -2 * q @ dpv[b, :, i]
                dterm2 = 2 * dJ2 @ J3 - 2 * np.sum(dD * J3 * MJ3)
                dterm3 = 2 * np.sum(dD * Nvec)
                dll[b, i] = -0.5 * dterm3 - 0.5 * (dterm1 - dterm2)
    return ll, dll, fired


def loglik_from_prediction(mean, var, covxhalf, yvar, y, old_var=None, old_covxhalf=None):
    """directbayeswoodbury.loglik for ANY emulator, given that emulator's outputs     (directbayeswoodbury.py:271-306).

    mean, var (m, B), covxhalf (m, B, k) at the proposals; old_var (m, n), old_covxhalf (m, n, k) at the emulator's own
    training thetas, needed only when the batch-wide test fires.  Returns (ll (B, 1), fired)."""
    m, B = mean.shape
    obsvar = 1 * np.squeeze(yvar) * np.ones(m)
    fired = extravar_trigger(var, covxhalf)
    if fired:
        obsvar = obsvar + np.mean(np.abs(old_var - np.sum(np.square(old_covxhalf), 2)), 1)
    ll = np.zeros((B, 1))
    root = np.sqrt(obsvar)
    for b in range(B):
        z = (np.squeeze(y) - mean[:, b]) / root                 # standardised residual
        J = covxhalf[:, b, :] / root[:, None]
        W, V = np.linalg.eigh(np.eye(J.shape[1]) + J.T @ J)
        J2 = J.T @ z
        proj = V.T @ J2
        ll[b, 0] = -0.5 * (z @ z - np.sum(proj ** 2 / W)) - 0.5 * np.sum(np.log(W))
    return ll, fired

