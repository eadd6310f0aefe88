"""Minimal 2-state Gaussian HMM (Baum-Welch + Viterbi) -- TEST INFRASTRUCTURE.

The reference segments windows with ``hmmlearn.hmm.GaussianHMM(n_components=2, covariance_type="full")``
fitted on all window KLDs as one sequence (F:1539-1541) and decoded per scaffold (F:769).  hmmlearn is
not installed here (and not installable offline), and the reference does not seed it, so "identical HMM
calls" (north_star) is checked with this deterministic stand-in following hmmlearn's documented defaults:
uniform start/transition init, means initialised from the 2-quantile split of the data (deterministic
replacement for its k-means init), variances = data variance + 1e-3 (hmmlearn's min_covar), 10 EM
iterations (n_iter=10, tol=1e-2), Viterbi decoding.
"""
import numpy as np


def _log_gauss(x, mean, var):
    return -0.5 * (np.log(2 * np.pi * var)[None, :] + (x[:, None] - mean[None, :]) ** 2 / var[None, :])


def fit(x, n_iter=10, tol=1e-2, min_covar=1e-3):
    x = np.asarray(x, float)
    order = np.sort(x)
    half = len(x) // 2
    mean = np.array([order[:half].mean(), order[half:].mean()])
    var = np.full(2, x.var() + min_covar)
    start = np.full(2, 0.5)
    trans = np.full((2, 2), 0.5)
    prev = -np.inf
    for _ in range(n_iter):
        logb = _log_gauss(x, mean, var)
        n = len(x)
        la = np.zeros((n, 2)); lb = np.zeros((n, 2))
        la[0] = np.log(start) + logb[0]
        lt = np.log(trans)
        for t in range(1, n):
            la[t] = logb[t] + np.logaddexp(la[t - 1, 0] + lt[0], la[t - 1, 1] + lt[1])
        for t in range(n - 2, -1, -1):
            lb[t] = np.logaddexp(lt[:, 0] + logb[t + 1, 0] + lb[t + 1, 0], lt[:, 1] + logb[t + 1, 1] + lb[t + 1, 1])
        ll = np.logaddexp(la[-1, 0], la[-1, 1])
        gamma = np.exp(la + lb - ll)
        xi = np.exp(la[:-1, :, None] + lt[None] + (logb[1:] + lb[1:])[:, None, :] - ll).sum(0)
        start = gamma[0] / gamma[0].sum()
        trans = xi / xi.sum(1, keepdims=True)
        w = gamma.sum(0)
        mean = (gamma * x[:, None]).sum(0) / w
        var = (gamma * (x[:, None] - mean[None]) ** 2).sum(0) / w + min_covar
        if ll - prev < tol:
            break
        prev = ll
    return start, trans, mean, var


def predict(model, x):
    start, trans, mean, var = model
    x = np.asarray(x, float)
    logb = _log_gauss(x, mean, var)
    lt = np.log(trans)
    n = len(x)
    delta = np.log(start) + logb[0]
    back = np.zeros((n, 2), int)
    for t in range(1, n):
        cand = delta[:, None] + lt
        back[t] = cand.argmax(0)
        delta = cand.max(0) + logb[t]
    path = np.zeros(n, int)
    path[-1] = int(delta.argmax())
    for t in range(n - 1, 0, -1):
        path[t - 1] = back[t, path[t]]
    return path
