"""Batched rollouts through the implicit dynamics — the caller of `f` in the reference's outer solver.

    rollout(im_dyn, x1, ū)                     iLQR.rollout(model, x1, ū)  (reference examples/cartpole.jl:79, acrobot.jl:92,
                                               planar_push.jl:113, hopper.jl:272): x[t+1] = f(x[t], ū[t])
    rollout_batch(im_dyn, x1, ū, x̄, K, k, α)   IterativeLQR's forward pass for R step sizes / rollouts at once:
                                               u[t] = ū[t] + α k[t] + K[t] (x[t] − x̄[t]),  x[t+1] = f(x[t], u[t])
One kernel launch runs all T−1 sequential contact solves of every rollout (csrc/contact_ip.cuh: contact_rollout_kernel); the
reference makes (T−1)·R calls of `f`.  No CPU fallback.
"""
import numpy as np

from . import _lib
from .dynamics import ImplicitDynamics, _dp, _ip


def rollout_batch(im_dyn: ImplicitDynamics, x1, ubar, xbar=None, K=None, k=None, alpha=None, return_status=False):
    """x1: [R, 2nq] (or [2nq] broadcast over len(alpha)); ubar: [T-1, nu] shared or [R, T-1, nu]; xbar: [T, 2nq];
    K: [T-1, nu, 2nq]; k: [T-1, nu]; alpha: [R].  Returns X [R, T, 2nq], U [R, T-1, nu] (and status [R, T-1])."""
    nq, nu = im_dyn.nq, im_dyn.nu
    nx = 2 * nq
    ubar = np.ascontiguousarray(ubar, dtype=np.float64)
    per_rollout = ubar.ndim == 3
    S = ubar.shape[-2]
    T = S + 1
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    if x1.ndim == 1:
        R = len(alpha) if alpha is not None else (ubar.shape[0] if per_rollout else 1)
        x1 = np.ascontiguousarray(np.broadcast_to(x1, (R, nx)))
    R = x1.shape[0]
    if x1.shape != (R, nx) or ubar.shape[-1] != nu or (per_rollout and ubar.shape[0] != R):
        raise ValueError("rollout_batch: inconsistent shapes")

    def opt(a, shape, name):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != shape:
            raise ValueError("rollout_batch: %s must have shape %r" % (name, shape))
        return a
    xbar = opt(xbar, (T, nx), "xbar"); K = opt(K, (S, nu, nx), "K"); k = opt(k, (S, nu), "k"); alpha = opt(alpha, (R,), "alpha")
    X = np.empty((R, T, nx)); U = np.empty((R, S, nu)); st = np.empty((R, S), dtype=np.int32)
    p = lambda a: None if a is None else _dp(a)
    _lib.check(_lib.lib().od_rollout_batch(im_dyn._handle(), R, T, _dp(x1), _dp(ubar), int(per_rollout), p(xbar), p(K), p(k), p(alpha),
                                           _dp(X), _dp(U), _ip(st)))
    return (X, U, st) if return_status else (X, U)


def rollout(im_dyn: ImplicitDynamics, x1, ubar):
    """iLQR.rollout(model, x1, ū) → list of T states x[t] = [q1; q2]."""
    X, _ = rollout_batch(im_dyn, np.asarray(x1, dtype=np.float64)[None], np.asarray(ubar, dtype=np.float64))
    return [X[0, t].copy() for t in range(X.shape[1])]
