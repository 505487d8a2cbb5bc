"""Batched Riccati backward pass on the device — the consumer of fx / fu inside the reference's outer solver.

IterativeLQR's backward pass (run by iLQR.solve!, reference examples/hopper.jl:292) is sequential in t and needs every knot
point's Jacobians.  `backward_pass_batch` runs NT trajectories side by side (one warp each) directly on the packed rows that
`ImplicitDynamics.step_grad_packed` / the fused all-gather produced, and returns gains in the layout `rollout_batch` consumes.
No CPU fallback.
"""
import numpy as np

from . import _lib
from .dynamics import ImplicitDynamics, _dp, _ip


def backward_pass_batch(im_dyn: ImplicitDynamics, jac, lx, lu, lxx, luu, lux=None, reg=0.0):
    """jac: [NT, T-1, nq + nq(2nq+nu)] packed rows; lx [NT,T,n], lu [NT,T-1,m], lxx [NT,T,n,n], luu [NT,T-1,m,m], lux [NT,T-1,m,n].
    Returns K [NT,T-1,m,n], k [NT,T-1,m], dV [NT,2], status [NT]."""
    nq, m = im_dyn.nq, im_dyn.nu
    n = 2 * nq
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    jac, lx, lu, lxx, luu = c(jac), c(lx), c(lu), c(lxx), c(luu)
    NT, S = jac.shape[0], jac.shape[1]
    T = S + 1
    if jac.shape != (NT, S, nq + nq * (n + m)) or lx.shape != (NT, T, n) or lu.shape != (NT, S, m) or lxx.shape != (NT, T, n, n) \
            or luu.shape != (NT, S, m, m):
        raise ValueError("backward_pass_batch: inconsistent shapes")
    if lux is not None:
        lux = c(lux)
        if lux.shape != (NT, S, m, n):
            raise ValueError("backward_pass_batch: lux must be [NT, T-1, m, n]")
    K = np.empty((NT, S, m, n)); k = np.empty((NT, S, m)); dV = np.empty((NT, 2)); st = np.empty(NT, dtype=np.int32)
    _lib.check(_lib.lib().od_riccati_batch(im_dyn._handle(), NT, T, _dp(jac), _dp(lx), _dp(lu), _dp(lxx), _dp(luu),
                                           None if lux is None else _dp(lux), float(reg), _dp(K), _dp(k), _dp(dV), _ip(st)))
    return K, k, dV, st
