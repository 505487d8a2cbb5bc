"""Gradient bundle (zeroth-order Jacobian) — mirror of reference src/gradient_bundle.jl and src/ls.jl.

    GradientBundle(model; N=100, ϵ=1e-4)        reference src/gradient_bundle.jl:26-85
    gradient!(sim, gb, q1, q2, u1)              reference src/gradient_bundle.jl:87-104   → `gradient(im_dyn, gb, q1, q2, u1)`
    fx_gb / fu_gb                               reference src/gradient_bundle.jl:109-147

The N perturbed steps and the nominal one run as ONE kernel launch ((N+1)·B problems); the least-squares fit of src/ls.jl
(one Newton step on a quadratic) is done in closed form on the device.  The reference's `zeros(nq)` bug (module-global nq,
src/gradient_bundle.jl:79-80) is not reproduced: buffers are sized from model.nq.
"""
import numpy as np

from . import _lib
from .dynamics import _dp, _ip, _f64


class GradientBundle:
    def __init__(self, model, N=100, ϵ=1.0e-4, eps=None, rng=None, eta=None):
        eps = ϵ if eps is None else eps
        self.ny = model.nq
        self.nz = 2 * model.nq + model.nu
        self.N = N
        if eta is None:
            # N fixed one-hot perturbations: coordinate rand(1:nz), magnitude ϵ·randn()  (src/gradient_bundle.jl:49-54)
            rng = np.random.default_rng() if rng is None else rng
            eta = np.zeros((N, self.nz))
            for i in range(N):
                eta[i, rng.integers(0, self.nz)] = eps * rng.normal()
        self.eta = np.ascontiguousarray(eta, dtype=np.float64).reshape(-1, self.nz)
        self.N = self.eta.shape[0]
        self.eps = eps
        self.dz = np.zeros((self.ny, self.nz))

    def unsampled(self):
        """Coordinates of [q1; q2; u1] that no perturbation touches: the fit's normal matrix Σ ηηᵀ is singular in them (the reference
        draws the N coordinates at random once, src/gradient_bundle.jl:49-54, and then fails in its LU; here the fit reports it)."""
        return np.flatnonzero(~(self.eta != 0.0).any(axis=0))

    def resample(self, rng=None, cover=True):
        """Draw the N one-hot perturbations again (same law as the constructor).  cover=True (needs N ≥ nz): the first nz draws take
        the coordinates 0 … nz−1 once each, so every column of the Jacobian is determined."""
        rng = np.random.default_rng() if rng is None else rng
        eta = np.zeros((self.N, self.nz))
        for i in range(self.N):
            j = i if (cover and self.N >= self.nz and i < self.nz) else rng.integers(0, self.nz)
            w = self.eps * rng.normal()
            eta[i, j] = w if w != 0.0 else self.eps
        self.eta = eta
        return self


def gradient_batch(im_dyn, gb, q1, q2, u):
    """dz[B, nq, 2nq+nu] ≈ ∂q3/∂[q1; q2; u1] for B problems, status[B]."""
    nq, nu = im_dyn.nq, im_dyn.nu
    q1 = _f64(q1, (-1, nq)); B = q1.shape[0]
    q2 = _f64(q2, (B, nq)); u = _f64(u, (B, nu))
    dz = np.empty((B, gb.nz, nq)); st = np.empty(B, dtype=np.int32)
    _lib.check(_lib.lib().od_bundle_batch(im_dyn._handle(), B, gb.N, _dp(gb.eta), _dp(q1), _dp(q2), _dp(u), _dp(dz), _ip(st)))
    return dz.transpose(0, 2, 1), st


def gradient(im_dyn, gb, q1, q2, u1):
    """gradient!(sim, gb, q1, q2, u1) → gb.dz (nq × (2nq+nu))."""
    dz, _ = gradient_batch(im_dyn, gb, np.asarray(q1)[None], np.asarray(q2)[None], np.asarray(u1)[None])
    gb.dz[...] = dz[0]
    return gb.dz


def fx_gb(dx, model, x, u, w):
    """reference src/gradient_bundle.jl:109-126 (model.info is the GradientBundle)."""
    x = np.asarray(x, dtype=np.float64); u = np.asarray(u, dtype=np.float64)
    q1 = x[model.idx_q1]; q2 = x[model.idx_q2]
    nq = model.nq
    for i in range(nq):
        dx[model.idx_q1[i], model.idx_q2[i]] = 1.0
    dz = gradient(model, model.info, q1, q2, u[model.idx_u1])
    dx[np.ix_(model.idx_q2, model.idx_q1)] = dz[:, :nq]
    dx[np.ix_(model.idx_q2, model.idx_q2)] = dz[:, nq:2 * nq]
    return dx


def fu_gb(du, model, x, u, w):
    """reference src/gradient_bundle.jl:136-147."""
    x = np.asarray(x, dtype=np.float64); u = np.asarray(u, dtype=np.float64)
    q1 = x[model.idx_q1]; q2 = x[model.idx_q2]
    nq = model.nq
    dz = gradient(model, model.info, q1, q2, u[model.idx_u1])
    du[model.idx_q2, :] = dz[:, 2 * nq:]
    return du
