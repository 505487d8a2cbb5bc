"""ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (oracle/README.md).

ctypes front-end of oracle/liboracle.so (the C++ fp64 restatement of the reference's CPU path).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the product package
(optimization_dynamics_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODELS = {"acrobot_impact": 0, "acrobot_nominal": 1, "cartpole_friction": 2, "cartpole_frictionless": 3,
          "planar_push": 4, "hopper": 5, "rocket": 6, "rocket_proj": 7}


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_api.cpp", "ip.hpp", "models.hpp", "dual.hpp")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def set_variant(tau_rule=0, apply_reg=0, mu_mode=0, soc_tau_cap=0):
    """Select a reading of the solver choices the reference tree does not pin (oracle/ip.hpp Options).  All zero = default."""
    lib().od_oracle_set_variant(int(tau_rule), int(apply_reg), int(mu_mode), int(soc_tau_cap))


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def dims(model):
    nq, nu, nz, nth = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    assert lib().od_oracle_dims(MODELS[model], C.byref(nq), C.byref(nu), C.byref(nz), C.byref(nth)) == 0
    return nq.value, nu.value, nz.value, nth.value


def residual(model, z, th, kappa=0.0):
    mid = MODELS[model]
    nz, nth = len(z), len(th)
    z = np.ascontiguousarray(z, dtype=np.float64); th = np.ascontiguousarray(th, dtype=np.float64)
    r = np.zeros(nz); rz = np.zeros((nz, nz)); rth = np.zeros((nz, nth))
    assert lib().od_oracle_residual(mid, _p(z), _p(th), C.c_double(kappa), _p(r), _p(rz), _p(rth)) == 0
    return r, rz, rth


def step_batch(model, q1, q2, u, h, kappa_tol, diff, fric=None, r_tol=1e-8, nthreads=0, full=False, diagnostics=True):
    """Returns dict(q3, dq1, dq2, du (column-major per problem, as (B, ncol, nq) arrays), iters, status, ls, r_vio, k_vio, margin)."""
    mid = MODELS[model]
    nq, nu, nz, nth = dims(model)
    q1 = np.ascontiguousarray(q1, dtype=np.float64).reshape(-1, nq); B = q1.shape[0]
    q2 = np.ascontiguousarray(q2, dtype=np.float64).reshape(B, nq)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(B, nu)
    fr = None if fric is None else np.ascontiguousarray(fric, dtype=np.float64)
    q3 = np.zeros((B, nq)); info = np.zeros((B, 4), dtype=np.int32); vio = np.zeros((B, 5)) if diagnostics else None
    dq1 = dq2 = du = dzf = None
    if diff:
        dq1 = np.zeros((B, nq, nq)); dq2 = np.zeros((B, nq, nq)); du = np.zeros((B, nu, nq))
        if full:
            dzf = np.zeros((B, nz, nth))
    zout = np.zeros((B, nz))
    rc = lib().od_oracle_step_batch(mid, B, _p(q1), _p(q2), _p(u), _p(fr), C.c_double(h), C.c_double(r_tol), C.c_double(kappa_tol),
                                    int(bool(diff)), _p(q3), _p(dq1), _p(dq2), _p(du), _p(dzf), _p(zout), _p(info, C.c_int), _p(vio), int(nthreads))
    assert rc == 0
    out = dict(q3=q3, dq1=dq1, dq2=dq2, du=du, dz_full=dzf, z=zout, iters=info[:, 0].copy(), status=info[:, 1].copy(), ls=info[:, 2].copy())
    if diagnostics:
        out.update(r_vio=vio[:, 0].copy(), k_vio=vio[:, 1].copy(), margin=vio[:, 2].copy(), ift_spread=vio[:, 3].copy(), q_uncertainty=vio[:, 4].copy())
    return out


def rocket_batch(x, u, h, u_max, proj, diff, nthreads=0):
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 12); B = x.shape[0]
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(B, 3)
    y = np.zeros((B, 12)); dx = np.zeros((B, 12, 12)) if diff else None; du = np.zeros((B, 3, 12)) if diff else None
    up = np.zeros((B, 3)); info = np.zeros((B, 4), dtype=np.int32); vio = np.zeros((B, 3))
    rc = lib().od_oracle_rocket_batch(B, _p(x), _p(u), C.c_double(h), C.c_double(u_max), int(bool(proj)), int(bool(diff)),
                                      _p(y), _p(dx), _p(du), _p(up), _p(info, C.c_int), _p(vio), int(nthreads))
    assert rc == 0
    return dict(y=y, dx=dx, du=du, uproj=up, iters=info[:, 0].copy(), status=info[:, 1].copy(), proj_iters=info[:, 3].copy(), margin=vio[:, 2].copy(),
                r_vio=vio[:, 0].copy())


def rocket_projection_batch(u, u_max, diff=True):
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, 3); B = u.shape[0]
    up = np.zeros((B, 3)); dp = np.zeros((B, 3, 3)); info = np.zeros((B, 4), dtype=np.int32); vio = np.zeros((B, 3))
    rc = lib().od_oracle_rocket_projection_batch(B, _p(u), C.c_double(u_max), int(bool(diff)), _p(up), _p(dp), _p(info, C.c_int), _p(vio))
    assert rc == 0
    return dict(up=up, dproj=dp, iters=info[:, 0].copy(), status=info[:, 1].copy(), margin=vio[:, 2].copy(), k_vio=vio[:, 1].copy())


def bundle_batch(model, eta, q1, q2, u, h, kappa_tol, fric=None, r_tol=1e-8, nthreads=0):
    mid = MODELS[model]
    nq, nu, nz, nth = dims(model)
    ncol = 2 * nq + nu
    eta = np.ascontiguousarray(eta, dtype=np.float64).reshape(-1, ncol); N = eta.shape[0]
    q1 = np.ascontiguousarray(q1, dtype=np.float64).reshape(-1, nq); B = q1.shape[0]
    q2 = np.ascontiguousarray(q2, dtype=np.float64).reshape(B, nq)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(B, nu)
    fr = None if fric is None else np.ascontiguousarray(fric, dtype=np.float64)
    dz = np.zeros((B, ncol, nq)); st = np.zeros(B, dtype=np.int32)
    rc = lib().od_oracle_bundle_batch(mid, B, N, _p(eta), _p(q1), _p(q2), _p(u), _p(fr), C.c_double(h), C.c_double(r_tol), C.c_double(kappa_tol),
                                      _p(dz), _p(st, C.c_int), int(nthreads))
    assert rc == 0
    return dict(dz=dz, status=st)


def least_squares(fz, feta, eta):
    fz = np.ascontiguousarray(fz, dtype=np.float64); ny = fz.shape[0]
    feta = np.ascontiguousarray(feta, dtype=np.float64).reshape(-1, ny); N = feta.shape[0]
    eta = np.ascontiguousarray(eta, dtype=np.float64).reshape(N, -1); nz = eta.shape[1]
    theta = np.zeros(ny * nz)
    rc = lib().od_oracle_least_squares(N, ny, nz, _p(fz), _p(feta), _p(eta), _p(theta))
    return rc, theta.reshape(nz, ny).T   # reshape(θ, ny, nz) column-major


def rollout_batch(model, x1, ubar, h, kappa_tol, xbar=None, K=None, k=None, alpha=None, fric=None, r_tol=1e-8):
    """Restatement of the reference's rollout loop: iLQR.rollout(model, x1, ū) (x[t+1] = f(x[t], ū[t]); reference
    examples/cartpole.jl:79) and the forward pass of IterativeLQR (u[t] = ū[t] + α k[t] + K[t](x[t] − x̄[t]); external package,
    step sizes α ≥ α_min of examples/cartpole.jl:86), with f of reference src/dynamics.jl:81-94 — one eval-simulator step per
    knot point.  x1: [R, 2nq]; ubar: [T-1, nu] or [R, T-1, nu].  Returns X [R, T, 2nq], U [R, T-1, nu], status [R, T-1]."""
    nq, nu, nz, nth = dims(model)
    x1 = np.ascontiguousarray(x1, dtype=np.float64).reshape(-1, 2 * nq); R = x1.shape[0]
    ubar = np.asarray(ubar, dtype=np.float64)
    S = ubar.shape[-2]; T = S + 1
    X = np.zeros((R, T, 2 * nq)); U = np.zeros((R, S, nu)); st = np.zeros((R, S), dtype=np.int32)
    al = np.ones(R) if alpha is None else np.asarray(alpha, dtype=np.float64)
    X[:, 0] = x1
    for t in range(S):
        u = np.broadcast_to(ubar[t] if ubar.ndim == 2 else ubar[:, t], (R, nu)).copy()
        if k is not None:
            u = u + al[:, None] * np.asarray(k)[t][None, :]
        if K is not None:
            u = u + np.einsum("uj,rj->ru", np.asarray(K)[t], X[:, t] - np.asarray(xbar)[t][None, :])
        U[:, t] = u
        o = step_batch(model, X[:, t, :nq], X[:, t, nq:], u, h, kappa_tol, False, fric=fric, r_tol=r_tol, diagnostics=False)
        X[:, t + 1, :nq] = X[:, t, nq:]
        X[:, t + 1, nq:] = o["q3"]
        st[:, t] = o["status"]
    return X, U, st


def fx_fu_from_packed(row, nq, nu):
    """fx = [0 I; ∂q3∂q1 ∂q3∂q2], fu = [0; ∂q3∂u1] from one packed output row (reference src/dynamics.jl:105-111,125)."""
    n = 2 * nq
    d1 = row[nq:nq + nq * nq].reshape(nq, nq).T
    d2 = row[nq + nq * nq:nq + 2 * nq * nq].reshape(nq, nq).T
    du = row[nq + 2 * nq * nq:].reshape(nu, nq).T
    fx = np.zeros((n, n)); fx[:nq, nq:] = np.eye(nq); fx[nq:, :nq] = d1; fx[nq:, nq:] = d2
    fu = np.zeros((n, nu)); fu[nq:] = du
    return fx, fu


def backward_pass(jac, lx, lu, lxx, luu, lux, nq, nu, reg=0.0):
    """Restatement of the iLQR backward pass (IterativeLQR.jl, external to the reference tree; textbook Riccati recursion without
    regularisation schedule) for ONE trajectory: returns K [T-1,m,n], k [T-1,m], dV [2], status."""
    S = jac.shape[0]; n = 2 * nq
    P = lxx[S].copy(); p = lx[S].copy()
    K = np.zeros((S, nu, n)); k = np.zeros((S, nu)); dV = np.zeros(2); status = 0
    for t in range(S - 1, -1, -1):
        fx, fu = fx_fu_from_packed(jac[t], nq, nu)
        Qx = lx[t] + fx.T @ p; Qu = lu[t] + fu.T @ p
        Qxx = lxx[t] + fx.T @ P @ fx
        Quu = luu[t] + fu.T @ P @ fu + reg * np.eye(nu)
        Qux = (0.0 if lux is None else lux[t]) + fu.T @ P @ fx
        try:
            np.linalg.cholesky(Quu)
            Kt = -np.linalg.solve(Quu, Qux); kt = -np.linalg.solve(Quu, Qu)
        except np.linalg.LinAlgError:
            status = 1; Kt = np.zeros((nu, n)); kt = np.zeros(nu)
        K[t] = Kt; k[t] = kt
        dV += [kt @ Qu, 0.5 * kt @ Quu @ kt]
        P = Qxx + Kt.T @ Quu @ Kt + Kt.T @ Qux + Qux.T @ Kt
        p = Qx + Kt.T @ Quu @ kt + Kt.T @ Qu + Qux.T @ kt
    return K, k, dV, status


def num_threads():
    return lib().od_oracle_num_threads()
