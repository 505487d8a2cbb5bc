"""TEST INFRASTRUCTURE: ctypes front-end of tests/_build/libhostcheck.so (product solver templates compiled for the host CPU)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = None
MODELS = {"acrobot_impact": 0, "acrobot_nominal": 1, "cartpole_friction": 2, "cartpole_frictionless": 3, "planar_push": 4, "hopper": 5}
DIMS = {"acrobot_impact": (2, 1), "acrobot_nominal": (2, 1), "cartpole_friction": (2, 1), "cartpole_frictionless": (2, 1),
        "planar_push": (5, 2), "hopper": (4, 2)}


def build(flags=None, tag=""):
    """flags: the -D… switches of the build (None = the shipped library's, _lib.DEFAULT_DEFINES; a variant is built as a separate
    library, tag = its file-name suffix)."""
    if flags is None:
        from optimization_dynamics_b200 import _lib
        flags = _lib.DEFAULT_DEFINES
    so = os.path.join(_HERE, "_build", "libhostcheck%s.so" % tag)
    csrc = os.path.join(_ROOT, "optimization_dynamics_b200", "csrc")
    srcs = [os.path.join(_HERE, "host_check.cu")] + [os.path.join(dp, f) for dp, _, fs in os.walk(csrc) for f in fs]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared"] + list(flags) +
                              ["-o", so, os.path.join(_HERE, "host_check.cu")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


class use_variant:
    """with use_variant(["-DOD_EXTRACT_SMEM=1"], "_extract"): …   — the calls inside run a differently built copy of the templates."""

    def __init__(self, flags, tag):
        self.flags, self.tag = flags, tag

    def __enter__(self):
        global _LIB
        self.saved = _LIB
        _LIB = C.CDLL(build(self.flags, self.tag))
        return self

    def __exit__(self, *exc):
        global _LIB
        _LIB = self.saved


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def step(model, q1, q2, u, h, k_eval=1e-4, k_grad=1e-3, fric=None, want_eval=True, want_grad=True, eta=None, r_tol=1e-8, reg=False):
    """reg: False = shared-memory-LU templates, True / 1 = register path with one lane, 4 / 8 / 16 = register path with that many
    cooperating lanes per problem, run by a team of 32 lock-stepped host threads per emulated warp (cone models only)."""
    nq, nu = DIMS[model]
    q1 = np.ascontiguousarray(q1, dtype=np.float64).reshape(-1, nq); B0 = q1.shape[0]
    q2 = np.ascontiguousarray(q2, dtype=np.float64).reshape(B0, nq)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(B0, nu)
    n_eta = 0
    if eta is not None:
        eta = np.ascontiguousarray(eta, dtype=np.float64); n_eta = eta.shape[0]
    B = B0 * (n_eta + 1) if eta is not None else B0
    fr = np.zeros(4)
    default = {"hopper": [0.5, 0.5], "cartpole_friction": [0.1, 0.1]}.get(model)   # same defaults as od_create
    if fric is None and default is not None:
        fric = default
    if fric is not None:
        fr[:len(fric)] = fric
    q3 = np.zeros((B, nq)); dq1 = np.zeros((B, nq, nq)); dq2 = np.zeros((B, nq, nq)); du = np.zeros((B, nu, nq))
    st = np.zeros(B, dtype=np.int32); it = np.zeros(B, dtype=np.int32)
    rc = lib().hc_contact_step(MODELS[model], B, _p(q1), _p(q2), _p(u), nq, nu, C.c_double(h), _p(fr), C.c_double(r_tol), C.c_double(k_eval), C.c_double(k_grad),
                               100, 25, int(want_eval), int(want_grad), _p(eta), n_eta, _p(q3), _p(dq1), _p(dq2), _p(du), _p(st, C.c_int), _p(it, C.c_int), int(reg))
    assert rc == 0
    return dict(q3=q3, dq1=dq1, dq2=dq2, du=du, status=st, st_eval=st & 15, st_grad=(st >> 4) & 15, it_eval=it & 0xFFFF, it_grad=(it >> 16) & 0xFFFF)


def rollout(model, x1, ubar, h, xbar=None, K=None, kff=None, alpha=None, k_eval=1e-4, fric=None, reg=False):
    """Product rollout template on the host: returns X [R,T,2nq], U [R,T-1,nu], status [R,T-1]."""
    nq, nu = DIMS[model]; nx = 2 * nq
    x1 = np.ascontiguousarray(x1, dtype=np.float64).reshape(-1, nx); R = x1.shape[0]
    ubar = np.ascontiguousarray(ubar, dtype=np.float64)
    per = ubar.ndim == 3
    S = ubar.shape[-2]; T = S + 1
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    xbar, K, kff, alpha = c(xbar), c(K), c(kff), c(alpha)
    fr = np.zeros(4)
    default = {"hopper": [0.5, 0.5], "cartpole_friction": [0.1, 0.1]}.get(model)
    if fric is None and default is not None:
        fric = default
    if fric is not None:
        fr[:len(fric)] = fric
    X = np.zeros((R, T, nx)); U = np.zeros((R, S, nu)); st = np.zeros((R, S), dtype=np.int32)
    rc = lib().hc_contact_rollout(MODELS[model], R, T, _p(x1), _p(ubar), C.c_longlong(S * nu if per else 0), _p(xbar), _p(K), _p(kff), _p(alpha),
                                  C.c_double(h), _p(fr), C.c_double(1e-8), C.c_double(k_eval), _p(X), _p(U), _p(st, C.c_int), int(reg))
    assert rc == 0
    return X, U, st


def riccati(jac, lx, lu, lxx, luu, lux, nq, nu, reg=0.0):
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    jac, lx, lu, lxx, luu, lux = c(jac), c(lx), c(lu), c(lxx), c(luu), c(lux)
    NT, S = jac.shape[0], jac.shape[1]; n = 2 * nq
    K = np.zeros((NT, S, nu, n)); k = np.zeros((NT, S, nu)); dV = np.zeros((NT, 2)); st = np.zeros(NT, dtype=np.int32)
    rc = lib().hc_riccati(NT, S + 1, nq, nu, _p(jac), _p(lx), _p(lu), _p(lxx), _p(luu), _p(lux), C.c_double(reg), _p(K), _p(k), _p(dV), _p(st, C.c_int))
    assert rc == 0
    return K, k, dV, st


def rocket(x, u, h, u_max, proj, want_grad=True, proj_only=False, reg=False):
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 12); B = x.shape[0]
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(B, 3)
    y = np.zeros((B, 12)); dx = np.zeros((B, 12, 12)); du = np.zeros((B, 3, 12)); up = np.zeros((B, 3)); dup = np.zeros((B, 3, 3))
    st = np.zeros(B, dtype=np.int32); it = np.zeros(B, dtype=np.int32)
    rc = lib().hc_rocket(B, _p(x), _p(u), C.c_double(h), C.c_double(u_max), int(proj), int(want_grad), int(proj_only), _p(y), _p(dx), _p(du), _p(up), _p(dup),
                         _p(st, C.c_int), _p(it, C.c_int), int(reg))
    assert rc == 0
    return dict(y=y, dx=dx, du=du, uproj=up, duproj=dup, status=st, iters=it)


def residual_and_direction(model, z, th):
    """Block residual [d|rs|rpsi|rv|rgam|rc0|rc1] at κ=0 and the condensed Newton direction rz⁻¹ r (oracle z ordering)."""
    z = np.ascontiguousarray(z, dtype=np.float64); th = np.ascontiguousarray(th, dtype=np.float64)
    out = np.zeros(len(z)); d = np.zeros(len(z))
    assert lib().hc_contact_residual(MODELS[model], _p(z), _p(th), _p(out), _p(d)) == 0
    return out, d


def sincos(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    s = np.empty_like(x); c = np.empty_like(x)
    assert lib().hc_sincos(x.size, _p(x), _p(s), _p(c)) == 0
    return s, c


def layout(model, lanes):
    """(NR, row pitch PW in doubles, workspace doubles per problem, rows per lane) of the register path for one configuration."""
    out = (C.c_int * 4)()
    assert lib().hc_layout(MODELS[model], int(lanes), out) == 0
    return tuple(out)
