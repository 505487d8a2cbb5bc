"""Rocket wrappers — mirror of reference src/models/rocket/dynamics.jl.

    RocketInfo(rocket, u_max, h, r, rz, rθ, r_p, rz_p, rθ_p)     :13-99     (the six generated functions are ignored)
    f_rocket / fx_rocket / fu_rocket                             :101-163
    soc_projection / soc_projection_gradient                     :168-210
    f_rocket_proj / fx_rocket_proj / fu_rocket_proj              :215-269
"""
import ctypes as C

import numpy as np

from . import _lib
from .dynamics import MODEL_IDS, _dp, _ip, _f64, rocket as rocket_model


class RocketInfo:
    def __init__(self, rocket=rocket_model, u_max=12.5, h=0.05, *generated, device=0):
        L = _lib.lib()
        self.h, self.u_max = float(h), float(u_max)
        opts = _lib.od_options()
        _lib.check(L.od_default_options(MODEL_IDS["rocket"], C.byref(opts)))
        p = np.array([self.u_max], dtype=np.float64)
        self._hd = L.od_create(MODEL_IDS["rocket"], self.h, C.byref(opts), _dp(p), 1, device)
        if not self._hd:
            raise RuntimeError("optdyn_b200: " + L.od_last_error().decode())
        self._memo_key, self._memo = None, None

    def __del__(self):
        try:
            if self._hd:
                _lib.lib().od_destroy(self._hd)
                self._hd = None
        except Exception:
            pass

    def step_batch(self, x, u, proj, grad=True, out=None):
        """y[B,12], dx[B,12,12], du[B,12,3], status[B] for B problems.  `out` = (y, dx_cm, du_cm, status): caller-owned C-contiguous
        buffers of shapes (B,12), (B,12,12), (B,3,12), (B,) — e.g. views of pinned host memory, which the copies then run from at
        full PCIe rate; the Jacobian buffers hold the ABI's column-major blocks (the returned views are their transposes)."""
        x = _f64(x, (-1, 12)); B = x.shape[0]; u = _f64(u, (B, 3))
        if out is not None:
            y, dx, du, st = out
            if not grad:
                dx = du = None
        else:
            y = np.empty((B, 12)); st = np.empty(B, dtype=np.int32)
            dx = np.empty((B, 12, 12)) if grad else None
            du = np.empty((B, 3, 12)) if grad else None
        _lib.check(_lib.lib().od_rocket_batch(self._hd, B, _dp(x), _dp(u), int(bool(proj)), _dp(y), None if dx is None else _dp(dx),
                                              None if du is None else _dp(du), _ip(st)))
        if not grad:
            return y, None, None, st
        return y, dx.transpose(0, 2, 1), du.transpose(0, 2, 1), st

    def projection_batch(self, u, grad=True):
        u = _f64(u, (-1, 3)); B = u.shape[0]
        up = np.empty((B, 3)); dp = np.empty((B, 3, 3)) if grad else None; st = np.empty(B, dtype=np.int32)
        _lib.check(_lib.lib().od_rocket_projection_batch(self._hd, B, _dp(u), _dp(up), None if dp is None else _dp(dp), _ip(st)))
        return up, (None if dp is None else dp.transpose(0, 2, 1)), st

    def launch_count(self):
        return int(_lib.lib().od_launch_count(self._hd))

    def _grad(self, x, u, proj):
        x = np.asarray(x, dtype=np.float64); u = np.asarray(u, dtype=np.float64)
        key = (x.tobytes(), u.tobytes(), bool(proj))
        if key != self._memo_key:
            y, dx, du, st = self.step_batch(x[None], u[None], proj, grad=True)
            self._memo_key, self._memo = key, (y[0], dx[0], du[0], int(st[0]))
        return self._memo


def f_rocket(d, info, x, u, w):
    y, _, _, _ = info.step_batch(np.asarray(x)[None], np.asarray(u)[None], False, grad=False)
    d[...] = y[0]
    return d


def fx_rocket(dx, info, x, u, w):
    dx[...] = info._grad(x, u, False)[1]
    return dx


def fu_rocket(du, info, x, u, w):
    du[...] = info._grad(x, u, False)[2]
    return du


def soc_projection(x, info):
    return info.projection_batch(np.asarray(x)[None], grad=False)[0][0]


def soc_projection_gradient(x, info):
    return info.projection_batch(np.asarray(x)[None], grad=True)[1][0]


def f_rocket_proj(d, info, x, u, w):
    y, _, _, _ = info.step_batch(np.asarray(x)[None], np.asarray(u)[None], True, grad=False)
    d[...] = y[0]
    return d


def fx_rocket_proj(dx, info, x, u, w):
    dx[...] = info._grad(x, u, True)[1]
    return dx


def fu_rocket_proj(du, info, x, u, w):
    du[...] = info._grad(x, u, True)[2]
    return du
