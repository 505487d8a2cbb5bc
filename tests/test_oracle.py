"""CPU tier: the oracle against algorithm-independent acceptance checks and the committed golden vectors.

The reference has no tests and cannot be run here (no Julia; RoboDojo.jl not in tree) ⇒ PARITY UNPINNED: these checks pin the
oracle to the mathematics of the in-tree residuals and to the only known-answer fixture in the reference (src/ls.jl:62-144)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from optimization_dynamics_b200 import workloads as W
from common import CONFIGS

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def theta_of(name, q1, q2, u, h, fric):
    th = [q2 - h * ((q2 - q1) / h), q2, u]
    if fric is not None:
        th.append(np.asarray(fric, dtype=float))
    elif name == "hopper":
        th.append(np.array([0.5, 0.5]))
    th.append(np.array([h]))
    return np.concatenate(th)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_convergence_feasibility_and_ift_identity(name):
    gen, h, ke, kg, fric, _ = CONFIGS[name]
    q1, q2, u = gen(256, h=h, seed=3)
    g = O.step_batch(name, q1, q2, u, h, kg, True, fric=fric, full=True)
    ok = g["status"] == 0
    assert ok.mean() > 0.97
    nq, nu, nz, nth = O.dims(name)
    for i in np.where(ok)[0][::7]:
        th = theta_of(name, q1[i], q2[i], u[i], h, fric)
        r, rz, rth = O.residual(name, g["z"][i], th)
        assert np.abs(r).max() < max(kg, 1e-8)                       # (i) converged: equality rows < r_tol, bilinear rows < κ_tol
        assert g["r_vio"][i] < 1e-8 and g["k_vio"][i] < kg
        ident = np.abs(rz @ g["dz_full"][i] + rth).max()             # (iii) IFT identity
        scale = max(1.0, np.abs(rth).max(), np.abs(g["dz_full"][i]).max())
        assert ident <= 1e-9 * scale, (name, i, ident)
    # (ii) cone feasibility of z*
    z = g["z"][ok]
    if name in ("hopper", "acrobot_impact", "planar_push"):
        sl = {"hopper": slice(4, 12), "acrobot_impact": slice(2, 6), "planar_push": slice(5, 7)}[name]
        assert (z[:, sl] > -1e-12).all()
    if name == "hopper":
        assert (z[:, 12:14] >= np.abs(z[:, 14:16]) - 1e-9).all() and (z[:, 16:18] >= np.abs(z[:, 18:20]) - 1e-9).all()


def test_eval_and_grad_solves_share_one_iterate_sequence():
    """ImplicitDynamics' two simulators differ only in κ_tol (src/dynamics.jl:60-64): the looser solve is a prefix of the tighter."""
    q1, q2, u = W.hopper_batch(512, seed=5)
    e = O.step_batch("hopper", q1, q2, u, 0.05, 1e-4, False)
    g = O.step_batch("hopper", q1, q2, u, 0.05, 1e-3, False)
    assert (g["iters"] <= e["iters"]).all()
    same = g["iters"] == e["iters"]
    assert np.array_equal(e["q3"][same], g["q3"][same])


@pytest.mark.parametrize("name", ["acrobot_nominal", "cartpole_frictionless"])
def test_cone_free_variants_agree_with_fsolve(name):
    """(vi) no cones ⇒ plain Newton; an independent root finder must land on the same q3."""
    from scipy.optimize import fsolve
    gen, h, ke, kg, fric, _ = CONFIGS[name]
    q1, q2, u = gen(24, h=h, seed=11)
    e = O.step_batch(name, q1, q2, u, h, ke, False)
    for i in range(24):
        th = theta_of(name, q1[i], q2[i], u[i], h, None)
        sol = fsolve(lambda z: O.residual(name, z, th)[0], q2[i], xtol=1e-13)
        assert np.abs(O.residual(name, sol, th)[0]).max() < 1e-9
        if e["status"][i] == 0 and np.abs(sol - e["q3"][i]).max() > 1e-7:
            # a different root of the nonlinear system is legitimate only if fsolve wandered; the oracle's must still be a root
            assert np.abs(O.residual(name, e["q3"][i], th)[0]).max() < 1e-8
        elif e["status"][i] == 0:
            assert np.abs(sol - e["q3"][i]).max() < 1e-7


def test_least_squares_known_answer():
    """The reference's only known-answer fixture (src/ls.jl:62-144): f(z) = A x + B u, ±ε one-hot perturbations ⇒ θ = [A B]."""
    A = np.array([[1.0, 1.0], [0.0, 1.0]]); Bm = np.array([0.0, 1.0])
    f = lambda z: A @ z[:2] + Bm * z[2]
    nz, eps = 3, 0.1
    eta = np.zeros((2 * nz, nz))
    for i in range(nz):
        eta[i, i] = eps; eta[i + nz, i] = -eps
    z0 = np.random.default_rng(0).random(nz)
    rc, M = O.least_squares(f(z0), np.array([f(z0 + e) for e in eta]), eta)
    assert rc == 0
    assert np.allclose(M, np.hstack([A, Bm[:, None]]), atol=1e-12)


def test_rocket_projection_feasible_and_identity_inside_cone():
    x, u = W.rocket_batch(256, seed=2)
    p = O.rocket_projection_batch(u, 12.5)
    assert (p["status"] == 0).all()
    up = p["up"]
    assert (np.hypot(up[:, 0], up[:, 1]) <= up[:, 2] + 1e-6).all() and (up[:, 2] <= 12.5 + 1e-6).all()      # examples/rocket.jl:151
    inside = (np.hypot(u[:, 0], u[:, 1]) < u[:, 2] - 0.5) & (u[:, 2] < 12.0) & (u[:, 2] > 0.5)
    assert inside.sum() > 10
    assert np.abs(up[inside] - u[inside]).max() < 2e-3          # κ_tol = 1e-4 smoothing
    assert np.abs(p["dproj"][inside] - np.eye(3)).max() < 2e-2


def test_rocket_dynamics_is_implicit_midpoint():
    x, u = W.rocket_batch(64, seed=4)
    r = O.rocket_batch(x, u, 0.05, 12.5, False, True)
    assert (r["status"] == 0).all()
    for i in range(0, 64, 9):
        th = np.concatenate([x[i], u[i], [0.05]])
        res, rz, rth = O.residual("rocket", r["y"][i], th)
        assert np.abs(res).max() < 1e-8
        assert np.abs(rz @ (-np.linalg.solve(rz, rth)) + rth).max() < 1e-9
        assert np.allclose(r["dx"][i].T, -np.linalg.solve(rz, rth)[:, :12], atol=1e-9)


def test_hopper_closed_form_bias_matches_lagrangian():
    """The oracle's hand-derived M(q), C(q,q̇) (oracle/models.hpp) against sympy differentiation of the Lagrangian."""
    import sympy as sp
    q = sp.symbols("x z t r"); v = sp.symbols("vx vz vt vr")
    mb, Ib, mf, g = 3.0, 0.75, 1.0, 9.81
    foot = [q[0] + q[3] * sp.sin(q[2]), q[1] - q[3] * sp.cos(q[2])]
    vf = [sum(sp.diff(foot[k], q[i]) * v[i] for i in range(4)) for k in range(2)]
    L = 0.5 * mb * (v[0] ** 2 + v[1] ** 2) + 0.5 * Ib * v[2] ** 2 + 0.5 * mf * (vf[0] ** 2 + vf[1] ** 2) - mb * g * q[1] - mf * g * foot[1]
    Cs = sp.Matrix(4, 4, lambda i, j: sp.diff(L, v[i], q[j])) * sp.Matrix(v) - sp.Matrix([sp.diff(L, qi) for qi in q])
    Ms = sp.Matrix(4, 4, lambda i, j: sp.diff(L, v[i], v[j]))
    fC = sp.lambdify(q + v, Cs); fM = sp.lambdify(q + v, Ms * sp.Matrix(v))
    rng = np.random.default_rng(0)
    h = 0.05
    for _ in range(5):
        q0, q1, q2 = rng.normal(size=4), rng.normal(size=4), rng.normal(size=4)
        z = np.concatenate([q2, np.zeros(16)]); th = np.concatenate([q0, q1, [0, 0], [0.5, 0.5], [h]])
        d = O.residual("hopper", z, th)[0][:4]
        qm1, vm1, qm2, vm2 = 0.5 * (q0 + q1), (q1 - q0) / h, 0.5 * (q1 + q2), (q2 - q1) / h
        ref = (0.5 * h * (-fC(*qm1, *vm1)) + fM(*qm1, *vm1) + 0.5 * h * (-fC(*qm2, *vm2)) - fM(*qm2, *vm2)).ravel()
        assert np.allclose(d, ref, atol=1e-9)


@pytest.mark.parametrize("name", list(CONFIGS) + ["rocket", "rocket_proj"])
def test_golden_vectors(name):
    """tests/golden/<model>.npz were produced by tests/golden/make_golden.py from this oracle (the reference cannot be run here);
    they pin the oracle — and through the GPU tests the CUDA path — against silent changes."""
    path = os.path.join(GOLDEN, name + ".npz")
    gold = np.load(path)
    if name.startswith("rocket"):
        r = O.rocket_batch(gold["x"], gold["u"], 0.05, 12.5, name == "rocket_proj", True)
        assert np.abs(r["y"] - gold["y"]).max() < 1e-11 and np.abs(r["dx"] - gold["dx"]).max() < 1e-9 and np.abs(r["du"] - gold["du"]).max() < 1e-9
        return
    gen, h, ke, kg, fric, _ = CONFIGS[name]
    e = O.step_batch(name, gold["q1"], gold["q2"], gold["u"], h, ke, False, fric=fric)
    g = O.step_batch(name, gold["q1"], gold["q2"], gold["u"], h, kg, True, fric=fric)
    assert np.array_equal(e["status"], gold["status_eval"]) and np.array_equal(g["status"], gold["status_grad"])
    ok = (e["status"] == 0) & (g["status"] == 0)
    assert np.abs(e["q3"] - gold["q3"])[ok].max() < 1e-11
    for k in ("dq1", "dq2", "du"):
        assert np.abs(g[k] - gold[k])[ok].max() < 1e-8
