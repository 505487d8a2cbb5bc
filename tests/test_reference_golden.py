"""Parity against the REAL reference (Julia + RoboDojo.jl), when its outputs have been dumped.

julia/dump_reference_golden.jl runs the unmodified reference's f / fx / fu (src/dynamics.jl:81-128) and the rocket wrappers
(src/models/rocket/dynamics.jl:101-269) on tests/golden/reference_inputs/*.csv and writes tests/golden/reference/*.csv.  No Julia
runtime exists in the build image, so those files are absent until a maintainer runs the script once; the tests below then hold
BOTH the oracle (CPU tier) and the CUDA path (GPU tier) to the reference's own numbers — 1e-8 on q3, 1e-6 on the sensitivities.
While the files are absent every test here SKIPS with the reason below and parity stays "unpinned" (DESIGN.md §5)."""
import os

import numpy as np
import pytest

from common import CONFIGS, Q3_TOL, GRAD_TOL

HERE = os.path.dirname(os.path.abspath(__file__))
REF_IN = os.path.join(HERE, "golden", "reference_inputs")
REF_OUT = os.path.join(HERE, "golden", "reference")
WHY = ("PARITY UNPINNED: tests/golden/reference/%s.csv is missing — run `julia --project=<reference checkout> "
       "julia/dump_reference_golden.jl` on a machine with the reference installed and commit its output")
DIMS = {"acrobot_impact": (2, 1), "acrobot_nominal": (2, 1), "cartpole_friction": (2, 1), "cartpole_frictionless": (2, 1),
        "planar_push": (5, 2), "hopper": (4, 2)}


def load(name, inputs_name=None):
    path = os.path.join(REF_OUT, name + ".csv")
    if not os.path.exists(path):
        pytest.skip(WHY % name)
    return np.loadtxt(os.path.join(REF_IN, (inputs_name or name) + ".csv"), delimiter=",", ndmin=2), np.loadtxt(path, delimiter=",", ndmin=2)


def split_contact(name, X, R):
    nq, nu = DIMS[name]
    q1, q2, u = X[:, :nq], X[:, nq:2 * nq], X[:, 2 * nq:]
    o = nq
    q3 = R[:, :nq]
    d1 = R[:, o:o + nq * nq].reshape(-1, nq, nq).transpose(0, 2, 1); o += nq * nq      # column-major blocks → [B, row, col]
    d2 = R[:, o:o + nq * nq].reshape(-1, nq, nq).transpose(0, 2, 1); o += nq * nq
    du = R[:, o:o + nq * nu].reshape(-1, nu, nq).transpose(0, 2, 1); o += nq * nu
    return q1, q2, u, q3, d1, d2, du, R[:, o].astype(int), R[:, o + 1].astype(int)


def report(tag, name, q3, d1, d2, du, rq3, r1, r2, ru):
    eq = np.abs(q3 - rq3).max(1)
    eg = np.maximum.reduce([np.abs(d1 - r1).reshape(len(eq), -1).max(1), np.abs(d2 - r2).reshape(len(eq), -1).max(1),
                            np.abs(du - ru).reshape(len(eq), -1).max(1)])
    print("%s vs REFERENCE, %s: max|q3| %.3e (%.1f %% within %g), max|grad| %.3e (%.1f %% within %g)" % (
        tag, name, eq.max(), 100 * (eq <= Q3_TOL).mean(), Q3_TOL, eg.max(), 100 * (eg <= GRAD_TOL).mean(), GRAD_TOL))
    return eq, eg


@pytest.mark.parametrize("name", list(CONFIGS))
def test_oracle_matches_the_reference(name):
    X, R = load(name)
    from oracle import oracle as O
    gen, h, ke, kg, fric, _ = CONFIGS[name]
    q1, q2, u, rq3, r1, r2, ru, it_e, it_g = split_contact(name, X, R)
    e = O.step_batch(name, q1, q2, u, h, ke, False, fric=fric)
    g = O.step_batch(name, q1, q2, u, h, kg, True, fric=fric)
    eq, eg = report("oracle", name, e["q3"], g["dq1"].transpose(0, 2, 1), g["dq2"].transpose(0, 2, 1), g["du"].transpose(0, 2, 1), rq3, r1, r2, ru)
    if (it_e >= 0).all():
        print("  iteration counts equal: eval %.3f, grad %.3f" % ((it_e == e["iters"]).mean(), (it_g == g["iters"]).mean()))
    ok = (e["status"] == 0) & (g["status"] == 0)
    assert (eq[ok] <= Q3_TOL).mean() >= 0.98 and (eg[ok] <= GRAD_TOL).mean() >= 0.98


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CONFIGS))
def test_cuda_path_matches_the_reference(built, name):
    X, R = load(name)
    import optimization_dynamics_b200 as od
    gen, h, ke, kg, fric, attr = CONFIGS[name]
    model = getattr(od, attr)
    if fric is not None:
        model.friction[:] = fric
    dyn = od.ImplicitDynamics(model, h, r_tol=1e-8, κ_eval_tol=ke, κ_grad_tol=kg)
    q1, q2, u, rq3, r1, r2, ru, _, _ = split_contact(name, X, R)
    q3, d1, d2, du, st = dyn.step_grad_batch(q1, q2, u)
    eq, eg = report("CUDA", name, q3, d1, d2, du, rq3, r1, r2, ru)
    ok = st == 0
    assert (eq[ok] <= Q3_TOL).mean() >= 0.98 and (eg[ok] <= GRAD_TOL).mean() >= 0.98


@pytest.mark.parametrize("name,proj", [("rocket", False), ("rocket_proj", True)])
def test_oracle_rocket_matches_the_reference(name, proj):
    X, R = load(name, "rocket")
    from oracle import oracle as O
    r = O.rocket_batch(X[:, :12], X[:, 12:], 0.05, 12.5, proj, True)
    ey = np.abs(r["y"] - R[:, :12]).max()
    edx = np.abs(r["dx"].reshape(len(X), -1) - R[:, 12:156]).max()
    edu = np.abs(r["du"].reshape(len(X), -1) - R[:, 156:]).max()
    print("oracle vs REFERENCE, %s: max|y| %.3e  max|dx| %.3e  max|du| %.3e" % (name, ey, edx, edu))
    assert ey <= Q3_TOL and edx <= GRAD_TOL and edu <= GRAD_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name,proj", [("rocket", False), ("rocket_proj", True)])
def test_cuda_rocket_matches_the_reference(built, name, proj):
    X, R = load(name, "rocket")
    import optimization_dynamics_b200 as od
    info = od.RocketInfo(od.rocket, 12.5, 0.05)
    y, dx, du, st = info.step_batch(X[:, :12], X[:, 12:], proj=proj)
    B = len(X)
    ey = np.abs(y - R[:, :12]).max()
    edx = np.abs(dx.transpose(0, 2, 1).reshape(B, -1) - R[:, 12:156]).max()
    edu = np.abs(du.transpose(0, 2, 1).reshape(B, -1) - R[:, 156:]).max()
    print("CUDA vs REFERENCE, %s: max|y| %.3e  max|dx| %.3e  max|du| %.3e" % (name, ey, edx, edu))
    assert ey <= Q3_TOL and edx <= GRAD_TOL and edu <= GRAD_TOL


def test_reference_inputs_are_the_golden_inputs():
    """The text inputs handed to the Julia script are bit-for-bit the inputs of tests/golden/*.npz."""
    for name in list(CONFIGS) + ["rocket"]:
        X = np.loadtxt(os.path.join(REF_IN, name + ".csv"), delimiter=",", ndmin=2)
        g = np.load(os.path.join(HERE, "golden", name + ".npz"))
        want = np.concatenate([g["x"], g["u"]], axis=1) if name == "rocket" else np.concatenate([g["q1"], g["q2"], g["u"]], axis=1)
        assert np.array_equal(X, want), name
