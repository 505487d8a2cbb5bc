"""CPU tier: (1) the product's solver templates, compiled for the host by tests/host_check.cu, against the oracle — the same
template code the CUDA kernels instantiate, so the reduced-system algebra is checked without a GPU; (2) the C ABI loads and
exports every symbol include/optdyn_b200.h declares and fails loudly without a CUDA device; (3) host-side helpers."""
import ctypes as C
import os
import sys
import re

import numpy as np
import pytest

import hostcheck as H
from oracle import oracle as O
from optimization_dynamics_b200 import workloads as W
from common import CONFIGS, compare, oracle_pair

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


REG_MODELS = ["hopper", "acrobot_impact", "cartpole_friction", "planar_push"]     # models whose latency configuration uses csrc/group_gj.cuh


@pytest.mark.parametrize("name,reg", [(n, False) for n in CONFIGS] + [(n, True) for n in REG_MODELS])
def test_solver_templates_match_oracle(name, reg):
    """reg=False: shared-memory LU path (one thread per problem); reg=True: the register-resident Gauss–Jordan algebra of the
    cooperative-lane path, run with one lane (shuffles are identities on the host)."""
    gen, h, ke, kg, fric, _ = CONFIGS[name]
    B = 1024 if name != "planar_push" else 512
    q1, q2, u = gen(B, h=h, seed=1)
    e, g = oracle_pair(O, name, q1, q2, u)
    r = H.step(name, q1, q2, u, h, ke, kg, fric=fric, reg=reg)
    tr = lambda a: a.transpose(0, 2, 1)
    eq, eg = compare(name, e, g, r["q3"], tr(r["dq1"]), tr(r["dq2"]), tr(r["du"]), r["st_eval"], r["st_grad"],
                     grad_outlier_fraction=0.0)
    # same iterate sequence ⇒ same iteration counts on every comparable sample
    ok = (e["status"] == 0) & (r["st_eval"] == 0) & (e["margin"] > 1e-6)
    assert (e["iters"][ok] != r["it_eval"][ok]).mean() <= 0.002


def test_eval_only_and_grad_only_paths_agree_with_combined():
    q1, q2, u = W.hopper_batch(128, seed=9)
    both = H.step("hopper", q1, q2, u, 0.05)
    ev = H.step("hopper", q1, q2, u, 0.05, want_grad=False)
    gr = H.step("hopper", q1, q2, u, 0.05, want_eval=False)
    assert np.array_equal(both["q3"], ev["q3"])
    assert np.array_equal(both["dq1"], gr["dq1"]) and np.array_equal(both["du"], gr["du"])


def test_residual_blocks_and_newton_direction_match_dense_oracle():
    rng = np.random.default_rng(1)
    for _ in range(10):
        z = np.concatenate([[0.1, 0.6, 0.2, 0.5] + 0.1 * rng.normal(size=4), rng.uniform(0.5, 1.5, 8), rng.uniform(1.0, 1.5, 2), rng.uniform(-0.3, 0.3, 2),
                            rng.uniform(1.0, 1.5, 2), rng.uniform(-0.3, 0.3, 2)])
        th = np.concatenate([[0.08, 0.61, 0.18, 0.52], [0.09, 0.6, 0.19, 0.51], rng.normal(size=2), [0.5, 0.6], [0.05]])
        r, rz, rth = O.residual("hopper", z, th)
        blocks, d = H.residual_and_direction("hopper", z, th)
        perm = list(range(16)) + [16, 18, 17, 19]       # oracle interleaves (cone row 0, cone row 1) per cone
        assert np.abs(blocks - r[perm]).max() < 1e-13
        assert np.abs(d - np.linalg.solve(rz, r)).max() < 1e-11


@pytest.mark.parametrize("proj,reg", [(False, False), (True, False), (False, True), (True, True)])
def test_rocket_templates_match_oracle(proj, reg):
    """reg=True: the cooperative-lane solver (dense_ipg.cuh, register Gauss–Jordan) run with one lane."""
    x, u = W.rocket_batch(512, seed=1)
    o = O.rocket_batch(x, u, 0.05, 12.5, proj, True)
    r = H.rocket(x, u, 0.05, 12.5, proj, reg=reg)
    ok = (o["status"] == 0) & (r["status"] == 0) & (o["margin"] > 1e-6)
    assert ok.mean() > 0.9
    assert np.abs(o["y"] - r["y"])[ok].max() < 1e-8
    assert np.abs(o["dx"] - r["dx"])[ok].max() < 1e-6 and np.abs(o["du"] - r["du"])[ok].max() < 1e-6


@pytest.mark.parametrize("name", ["hopper", "cartpole_friction"])
def test_bundle_closed_form_matches_reference_newton_fit(name):
    """(N+1)·B perturbed steps + normal equations == the oracle's restatement of gradient! + LeastSquares.update!."""
    gen, h, ke, kg, fric, _ = CONFIGS[name]
    nq, nu = H.DIMS[name]
    ncol = 2 * nq + nu
    q1, q2, u = gen(12, h=h, seed=3)
    eta = W.bundle_perturbations(ncol, N=64, seed=1)
    ob = O.bundle_batch(name, eta, q1, q2, u, h, ke, fric=fric)
    r = H.step(name, q1, q2, u, h, ke, kg, fric=fric, want_grad=False, eta=eta)
    f = r["q3"].reshape(12, 65, nq)
    M = np.einsum("bki,kj->bij", f[:, 1:] - f[:, :1], eta) @ np.linalg.inv(eta.T @ eta)
    assert (ob["status"] == 0).all()
    assert np.abs(M - ob["dz"].transpose(0, 2, 1)).max() < 1e-6


def test_c_abi_exports_every_declared_symbol(built):
    from optimization_dynamics_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "optdyn_b200.h")).read()
    declared = set(re.findall(r"\b(od_[a-z_0-9]+)\s*\(", hdr)) - {"od_handle", "od_options", "od_model"}
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    L = C.CDLL(_lib.SO_PATH)
    for s in declared:
        assert hasattr(L, s), s
    L2 = _lib.lib()
    nq, nu, nz, nth = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    for mid, want in ((5, (4, 2, 20, 13)), (4, (5, 2, 35, 13)), (2, (2, 1, 10, 8)), (0, (2, 1, 6, 6)), (6, (12, 3, 12, 16))):
        assert L2.od_model_dims(mid, C.byref(nq), C.byref(nu), C.byref(nz), C.byref(nth)) == 0
        assert (nq.value, nu.value, nz.value, nth.value) == want
    assert L2.od_model_dims(99, None, None, None, None) != 0
    o = _lib.od_options()
    assert L2.od_default_options(5, C.byref(o)) == 0 and o.r_tol == 1e-8 and o.max_ls == 25 and o.kappa_grad_tol == 1e-3


def test_no_cpu_fallback(built):
    """Without a CUDA device construction must raise — the product never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import optimization_dynamics_b200 as od
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        od.ImplicitDynamics(od.hopper, 0.05, κ_eval_tol=1e-4, κ_grad_tol=1e-3)
    with pytest.raises(RuntimeError):
        od.RocketInfo(od.rocket, 12.5, 0.05)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "optimization_dynamics_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.replace("against the oracle", "").replace("the oracle", "").lower() or f in ("contact_ip.cuh",), (dp, f)


def test_shard_ranges_cover_the_batch():
    from optimization_dynamics_b200.device import shard_range, shard_sizes
    for B in (0, 1, 7, 4096, 4097):
        for w in (1, 2, 3, 8):
            spans = [shard_range(B, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(shard_sizes(B, w)) - min(shard_sizes(B, w)) <= 1


def test_state_to_configuration():
    from optimization_dynamics_b200 import state_to_configuration
    x = [np.array([1.0, 2.0, 3.0, 4.0]), np.array([3.0, 4.0, 5.0, 6.0])]
    q = state_to_configuration(x)
    assert [list(v) for v in q] == [[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]]


def _rollout_agreement(name, X, U, st, Xo, Uo, sto, h, ke, fric):
    """One-step consistency along the product's own trajectory (1e-8), and whole-trajectory agreement with the oracle rollout."""
    nq = X.shape[2] // 2
    assert (st == 0).mean() > 0.99 and (sto == 0).mean() > 0.99
    R, T, _ = X.shape
    assert np.array_equal(X[:, 1:, :nq], X[:, :-1, nq:])                  # d[1:nq] = q2  (src/dynamics.jl:90)
    Xf, Uf = X[:, :-1].reshape(-1, 2 * nq), U.reshape(R * (T - 1), -1)
    o = O.step_batch(name, Xf[:, :nq], Xf[:, nq:], Uf, h, ke, False, fric=fric)
    ok = (o["status"] == 0) & (st.reshape(-1) == 0) & (o["margin"] > 1e-6) & (o["iters"] <= 30)
    assert ok.mean() > 0.95
    bad = ok & ~(np.abs(o["q3"] - X[:, 1:, nq:].reshape(-1, nq)).max(1) <= 1e-8)
    assert not (bad & ~(o["q_uncertainty"] > 1e-7)).any() and bad.mean() <= 0.005
    return float(np.abs(X - Xo).max()), float(np.abs(U - Uo).max())


@pytest.mark.parametrize("reg", [False, True])
def test_rollout_template_matches_oracle_rollout(reg):
    """contact_rollout_one (the code of contact_rollout_kernel) on the host: iLQR.rollout + closed-loop forward pass, hopper."""
    h = 0.05
    x1, ubar, K, k, alpha = W.hopper_rollout_inputs(8, T=21, h=h, seed=3)
    alpha[-1] = 1.0e-5                                                                 # α_min of examples/hopper.jl:277
    Xn, _, _ = O.rollout_batch("hopper", x1[None], ubar, h, 1e-4)                      # nominal open-loop rollout = x̄
    xbar = Xn[0]
    x1s = np.tile(x1, (8, 1))
    X, U, st = H.rollout("hopper", x1s, ubar, h, xbar=xbar, K=K, kff=k, alpha=alpha, reg=reg)
    Xo, Uo, sto = O.rollout_batch("hopper", x1s, ubar, h, 1e-4, xbar=xbar, K=K, k=k, alpha=alpha)
    ex, eu = _rollout_agreement("hopper", X, U, st, Xo, Uo, sto, h, 1e-4, None)
    assert ex < 1e-6 and eu < 1e-6, (ex, eu)
    # α → 0 reproduces the nominal trajectory
    assert np.abs(X[-1] - xbar).max() < 1e-4
    # open loop with the shared controls: every rollout identical to the nominal one
    X0, U0, st0 = H.rollout("hopper", x1s[:2], ubar, h, reg=reg)
    assert np.abs(X0[0] - xbar).max() < 1e-8 and np.array_equal(X0[0], X0[1]) and np.array_equal(U0[0], ubar)


def _riccati_case(NT=6, T=21, seed=2):
    """Jacobians along closed-loop hopper trajectories (oracle) + the tracking-cost expansion."""
    h = 0.05
    x1, ubar, K, k, alpha = W.hopper_rollout_inputs(NT, T=T, h=h, seed=seed)
    X, U, st = O.rollout_batch("hopper", np.tile(x1, (NT, 1)), ubar, h, 1e-4, k=k, alpha=alpha)
    Xf, Uf = X[:, :-1].reshape(-1, 8), U.reshape(-1, 2)
    g = O.step_batch("hopper", Xf[:, :4], Xf[:, 4:], Uf, h, 1e-3, True, diagnostics=False)
    jac = np.concatenate([g["q3"], g["dq1"].reshape(-1, 16), g["dq2"].reshape(-1, 16), g["du"].reshape(-1, 8)], axis=1).reshape(NT, T - 1, 44)
    x_goal = np.concatenate([[1.0, 0.55, 0.0, 0.5]] * 2)
    lx, lu, lxx, luu, lux = W.quadratic_cost_expansion(X, U, x_goal, 1.0e-1, 1.0e-1, 10.0, seed=seed)
    return jac, lx, lu, lxx, luu, lux


def test_riccati_template_matches_oracle_backward_pass():
    """riccati_one (the code of riccati_kernel) on the host against the numpy restatement of the iLQR backward pass."""
    jac, lx, lu, lxx, luu, lux = _riccati_case()
    for cross, reg in ((lux, 0.0), (None, 1.0e-3)):
        K, k, dV, st = H.riccati(jac, lx, lu, lxx, luu, cross, 4, 2, reg)
        for a in range(jac.shape[0]):
            Ko, ko, dVo, sto = O.backward_pass(jac[a], lx[a], lu[a], lxx[a], luu[a], None if cross is None else cross[a], 4, 2, reg)
            assert sto == 0 and st[a] == 0
            assert np.abs(K[a] - Ko).max() <= 1e-9 * max(1.0, np.abs(Ko).max()) and np.abs(k[a] - ko).max() <= 1e-9 * max(1.0, np.abs(ko).max())
            assert np.abs(dV[a] - dVo).max() <= 1e-9 * max(1.0, np.abs(dVo).max())
    # an indefinite Quu is reported, not silently inverted
    luu_bad = luu.copy(); luu_bad[0, 5] = -1.0e3 * np.eye(2)
    _, _, _, st = H.riccati(jac, lx, lu, lxx, luu_bad, None, 4, 2, 0.0)
    assert st[0] == 1 and (st[1:] == 0).all()
    assert O.backward_pass(jac[0], lx[0], lu[0], lxx[0], luu_bad[0], None, 4, 2)[3] == 1


def test_fast_math_accuracy():
    """csrc/fastmath.cuh od_sincos (Cody-Waite + fdlibm kernels) against libm: <= 2.5e-16 absolute over the ranges any model angle
    can take, exact at 0, NaN in -> NaN out (the solver's non-finite check relies on that)."""
    rng = np.random.default_rng(0)
    for span in (1.0, 10.0, 1e3, 1e6):
        x = rng.uniform(-span, span, 200000)
        s, c = H.sincos(x)
        assert np.abs(s - np.sin(x)).max() < 2.5e-16 and np.abs(c - np.cos(x)).max() < 2.5e-16, span
    x = np.array([0.0, np.pi / 4, -np.pi / 4, np.pi / 2, np.pi, -3 * np.pi / 4, 2 * np.pi, np.nan, np.inf])
    s, c = H.sincos(x)
    assert s[0] == 0.0 and c[0] == 1.0
    assert np.abs(s[:7] - np.sin(x[:7])).max() < 2.5e-16 and np.abs(c[:7] - np.cos(x[:7])).max() < 2.5e-16
    assert np.isnan(s[7]) and np.isnan(c[7]) and np.isnan(s[8]) and np.isnan(c[8])


def test_generated_model_code_is_up_to_date(tmp_path):
    """csrc/gen/*.cuh are committed generator output (tools/codegen/gen_models.py, the successor of the reference's
    deps/build.jl): regenerating must reproduce them byte for byte — including the tabulated tenth roots of the planar push."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_models", os.path.join(root, "tools", "codegen", "gen_models.py"))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    G.OUT = str(tmp_path)
    argv = sys.argv
    try:
        sys.argv = ["gen_models.py"]
        G.main()
    finally:
        sys.argv = argv
    gen_dir = os.path.join(root, "optimization_dynamics_b200", "csrc", "gen")
    names = sorted(f for f in os.listdir(gen_dir) if f.endswith(".cuh"))
    assert names == sorted(os.listdir(str(tmp_path)))
    for f in names:
        assert open(os.path.join(gen_dir, f)).read() == open(os.path.join(str(tmp_path), f)).read(), f
    pp = open(os.path.join(gen_dir, "model_planar_push.cuh")).read()
    assert pp.count("pow(") == 1 and "ipow<" in pp            # one tenth root per evaluation point, integer powers by squaring


@pytest.mark.parametrize("name,lanes,B", [("hopper", 8, 22), ("hopper", 4, 13), ("hopper", 16, 5), ("cartpole_friction", 4, 19),
                                          ("acrobot_impact", 4, 11), ("planar_push", 8, 6)])
def test_cooperative_lanes_on_the_host_match_one_lane(name, lanes, B):
    """The multi-lane register path (rows spread over the lanes of a group, pivot search by shuffles, pivot rows through the
    shared-memory mirror, warp votes of the lock-stepped state machine) run on the CPU by a team of 32 host threads per emulated
    warp (HostLaneTeam, csrc/group_gj.cuh): identical results to the one-lane run of the same templates — partial last warps,
    problems of one warp finishing at different iterations and the rank-revealing IFT of the planar push included — and parity
    with the oracle."""
    gen, h, ke, kg, fric, _ = CONFIGS[name]
    q1, q2, u = gen(B, h=h, seed=11)
    one = H.step(name, q1, q2, u, h, ke, kg, fric=fric, reg=1)
    many = H.step(name, q1, q2, u, h, ke, kg, fric=fric, reg=lanes)
    assert np.array_equal(one["status"], many["status"]) and np.array_equal(one["it_eval"], many["it_eval"])
    ok = one["status"] == 0
    assert ok.mean() > 0.8
    for k in ("q3", "dq1", "dq2", "du"):
        assert np.array_equal(one[k][ok], many[k][ok]), k
    e, g = oracle_pair(O, name, q1, q2, u)
    good = ok & (e["status"] == 0) & (g["status"] == 0) & (e["margin"] > 1e-6) & (g["margin"] > 1e-6) & (e["iters"] <= 30)
    assert np.abs(many["q3"] - e["q3"])[good].max() < 1e-8


@pytest.mark.parametrize("name,lanes,B", [("hopper", 8, 22), ("hopper", 4, 13), ("hopper", 16, 5), ("cartpole_friction", 4, 19),
                                          ("acrobot_impact", 4, 11), ("planar_push", 16, 4), ("planar_push", 8, 6)])
@pytest.mark.parametrize("flags,tag", [(["-DOD_EXTRACT_SMEM=0"], "_shuffle_gather")] + (
    [(["-DOD_EXTRACT_SMEM=1", "-DOD_INPLACE_Z=1"], "_v2z")] if os.environ.get("OD_TEST_ALL_VARIANTS") else []))
def test_prepared_mirror_variant_matches_shipped_path(name, lanes, B, flags, tag):
    """The shipped build (-DOD_EXTRACT_SMEM=1: inverse pivots and solutions through the shared-memory mirror, zero multiplier in pivot
    rows) against the round-1 path it replaced (-DOD_EXTRACT_SMEM=0: solution gathered by shuffles) and against the in-place iterate
    update on top of it (-DOD_INPLACE_Z=1, measured, not shipped): bit for bit, with one lane and with cooperative lanes.  (The in-place variant differs by rounding
    only on a retried line-search step, which these batches do not contain except, rarely, on the planar push.)"""
    gen, h, ke, kg, fric, _ = CONFIGS[name]
    q1, q2, u = gen(B, h=h, seed=12)
    base = H.step(name, q1, q2, u, h, ke, kg, fric=fric, reg=lanes)
    with H.use_variant(flags, tag):
        var1 = H.step(name, q1, q2, u, h, ke, kg, fric=fric, reg=1)
        var = H.step(name, q1, q2, u, h, ke, kg, fric=fric, reg=lanes)
    exact = not ("-DOD_INPLACE_Z=1" in flags and name == "planar_push")
    for other in (var1, var):
        assert np.array_equal(base["status"], other["status"]) and np.array_equal(base["it_eval"], other["it_eval"])
        for k in ("q3", "dq1", "dq2", "du"):
            if exact:
                assert np.array_equal(base[k], other[k], equal_nan=True), k
            else:
                assert np.allclose(base[k], other[k], rtol=0, atol=1e-9, equal_nan=True), k


@pytest.mark.parametrize("name,lanes,B", [("hopper", 8, 22), ("hopper", 4, 13), ("cartpole_friction", 4, 19), ("acrobot_impact", 4, 11), ("planar_push", 8, 6)])
@pytest.mark.skipif(not os.environ.get("OD_TEST_ALL_VARIANTS"), reason="measured-and-rejected A/B variant: set OD_TEST_ALL_VARIANTS=1 to build and check it (≈ 80 s of nvcc)")
def test_shared_memory_resident_elimination_matches_register_path(name, lanes, B):
    """GroupGJS (matrix resident in shared memory, rolled step loop — an A/B option, ContactIP::LASM; measured slower, off) against the
    register-resident GroupGJ: same arithmetic in the same order, so bit-identical, with one lane and with cooperative lanes.
    -DOD_LA_SMEM_MIN_NR=1 builds every model on GroupGJS; the shipped build (99) none."""
    gen, h, ke, kg, fric, _ = CONFIGS[name]
    q1, q2, u = gen(B, h=h, seed=13)
    base = H.step(name, q1, q2, u, h, ke, kg, fric=fric, reg=lanes)          # the shipped build: register-resident elimination
    with H.use_variant(["-DOD_EXTRACT_SMEM=1", "-DOD_LA_SMEM_MIN_NR=1"], "_lasm"):
        var1 = H.step(name, q1, q2, u, h, ke, kg, fric=fric, reg=1)
        var = H.step(name, q1, q2, u, h, ke, kg, fric=fric, reg=lanes)
    assert (base["status"] == 0).mean() > 0.7
    for other in (var1, var):
        assert np.array_equal(base["status"], other["status"]) and np.array_equal(base["it_eval"], other["it_eval"])
        for k in ("q3", "dq1", "dq2", "du"):
            assert np.array_equal(base[k], other[k], equal_nan=True), k


@pytest.mark.parametrize("variant", [None, "-DOD_EXTRACT_SMEM=0"])       # (OD_INPLACE_Z only touches the contact state machine)
def test_cooperative_lanes_rocket_and_rollouts_on_the_host(variant):
    """rocket_kernel_g (dense 12×12 dynamics + 10×10 cone projection + chain rule) and the closed-loop rollout template with
    cooperating lanes on emulated warps: identical to their one-lane runs, for the shipped path and the prepared variant."""
    import contextlib
    ctx = H.use_variant([variant], "_shuffle_gather") if variant else contextlib.nullcontext()
    x, u = W.rocket_batch(9, seed=4)
    q1, q2, _ = W.hopper_batch(6, h=0.05, seed=5)
    x1 = np.concatenate([q1, q2], axis=1)
    ubar = np.tile(np.array([0.0, 9.81 * 3.0 * 0.5 * 0.05]), (6, 4, 1)) + 0.1 * np.random.default_rng(6).standard_normal((6, 4, 2))
    with ctx:
        for proj in (False, True):
            one = H.rocket(x, u, 0.05, 12.5, proj, reg=1)
            for lanes in (8, 4):
                many = H.rocket(x, u, 0.05, 12.5, proj, reg=lanes)
                assert np.array_equal(one["status"], many["status"]) and np.array_equal(one["iters"], many["iters"])
                for k in ("y", "dx", "du", "uproj", "duproj"):
                    assert np.array_equal(one[k], many[k]), (proj, lanes, k)
        X1, U1, s1 = H.rollout("hopper", x1, ubar, 0.05, reg=1)
        X8, U8, s8 = H.rollout("hopper", x1, ubar, 0.05, reg=8)
        assert np.array_equal(s1, s8) and np.array_equal(X1, X8) and np.array_equal(U1, U8)


def test_pathological_inputs_agree_across_lanes_and_variants():
    """NaN / inf / absurd inputs and problems that run into the iteration cap: same status, iteration counts and (NaN-aware)
    outputs for one lane, eight emulated lanes, and the prepared variant — failures are reported, never hidden or hung."""
    q1, q2, u = W.hopper_batch(8, h=0.05, seed=2)
    q1[1, 0] = np.nan; u[2, :] = 1e12; q2[3, 1] = -50.0; q1[4, :] = q2[4, :]; u[5, :] = np.inf; q2[6, 3] = 1e-30
    runs = [H.step("hopper", q1, q2, u, 0.05, reg=1), H.step("hopper", q1, q2, u, 0.05, reg=8)]
    with H.use_variant(["-DOD_EXTRACT_SMEM=0"], "_shuffle_gather"):
        runs += [H.step("hopper", q1, q2, u, 0.05, reg=1), H.step("hopper", q1, q2, u, 0.05, reg=8)]
    if os.environ.get("OD_TEST_ALL_VARIANTS"):
        with H.use_variant(["-DOD_EXTRACT_SMEM=1", "-DOD_INPLACE_Z=1"], "_v2z"):
            runs += [H.step("hopper", q1, q2, u, 0.05, reg=1), H.step("hopper", q1, q2, u, 0.05, reg=8)]
    base = runs[0]
    assert base["st_eval"][1] == 2 and base["st_eval"][5] == 2            # non-finite inputs: ST_FAIL
    assert set(base["st_eval"][[2, 3]]) <= {1, 2}                          # absurd inputs: iteration cap or failure, reported
    assert base["status"][0] == 0 and base["status"][4] == 0
    conv = base["status"] == 0
    for n, other in enumerate(runs[1:]):
        assert np.array_equal(base["status"], other["status"]) and np.array_equal(base["it_eval"], other["it_eval"])
        for k in ("q3", "dq1", "dq2", "du"):
            assert np.array_equal(base[k][conv], other[k][conv], equal_nan=True), k
            if n < 3:      # shipped path and the mirror variant: bit-identical even on the wandering problems
                assert np.array_equal(base[k], other[k], equal_nan=True), k
            else:          # in-place iterate: the retried line-search steps of the two capped problems differ by rounding
                assert np.allclose(base[k], other[k], rtol=0, atol=1e-6, equal_nan=True), k


def test_user_model_front_end(tmp_path):
    """a model written as a specification file (tools/codegen/examples/particle_spec.py: point mass, ground
    contact, Coulomb friction) goes through `gen_models.py --spec` and the generated header + traits struct run under the
    product's solver templates on the host tier: free flight is the explicit variational step, a particle pressed on the ground
    stays on it and sticks inside the friction cone / slides outside, and the IFT sensitivities match finite differences."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path)
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "codegen", "gen_models.py"), "--spec",
                           os.path.join(root, "tools", "codegen", "examples", "particle_spec.py"), "--out", out],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    hdr = os.path.join(out, "model_particle.cuh")
    so = os.path.join(out, "libusermodel.so")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
                           "-DUSER_MODEL_HEADER=\"%s\"" % hdr, "-DUSER_MODEL=ParticleModel", "-o", so,
                           os.path.join(root, "tests", "user_model_check.cu")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)

    def step(q1, q2, u, mu=0.5, h=0.05, reg=0, ke=1e-6, kg=1e-6):
        q1, q2, u = (np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 2) for a in (q1, q2, u))
        B = q1.shape[0]
        q3 = np.zeros((B, 2)); d1 = np.zeros((B, 2, 2)); d2 = np.zeros((B, 2, 2)); du = np.zeros((B, 2, 2)); st = np.zeros(B, dtype=np.int32)
        fr = np.array([mu, 0, 0, 0.0])
        p = lambda a: a.ctypes.data_as(dp)
        assert L.um_step(B, p(q1), p(q2), p(u), C.c_double(h), p(fr), C.c_double(ke), C.c_double(kg), p(q3), p(d1), p(d2), p(du),
                         st.ctypes.data_as(C.POINTER(C.c_int)), reg) == 0
        return q3, d1.transpose(0, 2, 1), d2.transpose(0, 2, 1), du.transpose(0, 2, 1), st          # blocks are column-major
    m, g, h = 1.5, 9.81, 0.05
    for reg in (0, 1):
        # free flight: no contact force, q3 = 2 q2 − q1 + h² (u/m − g e_z)
        q1 = np.array([[0.0, 1.0]]); q2 = np.array([[0.01, 1.02]]); u = np.array([[0.3, 0.2]])
        q3, d1, d2, du, st = step(q1, q2, u, reg=reg)
        assert st[0] == 0 and np.allclose(q3[0], 2 * q2[0] - q1[0] + h * h * (u[0] / m - np.array([0.0, g])), atol=1e-6)
        # resting on the ground, small tangential push inside the friction cone (|u_x| < μ m g): sticks; a large one slides
        q1 = np.array([[0.0, 0.0]]); q2 = np.array([[0.0, 0.0]])
        q3, *_ , st = step(q1, q2, np.array([[0.2 * m * g, 0.0]]), reg=reg)
        assert st[0] == 0 and abs(q3[0, 1]) < 1e-5 and abs(q3[0, 0]) < 1e-5
        q3, *_ , st = step(q1, q2, np.array([[2.0 * m * g, 0.0]]), reg=reg)
        assert st[0] == 0 and abs(q3[0, 1]) < 1e-5 and abs(q3[0, 0] - h * h * (2.0 - 0.5) * g) < 1e-5
        # IFT sensitivities against central differences of the step itself (sliding contact: every block is exercised)
        q1 = np.array([[0.0, 0.02]]); q2 = np.array([[0.01, 0.005]]); u = np.array([[1.0 * m * g, -2.0]])
        q3, d1, d2, du, st = step(q1, q2, u, reg=reg, ke=1e-9, kg=1e-9)
        assert st[0] == 0
        eps = 1e-6
        for blk, arr in ((d1, q1), (d2, q2), (du, u)):
            for j in range(2):
                ap = arr.copy(); am = arr.copy(); ap[0, j] += eps; am[0, j] -= eps
                args_p = [ap if a is arr else a for a in (q1, q2, u)]; args_m = [am if a is arr else a for a in (q1, q2, u)]
                fd = (step(*args_p, reg=reg, ke=1e-9, kg=1e-9)[0][0] - step(*args_m, reg=reg, ke=1e-9, kg=1e-9)[0][0]) / (2 * eps)
                assert np.allclose(blk[0][:, j], fd, atol=2e-4), (reg, j, blk[0][:, j], fd)


def test_bundle_prepare_inverts_the_normal_matrix_and_reports_singularity(built):
    """od_bundle_prepare (host side of the gradient bundle, no GPU needed): (Σ η ηᵀ)⁻¹ for the shared perturbations; a coordinate that
    no perturbation touches — where the reference's LU would silently divide by zero (src/ls.jl:52) — is reported as an error."""
    from optimization_dynamics_b200 import _lib, workloads as W
    L = _lib.lib()
    ncol, N = 10, 64
    eta = W.bundle_perturbations(ncol, N=N, eps=1e-4, seed=3)
    hinv = np.zeros((ncol, ncol))
    assert L.od_bundle_prepare(ncol, N, eta.ctypes.data_as(_lib.c_double_p), hinv.ctypes.data_as(_lib.c_double_p)) == 0
    H = eta.T @ eta
    assert np.allclose(hinv @ H, np.eye(ncol), atol=1e-9)
    eta[:, 4] = 0.0                                                     # coordinate 4 never perturbed
    assert L.od_bundle_prepare(ncol, N, eta.ctypes.data_as(_lib.c_double_p), hinv.ctypes.data_as(_lib.c_double_p)) != 0
    assert b"singular" in L.od_last_error()
    assert L.od_bundle_prepare(17, N, eta.ctypes.data_as(_lib.c_double_p), hinv.ctypes.data_as(_lib.c_double_p)) != 0     # 2nq+nu > 16 unsupported


@pytest.mark.parametrize("name,lanes", [("hopper", 4), ("hopper", 8), ("hopper", 16), ("cartpole_friction", 4), ("acrobot_impact", 4),
                                         ("planar_push", 8), ("planar_push", 16), ("planar_push", 32)])
def test_row_moves_of_the_register_path_are_bank_conflict_free(name, lanes):
    """The lanes of a group move their own matrix rows between registers and the shared-memory mirror as 16-byte words
    (group_gj.cuh: factor_v2, fetch_rows); one wavefront serves a quarter-warp (8 lanes x 16 bytes = the 32 banks).  Replays those
    accesses for every shipped (model, lanes) configuration against the bank map: with the row pitch / workspace size rules of
    ContactIP::PW each quarter-warp must need exactly ONE wavefront.  (Planar push used to sit at a 256-byte pitch — every row on
    the same four banks, 59 % of the kernel's shared-memory wavefronts were replays; hopper 4 lanes at 2-way conflicts.)"""
    NR, PW, WS, RPL = H.layout(name, lanes)
    assert PW % 2 == 0 and WS % 2 == 0
    for s in range(RPL):
        for j in (0, 2, PW - 2):
            for q0 in range(0, 32, 8):                       # a quarter-warp
                per_bank = {}
                for lane in range(q0, q0 + 8):
                    slot, g = divmod(lane, lanes)
                    r = s * lanes + g
                    if r >= NR:
                        continue                             # padding rows are not moved
                    word = 2 * (slot * WS + r * PW + j)      # 4-byte words; a lane touches 4 consecutive banks
                    for w in range(4):
                        per_bank.setdefault((word + w) % 32, set()).add((word + w) // 32)
                assert all(len(v) == 1 for v in per_bank.values()), (name, lanes, PW, WS, s, j, q0)


def test_gradient_bundle_resampling_covers_every_coordinate():
    """GradientBundle(model; N, ϵ) draws N one-hot perturbations at random once (reference src/gradient_bundle.jl:49-54); with N not much
    larger than nz some coordinate is often never drawn and the reference's fit is singular.  unsampled() names those coordinates,
    resample(cover=True) redraws with every coordinate taken at least once."""
    import optimization_dynamics_b200 as od
    rng = np.random.default_rng(0)
    gb = od.GradientBundle(od.cartpole_friction, N=6, ϵ=1e-4, rng=rng)          # nz = 5: six random draws rarely cover all five
    assert gb.eta.shape == (6, 5) and ((gb.eta != 0).sum(1) == 1).all()
    gb.eta[:, 3] = 0.0
    assert 3 in gb.unsampled()
    gb.resample(rng=rng)
    assert gb.unsampled().size == 0 and ((gb.eta != 0).sum(1) == 1).all() and np.abs(gb.eta).max() < 1e-3
    few = od.GradientBundle(od.cartpole_friction, N=3, ϵ=1e-4, rng=rng).resample(rng=rng)     # N < nz: cannot cover, law unchanged
    assert few.eta.shape == (3, 5) and few.unsampled().size >= 2
