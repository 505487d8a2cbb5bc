"""Small launches of every kernel family, meant to run under `compute-sanitizer --tool memcheck` (out-of-bounds / misaligned global,
shared and local accesses): hopper 4 / 8 lanes, cartpole, acrobot, planar push per-warp and persistent sweep + IFT kernel, rocket with
projection, gradient bundle, rollouts, Riccati pass, the user-model library."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import optimization_dynamics_b200 as od
from common import CONFIGS
for name, B in (("hopper", 700), ("hopper", 3000), ("cartpole_friction", 300), ("acrobot_impact", 300), ("acrobot_nominal", 100), ("cartpole_frictionless", 100),
                ("planar_push", 200), ("planar_push", 4200)):
    gen, h, ke, kg, fric, attr = CONFIGS[name]
    m = getattr(od, attr)
    if fric is not None: m.friction[:] = fric
    dyn = od.ImplicitDynamics(m, h, r_tol=1e-8, κ_eval_tol=ke, κ_grad_tol=kg)
    q1, q2, u = gen(B, h=h, seed=1)
    q3, d1, d2, du, st = dyn.step_grad_batch(q1, q2, u)
    q3b, stb = dyn.step_batch(q1[:65], q2[:65], u[:65])
    print(name, B, "converged", float((st == 0).mean()), "launches", dyn.launch_count())
info = od.RocketInfo(od.rocket, 12.5, 0.05)
x, u = od.workloads.rocket_batch(300, seed=2)
y, dx, du, st = info.step_batch(x, u, proj=True); print("rocket ok", float((st == 0).mean()))
dyn = od.ImplicitDynamics(od.cartpole_friction, 0.05, r_tol=1e-8, κ_eval_tol=1e-4, κ_grad_tol=1e-3)
gb = od.GradientBundle(od.cartpole_friction, eta=od.workloads.bundle_perturbations(5, N=16, eps=1e-4, seed=0))
q1, q2, u = od.workloads.cartpole_batch(20, seed=3)
dz, st = od.gradient_batch(dyn, gb, q1, q2, u); print("bundle ok", float((st == 0).mean()))
dyn = od.ImplicitDynamics(od.hopper, 0.05, r_tol=1e-8, κ_eval_tol=1e-4, κ_grad_tol=1e-3)
x1, ubar, K, k, alpha = od.workloads.hopper_rollout_inputs(12, T=9, h=0.05, seed=3)
xbar = np.stack(od.rollout(dyn, x1, ubar)); X, U = od.rollout_batch(dyn, x1, ubar, xbar=xbar, K=K, k=k, alpha=alpha); print("rollouts ok", X.shape)
from optimization_dynamics_b200.user_model import build_user_model, UserModelDynamics
um = UserModelDynamics(build_user_model(os.path.join(ROOT, "tools", "codegen", "examples", "particle_spec.py")), 0.05, κ_eval_tol=1e-4, κ_grad_tol=1e-3, friction=[0.5])
r = um.step_grad_batch(np.zeros((33, 2)), np.zeros((33, 2)), np.ones((33, 2))); print("user model ok", float((r[4] == 0).mean()))
