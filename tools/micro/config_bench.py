"""Every BASELINE.json config at its own size through the host API on the GPU box, next to the CPU oracle (all host threads).
Not a bench.py line (the headline is configs[3]); a table for DESIGN.md."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import optimization_dynamics_b200 as od
from oracle import oracle as O
from common import CONFIGS

def timeit(f, n=20, warm=3):
    for _ in range(warm): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3

def dyn_of(name):
    gen, h, ke, kg, fric, attr = CONFIGS[name]
    m = getattr(od, attr)
    if fric is not None: m.friction[:] = fric
    return od.ImplicitDynamics(m, h, r_tol=1e-8, κ_eval_tol=ke, κ_grad_tol=kg), gen, h, ke, kg, fric

rows = []
# cfg 0: acrobot with joint limits, T = 51 derivative sweep of one rollout (50 problems)
d, gen, h, ke, kg, fric = dyn_of("acrobot_impact"); q1, q2, u = gen(50, h=h, seed=0)
g = timeit(lambda: d.step_grad_batch(q1, q2, u)); c = timeit(lambda: (O.step_batch("acrobot_impact", q1, q2, u, h, ke, False, diagnostics=False), O.step_batch("acrobot_impact", q1, q2, u, h, kg, True, diagnostics=False)), n=5)
rows.append(("cfg0 acrobot impact, T=51 sweep (50 step+grad)", 50, g, c))
# cfg 1: cartpole friction, T = 51, gradient bundle N = 64
d, gen, h, ke, kg, fric = dyn_of("cartpole_friction"); q1, q2, u = gen(50, h=h, seed=0)
gb = od.GradientBundle(od.cartpole_friction, eta=od.workloads.bundle_perturbations(5, N=64, seed=0))
g = timeit(lambda: od.gradient_batch(d, gb, q1, q2, u)); c = timeit(lambda: O.bundle_batch("cartpole_friction", gb.eta, q1, q2, u, h, ke, fric=fric), n=5)
rows.append(("cfg1 cartpole friction, T=51, bundle N=64 (50x65 solves + fit)", 50 * 65, g, c))
# cfg 2: planar push, T = 26, 1024 rollouts: derivative sweep of all knot points + the rollouts themselves
d, gen, h, ke, kg, fric = dyn_of("planar_push"); q1, q2, u = gen(1024 * 25, h=h, seed=0)
g = timeit(lambda: d.step_grad_batch(q1, q2, u), n=5); c = timeit(lambda: (O.step_batch("planar_push", q1, q2, u, h, ke, False, diagnostics=False), O.step_batch("planar_push", q1, q2, u, h, kg, True, diagnostics=False)), n=1, warm=1)
rows.append(("cfg2 planar push, 1024 rollouts x 25 knot points step+grad", 1024 * 25, g, c))
x1, ub = od.workloads.planar_push_rollout_inputs(1024, T=26, h=h, seed=1)
g = timeit(lambda: od.rollout_batch(d, x1, ub), n=5); c = timeit(lambda: O.rollout_batch("planar_push", x1, ub, h, ke), n=1, warm=1)
rows.append(("cfg2 planar push, 1024 rollouts of T=26 (25600 sequential f)", 1024 * 25, g, c))
# cfg 3: hopper 4096 (the bench line) through the unpacked API
d, gen, h, ke, kg, fric = dyn_of("hopper"); q1, q2, u = gen(4096, h=h, seed=0)
g = timeit(lambda: d.step_grad_batch(q1, q2, u)); c = timeit(lambda: (O.step_batch("hopper", q1, q2, u, h, ke, False, diagnostics=False), O.step_batch("hopper", q1, q2, u, h, kg, True, diagnostics=False)), n=3)
rows.append(("cfg3 hopper 4096 step+grad (separate pageable arrays)", 4096, g, c))
# cfg 4: rocket belly flop with SOC projection, batch 8192
info = od.RocketInfo(od.rocket, 12.5, 0.05); x, uu = od.workloads.rocket_batch(8192, seed=0)
g = timeit(lambda: info.step_batch(x, uu, True), n=5); c = timeit(lambda: O.rocket_batch(x, uu, 0.05, 12.5, True, True), n=1, warm=1)
rows.append(("cfg4 rocket + SOC projection, 8192 step+grad", 8192, g, c))
print("%-66s %9s %11s %11s %8s" % ("config (host API, pageable numpy arrays in/out)", "units", "GPU ms", "CPU ms", "ratio"))
for name, n, g, c in rows:
    print("%-66s %9d %11.3f %11.1f %8.0f" % (name, n, g, c, c / g))
