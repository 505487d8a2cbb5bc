"""Rollout timing on the GPU box: one contact_rollout_kernel launch vs (T−1) batched step launches vs the CPU oracle loop."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import optimization_dynamics_b200 as od
from oracle import oracle as O

def timeit(f, n=20):
    for _ in range(3): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3

h = 0.05
for R in (16, 256, 1024):
    x1, ubar, K, k, alpha = od.workloads.hopper_rollout_inputs(R, T=21, h=h, seed=3)
    alpha = np.maximum(alpha, 1e-5)
    dyn = od.ImplicitDynamics(od.hopper, h, r_tol=1e-8, κ_eval_tol=1e-4, κ_grad_tol=1e-3, nc=4, nb=2)
    xbar = np.stack(od.rollout(dyn, x1, ubar))
    t_roll = timeit(lambda: od.rollout_batch(dyn, x1, ubar, xbar=xbar, K=K, k=k, alpha=alpha))
    def stepwise():
        X = np.tile(x1, (R, 1))
        for t in range(20):
            u = ubar[t][None] + alpha[:, None] * k[t][None] + (X - xbar[t][None]) @ K[t].T
            q3, _ = dyn.step_batch(X[:, :4], X[:, 4:], u)
            X = np.concatenate([X[:, 4:], q3], axis=1)
    t_step = timeit(stepwise)
    t_cpu = timeit(lambda: O.rollout_batch("hopper", np.tile(x1, (R, 1)), ubar, h, 1e-4, xbar=xbar, K=K, k=k, alpha=alpha), n=3)
    print("hopper T=21 R=%5d: rollout kernel %.3f ms | 20 batched f launches %.3f ms | CPU oracle loop %.1f ms  (%.0f f-calls/s GPU)" % (
        R, t_roll, t_step, t_cpu, R * 20 / (t_roll * 1e-3)))
