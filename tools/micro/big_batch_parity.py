"""Parity at the saturating batch: hopper, 262 144 problems (seed 123), CUDA path vs the oracle on all host threads, by the rule of tests/common.py."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import optimization_dynamics_b200 as od
from oracle import oracle as O
from common import compare, oracle_pair
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
q1, q2, u = od.workloads.hopper_batch(B, h=0.05, seed=123)
dyn = od.ImplicitDynamics(od.hopper, 0.05, r_tol=1e-8, κ_eval_tol=1e-4, κ_grad_tol=1e-3)
t0 = time.perf_counter(); q3, d1, d2, du, st = dyn.step_grad_batch(q1, q2, u); t_gpu = time.perf_counter() - t0
t0 = time.perf_counter(); e, g = oracle_pair(O, "hopper", q1, q2, u); t_cpu = time.perf_counter() - t0
eq, eg = compare("hopper", e, g, q3, d1, d2, du, st & 15, (st >> 4) & 15)
print("hopper B=%d: max|q3 - oracle| %.2e  max|grad - oracle| %.2e  status agreement %.6f  iteration counts: GPU n/a, oracle mean %.2f max %d;  host call %.1f ms, oracle (with diagnostics, %d threads) %.1f s" % (
    B, eq, eg, float(((e["status"] == 0) == ((st & 15) == 0)).mean()), e["iters"].mean(), e["iters"].max(), t_gpu * 1e3, O.num_threads(), t_cpu))
