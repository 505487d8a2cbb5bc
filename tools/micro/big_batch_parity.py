"""Parity at large batches (default: hopper, 262 144 problems, seed 123; usage: big_batch_parity.py [B] [config name]), CUDA path vs the oracle on all host threads, by the rule of tests/common.py."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import optimization_dynamics_b200 as od
from oracle import oracle as O
from common import compare, oracle_pair
from common import CONFIGS
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
NAME = sys.argv[2] if len(sys.argv) > 2 else "hopper"
gen, h, ke, kg, fric, attr = CONFIGS[NAME]
q1, q2, u = gen(B, h=h, seed=123)
model = getattr(od, attr)
if fric is not None:
    model.friction[:] = fric
dyn = od.ImplicitDynamics(model, h, r_tol=1e-8, κ_eval_tol=ke, κ_grad_tol=kg)
t0 = time.perf_counter(); q3, d1, d2, du, st = dyn.step_grad_batch(q1, q2, u); t_gpu = time.perf_counter() - t0
t0 = time.perf_counter(); e, g = oracle_pair(O, NAME, q1, q2, u); t_cpu = time.perf_counter() - t0
ok_e = (e["status"] == 0) & ((st & 15) == 0) & (e["margin"] > 1e-6) & (e["iters"] <= 30)
ok_g = (g["status"] == 0) & (((st >> 4) & 15) == 0) & (g["margin"] > 1e-6) & (g["ift_spread"] < 1e-8) & (g["iters"] <= 30)
errq = np.abs(q3 - e["q3"]).max(1)
errg = np.maximum.reduce([np.abs(d1 - g["dq1"].transpose(0, 2, 1)).reshape(B, -1).max(1), np.abs(d2 - g["dq2"].transpose(0, 2, 1)).reshape(B, -1).max(1),
                          np.abs(du - g["du"].transpose(0, 2, 1)).reshape(B, -1).max(1)])
vq = ok_e & (errq > 1e-8) & ~(e["q_uncertainty"] > 1e-7); vg = ok_g & (errg > 1e-6) & ~(g["q_uncertainty"] > 1e-7)
print("%s B=%d: comparable %.4f / %.4f;  q3: median %.1e p99.9 %.1e, %d samples (%.4f %%) outside 1e-8 with a well-determined iterate (max %.1e);  gradients: median %.1e p99.9 %.1e, %d samples (%.4f %%) outside 1e-6 (max %.1e)" % (
    NAME, B, ok_e.mean(), ok_g.mean(), np.median(errq[ok_e]), np.quantile(errq[ok_e], 0.999), vq.sum(), 100.0 * vq.mean(), errq[vq].max() if vq.any() else 0.0,
    np.median(errg[ok_g]), np.quantile(errg[ok_g], 0.999), vg.sum(), 100.0 * vg.mean(), errg[vg].max() if vg.any() else 0.0))
eq, eg = float(errq[ok_e & ~vq].max()), float(errg[ok_g & ~vg].max())
print(NAME + " B=%d: max|q3 - oracle| %.2e  max|grad - oracle| %.2e  status agreement %.6f  iteration counts: GPU n/a, oracle mean %.2f max %d;  host call %.1f ms, oracle (with diagnostics, %d threads) %.1f s" % (
    B, eq, eg, float(((e["status"] == 0) == ((st & 15) == 0)).mean()), e["iters"].mean(), e["iters"].max(), t_gpu * 1e3, O.num_threads(), t_cpu))
