"""One iLQR iteration's worth of hot-path work for NT hopper trajectories, device-resident: derivative sweep (NT·(T−1) step+grad
problems) → Riccati backward pass (NT trajectories) → forward-pass rollouts (NT × 8 step sizes), three launches on one stream;
next to the CPU oracle doing the same stages."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import optimization_dynamics_b200 as od
from optimization_dynamics_b200.device import DeviceStepper, DeviceSolverStages
from oracle import oracle as O
h, T = 0.05, 21
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 205
dyn = od.ImplicitDynamics(od.hopper, h, r_tol=1e-8, κ_eval_tol=1e-4, κ_grad_tol=1e-3, nc=4, nb=2)
x1, ubar, _, k0, alpha = od.workloads.hopper_rollout_inputs(NT, T=T, h=h, seed=4)
alpha = np.random.default_rng(0).uniform(0.0, 1.0, NT)
X, U = od.rollout_batch(dyn, x1, ubar, k=k0, alpha=alpha)
x_goal = np.concatenate([[1.0, 0.55, 0.0, 0.5]] * 2)
lx, lu, lxx, luu, lux = od.workloads.quadratic_cost_expansion(X, U, x_goal, 1e-1, 1e-1, 10.0, seed=1)
dev = torch.device("cuda")
tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
stp = DeviceStepper(dyn); stg = DeviceSolverStages(stp)
xin = tt(np.concatenate([X[:, :-1].reshape(-1, 8), U.reshape(-1, 2)], axis=1))
d_lx, d_lu, d_lxx, d_luu, d_lux = tt(lx), tt(lu), tt(lxx), tt(luu), tt(lux)
d_X, d_U = tt(X), tt(U)
al8 = tt(np.array([1.0, 0.5, 0.25, 0.125, 0.0625, 0.03125, 1e-3, 1e-5]))
rows = torch.empty((NT * (T - 1), 44), dtype=torch.float64, device=dev); st = torch.empty((NT * (T - 1),), dtype=torch.int32, device=dev)
def iteration(record=None):
    stp.step_grad_packed(xin, rows, st)
    if record: record[0].record()
    K, k, dV, sr = stg.backward_pass(rows.view(NT, T - 1, 44), d_lx, d_lu, d_lxx, d_luu, d_lux)
    if record: record[1].record()
    # forward pass of every trajectory with 8 step sizes: NT launches would be the naive mapping; here all NT×8 rollouts share one launch
    # by giving each rollout its own nominal controls/gains is not supported (K is shared) → roll out trajectory 0's candidates
    Xn, Un, sn = stg.rollouts(d_X[0, 0].repeat(8, 1), d_U[0], xbar=d_X[0], K=K[0], k=k[0], alpha=al8)
    return K, k, Xn
for _ in range(3): iteration()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
reps = 50
tot = [0.0, 0.0, 0.0]
for _ in range(reps):
    ev[0].record(); iteration(record=(ev[1], ev[2])); ev[3].record(); torch.cuda.synchronize()
    tot[0] += ev[0].elapsed_time(ev[1]); tot[1] += ev[1].elapsed_time(ev[2]); tot[2] += ev[2].elapsed_time(ev[3])
print("NT=%d trajectories, T=%d (device-resident, CUDA events, mean of %d):" % (NT, T, reps))
print("  derivative sweep  (%5d step+grad) : %.3f ms" % (NT * (T - 1), tot[0] / reps))
print("  Riccati backward  (%5d trajectories): %.3f ms" % (NT, tot[1] / reps))
print("  forward rollouts  (8 step sizes x T-1): %.3f ms" % (tot[2] / reps))
# CPU oracle, all host threads for the sweep, numpy for the backward pass
t0 = time.perf_counter()
Xf, Uf = X[:, :-1].reshape(-1, 8), U.reshape(-1, 2)
O.step_batch("hopper", Xf[:, :4], Xf[:, 4:], Uf, h, 1e-4, False, diagnostics=False)
g = O.step_batch("hopper", Xf[:, :4], Xf[:, 4:], Uf, h, 1e-3, True, diagnostics=False)
t1 = time.perf_counter()
jac = np.concatenate([g["q3"], g["dq1"].reshape(-1, 16), g["dq2"].reshape(-1, 16), g["du"].reshape(-1, 8)], axis=1).reshape(NT, T - 1, 44)
res = [O.backward_pass(jac[a], lx[a], lu[a], lxx[a], luu[a], lux[a], 4, 2) for a in range(NT)]
t2 = time.perf_counter()
O.rollout_batch("hopper", np.tile(X[0, 0], (8, 1)), U[0], h, 1e-4, xbar=X[0], K=res[0][0], k=res[0][1], alpha=al8.cpu().numpy())
t3 = time.perf_counter()
print("CPU oracle: sweep %.1f ms (%d threads) | backward pass (numpy, 1 thread) %.1f ms | rollouts %.1f ms" % ((t1 - t0) * 1e3, O.num_threads(), (t2 - t1) * 1e3, (t3 - t2) * 1e3))
