"""BASELINE configs[2] as the reference phrases it — planar push `rotate`, T = 26, 1024 rollouts: one contact_rollout_kernel launch (device
resident and through the host API) next to the CPU oracle's rollout loop."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import optimization_dynamics_b200 as od
from optimization_dynamics_b200.device import DeviceStepper, DeviceSolverStages
from oracle import oracle as O
h, T = 0.1, 26
for R in (64, 1024):
    x1, ubar = od.workloads.planar_push_rollout_inputs(R, T=T, h=h, seed=1)
    dyn = od.ImplicitDynamics(od.planarpush, h, r_tol=1e-8, κ_eval_tol=1e-4, κ_grad_tol=1e-2)
    stg = DeviceSolverStages(DeviceStepper(dyn))
    dx1, dub = torch.from_numpy(x1).cuda(), torch.from_numpy(ubar).cuda()
    X = torch.empty((R, T, 10), dtype=torch.float64, device="cuda"); U = torch.empty((R, T - 1, 2), dtype=torch.float64, device="cuda"); st = torch.empty((R, T - 1), dtype=torch.int32, device="cuda")
    for _ in range(2): stg.rollouts(dx1, dub, X=X, U=U, status=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): stg.rollouts(dx1, dub, X=X, U=U, status=st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    t0 = time.perf_counter(); od.rollout_batch(dyn, x1, ubar); t_host = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); Xo, Uo, so = O.rollout_batch("planar_push", x1, ubar, h, 1e-4); t_cpu = (time.perf_counter() - t0) * 1e3
    ok = (st.cpu().numpy() == 0).all(1) & (so == 0).all(1)
    err = np.abs(X.cpu().numpy() - Xo)[ok].max() if ok.any() else float("nan")
    print("planar push T=%d R=%5d: rollout kernel %.3f ms device-resident (%.2e f-calls/s) | host API %.2f ms | CPU oracle loop (%d threads) %.1f ms | max|X - oracle| on %d clean rollouts %.2e" % (
        T, R, ms, R * (T - 1) / (ms * 1e-3), t_host, O.num_threads(), t_cpu, ok.sum(), err))
