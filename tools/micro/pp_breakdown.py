"""Planar push (BASELINE configs[2]) kernel time as a function of the iteration cap and of the gradient request: separates the bulk
of the batch from the stragglers (problems that run to max_iter) and the interior-point loop from the rank-revealing IFT.
usage: python tools/micro/pp_breakdown.py [B=25600]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import optimization_dynamics_b200 as od
from optimization_dynamics_b200.device import DeviceStepper
from optimization_dynamics_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 25600
q1, q2, u = od.workloads.planar_push_batch(B, h=0.1, seed=0)
xin = torch.from_numpy(np.concatenate([q1, q2, u], axis=1)).cuda()
for max_iter in (100, 50, 25, 15):
    for want_grad in (True, False):
        dyn = od.ImplicitDynamics(od.planarpush, 0.1, r_tol=1e-8, κ_eval_tol=1e-4, κ_grad_tol=1e-2)
        dyn.opts.max_iter = max_iter; dyn._make_handle()
        st = DeviceStepper(dyn)
        out = torch.empty((B, st.out_width), dtype=torch.float64, device="cuda"); s = torch.empty((B,), dtype=torch.int32, device="cuda"); it = torch.empty((B,), dtype=torch.int32, device="cuda")
        for _ in range(2): st.step_grad_packed(xin, out, s, it, want_grad=want_grad)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): st.step_grad_packed(xin, out, s, it, want_grad=want_grad)
        e1.record(); torch.cuda.synchronize()
        itn = (it.cpu().numpy() & 0xFFFF)
        hist = np.bincount(np.minimum(itn, 100), minlength=101)
        print("B=%d max_iter=%3d grad=%d lanes=%s: %.4f ms  %.3e solves/s  converged %.4f  iters mean %.2f  p50 %d p90 %d p99 %d max %d  (>25: %d, >50: %d)" % (
            B, max_iter, want_grad, os.environ.get("OD_LANES", "auto"), e0.elapsed_time(e1) / 5, B / (e0.elapsed_time(e1) / 5 * 1e-3), float(((s & 15) == 0).float().mean()),
            itn.mean(), np.percentile(itn, 50), np.percentile(itn, 90), np.percentile(itn, 99), itn.max(), (itn > 25).sum(), (itn > 50).sum()))
