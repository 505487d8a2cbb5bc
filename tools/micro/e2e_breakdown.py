"""Where the end-to-end time of one od_step_grad_packed call goes (run on the GPU box)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import optimization_dynamics_b200 as od
from optimization_dynamics_b200 import _lib
import ctypes as C
H = 0.05
dyn = od.ImplicitDynamics(od.hopper, H, r_tol=1e-8, κ_eval_tol=1e-4, κ_grad_tol=1e-3, nc=4, nb=2)
B = 4096
q1, q2, u = od.workloads.hopper_batch(B, h=H, seed=0)
xin = torch.from_numpy(np.concatenate([q1, q2, u], axis=1)).pin_memory()
out = torch.empty((B, 44), dtype=torch.float64).pin_memory(); st = torch.empty((B,), dtype=torch.int32).pin_memory()
xn, on, sn = xin.numpy(), out.numpy(), st.numpy()
def timeit(f, n=300):
    for _ in range(20): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e6
print("full call B=4096            : %.1f us" % timeit(lambda: dyn.step_grad_packed(xn, on, sn)))
L = _lib.lib(); hd = dyn._handle()
pin, pout, pst = xn.ctypes.data_as(_lib.c_double_p), on.ctypes.data_as(_lib.c_double_p), sn.ctypes.data_as(_lib.c_int32_p)
print("raw ctypes call B=4096      : %.1f us" % timeit(lambda: L.od_step_grad_packed(hd, B, pin, pout, pst)))
print("raw ctypes call B=0 (no-op) : %.1f us" % timeit(lambda: L.od_step_grad_packed(hd, 0, pin, pout, pst)))
print("od_synchronize only         : %.1f us" % timeit(lambda: L.od_synchronize(hd)))
# device-resident launch + sync through the same handle (no host buffers)
xd = xin.cuda(); od_ = torch.empty((B, 44), dtype=torch.float64, device="cuda"); sd = torch.empty((B,), dtype=torch.int32, device="cuda")
def dev():
    L.od_step_grad_packed_device(hd, B, C.c_void_p(xd.data_ptr()), C.c_void_p(od_.data_ptr()), C.c_void_p(sd.data_ptr()), None, 1, 1); L.od_synchronize(hd)
print("device-resident launch+sync : %.1f us" % timeit(dev))
for z in (0, 1, 2):
    pass
