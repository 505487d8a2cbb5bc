"""Device-resident kernel time of the step+gradient kernel for any contact model / batch size (CUDA events, L2 not flushed).
usage: python tools/micro/kernel_time.py <config name> <B> [reps]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import optimization_dynamics_b200 as od
from optimization_dynamics_b200.device import DeviceStepper
from common import CONFIGS
name, B = sys.argv[1], int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
gen, h, ke, kg, fric, attr = CONFIGS[name]
m = getattr(od, attr)
if fric is not None: m.friction[:] = fric
dyn = od.ImplicitDynamics(m, h, r_tol=1e-8, κ_eval_tol=ke, κ_grad_tol=kg)
q1, q2, u = gen(B, h=h, seed=0)
st = DeviceStepper(dyn)
xin = torch.from_numpy(np.concatenate([q1, q2, u], axis=1)).cuda()
out = torch.empty((B, st.out_width), dtype=torch.float64, device="cuda"); s = torch.empty((B,), dtype=torch.int32, device="cuda"); it = torch.empty((B,), dtype=torch.int32, device="cuda")
for _ in range(3): st.step_grad_packed(xin, out, s, it)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): st.step_grad_packed(xin, out, s, it)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
itn = (it.cpu().numpy() & 0xFFFF)
ck = float(out[s == 0].abs().sum())
print("%s B=%d lanes=%s reg=%s: %.4f ms/launch  %.3e solves/s  converged %.4f  iters mean %.2f max %d  checksum %.12e" % (
    name, B, os.environ.get("OD_LANES", "auto"), os.environ.get("OD_REG", "1"), ms, B / (ms * 1e-3), float((s == 0).float().mean()), itn.mean(), itn.max(), ck))
