"""Device-resident time of the rocket kernel (CUDA events).  usage: python tools/micro/rocket_time.py [B]"""
import os, sys, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import optimization_dynamics_b200 as od
from optimization_dynamics_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
info = od.RocketInfo(od.rocket, 12.5, 0.05)
x, u = od.workloads.rocket_batch(B, seed=0)
dev = torch.device("cuda")
xd, ud = torch.from_numpy(x).to(dev), torch.from_numpy(u).to(dev)
y = torch.empty((B, 12), dtype=torch.float64, device=dev); dx = torch.empty((B, 144), dtype=torch.float64, device=dev); du = torch.empty((B, 36), dtype=torch.float64, device=dev)
st = torch.empty((B,), dtype=torch.int32, device=dev); it = torch.empty((B,), dtype=torch.int32, device=dev)
L = _lib.lib(); hd = info._hd if hasattr(info, "_hd") else info._handle()
L.od_set_stream(hd, C.c_void_p(torch.cuda.current_stream().cuda_stream))
p = lambda t: C.c_void_p(t.data_ptr())
for proj in (0, 1):
    f = lambda: L.od_rocket_batch_device(hd, B, p(xd), p(ud), proj, p(y), p(dx), p(du), p(st), p(it))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    itn = it.cpu().numpy()
    print("rocket B=%d proj=%d: %.4f ms/launch  %.3e solves/s  ok %.4f  dyn iters mean %.2f max %d  proj iters mean %.2f max %d" % (
        B, proj, ms, B / (ms * 1e-3), float((st == 0).float().mean()), (itn & 0xFFFF).mean(), (itn & 0xFFFF).max(), (itn >> 16).mean(), (itn >> 16).max()))
