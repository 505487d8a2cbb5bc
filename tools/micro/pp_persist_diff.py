"""Debug aid: the planar-push batch through the per-warp kernel (OD_PERSIST=0) and through the persistent sweep (OD_PERSIST=1), each in
its own process; per-problem comparison of q3, Jacobians, status and iteration counts."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    sys.path.insert(0, ROOT)
    import torch
    import optimization_dynamics_b200 as od
    from optimization_dynamics_b200.device import DeviceStepper
    B = int(sys.argv[2])
    q1, q2, u = od.workloads.planar_push_batch(B, h=0.1, seed=17)
    dyn = od.ImplicitDynamics(od.planarpush, 0.1, r_tol=1e-8, κ_eval_tol=1e-4, κ_grad_tol=1e-2)
    st = DeviceStepper(dyn)
    xin = torch.from_numpy(np.concatenate([q1, q2, u], axis=1)).cuda()
    out = torch.zeros((B, st.out_width), dtype=torch.float64, device="cuda"); s = torch.zeros((B,), dtype=torch.int32, device="cuda"); it = torch.zeros((B,), dtype=torch.int32, device="cuda")
    st.step_grad_packed(xin, out, s, it); torch.cuda.synchronize()
    np.savez(sys.argv[3], out=out.cpu().numpy(), st=s.cpu().numpy(), it=it.cpu().numpy())
    sys.exit(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4700
res = []
for p in (0, 1):
    f = "/tmp/pp_persist_%d.npz" % p
    subprocess.check_call([sys.executable, os.path.abspath(__file__), "worker", str(B), f], env=dict(os.environ, OD_PERSIST=str(p), OD_LANES="8"))
    res.append(np.load(f))
a, b = res
ite_a, ite_b = a["it"] & 0xFFFF, b["it"] & 0xFFFF
itg_a, itg_b = a["it"] >> 16, b["it"] >> 16
print("status equal %.5f   eval iterations equal %.5f   grad iterations equal %.5f" % ((a["st"] == b["st"]).mean(), (ite_a == ite_b).mean(), (itg_a == itg_b).mean()))
dq = np.abs(a["out"][:, :5] - b["out"][:, :5]).max(1); dg = np.abs(a["out"][:, 5:] - b["out"][:, 5:]).max(1)
same_it = (ite_a == ite_b) & (itg_a == itg_b) & (a["st"] == 0) & (b["st"] == 0)
print("problems with equal iteration counts: q3 bit-identical %.5f, max|dq3| %.3e; grad bit-identical %.5f, max|dgrad| %.3e" % (
    (dq[same_it] == 0).mean(), dq[same_it].max(), (dg[same_it] == 0).mean(), dg[same_it].max()))
d = np.where(~(ite_a == ite_b))[0]
print("%d problems differ in eval iterations; first ones (index, per-warp it, persistent it, status a, status b):" % len(d))
for i in d[:12]:
    print("   ", i, ite_a[i], ite_b[i], a["st"][i], b["st"][i])
nz = np.where(same_it & (dq > 0))[0]
print("%d problems with equal counts but different q3 bits; iteration histogram of those:" % len(nz), np.bincount(ite_a[nz])[:40])
print("iteration histogram of the bit-identical ones:", np.bincount(ite_a[same_it & (dq == 0)])[:40])
d = dq[nz]
print("|dq3| of those: median %.2e  p90 %.2e  p99 %.2e  max %.2e;   relative to |q3|: median %.2e" % (np.median(d), np.quantile(d, .9), np.quantile(d, .99), d.max(),
      np.median(d / np.abs(a["out"][nz, :5]).max(1))))
print("differing fraction among even indices (pusher touching) %.3f, odd (gap) %.3f" % ((dq[same_it][0::2] > 0).mean(), (dq[same_it][1::2] > 0).mean()))
