"""Worker of tests/test_multi_gpu.py (launched by torch.distributed.run, one process per GPU, NCCL): every multi-GPU path of the
package against the single-GPU result computed on the same rank — bitwise."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import optimization_dynamics_b200 as od                                     # noqa: E402
from optimization_dynamics_b200 import device as D                          # noqa: E402
from common import CONFIGS                                                  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    report = []

    def make(name):
        gen, h, ke, kg, fric, attr = CONFIGS[name]
        m = getattr(od, attr)
        if fric is not None:
            m.friction[:] = fric
        dyn = od.ImplicitDynamics(m, h, r_tol=1e-8, κ_eval_tol=ke, κ_grad_tol=kg, device=local)
        return dyn, D.DeviceStepper(dyn), gen, h

    # ---- fused gather: even / ragged / fewer-problems-than-ranks shards, every barrier + store variant, register and non-register models
    for name, sizes in (("hopper", (4096, 1001, 1)), ("cartpole_friction", (513,)), ("planar_push", (130, 6600))):   # 6600: shards of ≥ 3072 run the persistent sweep + one forwarding kernel
        dyn, st, gen, h = make(name)
        for B_total in sizes:
            q1, q2, u = gen(B_total, h=h, seed=5)
            xall = torch.from_numpy(np.concatenate([q1, q2, u], axis=1)).to(dev)
            ref, st_ref = st.step_grad_packed(xall)                       # single-GPU result of the whole batch, on this rank
            lo, hi = D.shard_range(B_total, rank, world)
            for sync, mc in (("kernel", "auto"), ("kernel", False), ("launch", False)):
                if name == "planar_push" and sync == "kernel":
                    continue                                              # the in-kernel barrier needs the register path
                fg = D.FusedGather(st, B_total, sync=sync, multicast=mc)
                for rep in range(3):                                      # repeated steps: epochs advance, the two buffers alternate
                    got, stl = fg.step(xall[lo:hi].contiguous())
                    torch.cuda.synchronize(); dist.barrier()
                    ok = torch.equal(got, ref) and torch.equal(stl, st_ref[lo:hi])
                    report.append(("%s B=%d sync=%s multicast=%s(%s) rep=%d" % (name, B_total, fg.sync, mc, fg.multicast, rep), ok))
                del fg
            got = st.step_grad_sharded(xall[lo:hi].contiguous(), B_total)[0]    # kernel + ncclAllGather
            report.append(("%s B=%d nccl" % (name, B_total), torch.equal(got, ref)))
    # ---- a CUDA graph of fused steps replays correctly (device-side epoch)
    dyn, st, gen, h = make("hopper")
    q1, q2, u = gen(777, h=h, seed=9)
    xall = torch.from_numpy(np.concatenate([q1, q2, u], axis=1)).to(dev)
    ref, _ = st.step_grad_packed(xall)
    lo, hi = D.shard_range(777, rank, world)
    xl = xall[lo:hi].contiguous(); stl = torch.empty((hi - lo,), dtype=torch.int32, device=dev)
    fg = D.FusedGather(st, 777)
    fg.step(xl, stl); torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(4):
            fg.step(xl, stl)
    for rep in range(3):
        fg.buf.zero_(); torch.cuda.synchronize(); dist.barrier()
        g.replay(); torch.cuda.synchronize(); dist.barrier()
        report.append(("graph replay %d" % rep, torch.equal(fg.buf, ref)))
    # ---- gradient bundle, sample axis sharded
    dyn, st, gen, h = make("cartpole_friction")
    q1, q2, u = gen(50, h=h, seed=2)
    gb = od.GradientBundle(dyn.model, eta=od.workloads.bundle_perturbations(5, N=64, eps=1e-4, seed=0))
    bd = D.DeviceBundle(st, gb)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)          # noqa: E731
    dz1, s1 = bd.gradient_batch(t(q1), t(q2), t(u))
    dz2, s2 = bd.gradient_batch(t(q1), t(q2), t(u), sharded=True)
    report.append(("bundle sharded", torch.equal(dz1, dz2) and torch.equal(s1, s2)))
    # ---- rocket, sharded
    info = od.RocketInfo(od.rocket, 12.5, 0.05, device=local)
    rk = D.DeviceRocket(info)
    x, uu = od.workloads.rocket_batch(1025, seed=3)
    y, dx, du, s = rk.step(t(x), t(uu), True)
    lo, hi = D.shard_range(1025, rank, world)
    y2, dx2, du2, _ = rk.step_sharded(t(x[lo:hi]), t(uu[lo:hi]), 1025, True)
    report.append(("rocket sharded", torch.equal(y, y2) and torch.equal(dx, dx2) and torch.equal(du, du2)))
    # ---- host-facing sharded sweep
    dyn, st, gen, h = make("hopper")
    q1, q2, u = gen(1000, h=h, seed=11)
    xh = torch.from_numpy(np.concatenate([q1, q2, u], axis=1))
    ref, _ = st.step_grad_packed(xh.to(dev))
    lo, hi = D.shard_range(1000, rank, world)
    sh = D.ShardedHostSweep(st, 1000)
    outh = torch.empty((1000, st.out_width), dtype=torch.float64).pin_memory()
    sh.step(xh[lo:hi].contiguous().pin_memory(), outh)
    report.append(("sharded host sweep", torch.equal(outh, ref.cpu())))

    bad = [r for r in report if not r[1]]
    flag = torch.tensor([len(bad)], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        for r in report:
            print("%-70s %s" % (r[0], "ok" if r[1] else "FAILED"))
    if bad:
        print("rank %d FAILED: %r" % (rank, bad))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
