#!/usr/bin/env python
"""bench.py — contact-step + IFT-gradient solves/sec (hopper, batch 4096 per GPU), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N …

A "step" = one pass of the hot path over one batch: every rank solves its B (q1,q2,u) hopper problems (q3 at κ_eval = 1e-4 and
∂q3/∂(q1,q2,u1) at κ_grad = 1e-3 — what f + fx + fu deliver, reference src/dynamics.jl:81-128) with inputs resident in HBM and,
for N > 1, all-gathers the packed 352-B output rows so that every rank holds all Jacobians for the sequential Riccati pass
(weak scaling: B per GPU fixed).  Timed on the device with CUDA events around each step; the inputs of a step are kept out
of L2 by reading a different device copy of the batch every step from a pool larger than L2 (`--l2 pool`, default) or by a
256 MiB memset before every step, outside the event pair (`--l2 flush`).  `e2e` is the same metric through the public host API (ImplicitDynamics.step_grad_packed → C ABI) with pinned
HOST buffers, H2D + kernel + D2H inside the timed region.

`--impl reference` times the CPU restatement of the reference path (oracle/, kind "port": Julia and RoboDojo.jl are not
available here) with the reference's own call pattern (f, fx, fu = 3 interior-point solves per unit, src/dynamics.jl:88,103,123)
on all host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = 0.05
KAPPA_EVAL, KAPPA_GRAD, R_TOL = 1.0e-4, 1.0e-3, 1.0e-8      # reference examples/hopper.jl:42
BYTES_IN, BYTES_OUT = 80, 352                               # 10 + 44 fp64 words per unit
METRIC = "contact-step+IFT-gradient solves/sec (hopper, batch 4096)"
UNIT = "solves/s"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic():
    """Per-launch DRAM bytes of the step kernel from the committed ncu capture summary (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "latest_kernel_summary.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


def fp64_side():
    """Compute-side companion of the (mandated) HBM roofline: measured fp64 FMA peak of the GPU (tools/micro/fp64_peak.cu →
    profiles/fp64_peak.json) and the fp64-pipe utilisation of the step kernel in the committed ncu capture."""
    out = {}
    try:
        with open(os.path.join(ROOT, "profiles", "fp64_peak.json")) as f:
            out["fp64_fma_peak_tflops"] = float(json.load(f)["fp64_fma_tflops"])
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "latest_kernel_summary.json")) as f:
            l = json.load(f)["launches"][0]
        for k, v in l.items():
            if k.startswith("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"):
                out["fp64_pipe_busy_pct_ncu"] = float(v)
            if k.startswith("smsp__issue_active.avg.pct_of_peak_sustained_active"):
                out["issue_slots_busy_pct_ncu"] = float(v)
    except Exception:
        pass
    return out or None


class ClockSampler:
    """Samples SM clock and throttle reasons of the local GPU through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def run(self, period=0.02):
        def loop():
            while not self._stop.is_set():
                self.sample()
                time.sleep(period)
        self.t = threading.Thread(target=loop, daemon=True)
        self.t.start()

    def stop(self):
        self._stop.set()
        if hasattr(self, "t"):
            self.t.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_baseline(B, seconds=10.0, pattern="D", nthreads=0):
    """Oracle (port) on the host cores: repeated passes over the same hopper batch for about `seconds`.
    pattern D = eval solve + grad solve+IFT (2 solves/unit); pattern R = the reference's f, fx, fu (3 solves/unit)."""
    from oracle import oracle as O
    from optimization_dynamics_b200 import workloads as W
    q1, q2, u = W.hopper_batch(B, h=H, seed=0)
    cores = O.num_threads() if nthreads <= 0 else nthreads

    def one_pass():
        O.step_batch("hopper", q1, q2, u, H, KAPPA_EVAL, False, r_tol=R_TOL, nthreads=nthreads, diagnostics=False)
        O.step_batch("hopper", q1, q2, u, H, KAPPA_GRAD, True, r_tol=R_TOL, nthreads=nthreads, diagnostics=False)
        if pattern == "R":
            O.step_batch("hopper", q1, q2, u, H, KAPPA_GRAD, True, r_tol=R_TOL, nthreads=nthreads, diagnostics=False)
    one_pass()
    t0 = time.perf_counter(); n = 0
    while True:
        one_pass(); n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    return {"value": n * B / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d passes over the same %d-problem hopper batch in %.1f s, call pattern %s (%d interior-point solves per unit), "
                      "oracle/ C++ restatement (dense 20x20 LU, dual-number Jacobians), std::thread over %d host threads" % (
                          n, B, dt, pattern, 3 if pattern == "R" else 2, cores)}, dt / n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g   # builds the oracle if needed
    from oracle import oracle as O
    O.build()
    B = args.batch
    times = []
    from optimization_dynamics_b200 import workloads as W
    q1, q2, u = W.hopper_batch(B, h=H, seed=0)

    def step():
        O.step_batch("hopper", q1, q2, u, H, KAPPA_EVAL, False, r_tol=R_TOL, diagnostics=False)      # f
        O.step_batch("hopper", q1, q2, u, H, KAPPA_GRAD, True, r_tol=R_TOL, diagnostics=False)       # fx
        O.step_batch("hopper", q1, q2, u, H, KAPPA_GRAD, True, r_tol=R_TOL, diagnostics=False)       # fu (re-solves, src/dynamics.jl:123)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    cores = O.num_threads()
    sample = ("each step = one pass over the %d-problem hopper batch with the reference call pattern f+fx+fu (3 interior-point solves "
              "+ 2 IFTs per unit); oracle/ C++ port of the Julia path, %d host threads (the Julia reference itself is single-threaded)" % (B, cores))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": "hopper gait contact step + IFT gradient, batch %d, h=0.05 (CPU oracle port)" % B, "batch": B},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=4096, help="problems per GPU (weak scaling)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also report a saturating batch (262144 per GPU) in the JSON line")
    ap.add_argument("--l2", default="pool", choices=["pool", "flush"],
                    help="how inputs are kept out of L2 between timed steps: 'pool' = every step reads another copy of the batch from a "
                         "pool of device buffers larger than L2 (each buffer is used once per pass over the pool); 'flush' = a 256 MiB "
                         "memset before every step (it also evicts the kernel's code, which a solver loop would find cached)")
    ap.add_argument("--collective", default="fused", choices=["fused", "fused-kernel-barrier", "fused-launch-barrier", "nccl"],
                    help="N>1: 'fused' = all-gather fused into the kernel over NVLink peer memory, cross-rank barrier fused as well up to 4 ranks "
                         "(falls back to nccl if symmetric memory is unavailable); 'fused-kernel-barrier' / 'fused-launch-barrier' force the "
                         "barrier variant; "
                         "'nccl' = kernel + ncclAllGather")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 200 and args.warmup == 10:
            args.steps, args.warmup = 20, 3
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    import optimization_dynamics_b200 as od
    from optimization_dynamics_b200.device import DeviceStepper, FusedGather, all_gather_rows, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this benchmark has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    B_total = B * world

    dyn = od.ImplicitDynamics(od.hopper, H, r_tol=R_TOL, κ_eval_tol=KAPPA_EVAL, κ_grad_tol=KAPPA_GRAD, nc=4, nb=2, device=local)
    stepper = DeviceStepper(dyn)
    q1, q2, u = od.workloads.hopper_batch(B, h=H, seed=rank)              # a different seeded batch on every rank
    xin_host = torch.from_numpy(np.concatenate([q1, q2, u], axis=1)).pin_memory()
    xin = xin_host.to(dev)
    out = torch.empty((B, stepper.out_width), dtype=torch.float64, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    gathered = torch.empty((B_total, stepper.out_width), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # 2× the 126 MB L2
    # --l2 pool: P device copies of the batch, P·B·80 B ≥ 1.05 × L2 (B200: 126 MB), used round-robin — the rows a step reads were
    # last touched a whole pool (> L2) of traffic ago, so they come from HBM, while the kernel's instructions stay cached as they
    # would in a solver loop.  The pool is flushed once after it has been filled.
    pool = None
    if args.l2 == "pool":
        l2_bytes = torch.cuda.get_device_properties(dev).L2_cache_size
        P = max(2, int(1.05 * l2_bytes // (B * BYTES_IN)) + 1 + args.warmup)
        pool = xin.unsqueeze(0).repeat(P, 1, 1).contiguous()
        flush.zero_()
        torch.cuda.synchronize()

    def batch_in(k):               # inputs of timed step k (warm-up steps take the buffers from the end of the pool)
        return xin if pool is None else pool[k % pool.shape[0]]
    fused = None
    if world > 1 and args.collective != "nccl":
        okf = torch.ones(1, device=dev)
        try:
            fused = FusedGather(stepper, B_total, sync={"fused": "auto", "fused-kernel-barrier": "kernel", "fused-launch-barrier": "launch"}[args.collective])
        except Exception as ex:                       # symmetric memory unavailable: every rank must take the same path
            sys.stderr.write("rank %d: fused gather unavailable (%r), using ncclAllGather\n" % (rank, ex))
            okf.zero_()
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        if okf.item() == 0:
            fused = None

    def step(xin=xin):
        if fused is not None:
            fused.step(xin, status)
            return
        stepper.step_grad_packed(xin, out, status)
        if world > 1:
            all_gather_rows(out, B_total, gathered)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        if pool is None:
            flush.zero_(); step()
        else:
            step(batch_in(-1 - w))
    barrier()
    clocks = ClockSampler(local)
    clocks.run()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    kmid = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    n0 = dyn.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        if pool is None:
            flush.zero_()
        xk = batch_in(k)
        starts[k].record()
        if fused is not None:
            fused.launch(xk, status)
            kmid[k].record()
            fused.barrier()
        else:
            stepper.step_grad_packed(xk, out, status)
            kmid[k].record()
            if world > 1:
                all_gather_rows(out, B_total, gathered)
        ends[k].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = dyn.launch_count() - n0
    clk = clocks.stop()
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    ker_ms = sum(s.elapsed_time(e) for s, e in zip(starts, kmid))
    ok_frac = float((status == 0).float().mean().item())

    gather_check = None
    if world > 1:
        # the rows every rank now holds must be bit-identical to a plain ncclAllGather of the per-rank results
        stepper.step_grad_packed(xin, out, status)
        ref = all_gather_rows(out, B_total)
        if fused is not None:
            got, _ = fused.step(xin, status)
            torch.cuda.synchronize()
            gather_check = bool(torch.equal(got, ref))
            assert gather_check, "fused gather differs from ncclAllGather"
        lo, hi = shard_range(B_total, rank, world)
        assert torch.equal(ref[lo:hi], out)
        out_dev_result = out
    # ---- end-to-end through the public host API: pinned host in/out, H2D + kernel + D2H every step -----------------------------
    out_host = torch.empty((B, stepper.out_width), dtype=torch.float64).pin_memory()
    st_host = torch.empty((B,), dtype=torch.int32).pin_memory()
    xin_np, out_np, st_np = xin_host.numpy(), out_host.numpy(), st_host.numpy()
    dyn2 = od.ImplicitDynamics(od.hopper, H, r_tol=R_TOL, κ_eval_tol=KAPPA_EVAL, κ_grad_tol=KAPPA_GRAD, nc=4, nb=2, device=local)
    for _ in range(args.warmup):
        dyn2.step_grad_packed(xin_np, out_np, st_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dyn2.step_grad_packed(xin_np, out_np, st_np)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert np.array_equal(out_np, out.cpu().numpy()), "host API and device-resident path disagree"

    extra = None
    if args.extra:
        Bs = 262144
        a, b, c = od.workloads.hopper_batch(Bs, h=H, seed=100 + rank)
        xs = torch.from_numpy(np.concatenate([a, b, c], axis=1)).to(dev)
        os_ = torch.empty((Bs, stepper.out_width), dtype=torch.float64, device=dev); ss = torch.empty((Bs,), dtype=torch.int32, device=dev)
        for _ in range(3):
            stepper.step_grad_packed(xs, os_, ss)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            stepper.step_grad_packed(xs, os_, ss)
        e1.record(); torch.cuda.synchronize()
        extra = {"batch_per_gpu": Bs, "ms_per_launch": e0.elapsed_time(e1) / reps, "solves_per_s_per_gpu": Bs * reps / (e0.elapsed_time(e1) * 1e-3)}

    # max over ranks of the timed quantities
    t = torch.tensor([dev_ms, ker_ms, e2e_s, t_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, ker_ms, e2e_s, t_wall = [float(v) for v in t.tolist()]

    if rank == 0:
        peak, peak_src = measured_peak()
        ms_per_step = dev_ms / args.steps
        value = B_total * args.steps / (dev_ms * 1e-3)
        ker_ms_per = ker_ms / args.steps
        achieved = (BYTES_IN + BYTES_OUT) * B / (ker_ms_per * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "hopper gait contact step + IFT gradient (RoboDojo hopper, nq=4, nz=20), batch %d per GPU, h=0.05, "
                                   "kappa_eval=1e-4, kappa_grad=1e-3, r_tol=1e-8 (BASELINE.json configs[3])" % B,
                       "batch_per_gpu": B, "global_batch": B_total, "l2": ("flushed between timed steps (256 MiB memset outside the event pair)" if pool is None else
                              "inputs larger than L2: every timed step reads another device copy of the batch from a pool of %d buffers "
                              "(%.0f MB > %.0f MB L2; flushed once after filling), so its rows come from HBM; no per-step flush" % (
                                  pool.shape[0], pool.numel() * 8 / 1e6, torch.cuda.get_device_properties(dev).L2_cache_size / 1e6)),
                       "collective": ("none (1 GPU)" if world == 1 else
                                      ("all-gather and cross-rank barrier fused into the kernel: P2P stores of each finished 352-B row into every rank's buffer over NVLink, "
                                       "completion flags published by the last block of each rank" if fused.sync == "kernel" else
                                       "all-gather fused into the kernel: P2P stores of each finished 352-B row into every rank's buffer over NVLink + symmetric-memory barrier launch")
                                      if fused is not None else "kernel + ncclAllGather of 352-B rows"),
                       "gather_check_bitwise_equal_to_nccl": gather_check, "converged_fraction": ok_frac},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(),
                         "kernel": "od::contact_step_kernel<HopperModel, lanes=%d, problems/block=%d, register Gauss-Jordan>" % ((8, 4) if B <= 8192 else (4, 8)), "kernel_ms": ker_ms_per, "algorithmic_bytes_per_launch": (BYTES_IN + BYTES_OUT) * B, "compute_side": fp64_side(),
                         "peak_source": peak_src,
                         "note": "432 B vs ~1e5 fp64 flop per unit: the kernel is fp64-latency bound by construction (DESIGN.md §Roofline)"},
            "e2e": {"value": B_total * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": BYTES_IN * B, "d2h_bytes_per_step": (BYTES_OUT + 4) * B,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "api": "ImplicitDynamics.step_grad_packed -> od_step_grad_packed (pinned host buffers: inputs copied H2D, output rows written by the kernel straight into host memory over PCIe)"},
            "gpu_launches": int(launches) * world,
            "clocks": clk,
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
        }
        if extra:
            line["saturating_batch"] = extra
        if not args.no_cpu_baseline:
            cb, _ = cpu_baseline(B, seconds=args.cpu_seconds, pattern="D")
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
