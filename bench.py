#!/usr/bin/env python
"""bench.py — contact-step + IFT-gradient solves/sec (hopper, batch 4096), BASELINE.json's metric, and the other BASELINE configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config hopper|acrobot|cartpole_bundle|planar_push|rocket]
                    [--scaling strong|weak] [--batch B] [--impl ours|reference] [--extra]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N …

A "step" = one pass of the hot path over one batch (one derivative sweep): for every (q1, q2, u) problem q3 at κ_eval and
∂q3/∂(q1,q2,u1) at κ_grad — what f + fx + fu deliver (reference src/dynamics.jl:81-128) — with inputs resident in HBM.  At N > 1 the
batch is cut into contiguous shards, one per rank, and the packed output rows are all-gathered so that every rank holds all
Jacobians for the sequential Riccati pass.  `--scaling strong` (default; the metric is quoted on a GLOBAL batch of 4096): the same
4096 problems are split over the ranks; `--scaling weak`: 4096 per rank.  At N > 1 the line also carries the other scaling mode and
a saturating batch (262 144 per GPU) under `extra`.

Timing: the K timed steps are captured in ONE CUDA graph (a solver loop would do the same with its sweep → backward pass →
rollouts chain) and timed with two CUDA events around the graph launch, barrier + synchronize on both sides, max over ranks.
Every timed step reads another device copy of its inputs from a pool larger than L2, so input rows come from HBM.
`e2e` is the same metric through the public host API with pinned HOST buffers, H2D + kernel (+ gather) + D2H inside the timed region.

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

R_TOL = 1.0e-8
UNIT = "solves/s"
# name → workload description (BASELINE.json configs; algorithmic bytes per unit from BASELINE.md §4 / SURVEY.md §8 size table)
CONFIGS = {
    "hopper": dict(model="hopper", kind="contact", h=0.05, ke=1e-4, kg=1e-3, fric=None, batch=4096, bytes_in=80, bytes_out=352,
                   metric="contact-step+IFT-gradient solves/sec (hopper, batch 4096)", baseline="BASELINE.json configs[3] / north_star",
                   what="hopper gait contact step + IFT gradient (RoboDojo hopper, nq=4, nz=20), h=0.05, kappa_eval=1e-4, kappa_grad=1e-3, r_tol=1e-8"),
    "acrobot": dict(model="acrobot_impact", kind="contact", h=0.05, ke=1e-4, kg=1e-3, fric=None, batch=50, bytes_in=40, bytes_out=96,
                    metric="contact-step+IFT-gradient solves/sec (acrobot joint limits, T=51 single rollout: 50-timestep derivative sweep)",
                    baseline="BASELINE.json configs[0]",
                    what="acrobot with joint limits (nq=2, nz=6), one rollout of T=51: the 50 timesteps of its derivative sweep, h=0.05, kappa 1e-4/1e-3"),
    "cartpole_bundle": dict(model="cartpole_friction", kind="bundle", h=0.05, ke=1e-4, kg=1e-3, fric=[0.35, 0.35], batch=50, N=64, bytes_in=40, bytes_out=96,
                            metric="gradient-bundle Jacobians/sec (cartpole friction 0.35, T=51, N=64 samples: 50 x 65 eval solves + fit per sweep)",
                            baseline="BASELINE.json configs[1]",
                            what="cartpole joint friction mu=0.35 (nq=2, nz=10), T=51: gradient bundle with N=64 one-hot perturbations per timestep, eps=1e-4, kappa_eval=1e-4"),
    "planar_push": dict(model="planarpush", kind="contact", h=0.1, ke=1e-4, kg=1e-2, fric=None, batch=25600, bytes_in=96, bytes_out=520,
                        metric="contact-step+IFT-gradient solves/sec (planar push rotate, T=26, 1024 rollouts: 25600 solves per sweep)",
                        baseline="BASELINE.json configs[2]",
                        what="planar push (nq=5, nz=35), 1024 rollouts x 25 timesteps per derivative sweep, h=0.1, kappa_eval=1e-4, kappa_grad=1e-2"),
    "rocket": dict(model="rocket", kind="rocket", h=0.05, batch=8192, bytes_in=120, bytes_out=1536,
                   metric="rocket step+gradient solves/sec (SOC thrust projection + implicit midpoint, batch 8192)", baseline="BASELINE.json configs[4]",
                   what="rocket belly-flop with SOC thrust limits (12 states, 3 controls), u_max=12.5, h=0.05: projection solve + dynamics solve + both IFTs per unit"),
}
ORACLE_NAME = {"hopper": "hopper", "acrobot_impact": "acrobot_impact", "cartpole_friction": "cartpole_friction", "planarpush": "planar_push"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_summary(name):
    """Committed ncu capture summary of this config's dominant kernel (profiles/latest_<config>_kernel_summary.json), or {}."""
    for fn in ("latest_%s_kernel_summary.json" % name,) + (("latest_kernel_summary.json",) if name == "hopper" else ()):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                return json.load(f)
        except Exception:
            continue
    return {}


def compute_side(name):
    """Compute-side companion of the (mandated) HBM roofline: measured fp64 FMA peak of the GPU (tools/micro/fp64_peak.cu →
    profiles/fp64_peak.json) and the fp64-pipe / issue-slot utilisation of the kernel in the committed ncu capture."""
    out = {}
    try:
        with open(os.path.join(ROOT, "profiles", "fp64_peak.json")) as f:
            out["fp64_fma_peak_tflops"] = float(json.load(f)["fp64_fma_tflops"])
    except Exception:
        pass
    try:
        for k, v in ncu_summary(name)["launches"][0].items():
            if k.startswith("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"):
                out["fp64_pipe_busy_pct_ncu"] = float(v)
            if k.startswith("smsp__issue_active.avg.pct_of_peak_sustained_active"):
                out["issue_slots_busy_pct_ncu"] = float(v)
            if k.startswith("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"):
                out["tensor_pipe_busy_pct_ncu"] = float(v)
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

    def run(self, period=0.005):
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


def gen_inputs(cfg, B, seed):
    """Packed synthetic inputs [B, in_width] (numpy) of a config: contact / bundle rows [q1 | q2 | u], rocket rows [x | u]."""
    import numpy as np
    from optimization_dynamics_b200 import workloads as W
    if cfg["kind"] == "rocket":
        x, u = W.rocket_batch(B, seed=seed)
        return np.concatenate([x, u], axis=1)
    gen = {"hopper": W.hopper_batch, "acrobot_impact": W.acrobot_batch, "cartpole_friction": W.cartpole_batch, "planarpush": W.planar_push_batch}[cfg["model"]]
    q1, q2, u = gen(B, h=cfg["h"], seed=seed)
    return np.concatenate([q1, q2, u], axis=1)


def oracle_pass(cfg, xin, pattern, nthreads=0):
    """One pass of the CPU restatement over a batch.  pattern D = deduplicated (what the GPU path computes: eval solve + gradient solve
    + IFT), R = the reference's own call pattern f, fx, fu (3 interior-point solves per unit; src/dynamics.jl:88,103,123)."""
    from oracle import oracle as O
    if cfg["kind"] == "rocket":
        x, u = xin[:, :12], xin[:, 12:]
        O.rocket_batch(x, u, cfg["h"], 12.5, True, True, nthreads=nthreads)                       # fx (dx) — fu shares it in pattern D
        if pattern == "R":
            O.rocket_batch(x, u, cfg["h"], 12.5, True, False, nthreads=nthreads)                  # f
            O.rocket_batch(x, u, cfg["h"], 12.5, True, True, nthreads=nthreads)                   # fu re-solves (rocket/dynamics.jl:254-262)
        return
    name = ORACLE_NAME[cfg["model"]]
    nq = {"hopper": 4, "acrobot_impact": 2, "cartpole_friction": 2, "planar_push": 5}[name]
    q1, q2, u = xin[:, :nq], xin[:, nq:2 * nq], xin[:, 2 * nq:]
    if cfg["kind"] == "bundle":
        from optimization_dynamics_b200 import workloads as W
        eta = W.bundle_perturbations(2 * nq + (xin.shape[1] - 2 * nq), N=cfg["N"], eps=1e-4, seed=0)
        O.bundle_batch(name, eta, q1, q2, u, cfg["h"], cfg["ke"], fric=cfg["fric"], r_tol=R_TOL, nthreads=nthreads)     # fx_gb
        if pattern == "R":
            O.bundle_batch(name, eta, q1, q2, u, cfg["h"], cfg["ke"], fric=cfg["fric"], r_tol=R_TOL, nthreads=nthreads)  # fu_gb repeats it (gradient_bundle.jl:136-147)
        return
    O.step_batch(name, q1, q2, u, cfg["h"], cfg["ke"], False, fric=cfg["fric"], r_tol=R_TOL, nthreads=nthreads, diagnostics=False)     # f
    O.step_batch(name, q1, q2, u, cfg["h"], cfg["kg"], True, fric=cfg["fric"], r_tol=R_TOL, nthreads=nthreads, diagnostics=False)      # fx
    if pattern == "R":
        O.step_batch(name, q1, q2, u, cfg["h"], cfg["kg"], True, fric=cfg["fric"], r_tol=R_TOL, nthreads=nthreads, diagnostics=False)  # fu (re-solves, src/dynamics.jl:123)


def cpu_baseline(cfg, B, seconds=10.0, pattern="D", nthreads=0):
    """Oracle (port) on the host cores: repeated passes over the same batch for about `seconds`."""
    from oracle import oracle as O
    xin = gen_inputs(cfg, B, 0)
    cores = O.num_threads() if nthreads <= 0 else nthreads
    oracle_pass(cfg, xin, pattern, nthreads)
    t0 = time.perf_counter(); n = 0
    while True:
        oracle_pass(cfg, xin, pattern, nthreads); n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    return {"value": n * B / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d passes over the same %d-problem batch in %.1f s, call pattern %s (%s), oracle/ C++ restatement (dense LU, "
                      "dual-number Jacobians), std::thread over %d host threads" % (
                          n, B, dt, pattern, "reference call pattern f+fx+fu, 3 solves per unit" if pattern == "R" else "deduplicated: eval solve + gradient solve + IFT", cores)}


def config_dict(cfg, name, B_total, world, scaling):
    """Identical in both arms (ours / --impl reference) for the same command line."""
    return {"workload": "%s; global batch %d (%s)" % (cfg["what"], B_total, cfg["baseline"]), "config": name, "global_batch": B_total,
            "scaling": scaling, "n_gpus": world, "seed": 0,
            "l2": "GPU arm: inputs larger than L2 (every timed step reads another device copy of its batch from a pool larger than the L2, "
                  "flushed once after filling; no per-step flush); CPU arm: not applicable"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g   # noqa: F401  (builds the oracle if needed)
    from oracle import oracle as O
    O.build()
    cfg = CONFIGS[args.config]
    world = args.gpus
    B_total = (args.batch or cfg["batch"]) * (world if args.scaling == "weak" else 1)
    B = min(B_total, args.batch or cfg["batch"])          # bounded sample: one batch of the per-GPU size per step
    xin = gen_inputs(cfg, B, 0)
    for _ in range(args.warmup):
        oracle_pass(cfg, xin, "R")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_pass(cfg, xin, "R")
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    cores = O.num_threads()
    sample = ("each step = one pass over a %d-problem batch of this workload with the reference call pattern f+fx+fu (3 interior-point solves "
              "+ 2 IFTs per unit); oracle/ C++ port of the Julia path, %d host threads (the Julia reference itself is single-threaded)" % (B, cores))
    line = {"impl": "reference", "metric": cfg["metric"], "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(cfg, args.config, B_total, world, args.scaling),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


class Sweep:
    """One config's derivative sweep on this rank's GPU: device-resident step (+ exchange) and the host-API step for e2e."""

    def __init__(self, name, cfg, B_total, world, rank, dev, collective):
        import numpy as np
        import torch
        import optimization_dynamics_b200 as od
        from optimization_dynamics_b200 import device as D
        self.torch, self.np, self.od, self.D = torch, np, od, D
        self.name, self.cfg, self.B_total, self.world, self.rank, self.dev = name, cfg, B_total, world, rank, dev
        self.lo, self.hi = D.shard_range(B_total, rank, world)
        self.B = self.hi - self.lo
        self.collective = "none (1 GPU)"
        self.fused = None
        kind = cfg["kind"]
        if kind == "rocket":
            self.info = od.RocketInfo(od.rocket, 12.5, cfg["h"], device=dev.index)
            self.rk = D.DeviceRocket(self.info)
            self.in_width, self.out_width = 15, 12 + 144 + 36
            self.kernel = "od::rocket_kernel_g<4 lanes, 128-thread blocks phased at the phase boundaries> (SOC projection 10x10 + implicit midpoint 12x12, register Gauss-Jordan; 8 lanes below ~2000 problems)"
        else:
            model = getattr(od, cfg["model"])
            if cfg["fric"] is not None:
                model.friction[:] = cfg["fric"]
            self.dyn = od.ImplicitDynamics(model, cfg["h"], r_tol=R_TOL, κ_eval_tol=cfg["ke"], κ_grad_tol=cfg["kg"], device=dev.index)
            self.stepper = D.DeviceStepper(self.dyn)
            self.in_width, self.out_width = self.stepper.in_width, self.stepper.out_width
            self.kernel = "od::contact_step_kernel<%s> (cooperative lanes, register Gauss-Jordan)" % cfg["model"]
            if cfg["model"] == "planarpush":
                self.kernel = ("od::contact_sweep_kernel<planarpush, 8 lanes> (persistent, block-phased, queue-fed) + its resume launch for the problems parked "
                               "at 16 iterations (32 lanes) + od::contact_ift_kernel (rank-revealing IFT); per-warp od::contact_step_kernel below 4096 problems per GPU")
            if kind == "bundle":
                self.gb = od.GradientBundle(model, eta=od.workloads.bundle_perturbations(self.in_width, N=cfg["N"], eps=1e-4, seed=0))
                self.bundle = D.DeviceBundle(self.stepper, self.gb)
                self.out_width = self.in_width * self.dyn.nq
                self.kernel += " x (N+1) eval solves + od::bundle_fit_kernel"
        # inputs: the global batch is generated once (seed 0) and cut into contiguous shards
        full = gen_inputs(cfg, B_total, 0)
        self.x_host_full = full
        self.x_host = torch.from_numpy(np.ascontiguousarray(full[self.lo:self.hi])).pin_memory()
        self.x = self.x_host.to(dev)
        self.status = torch.empty((max(self.B, 1),), dtype=torch.int32, device=dev)[:self.B]
        if kind == "contact":
            self.out = torch.empty((self.B, self.out_width), dtype=torch.float64, device=dev)
            self.gathered = torch.empty((B_total, self.out_width), dtype=torch.float64, device=dev) if world > 1 else None
            if world > 1 and collective != "nccl":
                import torch.distributed as dist
                okf = torch.ones(1, device=dev)
                try:
                    self.fused = D.FusedGather(self.stepper, B_total, sync=("launch" if collective == "fused-launch-barrier" else "auto"),
                                               multicast=("auto" if collective == "fused" else False))
                except Exception as ex:                       # symmetric memory unavailable: every rank must take the same path
                    sys.stderr.write("rank %d: fused gather unavailable (%r), using ncclAllGather\n" % (rank, ex))
                    okf.zero_()
                dist.all_reduce(okf, op=dist.ReduceOp.MIN)
                if okf.item() == 0:
                    self.fused = None
            if world > 1:
                if self.fused is not None:
                    how = "one multimem.st per 16 B through the NVLink multicast alias (NVSwitch replicates)" if self.fused.multicast else "P2P stores into every rank's buffer over NVLink"
                    self.collective = ("all-gather fused into the kernel: each finished row leaves as %s; %s" % (
                        how, "cross-rank barrier fused as well (one system fence per block, flags published by the last block, device-side epoch)"
                        if self.fused.sync == "kernel" else "symmetric-memory barrier launch after the kernel"))
                    if cfg["model"] == "planarpush" and self.B >= 3072:
                        self.collective = ("persistent sweep into this rank's rows of its own gather buffer, one forwarding kernel stores them into every peer's "
                                           "buffer over NVLink (P2P), symmetric-memory barrier launch")
                else:
                    self.collective = "kernel + ncclAllGather of the packed rows"
        elif kind == "bundle":
            self.dz = torch.empty((B_total if world > 1 else self.B, self.in_width, self.dyn.nq), dtype=torch.float64, device=dev)
            self.bstatus = torch.empty((B_total if world > 1 else self.B,), dtype=torch.int32, device=dev)
            if world > 1:
                self.collective = "sample axis sharded: each rank solves a contiguous slice of the B x (N+1) eval solves, ncclAllGather of f_eta (nq doubles per solve), fit on every rank"
                self.x = torch.from_numpy(np.ascontiguousarray(full)).to(dev)       # every rank holds all nominal points; the SOLVES are sharded
        else:
            self.y = torch.empty((self.B, 12), dtype=torch.float64, device=dev)
            self.dx = torch.empty((self.B, 12, 12), dtype=torch.float64, device=dev)
            self.du = torch.empty((self.B, 3, 12), dtype=torch.float64, device=dev)
            if world > 1:
                self.collective = "kernel + 3 ncclAllGather (y, dx, du rows)"

    def make_pool(self, warmup):
        """P device copies of this rank's inputs, P·bytes ≥ 1.05 × L2, used round-robin: the rows a step reads were last touched a whole
        pool (> L2) of traffic ago, so they come from HBM, while the kernel's instructions stay cached as they would in a solver loop."""
        t = self.torch
        l2 = t.cuda.get_device_properties(self.dev).L2_cache_size
        per = max(self.x.numel() * 8, 1)
        P = min(max(2, int(1.05 * l2 // per) + 1 + warmup), 4096)
        self.pool = self.x.unsqueeze(0).repeat(P, 1, 1).contiguous()
        if self.cfg["kind"] == "rocket":                    # od_rocket_batch_device takes dense x and u arrays
            self.pool_x = self.pool[:, :, :12].contiguous(); self.pool_u = self.pool[:, :, 12:].contiguous()
        flush = t.empty(256 * 1024 * 1024, dtype=t.uint8, device=self.dev)
        flush.zero_()
        t.cuda.synchronize()
        del flush
        return P

    def step(self, k):
        """Device-resident step k (inputs = pool buffer k): kernel(s) + exchange on torch's current stream."""
        x = self.pool[k % self.pool.shape[0]]
        kind = self.cfg["kind"]
        if kind == "contact":
            if self.fused is not None:
                self.fused.step(x, self.status)
            else:
                self.stepper.step_grad_packed(x, self.out, self.status)
                if self.world > 1:
                    self.D.all_gather_rows(self.out, self.B_total, self.gathered)
        elif kind == "bundle":
            nq = self.dyn.nq
            self.bundle.gradient_batch(x[:, :nq], x[:, nq:2 * nq], x[:, 2 * nq:], self.dz, self.bstatus, sharded=self.world > 1)
        else:
            kk = k % self.pool.shape[0]
            self.rk.step(self.pool_x[kk], self.pool_u[kk], True, self.y, self.dx, self.du, self.status)
            if self.world > 1:
                self.D.all_gather_rows(self.y, self.B_total); self.D.all_gather_rows(self.dx.view(self.B, -1), self.B_total)
                self.D.all_gather_rows(self.du.view(self.B, -1), self.B_total)

    def graph_ok(self):
        # NCCL collectives and torch's symmetric-memory barrier are left out of graph capture; everything else is plain kernel launches
        if self.world == 1:
            return True
        return self.cfg["kind"] == "contact" and self.fused is not None and self.fused.sync == "kernel"

    def launches(self):
        if self.cfg["kind"] == "rocket":
            return self.info.launch_count()
        return self.dyn.launch_count()

    def converged(self):
        st = self.bstatus if self.cfg["kind"] == "bundle" else self.status
        return float((st == 0).float().mean().item()) if st.numel() else 1.0

    # ---- end to end through the public host API --------------------------------------------------------------------------------
    def e2e_setup(self):
        t, np = self.torch, self.np
        kind = self.cfg["kind"]
        if kind == "contact":
            if self.world == 1:
                self.e_out = t.empty((self.B, self.out_width), dtype=t.float64).pin_memory()
                self.e_st = t.empty((self.B,), dtype=t.int32).pin_memory()
                self.dyn2 = self.od.ImplicitDynamics(self.dyn.model, self.cfg["h"], r_tol=R_TOL, κ_eval_tol=self.cfg["ke"], κ_grad_tol=self.cfg["kg"], device=self.dev.index)
                self.e_np = (self.x_host.numpy(), self.e_out.numpy(), self.e_st.numpy())
                return ("ImplicitDynamics.step_grad_packed -> od_step_grad_packed (pinned host buffers; where the model's rows move coalesced the kernel reads its "
                        "input rows and writes its output rows in place over PCIe, else one H2D and one D2H copy)", self.in_width * 8 * self.B, (self.out_width * 8 + 4) * self.B)
            self.sh = self.D.ShardedHostSweep(self.stepper, self.B_total, collective="fused" if self.fused is not None else "nccl")
            self.e_out = t.empty((self.B_total, self.out_width), dtype=t.float64).pin_memory()
            return ("device.ShardedHostSweep.step: pinned host shard -> H2D -> kernel + %s -> D2H of ALL gathered rows to pinned host memory on every rank" % (
                "fused all-gather" if self.fused is not None else "ncclAllGather"), self.in_width * 8 * self.B, self.out_width * 8 * self.B_total)
        if kind == "bundle":
            nq = self.dyn.nq
            x = self.x_host_full if self.world > 1 else self.x_host.numpy()
            self.e_np = (np.ascontiguousarray(x[:, :nq]), np.ascontiguousarray(x[:, nq:2 * nq]), np.ascontiguousarray(x[:, 2 * nq:]))
            return ("gradient_batch -> od_bundle_batch (host arrays in, dz out)" + (" — every rank runs the whole sweep (replicas)" if self.world > 1 else ""),
                    self.in_width * 8 * len(x), (self.out_width * 8 + 4) * len(x))
        x = self.x_host.numpy()
        pin = lambda shape, dt: t.empty(shape, dtype=dt).pin_memory()        # noqa: E731
        self.e_pin = (t.from_numpy(np.ascontiguousarray(x[:, :12])).pin_memory(), t.from_numpy(np.ascontiguousarray(x[:, 12:])).pin_memory(),
                      pin((self.B, 12), t.float64), pin((self.B, 12, 12), t.float64), pin((self.B, 3, 12), t.float64), pin((self.B,), t.int32))
        self.e_np = tuple(a.numpy() for a in self.e_pin)
        return ("RocketInfo.step_batch -> od_rocket_batch (pinned host arrays in; y, dx, du, status out into pinned host arrays)", 15 * 8 * self.B, (192 * 8 + 4) * self.B)

    def e2e_step(self):
        kind = self.cfg["kind"]
        if kind == "contact":
            if self.world == 1:
                self.dyn2.step_grad_packed(*self.e_np)
            else:
                self.sh.step(self.x_host, self.e_out)
        elif kind == "bundle":
            self.od.gradient_batch(self.dyn, self.gb, *self.e_np)
        else:
            self.info.step_batch(self.e_np[0], self.e_np[1], proj=True, out=self.e_np[2:])


def single_call_latency(sw, reps=200):
    """BASELINE configs[0] is ONE rollout on the CPU: the reference's outer solver calls f / fx / fu once per timestep
    (examples/acrobot.jl:92,113).  Latency of exactly those calls through the host API (one problem per call = launch + synchronise)
    next to the CPU restatement on one thread, and the whole T-step rollout as one launch (od_rollout_batch) next to T−1 sequential f calls."""
    import numpy as np
    od, cfg = sw.od, sw.cfg
    from oracle import oracle as O
    name = ORACLE_NAME[cfg["model"]]
    nq, nu = sw.dyn.nq, sw.dyn.nu
    row = sw.x_host_full[0]
    x = np.ascontiguousarray(row[:2 * nq]); u0 = np.ascontiguousarray(row[2 * nq:]); w = np.zeros(0)
    d = np.zeros(2 * nq); dx = np.zeros((2 * nq, 2 * nq)); du = np.zeros((2 * nq, nu))
    us = [u0 + 1e-9 * k for k in range(reps)]                      # a new control every call: no memoised gradient
    for k in range(5):
        od.f(d, sw.dyn, x, us[k], w); od.fx(dx, sw.dyn, x, us[k], w); od.fu(du, sw.dyn, x, us[k], w)
    t0 = time.perf_counter()
    for k in range(reps):
        od.f(d, sw.dyn, x, us[k], w)
    t_f = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for k in range(reps):
        od.fx(dx, sw.dyn, x, us[k], w); od.fu(du, sw.dyn, x, us[k], w)
    t_g = (time.perf_counter() - t0) / reps
    q1, q2 = x[None, :nq], x[None, nq:]
    t0 = time.perf_counter()
    for k in range(reps):
        O.step_batch(name, q1, q2, us[k][None], cfg["h"], cfg["ke"], False, fric=cfg["fric"], r_tol=R_TOL, nthreads=1, diagnostics=False)
    c_f = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for k in range(reps):
        O.step_batch(name, q1, q2, us[k][None], cfg["h"], cfg["kg"], True, fric=cfg["fric"], r_tol=R_TOL, nthreads=1, diagnostics=False)   # fx
        O.step_batch(name, q1, q2, us[k][None], cfg["h"], cfg["kg"], True, fric=cfg["fric"], r_tol=R_TOL, nthreads=1, diagnostics=False)   # fu re-solves
    c_g = (time.perf_counter() - t0) / reps
    T = 51
    ubar = np.tile(u0, (T - 1, 1)) * 0.1
    od.rollout(sw.dyn, x, ubar)
    t0 = time.perf_counter()
    for _ in range(20):
        od.rollout(sw.dyn, x, ubar)
    t_roll = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for _ in range(5):
        xx = x.copy()
        for t in range(T - 1):
            od.f(d, sw.dyn, xx, ubar[t], w); xx = d.copy()
    t_seq = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    for _ in range(5):
        O.rollout_batch(name, x[None], ubar, cfg["h"], cfg["ke"], fric=cfg["fric"], r_tol=R_TOL)
    c_roll = (time.perf_counter() - t0) / 5
    return {"unit": "us per call", "f_gpu_host_api": 1e6 * t_f, "fx_plus_fu_gpu_host_api": 1e6 * t_g, "f_cpu_oracle_1_thread": 1e6 * c_f,
            "fx_plus_fu_cpu_oracle_1_thread_reference_pattern": 1e6 * c_g,
            "rollout_T51_one_launch_us": 1e6 * t_roll, "rollout_T51_as_50_sequential_f_calls_us": 1e6 * t_seq, "rollout_T51_cpu_oracle_us": 1e6 * c_roll,
            "note": "one problem per call = one kernel launch + synchronise: the GPU loses to one CPU core on a single tiny solve; the batched entry points "
                    "(all timesteps of the sweep, all line-search candidates of the forward pass) are what the device path is for"}


def timed_region(sweep, steps, barrier, use_graph):
    """K steps between two CUDA events, barrier + synchronize on both sides.  Returns (milliseconds, 'graph' | 'eager')."""
    torch = sweep.torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if use_graph:
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for k in range(steps):
                    sweep.step(k)
            barrier()
            g.replay()                                  # untimed: uploads the graph, warms the replay path
            barrier()
            e0.record(); g.replay(); e1.record()
            barrier()
            return e0.elapsed_time(e1), "graph"
        except Exception as ex:
            sys.stderr.write("CUDA graph capture unavailable (%r): timing an eager loop\n" % (ex,))
            torch.cuda.synchronize()
    barrier()
    e0.record()
    for k in range(steps):
        sweep.step(k)
    e1.record()
    barrier()
    return e0.elapsed_time(e1), "eager"


def measure(name, cfg, B_total, world, rank, dev, args, barrier, collective, clocks=None, e2e=True):
    """Device-timed value + e2e of one (config, batch, scaling) on all ranks; returns a dict (identical on every rank after the MAX)."""
    import torch
    import torch.distributed as dist
    sw = Sweep(name, cfg, B_total, world, rank, dev, collective)
    P = sw.make_pool(args.warmup)
    for w in range(args.warmup):
        sw.step(P - 1 - w)
    barrier()
    n0 = sw.launches()
    if clocks is not None:
        clocks.run()
    t_wall0 = time.perf_counter()
    ms, mode = timed_region(sw, args.steps, barrier, use_graph=(not args.no_graph) and sw.graph_ok())
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop() if clocks is not None else None
    per_step_launches = (sw.launches() - n0) // args.steps
    # graph mode: the launches are counted once, at capture; the timed replay runs exactly those
    launches_in_timed_region = (sw.launches() - n0)
    ok = sw.converged()
    res = {"ms": ms, "mode": mode, "launches": launches_in_timed_region, "per_step_launches": per_step_launches, "ok": ok, "pool": P, "clk": clk,
           "wall_s": t_wall, "sweep": sw}
    if e2e:
        api, h2d, d2h = sw.e2e_setup()
        for _ in range(args.warmup):
            sw.e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps or args.steps):
            sw.e2e_step()
        torch.cuda.synchronize()
        res["e2e_s"] = (time.perf_counter() - t0) / (args.e2e_steps or args.steps)
        res["e2e_api"], res["h2d"], res["d2h"] = api, h2d, d2h
    t = torch.tensor([res["ms"], res.get("e2e_s", 0.0), -ok], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["ms"], res["e2e_s"], res["ok"] = float(t[0]), float(t[1]), -float(t[2])
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--config", default="hopper", choices=list(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="global batch (strong scaling) / batch per GPU (weak scaling); 0 = the config's own (hopper: 4096)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N>1: strong = the metric's global batch split over the ranks (default: the metric is quoted on batch 4096); weak = that batch per rank")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time an eager launch loop instead of one CUDA graph of the K steps")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--extra", action="store_true", help="N=1: also report a saturating batch (262144) in the JSON line (always on for N>1)")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--collective", default="fused", choices=["fused", "fused-p2p", "fused-launch-barrier", "nccl"],
                    help="N>1 (contact configs): 'fused' = all-gather (NVLink multicast stores when available) and cross-rank barrier fused into the kernel; "
                         "'fused-p2p' = the same with per-peer stores; 'fused-launch-barrier' = fused rows + separate barrier launch; 'nccl' = kernel + ncclAllGather")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 200 and args.warmup == 10:
            args.steps, args.warmup = 20, 3
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    # Exactly ONE line on stdout: everything any library writes to file descriptor 1 during the run (NCCL's version banner at
    # NCCL_DEBUG=VERSION / WARN, torch warnings) is sent to stderr; the JSON line goes to the saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

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
            os.environ["NCCL_DEBUG"] = "WARN"          # the GPU boxes export NCCL_DEBUG=VERSION, whose banner goes to stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    name, cfg = args.config, CONFIGS[args.config]
    base = args.batch or cfg["batch"]
    B_total = base * (world if args.scaling == "weak" else 1)
    coll = args.collective
    clocks = ClockSampler(local)
    main_res = measure(name, cfg, B_total, world, rank, dev, args, barrier, coll, clocks=clocks, e2e=True)
    sw = main_res["sweep"]

    # bit-exactness of the exchange: the rows every rank now holds must equal a plain ncclAllGather of the per-rank results
    gather_check = None
    if world > 1 and cfg["kind"] == "contact":
        sw.stepper.step_grad_packed(sw.x, sw.out, sw.status)
        ref = sw.D.all_gather_rows(sw.out, B_total)
        if sw.fused is not None:
            got, _ = sw.fused.step(sw.x, sw.status)
            barrier()
            gather_check = bool(torch.equal(got, ref))
            assert gather_check, "fused gather differs from ncclAllGather"
        assert torch.equal(ref[sw.lo:sw.hi], sw.out)

    extra = {}
    want_extra = (world > 1 or args.extra) and not args.no_extra and name == "hopper"
    if want_extra:
        sub = argparse.Namespace(**vars(args)); sub.steps = max(20, min(args.steps, 50)); sub.e2e_steps = 5
        if world > 1:
            other = "weak" if args.scaling == "strong" else "strong"
            Bo = base * (world if other == "weak" else 1)
            r = measure(name, cfg, Bo, world, rank, dev, sub, barrier, coll, e2e=False)
            extra[other + "_scaling"] = {"global_batch": Bo, "batch_per_gpu": Bo // world, "ms_per_step": r["ms"] / sub.steps, "value": Bo * sub.steps / (r["ms"] * 1e-3),
                                         "unit": UNIT, "timed": r["mode"], "steps": sub.steps}
            del r
        sub2 = argparse.Namespace(**vars(sub)); sub2.steps = 10; sub2.warmup = 3
        Bs = 262144 * world
        r = measure(name, cfg, Bs, world, rank, dev, sub2, barrier, coll, e2e=False)
        extra["saturating_batch"] = {"global_batch": Bs, "batch_per_gpu": 262144, "ms_per_step": r["ms"] / sub2.steps, "value": Bs * sub2.steps / (r["ms"] * 1e-3),
                                     "unit": UNIT, "timed": r["mode"], "steps": sub2.steps}
        del r

    # every collective is behind us: the other ranks leave now, so that no GPU spins in an NCCL barrier while rank 0 times the CPU baseline
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        peak, peak_src = measured_peak()
        ms_per_step = main_res["ms"] / args.steps
        value = B_total * args.steps / (main_res["ms"] * 1e-3)
        bytes_unit = cfg["bytes_in"] + cfg["bytes_out"]
        units_per_launch = sw.B * (cfg.get("N", 0) + 1 if cfg["kind"] == "bundle" and world == 1 else 1)
        if cfg["kind"] == "bundle":
            # the dominant kernel runs (N+1) eval solves per problem; its algorithmic I/O per SOLVE is the input row + q3
            bytes_unit = cfg["bytes_in"] + 8 * sw.dyn.nq
            units_per_launch = (sw.B_total * (cfg["N"] + 1)) // world
        achieved = bytes_unit * units_per_launch / (ms_per_step * 1e-3) / 1e9
        summ = ncu_summary(name)
        line = {
            "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg, name, B_total, world, args.scaling),
            "details": {"batch_per_gpu": sw.B, "timed_region": ("one CUDA graph holding the %d steps, replayed once untimed then once between two CUDA events" % args.steps)
                        if main_res["mode"] == "graph" else "eager launch loop between two CUDA events",
                        "input_pool_buffers": main_res["pool"], "collective": sw.collective, "gather_check_bitwise_equal_to_nccl": gather_check,
                        "converged_fraction": main_res["ok"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": summ.get("dram_bytes_per_launch"),
                         "kernel": sw.kernel, "kernel_ms": ms_per_step, "algorithmic_bytes_per_launch": bytes_unit * units_per_launch,
                         "compute_side": compute_side(name), "peak_source": peak_src,
                         "note": "kernel_ms = timed region / steps (the region holds only this kernel%s); %d B vs ~1e5 fp64 flop per unit: "
                                 "fp64-latency bound by construction (DESIGN.md §6)" % (" and its fused exchange" if world > 1 else "", bytes_unit)},
            "e2e": {"value": B_total / main_res["e2e_s"], "unit": UNIT, "h2d_bytes_per_step": main_res["h2d"], "d2h_bytes_per_step": main_res["d2h"],
                    "ms_per_step": 1e3 * main_res["e2e_s"], "api": main_res["e2e_api"]},
            "gpu_launches": int(main_res["launches"]) * world,
            "clocks": main_res["clk"],
        }
        if name == "acrobot" and world == 1 and not args.no_cpu_baseline:
            extra["single_call_latency"] = single_call_latency(sw)
        if extra:
            line["extra"] = extra
        if not args.no_cpu_baseline and world == 1:          # (the CPU baseline is reported on rank 0 at N = 1 only)
            line["cpu_baseline"] = cpu_baseline(cfg, min(B_total, 4096), seconds=args.cpu_seconds, pattern="D")
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
