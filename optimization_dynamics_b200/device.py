"""Device-resident batched stepping (torch tensors are only the memory/stream plumbing) and single-node multi-GPU sharding.

The derivative sweep of iLQR evaluates fx/fu at every timestep independently; the Riccati backward pass then
needs all of them.  Each rank solves a contiguous slice of the batch and ONE NCCL all-gather of the packed output rows
(hopper: 44 doubles = 352 B per problem) gives every rank the full set — there is no other data-path communication.
"""
import ctypes as C

import numpy as np

from . import _lib
from .dynamics import ImplicitDynamics


def shard_range(B, rank, world):
    """Contiguous slice [lo, hi) of a batch of B problems owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(B, world):
    return [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]


def all_gather_rows(local, B_total, gathered=None):
    """All-gather contiguous row shards (shard_range) of a [B_total, W] matrix; one collective.  Ragged shards (B_total not a
    multiple of the world size) are padded by one row to the largest shard and compacted after the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    sizes = shard_sizes(B_total, world)
    W = local.shape[1]
    if gathered is None:
        gathered = torch.empty((B_total, W), dtype=local.dtype, device=local.device)
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(gathered, local)
        return gathered
    m = max(sizes)
    padded = torch.zeros((m, W), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    buf = torch.empty((world * m, W), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded)
    lo = 0
    for r, n in enumerate(sizes):
        gathered[lo:lo + n] = buf[r * m:r * m + n]
        lo += n
    return gathered


class DeviceStepper:
    """Runs the packed step+gradient kernel on torch CUDA tensors, on torch's current stream."""

    def __init__(self, im_dyn: ImplicitDynamics):
        import torch
        self.torch = torch
        self.dyn = im_dyn
        self.nq, self.nu = im_dyn.nq, im_dyn.nu
        self.in_width = 2 * self.nq + self.nu
        self.out_width = self.nq + self.nq * self.in_width

    def _bind_stream(self):
        s = self.torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().od_set_stream(self.dyn._handle(), C.c_void_p(s)))

    def step_grad_packed(self, xin, out=None, status=None, iters=None, want_eval=True, want_grad=True):
        """xin: [B, 2nq+nu] float64 CUDA tensor → out [B, nq + nq(2nq+nu)], status [B] int32.  Asynchronous."""
        t = self.torch
        assert xin.is_cuda and xin.dtype == t.float64 and xin.is_contiguous() and xin.shape[1] == self.in_width
        B = xin.shape[0]
        if out is None:
            out = t.empty((B, self.out_width), dtype=t.float64, device=xin.device)
        if status is None:
            status = t.empty((B,), dtype=t.int32, device=xin.device)
        self._bind_stream()
        _lib.check(_lib.lib().od_step_grad_packed_device(self.dyn._handle(), B, C.c_void_p(xin.data_ptr()), C.c_void_p(out.data_ptr()),
                                                         C.c_void_p(status.data_ptr()), C.c_void_p(iters.data_ptr()) if iters is not None else None,
                                                         int(want_eval), int(want_grad)))
        return out, status

    def step_grad_sharded(self, xin_local, B_total, gathered=None):
        """Each rank passes its slice (shard_range) of the packed inputs; returns the all-gathered [B_total, out_width] outputs."""
        out_local, status = self.step_grad_packed(xin_local)
        return all_gather_rows(out_local, B_total, gathered), status


class DeviceSolverStages:
    """The two stages either side of the derivative sweep, on torch CUDA tensors and torch's current stream (asynchronous):
    `backward_pass` (od_riccati_batch_device) and `rollouts` (od_rollout_batch_device).  With DeviceStepper.step_grad_packed in
    between, one iLQR iteration for NT trajectories is three launches with no host round trip."""

    def __init__(self, stepper: DeviceStepper):
        self.stepper = stepper
        self.torch = stepper.torch
        self.nq, self.nu = stepper.nq, stepper.nu

    def _p(self, t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def backward_pass(self, jac, lx, lu, lxx, luu, lux=None, reg=0.0, K=None, k=None, dV=None, status=None):
        t = self.torch
        NT, S = jac.shape[0], jac.shape[1]
        n, m = 2 * self.nq, self.nu
        for a in (jac, lx, lu, lxx, luu) + (() if lux is None else (lux,)):
            assert a.is_cuda and a.dtype == t.float64 and a.is_contiguous()
        K = t.empty((NT, S, m, n), dtype=t.float64, device=jac.device) if K is None else K
        k = t.empty((NT, S, m), dtype=t.float64, device=jac.device) if k is None else k
        dV = t.empty((NT, 2), dtype=t.float64, device=jac.device) if dV is None else dV
        status = t.empty((NT,), dtype=t.int32, device=jac.device) if status is None else status
        self.stepper._bind_stream()
        _lib.check(_lib.lib().od_riccati_batch_device(self.stepper.dyn._handle(), NT, S + 1, self._p(jac), self._p(lx), self._p(lu), self._p(lxx),
                                                      self._p(luu), self._p(lux), float(reg), self._p(K), self._p(k), self._p(dV), self._p(status)))
        return K, k, dV, status

    def rollouts(self, x1, ubar, xbar=None, K=None, k=None, alpha=None, X=None, U=None, status=None):
        """x1 [R, 2nq]; ubar [T-1, nu] (shared) or [R, T-1, nu]; xbar [T, 2nq], K [T-1, nu, 2nq], k [T-1, nu], alpha [R]."""
        t = self.torch
        R = x1.shape[0]
        per = ubar.dim() == 3
        S = ubar.shape[-2]
        n, m = 2 * self.nq, self.nu
        X = t.empty((R, S + 1, n), dtype=t.float64, device=x1.device) if X is None else X
        U = t.empty((R, S, m), dtype=t.float64, device=x1.device) if U is None else U
        status = t.empty((R, S), dtype=t.int32, device=x1.device) if status is None else status
        self.stepper._bind_stream()
        _lib.check(_lib.lib().od_rollout_batch_device(self.stepper.dyn._handle(), R, S + 1, self._p(x1), self._p(ubar), S * m if per else 0, self._p(xbar),
                                                      self._p(K), self._p(k), self._p(alpha), self._p(X), self._p(U), self._p(status), None))
        return X, U, status


class FusedGather:
    """All-gather fused into the step kernel over NVLink (torch symmetric memory supplies the peer-mapped and multicast mappings).

    Every rank owns [B_total, out_width] gather buffers; the kernel sends each finished 352-B row (hopper) to all of them while
    the other problems are still iterating — as ONE `multimem.st` per 16 bytes through the buffers' NVLink multicast alias
    (NVSwitch replicates; `multicast="auto"` uses it when the symmetric-memory handle has one), else as world−1 peer stores.
    `sync="kernel"` (default) also fuses the cross-rank barrier into the kernel: one system-scope fence per block, flags in
    symmetric memory published by the last block of each rank; the epoch lives on the device, so the launches can be captured in a
    CUDA graph and replayed; two gather buffers alternate so that a fast rank cannot overwrite rows a slower rank's consumer is
    still reading.  `sync="launch"` runs torch's symmetric-memory barrier as a separate launch after the kernel.
    Ragged shards (B_total not a multiple of the world size, or smaller than it) are fine: an empty shard still takes part in
    the barrier."""

    def __init__(self, stepper: DeviceStepper, B_total, group=None, sync="auto", multicast="auto"):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if sync == "auto":
            sync = "kernel"
        self.stepper, self.B_total, self.sync = stepper, B_total, sync
        if self.world > 8:
            raise RuntimeError("FusedGather: single node, at most 8 ranks")
        dev = torch.device("cuda", torch.cuda.current_device())
        grp = dist.group.WORLD if group is None else group
        self.bufs, self.handles, self.ptrs, self.mc = [], [], [], []
        reg_path = stepper.dyn.model.name in ("hopper", "cartpole_friction", "acrobot_impact")
        if sync == "kernel" and not reg_path:
            sync = self.sync = "launch"                  # the in-kernel barrier needs rows that leave the kernel coalesced
        for _ in range(2 if sync == "kernel" else 1):
            buf = symm_mem.empty((max(B_total, 1), stepper.out_width), dtype=torch.float64, device=dev)
            h = symm_mem.rendezvous(buf, grp)
            self.bufs.append(buf[:B_total]); self.handles.append(h)
            self.ptrs.append((C.c_uint64 * self.world)(*[int(p) for p in h.buffer_ptrs]))
            mc = 0
            if multicast in ("auto", True) and reg_path:
                try:
                    mc = int(h.multicast_ptr or 0)          # 0 when the fabric / driver has no multicast object for this allocation
                except Exception:
                    mc = 0
            self.mc.append(mc)
        if multicast is True and not all(self.mc):
            raise RuntimeError("FusedGather: NVLink multicast was requested but the symmetric-memory handle has no multicast pointer")
        self.multicast = bool(all(self.mc)) and multicast in ("auto", True)
        self.buf, self.handle = self.bufs[0], self.handles[0]
        self.row0 = shard_range(B_total, self.rank, self.world)[0]
        self.launches = 0
        if sync == "kernel":
            self.flags = symm_mem.empty((16,), dtype=torch.int64, device=dev)
            self.flags.zero_()
            self.flag_handle = symm_mem.rendezvous(self.flags, grp)
            self.flag_ptrs = (C.c_uint64 * self.world)(*[int(p) for p in self.flag_handle.buffer_ptrs])
            self.mc_flags = 0
            if self.multicast:
                try:
                    self.mc_flags = int(self.flag_handle.multicast_ptr or 0)
                except Exception:
                    self.mc_flags = 0
            self.counter = torch.zeros(4, dtype=torch.int64, device=dev)       # [0]: block counter (u32), [1]: epoch (u64), device side
            torch.cuda.synchronize()
            self.flag_handle.barrier(channel=0)          # every rank's flags are zero before anyone publishes epoch 1
            torch.cuda.synchronize()

    def step(self, xin_local, status=None, iters=None):
        """Asynchronous on torch's current stream; returns the gathered [B_total, out_width] tensor (complete when the work this
        call enqueues has finished) and the local status."""
        status = self.launch(xin_local, status, iters)
        self.barrier()
        return self.buf, status

    def barrier(self):
        if self.sync != "kernel":
            self.handle.barrier(channel=0)

    def launch(self, xin_local, status=None, iters=None):
        t = self.stepper.torch
        B = xin_local.shape[0]
        if status is None:
            status = t.empty((max(B, 1),), dtype=t.int32, device=xin_local.device)[:B]
        self.stepper._bind_stream()
        which = 0
        g = _lib.od_gather_desc()
        g.world, g.rank, g.row0 = self.world, self.rank, self.row0
        if self.sync == "kernel":
            self.launches += 1
            which = self.launches % 2
            g.flag_buffers = self.flag_ptrs
            g.block_counter = self.counter.data_ptr()
            g.epoch_dev = self.counter.data_ptr() + 8
            g.epoch = 0
            g.multicast_flags = self.mc_flags
        self.buf = self.bufs[which]
        g.gather_buffers = self.ptrs[which]
        g.multicast_buffer = self.mc[which] if self.multicast else 0
        _lib.check(_lib.lib().od_step_grad_packed_gather_ex_device(self.stepper.dyn._handle(), B, C.c_void_p(xin_local.data_ptr()), C.byref(g),
                                                                   C.c_void_p(status.data_ptr()), C.c_void_p(iters.data_ptr()) if iters is not None else None))
        return status


def unpack_outputs(out, nq, nu):
    """Split packed rows into q3 [B,nq] and Jacobians [B,nq,nq], [B,nq,nq], [B,nq,nu] (row = q3 component).  numpy or torch."""
    B = out.shape[0]
    o = nq
    q3 = out[:, :nq]
    d1 = out[:, o:o + nq * nq].reshape(B, nq, nq)
    d2 = out[:, o + nq * nq:o + 2 * nq * nq].reshape(B, nq, nq)
    du = out[:, o + 2 * nq * nq:].reshape(B, nu, nq)
    tr = (lambda a: a.transpose(0, 2, 1)) if isinstance(out, np.ndarray) else (lambda a: a.transpose(1, 2))
    return q3, tr(d1), tr(d2), tr(du)


class DeviceBundle:
    """Gradient bundle (reference src/gradient_bundle.jl:87-104 + src/ls.jl:44-60) on torch CUDA tensors: the (N+1)·B eval-sim steps
    in one launch and the closed-form fit in a second.  With torch.distributed initialised, `gradient_batch(..., sharded=True)` cuts
    the flattened (problem × perturbation) axis into contiguous slices, one per rank (shard the sample axis), all-gathers
    fη (nq doubles per solve) and fits on every rank — bit-identical to the single-GPU result."""

    def __init__(self, stepper: DeviceStepper, gb):
        t = stepper.torch
        self.stepper, self.torch = stepper, t
        self.nq, self.nu = stepper.nq, stepper.nu
        self.ncol, self.N = 2 * self.nq + self.nu, gb.N
        eta = np.ascontiguousarray(gb.eta, dtype=np.float64)
        hinv = np.zeros((self.ncol, self.ncol))
        _lib.check(_lib.lib().od_bundle_prepare(self.ncol, self.N, eta.ctypes.data_as(_lib.c_double_p), hinv.ctypes.data_as(_lib.c_double_p)))
        dev = t.device("cuda", t.cuda.current_device())
        self.eta = t.from_numpy(eta).to(dev)
        self.hinv = t.from_numpy(hinv).to(dev)

    def gradient_batch(self, q1, q2, u, dz=None, status=None, sharded=False):
        """q1, q2 [B,nq], u [B,nu] CUDA tensors (row-strided views of packed [q1 | q2 | u] rows are fine) → dz [B, 2nq+nu, nq]
        (column-major blocks, as the C ABI), status [B]."""
        t = self.torch
        B = q1.shape[0]
        assert q1.stride(1) == 1 and q2.stride(1) == 1 and u.stride(1) == 1 and q1.stride(0) == q2.stride(0)
        P = B * (self.N + 1)
        feta = t.empty((P, self.nq), dtype=t.float64, device=q1.device)
        stw = t.empty((P,), dtype=t.int32, device=q1.device)
        dz = t.empty((B, self.ncol, self.nq), dtype=t.float64, device=q1.device) if dz is None else dz
        status = t.empty((B,), dtype=t.int32, device=q1.device) if status is None else status
        self.stepper._bind_stream()
        L, hd = _lib.lib(), self.stepper.dyn._handle()
        p = lambda a: C.c_void_p(a.data_ptr())      # noqa: E731
        lo, hi = 0, P
        if sharded:
            import torch.distributed as dist
            lo, hi = shard_range(P, dist.get_rank(), dist.get_world_size())
        _lib.check(L.od_bundle_solve_device(hd, B, self.N, p(self.eta), p(q1), p(q2), p(u), q1.stride(0), u.stride(0), lo, hi - lo, p(feta), p(stw)))
        if sharded:
            feta = all_gather_rows(feta[lo:hi], P)
            stw = all_gather_rows(stw[lo:hi].view(-1, 1), P).view(-1)
        _lib.check(L.od_bundle_fit_device(hd, B, self.N, p(self.eta), p(self.hinv), p(feta), p(stw), p(dz), p(status)))
        return dz, status


class DeviceRocket:
    """f / fx / fu_rocket[_proj] (reference src/models/rocket/dynamics.jl:101-269) on torch CUDA tensors, one launch per batch; with
    torch.distributed, `step_sharded` solves a contiguous slice per rank and all-gathers y, dx, du (one NCCL all-gather each: the
    rows are 12 / 144 / 36 doubles)."""

    def __init__(self, info):
        import torch
        self.torch, self.info = torch, info

    def step(self, x, u, proj=True, y=None, dx=None, du=None, status=None):
        t = self.torch
        B = x.shape[0]
        y = t.empty((B, 12), dtype=t.float64, device=x.device) if y is None else y
        dx = t.empty((B, 12, 12), dtype=t.float64, device=x.device) if dx is None else dx
        du = t.empty((B, 3, 12), dtype=t.float64, device=x.device) if du is None else du
        status = t.empty((B,), dtype=t.int32, device=x.device) if status is None else status
        L = _lib.lib()
        _lib.check(L.od_set_stream(self.info._hd, C.c_void_p(t.cuda.current_stream().cuda_stream)))
        p = lambda a: C.c_void_p(a.data_ptr())      # noqa: E731
        _lib.check(L.od_rocket_batch_device(self.info._hd, B, p(x), p(u), int(bool(proj)), p(y), p(dx), p(du), p(status), None))
        return y, dx, du, status

    def step_sharded(self, x_local, u_local, B_total, proj=True):
        y, dx, du, st = self.step(x_local, u_local, proj)
        B = x_local.shape[0]
        return (all_gather_rows(y, B_total), all_gather_rows(dx.view(B, -1), B_total).view(B_total, 12, 12),
                all_gather_rows(du.view(B, -1), B_total).view(B_total, 3, 12), st)


class ShardedHostSweep:
    """Host-facing multi-GPU derivative sweep: each rank hands in its pinned host shard of the packed inputs and receives ALL packed
    output rows in pinned host memory — host in → H2D → kernel with the fused all-gather (or kernel + ncclAllGather) → D2H of the
    gathered rows — what a host-side Riccati pass on every rank consumes."""

    def __init__(self, stepper: DeviceStepper, B_total, collective="fused"):
        import torch
        import torch.distributed as dist
        self.torch, self.stepper, self.B_total = torch, stepper, B_total
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        lo, hi = shard_range(B_total, self.rank, self.world)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.xin = torch.empty((hi - lo, stepper.in_width), dtype=torch.float64, device=dev)
        self.status = torch.empty((hi - lo,), dtype=torch.int32, device=dev)
        self.fused = FusedGather(stepper, B_total) if (self.world > 1 and collective == "fused") else None
        self.out = None if self.fused is not None else torch.empty((hi - lo, stepper.out_width), dtype=torch.float64, device=dev)
        self.gathered = None if self.fused is not None or self.world == 1 else torch.empty((B_total, stepper.out_width), dtype=torch.float64, device=dev)

    def step(self, xin_host_local, out_host_all, status_host_local=None):
        t = self.torch
        self.xin.copy_(xin_host_local, non_blocking=True)
        if self.fused is not None:
            buf, _ = self.fused.step(self.xin, self.status)
        else:
            self.stepper.step_grad_packed(self.xin, self.out, self.status)
            buf = self.out if self.world == 1 else all_gather_rows(self.out, self.B_total, self.gathered)
        out_host_all.copy_(buf, non_blocking=True)
        if status_host_local is not None:
            status_host_local.copy_(self.status, non_blocking=True)
        t.cuda.current_stream().synchronize()
        return out_host_all
