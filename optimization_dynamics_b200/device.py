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
    """All-gather fused into the step kernel over NVLink peer memory (torch symmetric memory supplies the peer-mapped buffers).

    Every rank owns [B_total, out_width] gather buffers; the kernel stores each finished 352-B row (hopper) into all of them, so
    the exchange overlaps the compute of the problems still iterating.  `sync="kernel"` (default up to 4 ranks) also fuses the cross-rank
    barrier into the kernel (flags in symmetric memory, published by the last block of each rank; two gather buffers alternate so
    that a fast rank cannot overwrite rows a slower rank's consumer is still reading); `sync="launch"` runs torch's
    symmetric-memory barrier as a separate launch after the kernel."""

    def __init__(self, stepper: DeviceStepper, B_total, group=None, sync="auto"):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if sync == "auto":
            # measured on 8×B200 (hopper, 4096 problems per rank, ms per step): kernel-fused barrier 0.114 / 0.124 / 0.131 at
            # N = 2 / 4 / 8, separate barrier launch 0.117 / 0.127 / 0.127 — the per-warp system fence over 7 peers' outstanding
            # stores costs more than a launch at N = 8
            sync = "kernel" if self.world <= 4 else "launch"
        self.stepper, self.B_total, self.sync = stepper, B_total, sync
        if self.world > 8:
            raise RuntimeError("FusedGather: single node, at most 8 ranks")
        dev = torch.device("cuda", torch.cuda.current_device())
        grp = dist.group.WORLD if group is None else group
        self.bufs, self.handles, self.ptrs = [], [], []
        for _ in range(2 if sync == "kernel" else 1):
            buf = symm_mem.empty((B_total, stepper.out_width), dtype=torch.float64, device=dev)
            h = symm_mem.rendezvous(buf, grp)
            self.bufs.append(buf); self.handles.append(h)
            self.ptrs.append((C.c_uint64 * self.world)(*[int(p) for p in h.buffer_ptrs]))
        self.buf, self.handle = self.bufs[0], self.handles[0]
        self.row0 = shard_range(B_total, self.rank, self.world)[0]
        self.epoch = 0
        if sync == "kernel":
            self.flags = symm_mem.empty((16,), dtype=torch.int64, device=dev)
            self.flags.zero_()
            self.flag_handle = symm_mem.rendezvous(self.flags, grp)
            self.flag_ptrs = (C.c_uint64 * self.world)(*[int(p) for p in self.flag_handle.buffer_ptrs])
            self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
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
            status = t.empty((B,), dtype=t.int32, device=xin_local.device)
        self.stepper._bind_stream()
        L = _lib.lib()
        hd = self.stepper.dyn._handle()
        pit = C.c_void_p(iters.data_ptr()) if iters is not None else None
        if self.sync == "kernel":
            self.epoch += 1
            which = self.epoch % 2
            self.buf = self.bufs[which]
            _lib.check(L.od_step_grad_packed_gather_sync_device(hd, B, C.c_void_p(xin_local.data_ptr()), self.row0, self.world, self.rank,
                                                                self.ptrs[which], self.flag_ptrs, C.c_void_p(self.counter.data_ptr()),
                                                                self.epoch, C.c_void_p(status.data_ptr()), pit))
        else:
            _lib.check(L.od_step_grad_packed_gather_device(hd, B, C.c_void_p(xin_local.data_ptr()), self.row0, self.world, self.rank, self.ptrs[0],
                                                           C.c_void_p(status.data_ptr()), pit))
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
