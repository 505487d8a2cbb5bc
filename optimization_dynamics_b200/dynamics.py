"""Host-side mirror of the reference's implicit-dynamics API over the C ABI.

Same names, argument order and semantics as the reference's Julia code:
    ImplicitDynamics(model, h, r, rz, rθ; T, r_tol, κ_eval_tol, κ_grad_tol, no_impact, no_friction, n, m, d, nc, nb, info)
                                                             reference src/dynamics.jl:51-79
    f(d, model, x, u, w) / fx(dx, …) / fu(du, …)             reference src/dynamics.jl:81-128
    state_to_configuration(x)                                reference src/dynamics.jl:131-145
plus the batched calls the CUDA path is built for (`step_grad_batch`): all timesteps × samples × rollouts in one launch.
The three generated-function arguments of the reference constructor (r, rz, rθ) are accepted and ignored: the residual code
is compiled into the library (csrc/gen/).  There is no CPU fallback — construction raises without a CUDA device.
"""
import ctypes as C

import numpy as np

from . import _lib

MODEL_IDS = {"acrobot_impact": 0, "acrobot_nominal": 1, "cartpole_friction": 2, "cartpole_frictionless": 3, "planar_push": 4,
             "hopper": 5, "rocket": 6}


class Model:
    """Model singleton (reference: `acrobot_impact`, `cartpole_friction`, `planarpush`, `RoboDojo.hopper`, `rocket`, …)."""

    def __init__(self, name, nq, nu, nw, nc, friction=None, **consts):
        self.name, self.nq, self.nu, self.nw, self.nc = name, nq, nu, nw, nc
        self.friction = None if friction is None else np.array(friction, dtype=np.float64)   # mutable, like model.friction .= μ
        for k, v in consts.items():
            setattr(self, k, v)

    def __repr__(self):
        return "Model(%s, nq=%d, nu=%d)" % (self.name, self.nq, self.nu)


# reference src/models/acrobot/model.jl:159-163, cartpole/model.jl:131-132, planar_push/model.jl:196-200, rocket/model.jl:43-48
acrobot_impact = Model("acrobot_impact", 2, 1, 0, 2)
acrobot_nominal = Model("acrobot_nominal", 2, 1, 0, 0)
cartpole_friction = Model("cartpole_friction", 2, 1, 0, 2, friction=[0.1, 0.1], mc=1.0, mp=0.2, l=0.5, g=9.81)
cartpole_frictionless = Model("cartpole_frictionless", 2, 1, 0, 2, mc=1.0, mp=0.2, l=0.5, g=9.81)
planarpush = Model("planar_push", 5, 2, 0, 5)
hopper = Model("hopper", 4, 2, 0, 4, friction=[0.5, 0.5], mass_body=3.0, mass_foot=1.0, inertia_body=0.75, gravity=9.81,
               body_radius=0.1, foot_radius=0.05, leg_len_max=1.0, leg_len_min=0.25)
rocket = Model("rocket", 12, 3, 0, 0, mass=1.0, length=1.0)


def _dp(a):
    return a.ctypes.data_as(_lib.c_double_p)


def _ip(a):
    return a.ctypes.data_as(_lib.c_int32_p)


def _f64(a, shape):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim == len(shape) and all(s == -1 or s == n for s, n in zip(shape, a.shape)):
        return a                       # already the right shape: keep the caller's object (identity is used for call caching)
    return a.reshape(shape)


class SimulatorGrad:
    """`sim.grad` of a RoboDojo Simulator as the reference reads it (examples/hopper.jl:97-99,141-142,159; sized at
    src/dynamics.jl:39-46): 1-element lists (T = 1) of the column-major-in-Julia blocks, here [nq, ncol] arrays.
    Python identifiers cannot contain `∂`, so the fields are spelled dq3dq1 / dq3dq2 / dq3du1; the reference's names work through
    getattr(sim.grad, "∂q3∂q1")."""
    _ALIASES = {"∂q3∂q1": "dq3dq1", "∂q3∂q2": "dq3dq2", "∂q3∂u1": "dq3du1"}

    def __init__(self, nq, nu):
        self.dq3dq1 = [np.zeros((nq, nq))]
        self.dq3dq2 = [np.zeros((nq, nq))]
        self.dq3du1 = [np.zeros((nq, nu))]

    def __getattr__(self, name):
        alias = SimulatorGrad._ALIASES.get(name)
        if alias is None:
            raise AttributeError(name)
        return getattr(self, alias)


class Simulator:
    """What `model.eval_sim` / `model.grad_sim` are to the reference's callers (src/dynamics.jl:1-14,16-49): `.h`, `.model`
    (`.model.nq` …), `.grad.∂q3∂{q1,q2,u1}[1]`, and the target of `RoboDojo.step!(sim, q2, v1, u1, t)` (`robodojo.step`).  A proxy:
    both simulators of one ImplicitDynamics share its device handle; `diff_sol` selects κ_tol and whether the IFT runs."""

    def __init__(self, owner, diff_sol):
        self._owner = owner
        self.diff_sol = bool(diff_sol)
        self.model = owner.model
        self.h = owner.h
        self.grad = SimulatorGrad(owner.nq, owner.nu)
        self.status = 0

    @property
    def κ_tol(self):
        return self._owner.opts.kappa_grad_tol if self.diff_sol else self._owner.opts.kappa_eval_tol

    def step(self, q, v, u, t=1):
        """q3 = step!(sim, q, v, u, t) for one problem; the gradient simulator also refreshes `sim.grad`."""
        o = self._owner
        q = _f64(q, (1, o.nq)); v = _f64(v, (1, o.nq)); u = _f64(u, (1, o.nu))
        q3 = np.empty((1, o.nq)); st = np.empty(1, dtype=np.int32)
        L = _lib.lib()
        if self.diff_sol:
            d1 = np.empty((1, o.nq, o.nq)); d2 = np.empty((1, o.nq, o.nq)); du = np.empty((1, o.nu, o.nq))
            _lib.check(L.od_sim_step_batch(o._handle(), 1, 1, _dp(q), _dp(v), _dp(u), _dp(q3), _dp(d1), _dp(d2), _dp(du), _ip(st)))
            self.grad.dq3dq1[0] = d1[0].T; self.grad.dq3dq2[0] = d2[0].T; self.grad.dq3du1[0] = du[0].T
        else:
            _lib.check(L.od_sim_step_batch(o._handle(), 1, 0, _dp(q), _dp(v), _dp(u), _dp(q3), None, None, None, _ip(st)))
        self.status = int(st[0])
        return q3[0]

    def step_batch(self, q, v, u):
        """Batched step!: returns (q3[B,nq], status[B]) and, for the gradient simulator, also the three Jacobian stacks [B,nq,ncol]."""
        o = self._owner
        q = _f64(q, (-1, o.nq)); B = q.shape[0]
        v = _f64(v, (B, o.nq)); u = _f64(u, (B, o.nu))
        q3 = np.empty((B, o.nq)); st = np.empty(B, dtype=np.int32)
        L = _lib.lib()
        if not self.diff_sol:
            _lib.check(L.od_sim_step_batch(o._handle(), B, 0, _dp(q), _dp(v), _dp(u), _dp(q3), None, None, None, _ip(st)))
            return q3, st
        d1 = np.empty((B, o.nq, o.nq)); d2 = np.empty((B, o.nq, o.nq)); du = np.empty((B, o.nu, o.nq))
        _lib.check(L.od_sim_step_batch(o._handle(), B, 1, _dp(q), _dp(v), _dp(u), _dp(q3), _dp(d1), _dp(d2), _dp(du), _ip(st)))
        return q3, d1.transpose(0, 2, 1), d2.transpose(0, 2, 1), du.transpose(0, 2, 1), st


class ImplicitDynamics:
    def __init__(self, model, h, r_func=None, rz_func=None, rθ_func=None, T=1, r_tol=1.0e-8, κ_eval_tol=1.0e-6, κ_grad_tol=1.0e-6,
                 no_impact=False, no_friction=False, n=None, m=None, d=None, nc=None, nb=None, info=None, device=0,
                 kappa_eval_tol=None, kappa_grad_tol=None):
        L = _lib.lib()
        self.model = model
        self.h = float(h)
        self.n = 2 * model.nq if n is None else n
        self.m = model.nu if m is None else m
        self.d = model.nw if d is None else d
        self.info = info
        self.nq, self.nu = model.nq, model.nu
        self.idx_q1 = np.arange(model.nq)
        self.idx_q2 = model.nq + np.arange(model.nq)
        self.idx_u1 = np.arange(model.nu)
        opts = _lib.od_options()
        _lib.check(L.od_default_options(MODEL_IDS[model.name], C.byref(opts)))
        opts.r_tol = r_tol
        opts.kappa_eval_tol = κ_eval_tol if kappa_eval_tol is None else kappa_eval_tol
        opts.kappa_grad_tol = κ_grad_tol if kappa_grad_tol is None else kappa_grad_tol
        self.opts = opts
        self.device = device
        self._friction_seen = None
        self._hd = None
        self._make_handle()
        self._packed_cache = None
        self._lib = L
        self._memo_key = None      # (x, u) of the last gradient solve: fx and fu share one launch (reference solves twice)
        self._memo = None
        # the two simulators of the reference struct (src/dynamics.jl:2-3,60-64): eval_sim (diff_sol = false, κ_eval_tol) and
        # grad_sim (diff_sol = true, κ_grad_tol) — what examples/hopper.jl:52-160 drives through RoboDojo.step!
        self.eval_sim = Simulator(self, False)
        self.grad_sim = Simulator(self, True)

    # the reference mutates model.friction after construction (examples/cartpole.jl:21); re-create the handle if it changed
    def _make_handle(self):
        L = _lib.lib()
        fr = self.model.friction
        params = None if fr is None else _dp(np.ascontiguousarray(fr, dtype=np.float64))
        hd = L.od_create(MODEL_IDS[self.model.name], self.h, C.byref(self.opts), params, 0 if fr is None else len(fr), self.device)
        if not hd:
            raise RuntimeError("optdyn_b200: " + L.od_last_error().decode())
        if self._hd:
            L.od_destroy(self._hd)
        self._hd = hd
        self._friction_seen = None if fr is None else fr.tobytes()

    def _handle(self):
        fr = self.model.friction
        if fr is not None and fr.tobytes() != self._friction_seen:
            self._make_handle()
            self._memo_key = None
        return self._hd

    def __del__(self):
        try:
            if self._hd:
                _lib.lib().od_destroy(self._hd)
                self._hd = None
        except Exception:
            pass

    # ---- batched API -------------------------------------------------------------------------------------------------
    def step_batch(self, q1, q2, u):
        """q3 for B problems (eval simulator).  Returns (q3[B,nq], status[B])."""
        q1 = _f64(q1, (-1, self.nq)); B = q1.shape[0]
        q2 = _f64(q2, (B, self.nq)); u = _f64(u, (B, self.nu))
        q3 = np.empty((B, self.nq)); st = np.empty(B, dtype=np.int32)
        _lib.check(_lib.lib().od_step_batch(self._handle(), B, _dp(q1), _dp(q2), _dp(u), _dp(q3), _ip(st)))
        return q3, st

    def step_grad_batch(self, q1, q2, u, want_q3=True):
        """q3 and ∂q3/∂q1, ∂q3/∂q2, ∂q3/∂u1 for B problems.  Jacobians are returned as [B, nq, ncol] (row = q3 component)."""
        q1 = _f64(q1, (-1, self.nq)); B = q1.shape[0]
        q2 = _f64(q2, (B, self.nq)); u = _f64(u, (B, self.nu))
        q3 = np.empty((B, self.nq)) if want_q3 else None
        d1 = np.empty((B, self.nq, self.nq)); d2 = np.empty((B, self.nq, self.nq)); du = np.empty((B, self.nu, self.nq))
        st = np.empty(B, dtype=np.int32)
        _lib.check(_lib.lib().od_step_grad_batch(self._handle(), B, _dp(q1), _dp(q2), _dp(u), None if q3 is None else _dp(q3),
                                                 _dp(d1), _dp(d2), _dp(du), _ip(st)))
        # the ABI blocks are column-major; as C arrays they are the transposes
        return q3, d1.transpose(0, 2, 1), d2.transpose(0, 2, 1), du.transpose(0, 2, 1), st

    def step_grad_packed(self, xin, out=None, status=None):
        """Packed rows in [q1|q2|u] → out [q3 | ∂q3/∂q1 | ∂q3/∂q2 | ∂q3/∂u1] (blocks column-major).  Fewest transfers."""
        inw = 2 * self.nq + self.nu
        outw = self.nq + self.nq * inw
        c = self._packed_cache          # repeated calls on the same buffers (a solver's trajectory arrays) skip the numpy→ctypes conversions
        if c is not None and c[0] is xin and c[1] is out and c[2] is status:
            B, pin, pout, pst = c[3:]
        else:
            keep = xin, out, status
            xin = _f64(xin, (-1, inw)); B = xin.shape[0]
            out = np.empty((B, outw)) if out is None else out
            status = np.empty(B, dtype=np.int32) if status is None else status
            if out.shape != (B, outw) or out.dtype != np.float64 or not out.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous float64 array of shape (%d, %d)" % (B, outw))
            if status.shape != (B,) or status.dtype != np.int32:
                raise ValueError("status must be an int32 array of shape (%d,)" % B)
            pin, pout, pst = _dp(xin), _dp(out), _ip(status)
            if keep[0] is xin and keep[1] is out and keep[2] is status:
                self._packed_cache = (xin, out, status, B, pin, pout, pst)
        rc = self._lib.od_step_grad_packed(self._handle(), B, pin, pout, pst)
        if rc:
            _lib.check(rc)
        return out, status

    def launch_count(self):
        return int(_lib.lib().od_launch_count(self._hd))

    # ---- single-problem gradient with memoisation --------------------------------------------------------------------------
    def _grad(self, x, u):
        x = np.asarray(x, dtype=np.float64); u = np.asarray(u, dtype=np.float64)
        self._handle()             # a mutated model.friction (examples/cartpole.jl:21) re-creates the handle and drops the memo
        key = (x.tobytes(), u.tobytes())
        if key != self._memo_key:
            q1 = x[self.idx_q1]; q2 = x[self.idx_q2]
            _, d1, d2, du, st = self.step_grad_batch(q1[None], q2[None], u[self.idx_u1][None], want_q3=False)
            self._memo_key, self._memo = key, (d1[0], d2[0], du[0], int(st[0]))
        return self._memo


def f(d, model, x, u, w):
    """d = [q2; q3] — reference src/dynamics.jl:81-94."""
    x = np.asarray(x, dtype=np.float64)
    q1 = x[model.idx_q1]; q2 = x[model.idx_q2]
    q3, _ = model.step_batch(q1[None], q2[None], np.asarray(u, dtype=np.float64)[model.idx_u1][None])
    d[model.idx_q1] = q2
    d[model.idx_q2] = q3[0]
    return d


def fx(dx, model, x, u, w):
    """dx = [0 I; ∂q3∂q1 ∂q3∂q2]; only these blocks are written (reference src/dynamics.jl:96-114)."""
    d1, d2, _, _ = model._grad(x, u)
    nq = model.nq
    for i in range(nq):
        dx[model.idx_q1[i], model.idx_q2[i]] = 1.0
    dx[np.ix_(model.idx_q2, model.idx_q1)] = d1
    dx[np.ix_(model.idx_q2, model.idx_q2)] = d2
    return dx


def fu(du, model, x, u, w):
    """du[q2 rows, :] = ∂q3∂u1 (reference src/dynamics.jl:116-128)."""
    _, _, dU, _ = model._grad(x, u)
    du[model.idx_q2, :] = dU
    return du


def state_to_configuration(x):
    """[x[1][1:nq], x[1][nq+1:2nq], x[2][nq+1:2nq], …] — reference src/dynamics.jl:131-145."""
    nq = len(x[0]) // 2
    q = []
    for t, xt in enumerate(x):
        xt = np.asarray(xt)
        if t == 0:
            q.append(xt[:nq].copy())
        q.append(xt[nq:2 * nq].copy())
    return q
