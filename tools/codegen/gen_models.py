#!/usr/bin/env python
"""Build-time code generator: symbolic model definitions -> CUDA device functions.

Successor of the reference's Symbolics.jl code generation (reference deps/build.jl:27-49, src/models/*/codegen.jl): each
model's residual is written down symbolically (sympy), differentiated exactly, common sub-expressions are eliminated and the
result is printed as straight-line `__device__` code under optimization_dynamics_b200/csrc/gen/.  Unlike the reference,
which generates dense r / rz / rθ, the contact models are emitted in *block* form for the condensed Newton solver
(csrc/contact_ip.cuh):

    d(q, γ, b; θ)   dynamics rows                ϕ(q)      signed distances        ψ̂(γ; θ)  friction-cone radius targets
    vT(q; θ)        tangential velocities
    D = ∂d/∂q   Eγ = ∂d/∂γ   Eb = ∂d/∂b   N = ∂ϕ/∂q   V = ∂vT/∂q   Mψ = ∂ψ̂/∂γ        (jac)
    Dθ = ∂d/∂θ'   Vθ = ∂vT/∂θ'   with θ' = (q1, q2, u) — the columns f/fx/fu return    (jacth)

Run:  python tools/codegen/gen_models.py            (regenerates every built-in header; output is committed)
      python tools/codegen/gen_models.py --spec my_model.py --out DIR      (a user-written model specification → DIR/model_<name>.cuh
                                                                             with the traits struct; format and example:
                                                                             tools/codegen/examples/particle_spec.py)
"""
import os
import sys
import time

import sympy as sp
from sympy.printing.c import C99CodePrinter

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "optimization_dynamics_b200", "csrc", "gen")


class Printer(C99CodePrinter):
    def _print_Pow(self, expr):
        b, e = expr.base, expr.exp
        if e.is_Integer:
            n = int(e)
            if n == -1:
                return "(1.0/(%s))" % self._print(b)
            if 0 < n <= 3:
                s = self.parenthesize(b, 1000)
                return "(" + "*".join([s] * n) + ")"
            if -3 <= n < 0:
                s = self.parenthesize(b, 1000)
                return "(1.0/(" + "*".join([s] * (-n)) + "))"
            if 3 < n <= 64:                                   # square-and-multiply (fastmath.cuh): x^10 in 4 multiplications, not 9
                return "ipow<%d>(%s)" % (n, self._print(b))
            if -64 <= n < -3:
                return "(1.0/ipow<%d>(%s))" % (-n, self._print(b))
        if e == sp.Rational(1, 2):
            return "sqrt(%s)" % self._print(b)
        if e == -sp.Rational(1, 2):
            return "rsqrt_d(%s)" % self._print(b)
        if e.is_Rational and e.q == 2 and abs(e.p) <= 9:          # b^(n/2), n odd: integer power times one (r)sqrt instead of pow()
            n = (abs(int(e.p)) - 1) // 2
            sb = self.parenthesize(b, 1000)
            ip = "*".join([sb] * n)
            if e.p > 0:
                return "(%s*sqrt(%s))" % (ip, self._print(b))
            return "(rsqrt_d(%s)/(%s))" % (self._print(b), ip)
        return "pow(%s, %s)" % (self._print(b), self._print(sp.Float(e) if e.is_Rational else e))

    def _print_Rational(self, expr):
        return "(%d.0/%d.0)" % (expr.p, expr.q)

    def _print_Float(self, expr):
        return repr(float(expr))

    def _print_Integer(self, expr):
        return "%d.0" % int(expr)


PR = Printer()


def vec(name, n):
    return [sp.Symbol("%s%d" % (name, i), real=True) for i in range(n)]


def emit_function(fname, inputs, outputs, doc=""):
    """inputs: list of (array_name, symbols); outputs: list of (array_name, sympy Matrix/list flattened row-major)."""
    exprs, slots = [], []
    for oname, vals in outputs:
        flat = list(vals)
        for i, e in enumerate(flat):
            exprs.append(sp.sympify(e))
            slots.append((oname, i))
    t0 = time.time()
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("x"), optimizations="basic", order="none")
    lines = []
    used = set()
    for e in list(red) + [r for _, r in repl]:
        used |= e.free_symbols
    args = ", ".join(["const double* __restrict__ %s" % n for n, _ in inputs] + ["double* __restrict__ %s" % n for n, _ in outputs])
    lines.append("// %s" % doc if doc else "")
    lines.append("OD_HD void %s(%s) {" % (fname, args))
    for aname, syms in inputs:
        for i, s in enumerate(syms):
            if s in used:
                lines.append("    const double %s = %s[%d];" % (s.name, aname, i))
    # pair sin(a)/cos(a) of the same argument into one sincos() call (one range reduction instead of two)
    trig = {}
    for s_, e in repl:
        if e.func in (sp.sin, sp.cos):
            trig.setdefault(e.args[0], {})[e.func] = s_
    paired = {arg: d for arg, d in trig.items() if len(d) == 2}
    emitted = set()
    for s_, e in repl:
        if e.func in (sp.sin, sp.cos) and e.args[0] in paired:
            arg = e.args[0]
            if arg not in emitted:
                emitted.add(arg)
                d = paired[arg]
                lines.append("    double %s, %s; od_sincos(%s, &%s, &%s);" % (d[sp.sin].name, d[sp.cos].name, PR.doprint(arg), d[sp.sin].name, d[sp.cos].name))
            continue
        lines.append("    const double %s = %s;" % (s_.name, PR.doprint(e)))
    for (oname, i), e in zip(slots, red):
        lines.append("    %s[%d] = %s;" % (oname, i, PR.doprint(e)))
    lines.append("}")
    nops = sum(sp.count_ops(e) for _, e in repl) + sum(sp.count_ops(e) for e in red)
    sys.stderr.write("  %-28s %4d temporaries %6d ops  (%.1fs)\n" % (fname, len(repl), nops, time.time() - t0))
    return "\n".join(lines) + "\n", nops


def variational(Mv, C, q0, q1, q2, h):
    """d = h/2 D1L1 + D2L1 + h/2 D1L2 − D2L2, D1L = −C(q,v), D2L = M(q) v   (reference src/models/cartpole/model.jl:53-63)."""
    n = len(q0)
    qm1 = [(q0[i] + q1[i]) / 2 for i in range(n)]
    vm1 = [(q1[i] - q0[i]) / h for i in range(n)]
    qm2 = [(q1[i] + q2[i]) / 2 for i in range(n)]
    vm2 = [(q2[i] - q1[i]) / h for i in range(n)]
    C1, C2 = C(qm1, vm1), C(qm2, vm2)
    p1, p2 = Mv(qm1, vm1), Mv(qm2, vm2)
    return [h / 2 * (-C1[i]) + p1[i] + h / 2 * (-C2[i]) - p2[i] for i in range(n)], qm2, vm2


def lagrangian_MC(L, q, qd):
    """M = ∂²L/∂q̇², C = (∂²L/∂q̇∂q) q̇ − ∂L/∂q  (RoboDojo codegen convention)."""
    n = len(q)
    M = sp.Matrix(n, n, lambda i, j: sp.diff(L, qd[i], qd[j]))
    dLq = sp.Matrix([sp.diff(L, q[i]) for i in range(n)])
    ddLqdq = sp.Matrix(n, n, lambda i, j: sp.diff(L, qd[i], q[j]))
    C = ddLqdq * sp.Matrix(qd) - dLq
    return sp.simplify(M), sp.simplify(C)


# ---------------------------------------------------------------------------------------------------------------------
# model definitions: return dict(name, NQ, NU, NC, NP, NB, cone_dims, theta symbols, d, phi, psit, vT)
# ---------------------------------------------------------------------------------------------------------------------
def model_hopper():
    NQ, NU, NC, NP, NB = 4, 2, 4, 2, 2
    q, gam, b = vec("q", NQ), vec("g", NC), vec("b", NB)
    th = vec("t", 13)
    q0, q1, u, mu, h = th[0:4], th[4:8], th[8:10], th[10:12], th[12]
    mb, Ib, mf, grav = 3.0, 0.75, 1.0, 9.81
    rb, rf, lmax, lmin = 0.1, 0.05, 1.0, 0.25
    # Lagrangian (RoboDojo hopper)
    Q, Qd = vec("Q", 4), vec("V", 4)
    foot = [Q[0] + Q[3] * sp.sin(Q[2]), Q[1] - Q[3] * sp.cos(Q[2])]
    vfoot = [sum(sp.diff(foot[k], Q[i]) * Qd[i] for i in range(4)) for k in range(2)]
    L = (sp.Rational(1, 2) * mb * (Qd[0] ** 2 + Qd[1] ** 2) + sp.Rational(1, 2) * Ib * Qd[2] ** 2
         + sp.Rational(1, 2) * mf * (vfoot[0] ** 2 + vfoot[1] ** 2) - mb * grav * Q[1] - mf * grav * foot[1])
    Msym, Csym = lagrangian_MC(L, Q, Qd)

    def Mv(qq, vv):
        sub = dict(zip(Q + Qd, list(qq) + list(vv)))
        return list((Msym * sp.Matrix(Qd)).subs(sub))

    def C(qq, vv):
        sub = dict(zip(Q + Qd, list(qq) + list(vv)))
        return list(Csym.subs(sub))

    d, qm2, vm2 = variational(Mv, C, q0, q1, q, h)
    # B(qm2)' u, B = [0 0 1 0; −sin t  cos t  0  1]
    d[0] += -sp.sin(qm2[2]) * u[1]
    d[1] += sp.cos(qm2[2]) * u[1]
    d[2] += u[0]
    d[3] += u[1]
    st, ct = sp.sin(q[2]), sp.cos(q[2])
    Jfx = [1, 0, q[3] * ct, st]
    Jfz = [0, 1, q[3] * st, -ct]
    for i in range(4):
        d[i] += Jfx[i] * b[1] + Jfz[i] * gam[1]
    d[0] += b[0]
    d[1] += gam[0]
    d[2] += rb * b[0]
    d[3] += gam[2] - gam[3]
    phi = [q[1] - rb, q[1] - q[3] * ct - rf, q[3] - lmin, lmax - q[3]]
    psit = [mu[0] * gam[0], mu[1] * gam[1]]
    v = [(q[i] - q1[i]) / h for i in range(4)]
    vT = [v[0] + rb * v[2], sum(Jfx[i] * v[i] for i in range(4))]
    return dict(name="hopper", NQ=NQ, NU=NU, NC=NC, NP=NP, NB=NB, cone_dims=[1, 1], NTH=13, q=q, gam=gam, b=b, th=th,
                d=d, phi=phi, psit=psit, vT=vT)


def _acrobot(impact):
    NQ, NU = 2, 1
    NC = 2 if impact else 0
    q, gam, b = vec("q", NQ), vec("g", max(NC, 1)), vec("b", 1)
    th = vec("t", 6)
    q0, q1, u, h = th[0:2], th[2:4], th[4], th[5]
    m1, J1, l1, lc1, m2, J2, l2, lc2, g = 1.0, 0.333, 1.0, 0.5, 1.0, 0.333, 1.0, 0.5, 9.81

    def Mv(x, v):
        a = J1 + J2 + m2 * l1 * l1 + 2.0 * m2 * l1 * lc2 * sp.cos(x[1])
        bb = J2 + m2 * l1 * lc2 * sp.cos(x[1])
        c = J2
        return [a * v[0] + bb * v[1], bb * v[0] + c * v[1]]

    def C(x, v):
        k = m2 * l1 * lc2
        ca, cb, cc = -2.0 * k * sp.sin(x[1]) * v[1], -1.0 * k * sp.sin(x[1]) * v[1], k * sp.sin(x[1]) * v[0]
        ta = -1.0 * m1 * g * lc1 * sp.sin(x[0]) - m2 * g * (l1 * sp.sin(x[0]) + lc2 * sp.sin(x[0] + x[1]))
        tb = -1.0 * m2 * g * lc2 * sp.sin(x[0] + x[1])
        return [ca * v[0] + cb * v[1] - ta, cc * v[0] - tb]

    d, qm2, vm2 = variational(Mv, C, q0, q1, q, h)
    d[1] += u
    for i in range(2):
        d[i] += -h * sp.Rational(1, 2) * vm2[i]
    phi = []
    if impact:
        phi = [sp.pi / 2 - q[1], q[1] + sp.pi / 2]
        d[1] += gam[1] - gam[0]     # P' λ, P = ∂ϕ/∂q
    return dict(name="acrobot_impact" if impact else "acrobot_nominal", NQ=NQ, NU=NU, NC=NC, NP=0, NB=0, cone_dims=[], NTH=6,
                q=q, gam=gam, b=b, th=th, d=d, phi=phi, psit=[], vT=[])


def _cartpole(friction):
    NQ, NU = 2, 1
    NP = NB = 2 if friction else 0
    q, gam, b = vec("q", NQ), vec("g", 1), vec("b", max(NB, 1))
    th = vec("t", 8 if friction else 6)
    q0, q1, u = th[0:2], th[2:4], th[4]
    h = th[7] if friction else th[5]
    mc, mp, l, g = 1.0, 0.2, 0.5, 9.81

    def Mv(x, v):
        return [(mc + mp) * v[0] + mp * l * sp.cos(x[1]) * v[1], mp * l * sp.cos(x[1]) * v[0] + mp * l ** 2 * v[1]]

    def C(x, v):
        c12 = -1.0 * mp * v[1] * l * sp.sin(x[1])
        return [-(c12 * v[1]), mp * g * l * sp.sin(x[1])]

    d, qm2, vm2 = variational(Mv, C, q0, q1, q, h)
    d[0] += u
    psit, vT = [], []
    if friction:
        d[0] += b[0]
        d[1] += b[1]
        psit = [th[5] * (mp + mc) * g * h, th[6] * (mp * g * l) * h]
        vT = [(q[0] - q1[0]) / h, (q[1] - q1[1]) / h]
    return dict(name="cartpole_friction" if friction else "cartpole_frictionless", NQ=NQ, NU=NU, NC=0, NP=NP, NB=NB,
                cone_dims=[1, 1] if friction else [], NTH=len(th), q=q, gam=gam, b=b, th=th, d=d, phi=[], psit=psit, vT=vT)


def model_planar_push():
    NQ, NU, NC, NP, NB = 5, 2, 1, 5, 9
    q, gam, b = vec("q", NQ), vec("g", NC), vec("b", NB)
    th = vec("t", 13)
    q0, q1, u, h = th[0:5], th[5:10], th[10:12], th[12]
    r_dim, mu_s, mu_p, grav, m_b, m_p = 0.1, 0.5, 0.5, 9.81, 1.0, 10.0
    inertia = 1.0 / 12.0 * m_b * ((2.0 * r_dim) ** 2 + (2.0 * r_dim) ** 2)
    c, s = sp.cos(-q[2]), sp.sin(-q[2])
    dx, dy = q[3] - q[0], q[4] - q[1]
    D1, D2 = c * dx - s * dy, s * dx + c * dy
    phi = (D1 ** 10 + D2 ** 10) ** sp.Rational(1, 10) - r_dim
    N = [sp.diff(phi, q[i]) for i in range(5)]
    cc = [(r_dim, r_dim), (-r_dim, r_dim), (r_dim, -r_dim), (-r_dim, -r_dim)]
    ct, st = sp.cos(q[2]), sp.sin(q[2])
    P = []
    for (cx, cy) in cc:
        px = q[0] + ct * cx - st * cy
        py = q[1] + st * cx + ct * cy
        P.append([sp.diff(px, q[i]) for i in range(5)])
        P.append([sp.diff(py, q[i]) for i in range(5)])
    nn = sp.sqrt(N[3] ** 2 + N[4] ** 2)
    n1, n2 = N[3] / nn, N[4] / nn
    t1, t2 = -n2, n1
    r1, r2 = q[3] - q[0], q[4] - q[1]
    m = r1 * t2 - r2 * t1
    P.append([t1, t2, m, -t1, -t2])
    Md = [m_b, m_b, inertia, m_p, m_p]
    d = []
    for i in range(5):
        vm1, vm2 = (q1[i] - q0[i]) / h, (q[i] - q1[i]) / h
        e = Md[i] * vm1 - Md[i] * vm2
        if i == 3:
            e += u[0]
        if i == 4:
            e += u[1]
        e += N[i] * gam[0] + sum(P[k][i] * b[k] for k in range(9))
        d.append(e)
    psit = [mu_s * m_b * grav * h * sp.Rational(1, 4)] * 4 + [mu_p * gam[0]]
    vT = [sum(P[k][j] * (q[j] - q1[j]) for j in range(5)) / h for k in range(9)]
    return dict(name="planar_push", NQ=NQ, NU=NU, NC=NC, NP=NP, NB=NB, cone_dims=[2, 2, 2, 2, 1], NTH=13, q=q, gam=gam, b=b, th=th,
                d=d, phi=[phi], psit=psit, vT=vT)


def hoist_trig(expr_lists, qsyms):
    """Replace every sin(a)/cos(a) in the given expression lists by symbols.  Returns (new lists, const args, var args) where
    `const` arguments depend on θ only (evaluated once per problem) and `var` arguments depend on q (once per candidate point).
    Symbol for argument k of a class: <cls>S<k> / <cls>C<k>, stored in the array tr<cls>[2k], tr<cls>[2k+1]."""
    args = []
    for lst in expr_lists:
        for e in lst:
            for a in sp.sympify(e).atoms(sp.sin, sp.cos):
                if a.args[0] not in args:
                    args.append(a.args[0])
    args = sorted(args, key=lambda a: sp.default_sort_key(a))
    const = [a for a in args if not (a.free_symbols & set(qsyms))]
    var = [a for a in args if a.free_symbols & set(qsyms)]
    table = {}
    syms = {"c": [], "v": []}
    for cls, lst in (("c", const), ("v", var)):
        for k, a in enumerate(lst):
            S, Cc = sp.Symbol("tr%s%d" % (cls, 2 * k), real=True), sp.Symbol("tr%s%d" % (cls, 2 * k + 1), real=True)
            table[sp.sin(a)] = S
            table[sp.cos(a)] = Cc
            syms[cls] += [S, Cc]
    new_lists = [[sp.sympify(e).xreplace(table) for e in lst] for lst in expr_lists]
    return new_lists, const, var, syms


def hoist_roots(expr_lists, den=10):
    """Fractional powers b^(m/den) (the 10-norm signed distance of the planar push and its derivatives: exponents −9/10, −9/5,
    −19/10, −14/5, 1/10 of ONE base) are rewritten as b^k · W^r with W = b^(1/den) a table symbol and |r| ≤ den/2, so that an
    evaluation point costs one pow() (in trig_var) instead of one per distinct exponent (12 in the planar-push Jacobian code).
    The base b itself is tabulated too.  Returns (new lists, bases, root symbols, base symbols); the caller maps the symbols
    into the trv table."""
    bases = []
    for lst in expr_lists:
        for e in lst:
            for p in sp.sympify(e).atoms(sp.Pow):
                if p.exp.is_Rational and not p.exp.is_Integer and den % p.exp.q == 0 and p.exp.q != 2 and p.base not in bases:
                    bases.append(p.base)
    if not bases:
        return expr_lists, [], [], []
    bases = sorted(bases, key=lambda a: sp.default_sort_key(a))
    syms = [sp.Symbol("rt%d" % k, positive=True) for k in range(len(bases))]
    bsyms = [sp.Symbol("rb%d" % k, positive=True) for k in range(len(bases))]      # the base itself is tabulated next to its root

    def rewrite(p):
        if not (p.is_Pow and p.exp.is_Rational and not p.exp.is_Integer and den % p.exp.q == 0 and p.exp.q != 2 and p.base in bases):
            return p
        m = int(p.exp * den)
        k = int(sp.floor(sp.Rational(m, den) + sp.Rational(1, 2)))
        r = m - den * k
        return bsyms[bases.index(p.base)] ** k * syms[bases.index(p.base)] ** r

    new_lists = [[sp.sympify(e).replace(lambda x: x.is_Pow, rewrite) for e in lst] for lst in expr_lists]
    return new_lists, bases, syms, bsyms


def emit_trig(fname, inputs, args, arr, roots=(), root_den=10):
    lines = ["__host__ __device__ __forceinline__ void %s(%s, double* __restrict__ %s) {" % (
        fname, ", ".join("const double* __restrict__ %s" % n for n, _ in inputs), arr)]
    used = set()
    for a in list(args) + list(roots):
        used |= a.free_symbols
    for aname, syms in inputs:
        for i, sy in enumerate(syms):
            if sy in used:
                lines.append("    const double %s = %s[%d];" % (sy.name, aname, i))
    for k, a in enumerate(args):
        lines.append("    od_sincos(%s, &%s[%d], &%s[%d]);" % (PR.doprint(a), arr, 2 * k, arr, 2 * k + 1))
    if roots:                                         # b^(1/den) of the hoisted bases; they may use the sin/cos entries above
        for k in range(2 * len(args)):
            lines.append("    const double %s%d = %s[%d];" % (arr, k, arr, k))
        for k, bexpr in enumerate(roots):
            o = 2 * len(args) + 2 * k
            lines.append("    %s[%d] = %s;" % (arr, o + 1, PR.doprint(bexpr)))
            lines.append("    %s[%d] = pow(%s[%d], %s);" % (arr, o, arr, o + 1, repr(1.0 / root_den)))
    if not args and not roots:
        lines.append("    (void)%s;" % arr)
    lines.append("}")
    return "\n".join(lines) + "\n"


def gen_contact(m):
    NQ, NU, NC, NP, NB = m["NQ"], m["NU"], m["NC"], m["NP"], m["NB"]
    q, gam, b, th = m["q"], m["gam"], m["b"], m["th"]
    thp = th[0:2 * NQ + NU]
    d, phi, psit, vT = sp.Matrix(m["d"]), sp.Matrix(m["phi"]), sp.Matrix(m["psit"]), sp.Matrix(m["vT"])

    def J(f, x, n):
        return list(f.jacobian(sp.Matrix(x[:n]))) if (len(f) and n) else []

    groups = [list(d), list(phi), list(psit), list(vT),
              J(d, q, NQ), J(d, gam, NC), J(d, b, NB), J(phi, q, NQ), J(vT, q, NQ), J(psit, gam, NC),
              J(d, thp, len(thp)), J(vT, thp, len(thp))]
    groups, targs_c, targs_v, tsyms = hoist_trig(groups, q)
    groups, rbases, rsyms, rbsyms = hoist_roots(groups)
    assert all(bx.free_symbols & (set(q) | set(tsyms["v"])) for bx in rbases), "a θ-only root would belong in trig_const"
    rmap = {}
    for k in range(len(rbases)):                       # table layout after the sin/cos pairs: [root, base] per hoisted base
        rmap[rsyms[k]] = sp.Symbol("trv%d" % (2 * len(targs_v) + 2 * k), real=True)
        rmap[rbsyms[k]] = sp.Symbol("trv%d" % (2 * len(targs_v) + 2 * k + 1), real=True)
        tsyms["v"] = tsyms["v"] + [rmap[rsyms[k]], rmap[rbsyms[k]]]
    if rmap:
        groups = [[e.xreplace(rmap) for e in lst] for lst in groups]
    NTC, NTV = 2 * len(targs_c), 2 * len(targs_v) + 2 * len(rbases)
    ins = [("q", q), ("gam", gam), ("b", b), ("th", th), ("trc", tsyms["c"]), ("trv", tsyms["v"])]

    out = ["// GENERATED by tools/codegen/gen_models.py — do not edit.  Model: %s" % m["name"],
           "// Block-form residual pieces of the contact-implicit step (see csrc/contact_ip.cuh for the layout).",
           "// sin/cos (and tenth roots) are hoisted: trig_const(θ) once per problem, trig_var(q,θ) once per candidate point; eq/jac/jacth take the tables.",
           "#pragma once", "namespace od { namespace gen_%s {" % m["name"],
           "constexpr int NQ = %d, NU = %d, NC = %d, NP = %d, NB = %d, NTH = %d, NTC = %d, NTV = %d;" % (NQ, NU, NC, NP, NB, m["NTH"], NTC, NTV), ""]
    out.append(emit_trig("trig_const", [("th", th)], targs_c, "trc"))
    out.append(emit_trig("trig_var", [("q", q), ("th", th)], targs_v, "trv", roots=rbases))
    total = {}
    src, total["eq"] = emit_function("eq", ins, [("d", groups[0]), ("phi", groups[1]), ("psit", groups[2]), ("vT", groups[3])],
                                     "d(q,γ,b;θ), ϕ(q), ψ̂(γ;θ), vT(q;θ)")
    out.append(src)
    src, total["jac"] = emit_function("jac", ins, [("D", groups[4]), ("Eg", groups[5]), ("Eb", groups[6]), ("N", groups[7]),
                                                   ("V", groups[8]), ("Mpsi", groups[9])],
                                      "row-major D[NQ×NQ], Eg[NQ×NC], Eb[NQ×NB], N[NC×NQ], V[NB×NQ], Mpsi[NP×NC]")
    out.append(src)
    src, total["jacth"] = emit_function("jacth", ins, [("Dth", groups[10]), ("Vth", groups[11])],
                                        "row-major Dth[NQ×(2NQ+NU)], Vth[NB×(2NQ+NU)]  (θ' = q1,q2,u)")
    out.append(src)
    out.append("constexpr int OPS_EQ = %d, OPS_JAC = %d, OPS_JACTH = %d;" % (total["eq"], total["jac"], total["jacth"]))
    out.append("} }  // namespace od::gen_%s" % m["name"])
    return "\n".join(out) + "\n"


# ---------------------------------------------------------------------------------------------------------------------
# dense models (rocket dynamics, rocket thrust projection): r, rz (row-major dense), rθ' columns
# ---------------------------------------------------------------------------------------------------------------------
def model_rocket():
    z, th = vec("z", 12), vec("t", 16)
    x, u, h = th[0:12], th[12:15], th[15]
    mass, length = 1.0, 1.0
    Jd = [1.0 / 12.0, 1.0 / 12.0, 1.0e-5]
    g = [0.0, 0.0, -9.81]

    def f(s):
        r, v, w = s[3:6], s[6:9], s[9:12]
        rr = sum(a * a for a in r)
        wr = sum(a * bb for a, bb in zip(w, r))
        cx = [w[1] * r[2] - w[2] * r[1], w[2] * r[0] - w[0] * r[2], w[0] * r[1] - w[1] * r[0]]
        rd = [sp.Rational(1, 4) * ((1 - rr) * w[i] - 2 * cx[i] + 2 * wr * r[i]) for i in range(3)]
        den = (1 + rr) ** 2
        rF = [r[1] * u[2] - r[2] * u[1], r[2] * u[0] - r[0] * u[2], r[0] * u[1] - r[1] * u[0]]
        rrF = [r[1] * rF[2] - r[2] * rF[1], r[2] * rF[0] - r[0] * rF[2], r[0] * rF[1] - r[1] * rF[0]]
        vd = [g[i] + (1.0 / mass) * (u[i] + 8 * rrF[i] / den + 4 * (1 - rr) * rF[i] / den) for i in range(3)]
        tau = [length * u[1], -length * u[0], 0]
        Jw = [Jd[i] * w[i] for i in range(3)]
        wJw = [w[1] * Jw[2] - w[2] * Jw[1], w[2] * Jw[0] - w[0] * Jw[2], w[0] * Jw[1] - w[1] * Jw[0]]
        wd = [(1.0 / Jd[i]) * (tau[i] - wJw[i]) for i in range(3)]
        return list(v) + rd + vd + wd

    xm = [(x[i] + z[i]) / 2 for i in range(12)]
    fm = f(xm)
    r = [z[i] - (x[i] + h * fm[i]) for i in range(12)]
    return dict(name="rocket", NZ=12, NTH=16, NTHP=15, z=z, th=th, r=r, kappa_rows=[])


def model_rocket_proj():
    z, th = vec("z", 10), vec("t", 4)
    u, p, s, w, y, v = z[0:3], z[3], z[4], z[5], z[6], z[7:10]
    r = [u[0] - th[0] - v[0], u[1] - th[1] - v[1], u[2] - th[2] - v[2] - (y + p), th[3] - u[2] - s, -y - w,
         w * s, p * u[2],
         u[2] * v[2] + u[0] * v[0] + u[1] * v[1], u[2] * v[0] + v[2] * u[0], u[2] * v[1] + v[2] * u[1]]
    return dict(name="rocket_proj", NZ=10, NTH=4, NTHP=3, z=z, th=th, r=r, kappa_rows=[5, 6, 7])


def gen_dense(m):
    z, th = m["z"], m["th"]
    r = sp.Matrix(m["r"])
    ins = [("z", z), ("th", th)]
    out = ["// GENERATED by tools/codegen/gen_models.py — do not edit.  Model: %s" % m["name"],
           "// Dense residual r(z;θ,κ=0), Jacobian rz (row-major NZ×NZ) and rθ' (row-major NZ×NTHP).",
           "#pragma once", "namespace od { namespace gen_%s {" % m["name"],
           "constexpr int NZ = %d, NTH = %d, NTHP = %d;" % (m["NZ"], m["NTH"], m["NTHP"]), ""]
    src, o1 = emit_function("res", ins, [("r", r)], "r(z;θ) without the −κ shift of the bilinear rows")
    out.append(src)
    src, o2 = emit_function("jac", ins, [("rz", list(r.jacobian(sp.Matrix(z))))], "rz, row-major")
    out.append(src)
    src, o3 = emit_function("jacth", ins, [("rth", list(r.jacobian(sp.Matrix(th[:m["NTHP"]]))))], "rθ', row-major")
    out.append(src)
    out.append("constexpr int OPS_RES = %d, OPS_JAC = %d, OPS_JACTH = %d;" % (o1, o2, o3))
    out.append("} }  // namespace od::gen_%s" % m["name"])
    return "\n".join(out) + "\n"


TRAITS = '''
// Traits binding od::gen_{name} to the solver templates (csrc/contact_ip.cuh) — same shape as the built-in models of csrc/models.cuh.
namespace od {{
struct {cls} {{
    static constexpr int NQ = gen_{name}::NQ, NU = gen_{name}::NU, NC = gen_{name}::NC, NP = gen_{name}::NP, NB = gen_{name}::NB;
    static constexpr int NTH = gen_{name}::NTH, NF = {nf}, NTC = gen_{name}::NTC, NTV = gen_{name}::NTV;
    static constexpr bool ROBUST_IFT = false;          // set to true for models with redundant contact constraints (planar push)
    __host__ __device__ static constexpr int cone_off(int k) {{ return {off}; }}
    __host__ __device__ static constexpr int cone_dim(int k) {{ return {dim}; }}
    OD_HD static void trig_const(const double* th, double* trc) {{ gen_{name}::trig_const(th, trc); }}
    OD_HD static void trig_var(const double* q, const double* th, double* trv) {{ gen_{name}::trig_var(q, th, trv); }}
    OD_HD static void eq(const double* q, const double* g, const double* b, const double* th, const double* trc, const double* trv,
                         double* d, double* phi, double* psit, double* vT) {{ gen_{name}::eq(q, g, b, th, trc, trv, d, phi, psit, vT); }}
    OD_HD static void jac(const double* q, const double* g, const double* b, const double* th, const double* trc, const double* trv,
                          double* D, double* Eg, double* Eb, double* N, double* V, double* Mpsi) {{
        gen_{name}::jac(q, g, b, th, trc, trv, D, Eg, Eb, N, V, Mpsi); }}
    OD_HD static void jacth(const double* q, const double* g, const double* b, const double* th, const double* trc, const double* trv,
                            double* Dth, double* Vth) {{ gen_{name}::jacth(q, g, b, th, trc, trv, Dth, Vth); }}
}};
}}  // namespace od
'''


def gen_user_model(spec_path, out_dir):
    """--spec: a user-written model specification (see tools/codegen/examples/particle_spec.py for the format) → one header with
    the generated device code and the traits struct od::<Name>Model that the solver templates take."""
    import importlib.util
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("od_user_spec", spec_path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    m = mod.model()
    need = ["name", "NQ", "NU", "NC", "NP", "NB", "cone_dims", "NTH", "q", "gam", "b", "th", "d", "phi", "psit", "vT"]
    missing = [k for k in need if k not in m]
    if missing:
        raise SystemExit("model(): missing keys %s" % missing)
    NQ, NU, NC, NP, NB = m["NQ"], m["NU"], m["NC"], m["NP"], m["NB"]
    if len(m["d"]) != NQ or len(m["phi"]) != NC or len(m["psit"]) != NP or len(m["vT"]) != NB or sum(m["cone_dims"]) != NB or len(m["cone_dims"]) != NP:
        raise SystemExit("model(): d / phi / psit / vT / cone_dims do not match NQ / NC / NP / NB")
    nf = m["NTH"] - (2 * NQ + NU) - 1
    if nf < 0 or len(m["th"]) != m["NTH"]:
        raise SystemExit("model(): θ must be [q0 (NQ), q1 (NQ), u (NU), friction parameters, h]")
    m = dict(m)
    m["q"] = list(m["q"]); m["gam"] = list(m["gam"]) or vec("g", 1); m["b"] = list(m["b"]) or vec("b", 1)
    offs, o = [], 0
    for dmn in m["cone_dims"]:
        offs.append(o); o += dmn

    def chain(vals):                                   # value of entry k as a constexpr expression
        if not vals:
            return "0"
        e = str(vals[-1])
        for k in range(len(vals) - 2, -1, -1):
            e = "(k == %d ? %d : %s)" % (k, vals[k], e)
        return e

    cls = "".join(w.capitalize() for w in m["name"].split("_")) + "Model"
    src = gen_contact(m) + TRAITS.format(name=m["name"], cls=cls, nf=nf, off=chain(offs), dim=chain(list(m["cone_dims"])))
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "model_%s.cuh" % m["name"])
    open(path, "w").write(src)
    sys.stderr.write("wrote %s (traits struct od::%s)\n" % (path, cls))
    return path, cls


def main():
    if "--spec" in sys.argv:
        i = sys.argv.index("--spec")
        out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else "."
        gen_user_model(sys.argv[i + 1], out)
        return
    only = set(sys.argv[1:])
    os.makedirs(OUT, exist_ok=True)
    contact = [("hopper", model_hopper), ("acrobot_impact", lambda: _acrobot(True)), ("acrobot_nominal", lambda: _acrobot(False)),
               ("cartpole_friction", lambda: _cartpole(True)), ("cartpole_frictionless", lambda: _cartpole(False)),
               ("planar_push", model_planar_push)]
    dense = [("rocket", model_rocket), ("rocket_proj", model_rocket_proj)]
    for name, fn in contact:
        if only and name not in only:
            continue
        sys.stderr.write("[%s]\n" % name)
        open(os.path.join(OUT, "model_%s.cuh" % name), "w").write(gen_contact(fn()))
    for name, fn in dense:
        if only and name not in only:
            continue
        sys.stderr.write("[%s]\n" % name)
        open(os.path.join(OUT, "model_%s.cuh" % name), "w").write(gen_dense(fn()))


if __name__ == "__main__":
    main()
