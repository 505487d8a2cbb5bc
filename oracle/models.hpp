// ORACLE — TEST INFRASTRUCTURE ONLY (see dual.hpp header).  PARITY UNPINNED: the reference has no tests / golden
// vectors and its solver + hopper model live in RoboDojo.jl (not in /root/reference); see oracle/README.md.
//
// Residual functions r(z; θ, κ) of every model on the hot path, restated from the reference's Julia sources as
// scalar-type templates so that the same expression yields values (S=double) and exact Jacobians (S=Dual).
//
//   acrobot  : reference src/models/acrobot/model.jl:41-157
//   cartpole : reference src/models/cartpole/model.jl:28-129
//   planar push : reference src/models/planar_push/model.jl:24-187
//   rocket dynamics / SOC projection : reference src/models/rocket/model.jl:14-48, codegen.jl:14-22,45-64
//   hopper   : NOT IN TREE (RoboDojo.jl `hopper`); structure pinned by reference examples/hopper.jl:38-50,178,270 and
//              examples/comparisons/hopper.jl:22-37,74-77,152-155; constants are the oracle's documented choice
//              (SURVEY.md Appendix A.4).
//
// θ layout for the Lagrangian models is RoboDojo's [q1; q2; u; w; friction; h] (evidence: cartpole/model.jl:86-91);
// here the two past configurations are called q0,q1 and the unknown q2 exactly as in the model files.
#pragma once
#include "dual.hpp"

namespace od_oracle {

// Second-order-cone (Jordan) product, RoboDojo `cone_product` (call sites: cartpole/model.jl:111-112).
template <class S, int D>
inline void cone_product(const S* u, const S* v, S* out) {
    S acc = u[0] * v[0];
    for (int i = 1; i < D; ++i) acc = acc + u[i] * v[i];
    out[0] = acc;
    for (int i = 1; i < D; ++i) out[i] = u[0] * v[i] + v[0] * u[i];
}

// Discrete Lagrangian terms shared by all Lagrangian models: RoboDojo `lagrangian_derivatives`
// D1L = -C(q,v), D2L = M(q) v  (call sites: acrobot/model.jl:97-98, cartpole/model.jl:58-59, planar_push/model.jl:155-156)
// and the midpoint variational integrator  d = h/2 D1L1 + D2L1 + h/2 D1L2 - D2L2  (cartpole/model.jl:53-63).
template <class Model, class S, int NQ>
inline void variational_integrator(const Model& m, const S* q0, const S* q1, const S* q2, const S& h, S* d) {
    S qm1[NQ], vm1[NQ], qm2[NQ], vm2[NQ];
    for (int i = 0; i < NQ; ++i) {
        qm1[i] = (q0[i] + q1[i]) * 0.5;
        vm1[i] = (q1[i] - q0[i]) / h;
        qm2[i] = (q1[i] + q2[i]) * 0.5;
        vm2[i] = (q2[i] - q1[i]) / h;
    }
    S C1[NQ], C2[NQ], p1[NQ], p2[NQ];
    m.template bias<S>(qm1, vm1, C1);
    m.template bias<S>(qm2, vm2, C2);
    m.template momentum<S>(qm1, vm1, p1);
    m.template momentum<S>(qm2, vm2, p2);
    for (int i = 0; i < NQ; ++i) d[i] = (h * 0.5) * (-C1[i]) + p1[i] + (h * 0.5) * (-C2[i]) - p2[i];
}

// ===================================================================================================================
// Acrobot (double pendulum with elbow joint limits).  z = [q2(2), λ(2), s(2)], θ = [q0(2), q1(2), u(1), h].
// ===================================================================================================================
struct Acrobot {
    static constexpr int NQ = 2, NU = 1;
    double m1 = 1.0, J1 = 0.333, l1 = 1.0, lc1 = 0.5, m2 = 1.0, J2 = 0.333, l2 = 1.0, lc2 = 0.5, g = 9.81;  // model.jl:159-160
    bool impact = true;

    template <class S> void momentum(const S* q, const S* v, S* p) const {  // M_func, model.jl:41-51
        S c2 = od_cos(q[1]);
        S a = (J1 + J2 + m2 * l1 * l1) + (2.0 * m2 * l1 * lc2) * c2;
        S b = J2 + (m2 * l1 * lc2) * c2;
        S c = S(J2);
        p[0] = a * v[0] + b * v[1];
        p[1] = b * v[0] + c * v[1];
    }
    template <class S> void bias(const S* q, const S* v, S* C) const {  // C_func = c*q̇ - τ, model.jl:53-79
        S s2 = od_sin(q[1]);
        double k = m2 * l1 * lc2;
        S ca = (-2.0 * k) * s2 * v[1];
        S cb = (-1.0 * k) * s2 * v[1];
        S cc = k * s2 * v[0];
        S s1 = od_sin(q[0]), s12 = od_sin(q[0] + q[1]);
        S ta = (-1.0 * m1 * g * lc1) * s1 - (m2 * g) * (l1 * s1 + lc2 * s12);
        S tb = (-1.0 * m2 * g * lc2) * s12;
        C[0] = ca * v[0] + cb * v[1] - ta;
        C[1] = cc * v[0] - tb;
    }
    // impact: nz=6 ; nominal: nz=2.  nθ=6.
    template <class S> void residual(const S* z, const S* th, const S& kappa, S* r) const {
        const S* q0 = th; const S* q1 = th + 2; const S& u = th[4]; const S& h = th[5];
        const S* q2 = z;
        S d[2];
        variational_integrator<Acrobot, S, 2>(*this, q0, q1, q2, h, d);
        // + B(qm2) u + P(q2)' λ - h/2 vm2   (model.jl:100-103,116-118);  B=[0;1], P=[[0,-1],[0,1]] (model.jl:73-88)
        d[1] = d[1] + u;
        for (int i = 0; i < 2; ++i) d[i] = d[i] - (h * 0.5) * ((q2[i] - q1[i]) / h);
        if (!impact) { r[0] = d[0]; r[1] = d[1]; return; }
        const S* lam = z + 2; const S* s = z + 4;
        d[1] = d[1] + (lam[1] - lam[0]);
        r[0] = d[0]; r[1] = d[1];
        r[2] = s[0] - (0.5 * M_PI - q2[1]);   // ϕ = [π/2 - q2, q2 + π/2], model.jl:81-83
        r[3] = s[1] - (q2[1] + 0.5 * M_PI);
        r[4] = lam[0] * s[0] - kappa;
        r[5] = lam[1] * s[1] - kappa;
    }
};

// ===================================================================================================================
// Cartpole with Coulomb joint friction.  z = [q2(2), ψ(2), b(2), sψ(2), sb(2)], θ = [q0, q1, u, μ_slider, μ_angle, h].
// Frictionless: z = q2, θ = [q0, q1, u, h]  (model.jl:116-129).
// ===================================================================================================================
struct Cartpole {
    static constexpr int NQ = 2, NU = 1;
    double mc = 1.0, mp = 0.2, l = 0.5, g = 9.81;  // model.jl:131-132
    bool friction = true;

    template <class S> void momentum(const S* q, const S* v, S* p) const {  // model.jl:28-32
        S c = od_cos(q[1]);
        p[0] = (mc + mp) * v[0] + (mp * l) * c * v[1];
        p[1] = (mp * l) * c * v[0] + (mp * l * l) * v[1];
    }
    template <class S> void bias(const S* q, const S* v, S* C) const {  // -C*q̇ + G, model.jl:43-49
        S s = od_sin(q[1]);
        S c12 = (-1.0 * mp) * v[1] * l * s;
        C[0] = -(c12 * v[1]);
        C[1] = (mp * g * l) * s;
    }
    template <class S> void residual(const S* z, const S* th, const S& kappa, S* r) const {
        const S* q0 = th; const S* q1 = th + 2; const S& u = th[4];
        const S* q2 = z;
        if (!friction) {
            const S& h = th[5];
            S d[2];
            variational_integrator<Cartpole, S, 2>(*this, q0, q1, q2, h, d);
            r[0] = d[0] + u; r[1] = d[1];   // B=[1;0], model.jl:34-36
            return;
        }
        const S& mu_s = th[5]; const S& mu_a = th[6]; const S& h = th[7];
        const S* psi = z + 2; const S* b = z + 4; const S* spsi = z + 6; const S* sb = z + 8;
        S d[2];
        variational_integrator<Cartpole, S, 2>(*this, q0, q1, q2, h, d);
        r[0] = d[0] + u + b[0];   // P = I, λ = b (model.jl:38-41,102)
        r[1] = d[1] + b[1];
        r[2] = sb[0] - (q2[0] - q1[0]) / h;
        r[3] = psi[0] - mu_s * ((mp + mc) * g) * h;
        r[4] = sb[1] - (q2[1] - q1[1]) / h;
        r[5] = psi[1] - mu_a * (mp * g * l) * h;
        for (int i = 0; i < 2; ++i) {
            S uu[2] = {psi[i], b[i]}, vv[2] = {spsi[i], sb[i]}, cp[2];
            cone_product<S, 2>(uu, vv, cp);
            r[6 + 2 * i] = cp[0] - kappa;
            r[7 + 2 * i] = cp[1];
        }
    }
};

// ===================================================================================================================
// Planar push.  z = [q2(5), γ(1), s(1), ψ(5), b(9), sψ(5), sb(9)] (35), θ = [q0(5), q1(5), u(2), h] (13).
// ===================================================================================================================
struct PlanarPush {
    static constexpr int NQ = 5, NU = 2;
    double r_dim = 0.1, mu_surface = 0.5, mu_pusher = 0.5, gravity = 9.81, mass_block = 1.0, mass_pusher = 10.0;  // model.jl:24,42-46
    double inertia() const { return 1.0 / 12.0 * mass_block * ((2.0 * r_dim) * (2.0 * r_dim) + (2.0 * r_dim) * (2.0 * r_dim)); }  // :47

    // sd_2d_box(p_pusher, p_block), model.jl:26-31: Δ = R(-θ)(p - pos); (Δ1^10 + Δ2^10)^(1/10) - r_dim
    template <class S> S sdf(const S* q) const {
        S c = od_cos(-q[2]), s = od_sin(-q[2]);
        S dx = q[3] - q[0], dy = q[4] - q[1];
        S D1 = c * dx - s * dy;
        S D2 = s * dx + c * dy;
        S sum = od_ipow(D1, 10) + od_ipow(D2, 10);
        return od_pow(sum, 0.1) - r_dim;
    }
    // N = ∂ϕ/∂q (Symbolics.jacobian inside the residual, model.jl:82-85,143-144): exact derivative via an inner dual.
    template <class S> void sdf_grad(const S* q, S* N) const {
        typedef Dual<S, 5> D;
        D qd[5];
        for (int i = 0; i < 5; ++i) qd[i] = D::variable(q[i], i);
        D phi = sdf<D>(qd);
        for (int i = 0; i < 5; ++i) N[i] = phi.d[i];
    }
    // P (9x5): rows 1..8 = ∂p_corners/∂q, row 9 = pusher tangent (model.jl:87-119)
    template <class S> void P_func(const S* q, const S* N, S P[9][5]) const {
        const double cc[4][2] = {{r_dim, r_dim}, {-r_dim, r_dim}, {r_dim, -r_dim}, {-r_dim, -r_dim}};  // :34-39
        S c = od_cos(q[2]), s = od_sin(q[2]);
        for (int k = 0; k < 4; ++k) {
            // p = pos + R(θ) cc ; ∂/∂θ = [-s*cx - c*cy ; c*cx - s*cy]
            for (int j = 0; j < 5; ++j) { P[2 * k][j] = S(0.0); P[2 * k + 1][j] = S(0.0); }
            P[2 * k][0] = S(1.0);
            P[2 * k + 1][1] = S(1.0);
            P[2 * k][2] = (-cc[k][0]) * s - cc[k][1] * c;
            P[2 * k + 1][2] = cc[k][0] * c - cc[k][1] * s;
        }
        S nn = od_sqrt(N[3] * N[3] + N[4] * N[4]);
        S n1 = N[3] / nn, n2 = N[4] / nn;
        S t1 = -n2, t2 = n1;
        S r1 = q[3] - q[0], r2 = q[4] - q[1];
        S m = r1 * t2 - r2 * t1;
        P[8][0] = t1; P[8][1] = t2; P[8][2] = m; P[8][3] = -t1; P[8][4] = -t2;
    }
    template <class S> void residual(const S* z, const S* th, const S& kappa, S* r) const {
        const S* q0 = th; const S* q1 = th + 5; const S* u = th + 10; const S& h = th[12];
        const S* q2 = z; const S& gam = z[5]; const S& s1 = z[6];
        const S* psi = z + 7; const S* b = z + 12; const S* spsi = z + 21; const S* sb = z + 26;
        S phi = sdf<S>(q2);
        S N[5];
        sdf_grad<S>(q2, N);
        S P[9][5];
        P_func<S>(q2, N, P);
        // M = diag(mb, mb, I, mp, mp), C = 0 (model.jl:54-59):  d = M vm1 - M vm2 + B u + N γ + P' b   (model.jl:150-161)
        const double Md[5] = {mass_block, mass_block, inertia(), mass_pusher, mass_pusher};
        for (int i = 0; i < 5; ++i) {
            S vm1 = (q1[i] - q0[i]) / h, vm2 = (q2[i] - q1[i]) / h;
            // 0.5 h D1L1 + D2L1 + 0.5 h D1L2 - D2L2 with D1L = -C = 0
            S d = (h * 0.5) * S(0.0) + Md[i] * vm1 + (h * 0.5) * S(0.0) - Md[i] * vm2;
            if (i == 3) d = d + u[0];
            if (i == 4) d = d + u[1];
            d = d + N[i] * gam;
            for (int k = 0; k < 9; ++k) d = d + P[k][i] * b[k];
            r[i] = d;
        }
        r[5] = s1 - phi;
        for (int i = 0; i < 4; ++i) r[6 + i] = psi[i] - (mu_surface * mass_block * gravity) * h * 0.25;
        r[10] = psi[4] - mu_pusher * gam;
        for (int k = 0; k < 9; ++k) {
            S vT = S(0.0);
            for (int j = 0; j < 5; ++j) vT = vT + P[k][j] * (q2[j] - q1[j]);
            r[11 + k] = vT / h - sb[k];
        }
        r[20] = gam * s1 - kappa;
        for (int i = 0; i < 4; ++i) {
            S uu[3] = {psi[i], b[2 * i], b[2 * i + 1]}, vv[3] = {spsi[i], sb[2 * i], sb[2 * i + 1]}, cp[3];
            cone_product<S, 3>(uu, vv, cp);
            r[21 + 3 * i] = cp[0] - kappa; r[22 + 3 * i] = cp[1]; r[23 + 3 * i] = cp[2];
        }
        S uu[2] = {psi[4], b[8]}, vv[2] = {spsi[4], sb[8]}, cp[2];
        cone_product<S, 2>(uu, vv, cp);
        r[33] = cp[0] - kappa; r[34] = cp[1];
    }
};

// ===================================================================================================================
// Hopper (RoboDojo `hopper`, NOT IN TREE).  q = (x, z, t, r);  z = [q2(4), γ(4), s(4), ψ(2), b(2), sψ(2), sb(2)] (20),
// θ = [q0(4), q1(4), u(2), μ_body, μ_foot, h] (13).  Rows: [d(4); s-ϕ(4); ψ-μγ(2); vT-sb(2); γ∘s-κ(4); 2×SOC2(4)].
// ===================================================================================================================
struct Hopper {
    static constexpr int NQ = 4, NU = 2;
    double mass_body = 3.0, inertia_body = 0.75, mass_foot = 1.0, gravity = 9.81;
    double body_radius = 0.1, foot_radius = 0.05, leg_len_max = 1.0, leg_len_min = 0.25;

    // L = ½ mb |ṗ_body|² + ½ Ib ṫ² + ½ mf |ṗ_foot|² − mb g z − mf g z_foot, foot = body + r (sin t, −cos t)
    // (foot kinematics pinned by examples/hopper.jl:178: q=[0, 0.5+foot_radius, 0, 0.5] puts the foot on the ground).
    // M(q) = ∂²L/∂q̇², C(q,q̇) = (∂²L/∂q̇∂q) q̇ − ∂L/∂q — closed forms, checked against sympy differentiation of the Lagrangian in tests/test_oracle.py::test_hopper_closed_form_bias_matches_lagrangian.
    template <class S> void momentum(const S* q, const S* v, S* p) const {
        S s = od_sin(q[2]), c = od_cos(q[2]);
        const double mb = mass_body, mf = mass_foot;
        p[0] = (mb + mf) * v[0] + mf * (q[3] * c * v[2] + s * v[3]);
        p[1] = (mb + mf) * v[1] + mf * (q[3] * s * v[2] - c * v[3]);
        p[2] = (inertia_body + mf * q[3] * q[3]) * v[2] + mf * q[3] * (c * v[0] + s * v[1]);
        p[3] = mf * v[3] + mf * (s * v[0] - c * v[1]);
    }
    template <class S> void bias(const S* q, const S* v, S* C) const {
        S s = od_sin(q[2]), c = od_cos(q[2]);
        const double mb = mass_body, mf = mass_foot, g = gravity;
        C[0] = mf * (2.0 * c * v[2] * v[3] - q[3] * s * v[2] * v[2]);
        C[1] = mf * (2.0 * s * v[2] * v[3] + q[3] * c * v[2] * v[2]) + (mb + mf) * g;
        C[2] = 2.0 * mf * q[3] * v[3] * v[2] + (mf * g) * q[3] * s;
        C[3] = -(mf * q[3] * v[2] * v[2]) - (mf * g) * c;
    }
    template <class S> void residual(const S* z, const S* th, const S& kappa, S* r) const {
        const S* q0 = th; const S* q1 = th + 4; const S* u = th + 8; const S* mu = th + 10; const S& h = th[12];
        const S* q2 = z; const S* gam = z + 4; const S* s = z + 8;
        const S* psi = z + 12; const S* b = z + 14; const S* spsi = z + 16; const S* sb = z + 18;
        S d[4];
        variational_integrator<Hopper, S, 4>(*this, q0, q1, q2, h, d);
        // input Jacobian B(qm2) = [0 0 1 0; −sin t, cos t, 0, 1] (u = body torque, leg force)
        S tm = (q1[2] + q2[2]) * 0.5;
        S sm = od_sin(tm), cm = od_cos(tm);
        d[0] = d[0] - sm * u[1];
        d[1] = d[1] + cm * u[1];
        d[2] = d[2] + u[0];
        d[3] = d[3] + u[1];
        // contact impulses λ = J(q2)' [b_body; γ1; b_foot; γ2; γ3; γ4], λ[3] += body_radius b_body
        // (examples/comparisons/hopper.jl:25-30); J rows: body x, body z, foot x, foot z, +r, −r.
        S st = od_sin(q2[2]), ct = od_cos(q2[2]);
        S Jfx[4] = {S(1.0), S(0.0), q2[3] * ct, st};
        S Jfz[4] = {S(0.0), S(1.0), q2[3] * st, -ct};
        for (int i = 0; i < 4; ++i) d[i] = d[i] + Jfx[i] * b[1] + Jfz[i] * gam[1];
        d[0] = d[0] + b[0];
        d[1] = d[1] + gam[0];
        d[2] = d[2] + body_radius * b[0];
        d[3] = d[3] + gam[2] - gam[3];
        for (int i = 0; i < 4; ++i) r[i] = d[i];
        // signed distances: body–ground, foot–ground, leg length lower / upper bound
        r[4] = s[0] - (q2[1] - body_radius);
        r[5] = s[1] - (q2[1] - q2[3] * ct - foot_radius);
        r[6] = s[2] - (q2[3] - leg_len_min);
        r[7] = s[3] - (leg_len_max - q2[3]);
        r[8] = psi[0] - mu[0] * gam[0];
        r[9] = psi[1] - mu[1] * gam[1];
        // tangential velocities (examples/comparisons/hopper.jl:152-155)
        S v[4];
        for (int i = 0; i < 4; ++i) v[i] = (q2[i] - q1[i]) / h;
        S vT_body = v[0] + body_radius * v[2];
        S vT_foot = Jfx[0] * v[0] + Jfx[1] * v[1] + Jfx[2] * v[2] + Jfx[3] * v[3];
        r[10] = vT_body - sb[0];
        r[11] = vT_foot - sb[1];
        for (int i = 0; i < 4; ++i) r[12 + i] = gam[i] * s[i] - kappa;
        for (int i = 0; i < 2; ++i) {
            S uu[2] = {psi[i], b[i]}, vv[2] = {spsi[i], sb[i]}, cp[2];
            cone_product<S, 2>(uu, vv, cp);
            r[16 + 2 * i] = cp[0] - kappa;
            r[17 + 2 * i] = cp[1];
        }
    }
};

// ===================================================================================================================
// Rocket: implicit midpoint on a 12-state MRP rigid body.  z = y(12), θ = [x(12), u(3), h].
// ===================================================================================================================
struct Rocket {
    double mass = 1.0, len = 1.0, g[3] = {0.0, 0.0, -9.81};
    double J[3] = {1.0 / 12.0, 1.0 / 12.0, 1.0e-5};  // model.jl:40 (mass = len = 1)

    template <class S> void f(const S* x, const S* u, S* out) const {  // model.jl:14-33
        const S* r = x + 3; const S* v = x + 6; const S* w = x + 9;
        for (int i = 0; i < 3; ++i) out[i] = v[i];
        S rr = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        S wr = w[0] * r[0] + w[1] * r[1] + w[2] * r[2];
        S cx[3] = {w[1] * r[2] - w[2] * r[1], w[2] * r[0] - w[0] * r[2], w[0] * r[1] - w[1] * r[0]};  // ω × r
        for (int i = 0; i < 3; ++i) out[3 + i] = 0.25 * ((1.0 - rr) * w[i] - 2.0 * cx[i] + 2.0 * wr * r[i]);
        // Rotations.jl MRP → rotation matrix: R = I + (8 [r]×² + 4 (1 − r'r) [r]×) / (1 + r'r)²
        S den = (1.0 + rr) * (1.0 + rr);
        S a = 8.0 / den, bb = 4.0 * (1.0 - rr) / den;
        // [r]× F and [r]×([r]× F)
        S rF[3] = {r[1] * u[2] - r[2] * u[1], r[2] * u[0] - r[0] * u[2], r[0] * u[1] - r[1] * u[0]};
        S rrF[3] = {r[1] * rF[2] - r[2] * rF[1], r[2] * rF[0] - r[0] * rF[2], r[0] * rF[1] - r[1] * rF[0]};
        for (int i = 0; i < 3; ++i) out[6 + i] = g[i] + (1.0 / mass) * (u[i] + a * rrF[i] + bb * rF[i]);
        S tau[3] = {len * u[1], -len * u[0], S(0.0)};
        S Jw[3] = {J[0] * w[0], J[1] * w[1], J[2] * w[2]};
        S wJw[3] = {w[1] * Jw[2] - w[2] * Jw[1], w[2] * Jw[0] - w[0] * Jw[2], w[0] * Jw[1] - w[1] * Jw[0]};
        for (int i = 0; i < 3; ++i) out[9 + i] = (1.0 / J[i]) * (tau[i] - wJw[i]);
    }
    template <class S> void residual(const S* z, const S* th, const S& /*kappa*/, S* r) const {  // codegen.jl:14-22
        const S* x = th; const S* u = th + 12; const S& h = th[15];
        S xm[12], fm[12];
        for (int i = 0; i < 12; ++i) xm[i] = 0.5 * (x[i] + z[i]);
        f<S>(xm, u, fm);
        for (int i = 0; i < 12; ++i) r[i] = z[i] - (x[i] + h * fm[i]);
    }
};

// Rocket thrust-limit projection: z = [u(3), p, s, w, y, v(3)] (10), θ = [ū(3), u_max].  codegen.jl:45-64
struct RocketProjection {
    template <class S> void residual(const S* z, const S* th, const S& kappa, S* r) const {
        const S* u = z; const S& p = z[3]; const S& s = z[4]; const S& w = z[5]; const S& y = z[6]; const S* v = z + 7;
        r[0] = u[0] - th[0] - v[0];
        r[1] = u[1] - th[1] - v[1];
        r[2] = u[2] - th[2] - v[2] - (y + p);
        r[3] = th[3] - u[2] - s;
        r[4] = -y - w;
        r[5] = w * s - kappa;
        r[6] = p * u[2] - kappa;
        S uu[3] = {u[2], u[0], u[1]}, vv[3] = {v[2], v[0], v[1]}, cp[3];
        cone_product<S, 3>(uu, vv, cp);
        r[7] = cp[0] - kappa; r[8] = cp[1]; r[9] = cp[2];
    }
};

}  // namespace od_oracle
