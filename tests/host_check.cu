// TEST INFRASTRUCTURE — not part of liboptdyn_b200.so.
// Runs the product's solver templates (csrc/contact_ip.cuh, dense_ip.cuh, rocket.cuh — the exact code the CUDA kernels
// instantiate) on the host CPU, one problem after another, so that the CPU-only test tier (`pytest -m "not gpu"`, run in a
// container without a GPU) can check the condensed-Newton algebra against the oracle before any GPU time is spent.
// Built by tests/conftest.py with:  nvcc -O2 -std=c++17 -Xcompiler -fPIC -shared -o tests/_build/libhostcheck.so tests/host_check.cu
#include <string.h>
#include <stdint.h>
#include "../optimization_dynamics_b200/csrc/rocket.cuh"
#include "../optimization_dynamics_b200/csrc/riccati.cuh"
#include <vector>

using namespace od;

// ---- lane team: T host threads in lockstep stand in for the lanes of one warp (HostLaneTeam, csrc/group_gj.cuh) ---------------
#include <pthread.h>
#include <thread>
struct TeamShared {
    int T; pthread_barrier_t bar; double bufd[32]; unsigned bufu[32]; int bufp[32];
    explicit TeamShared(int t) : T(t) { pthread_barrier_init(&bar, nullptr, (unsigned)t); }
    ~TeamShared() { pthread_barrier_destroy(&bar); }
    void wait() { pthread_barrier_wait(&bar); }
};
struct LaneView : HostLaneTeam {
    TeamShared* S;
    double shfl_f64(double v, int src) override { S->bufd[lane] = v; S->wait(); const double r = S->bufd[src]; S->wait(); return r; }
    unsigned shfl_u32(unsigned v, int src) override { S->bufu[lane] = v; S->wait(); const unsigned r = S->bufu[src]; S->wait(); return r; }
    bool any(bool p) override { S->bufp[lane] = p ? 1 : 0; S->wait(); int r = 0; for (int l = 0; l < S->T; ++l) r |= S->bufp[l]; S->wait(); return r != 0; }
    void sync() override { S->wait(); }
};
// Runs fn(lane) on T lock-stepped threads with a lane team installed (every thread must make the same sequence of collective calls,
// which is what the warp-synchronous design of the kernels guarantees).
template <class F> static void run_team(int T, F fn) {
    TeamShared sh(T);
    std::vector<std::thread> th;
    for (int lane = 0; lane < T; ++lane)
        th.emplace_back([&sh, lane, &fn]() {
            LaneView v; v.S = &sh; v.lane = lane;
            host_lane_team() = &v;
            fn(lane);
            host_lane_team() = nullptr;
        });
    for (auto& t : th) t.join();
}
// The register path with G lanes per problem: one emulated warp (32 lanes = 32/G problems) after the other, exactly the index
// arithmetic of contact_step_kernel (padding lanes of the last warp repeat the last problem).
template <class M, int G> static void run_contact_lanes(const StepArgs& a) {
    constexpr int PPB = 32 / G;
    typedef ContactIP<M, G, PPB, true> IP;
    std::vector<double> ws((size_t)PPB * IP::WS_SLOT + 2);
    double* base = ws.data(); if (reinterpret_cast<uintptr_t>(base) & 15) ++base;        // 16-byte aligned like dynamic shared memory
    const int nwarps = (a.B + PPB - 1) / PPB;
    for (int w = 0; w < nwarps; ++w)
        run_team(32, [&](int lane) {
            const int slot = lane / G, g = lane % G;
            int i = w * PPB + slot; if (i >= a.B) i = a.B - 1;
            contact_step_one<M, G, PPB, true>(a, i, base + slot * IP::WS_SLOT, g, 0xffffffffu);
        });
}

template <class M> static void run_contact(const StepArgs& a, int reg) {
    if constexpr (M::NC + M::NP > 0) {           // cone models have the cooperative register path: reg = 4 / 8 / 16 selects the lane count
        if (reg == 4) { run_contact_lanes<M, 4>(a); return; }
        if (reg == 8) { run_contact_lanes<M, 8>(a); return; }
        if (reg == 16) { run_contact_lanes<M, 16>(a); return; }
    }
    if (reg) {   // register-path algebra (group_gj.cuh) with one lane: shuffles are identities, the Gauss–Jordan arithmetic is the same
        alignas(16) double ws[ContactIP<M, 1, 1, true>::WS];
        for (int i = 0; i < a.B; ++i) contact_step_one<M, 1, 1, true>(a, i, ws, 0, 0u);
    } else {
        alignas(16) double ws[ContactIP<M>::WS];
        for (int i = 0; i < a.B; ++i) contact_step_one<M, 1, 1>(a, i, ws, 0, 0u);
    }
}

extern "C" int hc_contact_step(int model, int B, const double* q1, const double* q2, const double* u, int nq, int nu, double h, const double* fric,
                               double r_tol, double k_eval, double k_grad, int max_iter, int max_ls, int want_eval, int want_grad,
                               const double* eta, int n_eta, double* q3, double* dq1, double* dq2, double* du, int* status, int* iters, int reg) {
    StepArgs a; memset(&a, 0, sizeof(a));
    a.B = B; a.q1 = q1; a.q2 = q2; a.u = u; a.in_stride_q = nq; a.in_stride_u = nu;
    a.q3 = q3; a.dq1 = want_grad ? dq1 : nullptr; a.dq2 = dq2; a.du = du;
    a.out_stride_q3 = nq; a.out_stride_dq = nq * nq; a.out_stride_du = nq * nu;
    a.status = status; a.iters = iters; a.h = h; a.want_eval = want_eval; a.want_grad = want_grad; a.eta = eta; a.n_eta = n_eta;
    for (int k = 0; k < 4; ++k) a.fric[k] = fric ? fric[k] : 0.0;
    a.opts.r_tol = r_tol; a.opts.kappa_eval_tol = k_eval; a.opts.kappa_grad_tol = k_grad; a.opts.ls_scale = 0.5; a.opts.max_iter = max_iter; a.opts.max_ls = max_ls;
    switch (model) {
        case 0: run_contact<AcrobotImpactModel>(a, reg); break;
        case 1: run_contact<AcrobotNominalModel>(a, reg); break;
        case 2: run_contact<CartpoleFrictionModel>(a, reg); break;
        case 3: run_contact<CartpoleFrictionlessModel>(a, reg); break;
        case 4: run_contact<PlanarPushModel>(a, reg); break;
        case 5: run_contact<HopperModel>(a, reg); break;
        default: return 1;
    }
    return 0;
}

template <class M, int G> static void run_rollout_lanes(const RolloutArgs& a) {
    constexpr int PPB = 32 / G;
    typedef ContactIP<M, G, PPB, true> IP;
    std::vector<double> ws((size_t)PPB * IP::WS_SLOT + 2);
    double* base = ws.data(); if (reinterpret_cast<uintptr_t>(base) & 15) ++base;
    const int nwarps = (a.R + PPB - 1) / PPB;
    for (int w = 0; w < nwarps; ++w)
        run_team(32, [&](int lane) {
            const int slot = lane / G, g = lane % G;
            int r = w * PPB + slot; if (r >= a.R) r = a.R - 1;
            contact_rollout_one<M, G, PPB, true>(a, r, base + slot * IP::WS_SLOT, g, 0xffffffffu);
        });
}
template <class M> static void run_rollout(const RolloutArgs& a, int reg) {
    if constexpr (M::NC + M::NP > 0) {
        if (reg == 4) { run_rollout_lanes<M, 4>(a); return; }
        if (reg == 8) { run_rollout_lanes<M, 8>(a); return; }
    }
    if (reg) { alignas(16) double ws[ContactIP<M, 1, 1, true>::WS]; for (int r = 0; r < a.R; ++r) contact_rollout_one<M, 1, 1, true>(a, r, ws, 0, 0u); }
    else { alignas(16) double ws[ContactIP<M>::WS]; for (int r = 0; r < a.R; ++r) contact_rollout_one<M, 1, 1, false>(a, r, ws, 0, 0u); }
}
extern "C" int hc_contact_rollout(int model, int R, int T, const double* x1, const double* ubar, long long ubar_stride, const double* xbar, const double* K,
                                  const double* kff, const double* alpha, double h, const double* fric, double r_tol, double k_eval,
                                  double* X, double* U, int* status, int reg) {
    RolloutArgs a; memset(&a, 0, sizeof(a));
    a.R = R; a.T = T; a.x1 = x1; a.ubar = ubar; a.ubar_stride = ubar_stride; a.xbar = xbar; a.K = K; a.kff = kff; a.alpha = alpha; a.X = X; a.U = U;
    a.status = status; a.h = h;
    for (int k = 0; k < 4; ++k) a.fric[k] = fric ? fric[k] : 0.0;
    a.opts.r_tol = r_tol; a.opts.kappa_eval_tol = k_eval; a.opts.kappa_grad_tol = k_eval; a.opts.ls_scale = 0.5; a.opts.max_iter = 100; a.opts.max_ls = 25;
    switch (model) {
        case 0: run_rollout<AcrobotImpactModel>(a, reg); break;
        case 2: run_rollout<CartpoleFrictionModel>(a, reg); break;
        case 4: run_rollout<PlanarPushModel>(a, reg); break;
        case 5: run_rollout<HopperModel>(a, reg); break;
        default: return 1;
    }
    return 0;
}

extern "C" int hc_riccati(int NT, int T, int nq, int nu, const double* jac, const double* lx, const double* lu, const double* lxx, const double* luu,
                          const double* lux, double reg, double* K, double* k, double* dV, int* status) {
    RiccatiArgs a; memset(&a, 0, sizeof(a));
    a.NT = NT; a.T = T; a.jac = jac; a.lx = lx; a.lu = lu; a.lxx = lxx; a.luu = luu; a.lux = lux; a.reg = reg;
    a.K = K; a.k = k; a.dV = dV; a.status = status;
    if (nq == 4 && nu == 2) { std::vector<double> ws(Riccati<4, 2, 1>::WS); for (int tr = 0; tr < NT; ++tr) Riccati<4, 2, 1>::run(a, tr, ws.data(), 0); }
    else if (nq == 2 && nu == 1) { std::vector<double> ws(Riccati<2, 1, 1>::WS); for (int tr = 0; tr < NT; ++tr) Riccati<2, 1, 1>::run(a, tr, ws.data(), 0); }
    else if (nq == 5 && nu == 2) { std::vector<double> ws(Riccati<5, 2, 1>::WS); for (int tr = 0; tr < NT; ++tr) Riccati<5, 2, 1>::run(a, tr, ws.data(), 0); }
    else return 1;
    return 0;
}

// rocket_kernel_g with G lanes per problem on emulated warps (same index arithmetic as the kernel)
template <int G> static void run_rocket_lanes(const RocketArgs& a) {
    constexpr int PPB = 32 / G;
    std::vector<double> ws((size_t)PPB * RocketG<G>::WS + 2);
    double* base = ws.data(); if (reinterpret_cast<uintptr_t>(base) & 15) ++base;
    const int nwarps = (a.B + PPB - 1) / PPB;
    for (int w = 0; w < nwarps; ++w)
        run_team(32, [&](int lane) {
            const int slot = lane / G, g = lane % G;
            int i = w * PPB + slot; if (i >= a.B) i = a.B - 1;
            RocketG<G>::run(a, i, base + slot * RocketG<G>::WS, g);
        });
}

extern "C" int hc_rocket(int B, const double* x, const double* u, double h, double u_max, int proj, int want_grad, int proj_only,
                         double* y, double* dx, double* du, double* uproj, double* duproj, int* status, int* iters, int reg) {
    RocketArgs a; memset(&a, 0, sizeof(a));
    a.B = B; a.x = x; a.u = u; a.y = y; a.dx = dx; a.du = du; a.uproj = uproj; a.duproj = duproj; a.status = status; a.iters = iters;
    a.h = h; a.u_max = u_max; a.proj = proj; a.want_grad = want_grad; a.proj_only = proj_only;
    a.opts.r_tol = 1e-8; a.opts.kappa_eval_tol = 1e-4; a.opts.kappa_grad_tol = 1e-4; a.opts.ls_scale = 0.5; a.opts.max_iter = 100; a.opts.max_ls = 25;
    if (reg == 4) run_rocket_lanes<4>(a);
    else if (reg == 8) run_rocket_lanes<8>(a);
    else if (reg) { alignas(16) double ws[RocketG<1>::WS]; for (int i = 0; i < B; ++i) RocketG<1>::run(a, i, ws, 0); }
    else for (int i = 0; i < B; ++i) rocket_one(a, i);
    return 0;
}

// residual blocks [d | rs | rpsi | rv | rgam | rc0 | rc1] and the condensed Newton direction for r at z (oracle z ordering q,γ,s,ψ,b,sψ,sb)
template <class M> static void resid_t(const double* zz, const double* th, double* out, double* dir) {
    typedef ContactIP<M> IP;
    typename IP::Z z; typename IP::R r; double rv, kv;
    const double* p = zz;
    for (int i = 0; i < M::NQ; ++i) z.q[i] = *p++;
    for (int i = 0; i < M::NC; ++i) z.gam[i] = *p++;
    for (int i = 0; i < M::NC; ++i) z.s[i] = *p++;
    for (int i = 0; i < M::NP; ++i) z.psi[i] = *p++;
    for (int i = 0; i < M::NB; ++i) z.b[i] = *p++;
    for (int i = 0; i < M::NP; ++i) z.spsi[i] = *p++;
    for (int i = 0; i < M::NB; ++i) z.sb[i] = *p++;
    double trc[IP::NTC1], trv[IP::NTV1];
    M::trig_const(th, trc); M::trig_var(z.q, th, trv);
    IP::residual(z, th, trc, trv, r, rv, kv);
    double* o = out;
    for (int i = 0; i < M::NQ; ++i) *o++ = r.d[i];
    for (int i = 0; i < M::NC; ++i) *o++ = r.rs[i];
    for (int i = 0; i < M::NP; ++i) *o++ = r.rpsi[i];
    for (int i = 0; i < M::NB; ++i) *o++ = r.rv[i];
    for (int i = 0; i < M::NC; ++i) *o++ = r.rgam[i];
    for (int i = 0; i < M::NP; ++i) *o++ = r.rc0[i];
    for (int i = 0; i < M::NB; ++i) *o++ = r.rc1[i];
    typename IP::Lin L; typename IP::Z D; double ws[IP::WS];
    L.ws = ws; L.g = 0; L.gmask = 0u;
    IP::linearize(z, th, trc, trv, r, L);
    IP::solve(L, z, r, D);
    o = dir;
    for (int i = 0; i < M::NQ; ++i) *o++ = D.q[i];
    for (int i = 0; i < M::NC; ++i) *o++ = D.gam[i];
    for (int i = 0; i < M::NC; ++i) *o++ = D.s[i];
    for (int i = 0; i < M::NP; ++i) *o++ = D.psi[i];
    for (int i = 0; i < M::NB; ++i) *o++ = D.b[i];
    for (int i = 0; i < M::NP; ++i) *o++ = D.spsi[i];
    for (int i = 0; i < M::NB; ++i) *o++ = D.sb[i];
}
extern "C" int hc_contact_residual(int model, const double* z, const double* th, double* out, double* dir) {
    switch (model) {
        case 0: resid_t<AcrobotImpactModel>(z, th, out, dir); break;
        case 1: resid_t<AcrobotNominalModel>(z, th, out, dir); break;
        case 2: resid_t<CartpoleFrictionModel>(z, th, out, dir); break;
        case 3: resid_t<CartpoleFrictionlessModel>(z, th, out, dir); break;
        case 4: resid_t<PlanarPushModel>(z, th, out, dir); break;
        case 5: resid_t<HopperModel>(z, th, out, dir); break;
        default: return 1;
    }
    return 0;
}

// IFT sensitivities of the product code at a given z (oracle ordering)
template <class M> static void sens_t(const double* zz, const double* th, double* dq1, double* dq2, double* du) {
    typedef ContactIP<M> IP;
    typename IP::Z z;
    const double* p = zz;
    for (int i = 0; i < M::NQ; ++i) z.q[i] = *p++;
    for (int i = 0; i < M::NC; ++i) z.gam[i] = *p++;
    for (int i = 0; i < M::NC; ++i) z.s[i] = *p++;
    for (int i = 0; i < M::NP; ++i) z.psi[i] = *p++;
    for (int i = 0; i < M::NB; ++i) z.b[i] = *p++;
    for (int i = 0; i < M::NP; ++i) z.spsi[i] = *p++;
    for (int i = 0; i < M::NB; ++i) z.sb[i] = *p++;
    typename IP::Lin L; double ws[IP::WS];
    L.ws = ws; L.g = 0; L.gmask = 0u;
    double trc[IP::NTC1], trv[IP::NTV1];
    M::trig_const(th, trc); M::trig_var(z.q, th, trv);
    IP::assemble(z, th, trc, trv, L); IP::factor(L);
    IP::sensitivities(L, z, th, trc, trv, dq1, dq2, du);
}
extern "C" int hc_contact_sens(int model, const double* z, const double* th, double* dq1, double* dq2, double* du) {
    switch (model) {
        case 0: sens_t<AcrobotImpactModel>(z, th, dq1, dq2, du); break;
        case 2: sens_t<CartpoleFrictionModel>(z, th, dq1, dq2, du); break;
        case 4: sens_t<PlanarPushModel>(z, th, dq1, dq2, du); break;
        case 5: sens_t<HopperModel>(z, th, dq1, dq2, du); break;
        default: return 1;
    }
    return 0;
}

template <class M> static void K_t(const double* zz, const double* th, double* Kout) {
    typedef ContactIP<M> IP;
    typename IP::Z z;
    const double* p = zz;
    for (int i = 0; i < M::NQ; ++i) z.q[i] = *p++;
    for (int i = 0; i < M::NC; ++i) z.gam[i] = *p++;
    for (int i = 0; i < M::NC; ++i) z.s[i] = *p++;
    for (int i = 0; i < M::NP; ++i) z.psi[i] = *p++;
    for (int i = 0; i < M::NB; ++i) z.b[i] = *p++;
    for (int i = 0; i < M::NP; ++i) z.spsi[i] = *p++;
    for (int i = 0; i < M::NB; ++i) z.sb[i] = *p++;
    typename IP::Lin L; double ws[IP::WS];
    L.ws = ws; L.g = 0; L.gmask = 0u;
    double trc[IP::NTC1], trv[IP::NTV1];
    M::trig_const(th, trc); M::trig_var(z.q, th, trv);
    IP::assemble(z, th, trc, trv, L);
    for (int i = 0; i < IP::NR; ++i) for (int j = 0; j < IP::NR; ++j) Kout[i * IP::NR + j] = L.K(i, j);
}
extern "C" int hc_contact_K(int model, const double* z, const double* th, double* K) {
    switch (model) {
        case 4: K_t<PlanarPushModel>(z, th, K); break;
        case 5: K_t<HopperModel>(z, th, K); break;
        default: return 1;
    }
    return 0;
}

// fastmath.cuh on the host: the same sin/cos code the kernels run (the reciprocal and the bit tricks have host twins there).
extern "C" int hc_sincos(int n, const double* x, double* s, double* c) {
    for (int i = 0; i < n; ++i) od_sincos(x[i], s + i, c + i);
    return 0;
}

// Shared-memory layout of the register path for one (model, lanes) configuration: out = {NR, PW (row pitch, doubles), workspace doubles
// per problem, rows per lane} — tests/test_host_logic.py replays the row moves of a quarter-warp against the 32 banks.
template <class M, int G>
static void layout_t(int* out) { typedef ContactIP<M, G, 32 / G, true> IP; out[0] = IP::NR; out[1] = IP::PW; out[2] = IP::WS_SLOT; out[3] = IP::RPL; }
extern "C" int hc_layout(int model, int lanes, int* out) {
    switch (model * 100 + lanes) {
        case 4: layout_t<AcrobotImpactModel, 4>(out); return 0;
        case 204: layout_t<CartpoleFrictionModel, 4>(out); return 0;
        case 404: layout_t<PlanarPushModel, 4>(out); return 0;
        case 408: layout_t<PlanarPushModel, 8>(out); return 0;
        case 416: layout_t<PlanarPushModel, 16>(out); return 0;
        case 432: layout_t<PlanarPushModel, 32>(out); return 0;
        case 504: layout_t<HopperModel, 4>(out); return 0;
        case 508: layout_t<HopperModel, 8>(out); return 0;
        case 516: layout_t<HopperModel, 16>(out); return 0;
    }
    return 1;
}
