// ORACLE — TEST INFRASTRUCTURE ONLY (see dual.hpp header).  PARITY UNPINNED (see oracle/README.md).
//
// C entry points (ctypes) around the restated solver: the semantic of RoboDojo.step! as driven by the reference's
// f / fx / fu (src/dynamics.jl:81-128), gradient bundle (src/gradient_bundle.jl:87-147 + src/ls.jl:20-60) and rocket
// wrappers (src/models/rocket/dynamics.jl:101-269).  Matrices are column-major (Julia layout), one problem after another.
#include <cstring>
#include <cstdio>
#include <vector>
#include <atomic>
#include <thread>
#include "ip.hpp"
#include "models.hpp"

using namespace od_oracle;

namespace {

// dynamic-chunk parallel loop over [0,B) on `nthreads` std::threads (nthreads<=0: all hardware threads)
template <class F>
void parallel_for(int B, int nthreads, int chunk, F&& body) {
    int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt == 1 || B <= chunk) { for (int i = 0; i < B; ++i) body(i); return; }
    std::atomic<int> next(0);
    auto worker = [&]() { for (;;) { int s = next.fetch_add(chunk); if (s >= B) return; int e = std::min(B, s + chunk); for (int i = s; i < e; ++i) body(i); } };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
}

std::vector<int> range(int a, int b) { std::vector<int> v; for (int i = a; i < b; ++i) v.push_back(i); return v; }

// ---- per-model static description --------------------------------------------------------------------------------
enum ModelId { ACROBOT_IMPACT = 0, ACROBOT_NOMINAL = 1, CARTPOLE_FRICTION = 2, CARTPOLE_FRICTIONLESS = 3, PLANAR_PUSH = 4, HOPPER = 5,
               ROCKET = 6, ROCKET_PROJ = 7 };

ConeIndex idx_acrobot_impact() {   // simulator_impact.jl:20-31
    ConeIndex c; c.ort_p = {2, 3}; c.ort_d = {4, 5}; c.equr = range(0, 4); c.ortr = {4, 5}; return c; }
ConeIndex idx_none(int n) { ConeIndex c; c.equr = range(0, n); return c; }
ConeIndex idx_cartpole_friction() {  // simulator_friction.jl:22-33
    ConeIndex c; c.soc_p = {{2, 4}, {3, 5}}; c.soc_d = {{6, 8}, {7, 9}}; c.equr = range(0, 6); c.socr = {{6, 7}, {8, 9}}; return c; }
ConeIndex idx_planar_push() {  // simulator.jl:19-49
    ConeIndex c; c.ort_p = {5}; c.ort_d = {6};
    for (int i = 0; i < 4; ++i) { c.soc_p.push_back({7 + i, 12 + 2 * i, 13 + 2 * i}); c.soc_d.push_back({21 + i, 26 + 2 * i, 27 + 2 * i}); c.socr.push_back({21 + 3 * i, 22 + 3 * i, 23 + 3 * i}); }
    c.soc_p.push_back({11, 20}); c.soc_d.push_back({25, 34}); c.socr.push_back({33, 34});
    c.equr = range(0, 20); c.ortr = {20}; return c; }
ConeIndex idx_hopper() {  // same pattern as planar push
    ConeIndex c; c.ort_p = {4, 5, 6, 7}; c.ort_d = {8, 9, 10, 11};
    c.soc_p = {{12, 14}, {13, 15}}; c.soc_d = {{16, 18}, {17, 19}}; c.socr = {{16, 17}, {18, 19}};
    c.equr = range(0, 12); c.ortr = {12, 13, 14, 15}; return c; }
ConeIndex idx_rocket_proj() {  // rocket/dynamics.jl:52-63 (1-based → 0-based)
    ConeIndex c; c.ort_p = {4, 2}; c.ort_d = {5, 3}; c.soc_p = {{2, 0, 1}}; c.soc_d = {{9, 7, 8}};
    c.equr = range(0, 5); c.ortr = {5, 6}; c.socr = {{7, 8, 9}}; return c; }

// Readings of the unpinned solver choices (ip.hpp Options::tau_rule …): process-wide, set by od_oracle_set_variant.  Tests only.
int g_variant[4] = {0, 0, 0, 0};
void apply_variant(Options& o) { o.tau_rule = g_variant[0]; o.apply_reg = g_variant[1]; o.mu_mode = g_variant[2]; o.soc_tau_cap = g_variant[3]; }

Options contact_opts(double r_tol, double kappa_tol, bool diff) {  // src/dynamics.jl:25-33
    Options o; apply_variant(o); o.undercut = std::numeric_limits<double>::infinity(); o.gamma_reg = 0.1; o.r_tol = r_tol; o.kappa_tol = kappa_tol;
    o.max_ls = 25; o.eps_min = 0.25; o.diff_sol = diff; return o; }

struct StepOut { double* q3; double* dq1; double* dq2; double* du; double* dz_full; SolveInfo info; };

// One RoboDojo.step!: q1' = q2 − h v1 with v1 = (q2 − q1)/h (src/dynamics.jl:82-88), initialize_z!, θ pack, solve, unpack.
template <class Model, int NZ, int NTH, int NQ, int NU>
SolveInfo step_generic(const Model& model, const ConeIndex& idx, const Options& opts, const double* q1, const double* q2, const double* u,
                       const double* fric, int nfric, double h, double cone_init_scalar, double cone_init_vec,
                       double* q3, double* dq1, double* dq2, double* du, double* dz_full, double* z_out) {
    InteriorPoint<Model, NZ, NTH> ip(model, idx, opts);
    ip.n_out_rows = NQ;
    for (int i = 0; i < NQ; ++i) {
        double v1 = (q2[i] - q1[i]) / h;
        ip.th[i] = q2[i] - h * v1;
        ip.th[NQ + i] = q2[i];
    }
    for (int i = 0; i < NU; ++i) ip.th[2 * NQ + i] = u[i];
    for (int i = 0; i < nfric; ++i) ip.th[2 * NQ + NU + i] = fric[i];
    ip.th[NTH - 1] = h;
    // initialize_z!: q ← q2 (current configuration), orthant/ψ/sψ ← 1, b/sb ← 0.1
    // (acrobot/simulator_impact.jl:34-38, cartpole/simulator_friction.jl:36-42, planar_push/simulator.jl:52-60)
    for (int i = 0; i < NZ; ++i) ip.z[i] = 0.0;
    for (int i = 0; i < NQ; ++i) ip.z[i] = q2[i];
    for (size_t k = 0; k < idx.ort_p.size(); ++k) { ip.z[idx.ort_p[k]] = cone_init_scalar; ip.z[idx.ort_d[k]] = cone_init_scalar; }
    for (size_t c = 0; c < idx.soc_p.size(); ++c) {
        ip.z[idx.soc_p[c][0]] = cone_init_scalar; ip.z[idx.soc_d[c][0]] = cone_init_scalar;
        for (size_t k = 1; k < idx.soc_p[c].size(); ++k) { ip.z[idx.soc_p[c][k]] = cone_init_vec; ip.z[idx.soc_d[c][k]] = cone_init_vec; }
    }
    SolveInfo info = ip.solve();
    if (q3) for (int i = 0; i < NQ; ++i) q3[i] = ip.z[i];
    if (z_out) for (int i = 0; i < NZ; ++i) z_out[i] = ip.z[i];
    if (opts.diff_sol && info.status != 2) {
        for (int i = 0; i < NQ; ++i) {
            for (int j = 0; j < NQ; ++j) {
                if (dq1) dq1[j * NQ + i] = ip.dz[i * NTH + j];
                if (dq2) dq2[j * NQ + i] = ip.dz[i * NTH + NQ + j];
            }
            for (int j = 0; j < NU; ++j) if (du) du[j * NQ + i] = ip.dz[i * NTH + 2 * NQ + j];
        }
        if (dz_full) for (int i = 0; i < NZ * NTH; ++i) dz_full[i] = ip.dz[i];   // row-major NZ×NTH, for the IFT identity test
    }
    return info;
}

struct Dims { int nq, nu, nz, nth, nfric; };
bool dims_of(int model, Dims* d) {
    switch (model) {
        case ACROBOT_IMPACT: *d = {2, 1, 6, 6, 0}; return true;
        case ACROBOT_NOMINAL: *d = {2, 1, 2, 6, 0}; return true;
        case CARTPOLE_FRICTION: *d = {2, 1, 10, 8, 2}; return true;
        case CARTPOLE_FRICTIONLESS: *d = {2, 1, 2, 6, 0}; return true;
        case PLANAR_PUSH: *d = {5, 2, 35, 13, 0}; return true;
        case HOPPER: *d = {4, 2, 20, 13, 2}; return true;
        case ROCKET: case ROCKET_PROJ: *d = {12, 3, 12, 16, 0}; return true;
    }
    return false;
}

SolveInfo step_model(int model, const Options& o, const double* q1, const double* q2, const double* u, const double* fric, double h,
                     double* q3, double* dq1, double* dq2, double* du, double* dzf, double* z_out) {
    static const ConeIndex ai = idx_acrobot_impact(), n2 = idx_none(2), cf = idx_cartpole_friction(), pp = idx_planar_push(), hp = idx_hopper();
    switch (model) {
        case ACROBOT_IMPACT: { Acrobot m; m.impact = true; return step_generic<Acrobot, 6, 6, 2, 1>(m, ai, o, q1, q2, u, nullptr, 0, h, 1.0, 0.1, q3, dq1, dq2, du, dzf, z_out); }
        case ACROBOT_NOMINAL: { Acrobot m; m.impact = false; return step_generic<Acrobot, 2, 6, 2, 1>(m, n2, o, q1, q2, u, nullptr, 0, h, 1.0, 0.1, q3, dq1, dq2, du, dzf, z_out); }
        case CARTPOLE_FRICTION: { Cartpole m; m.friction = true; static const double def[2] = {0.1, 0.1};
            return step_generic<Cartpole, 10, 8, 2, 1>(m, cf, o, q1, q2, u, fric ? fric : def, 2, h, 1.0, 0.1, q3, dq1, dq2, du, dzf, z_out); }
        case CARTPOLE_FRICTIONLESS: { Cartpole m; m.friction = false; return step_generic<Cartpole, 2, 6, 2, 1>(m, n2, o, q1, q2, u, nullptr, 0, h, 1.0, 0.1, q3, dq1, dq2, du, dzf, z_out); }
        case PLANAR_PUSH: { PlanarPush m; return step_generic<PlanarPush, 35, 13, 5, 2>(m, pp, o, q1, q2, u, nullptr, 0, h, 1.0, 0.1, q3, dq1, dq2, du, dzf, z_out); }
        case HOPPER: { Hopper m; static const double def[2] = {0.5, 0.5};
            return step_generic<Hopper, 20, 13, 4, 2>(m, hp, o, q1, q2, u, fric ? fric : def, 2, h, 1.0, 0.1, q3, dq1, dq2, du, dzf, z_out); }
    }
    SolveInfo bad; bad.status = 2; return bad;
}

// soc_projection / soc_projection_gradient (rocket/dynamics.jl:168-210)
SolveInfo rocket_projection(const double* u, double u_max, bool diff, double* up, double* dproj /*3×3 col-major*/) {
    static const ConeIndex ci = idx_rocket_proj();
    Options o; o.r_tol = 1e-8; o.kappa_tol = 1e-4; o.max_ls = 25; o.eps_min = 0.0; o.undercut = std::numeric_limits<double>::infinity();
    o.gamma_reg = 0.0; o.kappa_reg = 0.0; o.diff_sol = diff;   // rocket/dynamics.jl:77-86
    apply_variant(o);
    RocketProjection m;
    InteriorPoint<RocketProjection, 10, 4> ip(m, ci, o);
    ip.n_out_rows = 3;
    for (int i = 0; i < 10; ++i) ip.z[i] = 0.1;
    ip.z[2] += 1.0; ip.z[9] += 1.0; ip.z[6] = 0.0;
    for (int i = 0; i < 3; ++i) ip.th[i] = u[i];
    ip.th[3] = u_max;
    SolveInfo info = ip.solve();
    for (int i = 0; i < 3; ++i) up[i] = ip.z[i];
    if (diff && dproj && info.status != 2) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) dproj[j * 3 + i] = ip.dz[i * 4 + j];
    return info;
}

// f_rocket / fx_rocket / fu_rocket and the *_proj variants (rocket/dynamics.jl:101-163,215-269)
SolveInfo rocket_step(const double* x, const double* u, double h, double u_max, bool proj, bool diff,
                      double* y, double* dx /*12×12*/, double* du /*12×3*/, SolveInfo* proj_info) {
    static const ConeIndex ci = idx_none(12);
    Options o; o.r_tol = 1e-8; o.kappa_tol = 1.0; o.max_ls = 25; o.eps_min = 0.25; o.diff_sol = diff;   // rocket/dynamics.jl:21-27
    Rocket m;
    InteriorPoint<Rocket, 12, 16> ip(m, ci, o);
    double ue[3] = {u[0], u[1], u[2]}, dproj[9];
    if (proj) { SolveInfo pi = rocket_projection(u, u_max, diff, ue, dproj); if (proj_info) *proj_info = pi; }
    for (int i = 0; i < 12; ++i) { ip.z[i] = x[i]; ip.th[i] = x[i]; }
    for (int i = 0; i < 3; ++i) ip.th[12 + i] = ue[i];
    ip.th[15] = h;
    SolveInfo info = ip.solve();
    if (y) for (int i = 0; i < 12; ++i) y[i] = ip.z[i];
    if (diff && info.status != 2) {
        if (dx) for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) dx[j * 12 + i] = ip.dz[i * 16 + j];
        if (du) for (int i = 0; i < 12; ++i) for (int j = 0; j < 3; ++j) {
            if (!proj) du[j * 12 + i] = ip.dz[i * 16 + 12 + j];
            else { double s = 0.0; for (int k = 0; k < 3; ++k) s += ip.dz[i * 16 + 12 + k] * dproj[j * 3 + k]; du[j * 12 + i] = s; }   // mul!(du, du_dyn, du_proj), :267
        }
    }
    return info;
}

// LeastSquares.update! (src/ls.jl:44-60) on c(θ) = Σ_i ‖fη_i − fz − M η_i‖², M = reshape(θ, ny, nz)
int least_squares_fit(int N, int ny, int nz, const double* fz, const double* feta /*N×ny*/, const double* eta /*N×nz*/, double* theta /*ny*nz col-major*/) {
    const int nt = ny * nz;
    std::vector<double> g(nt), H(nt * nt), d(nt);
    std::vector<int> piv(nt);
    auto grad = [&]() {
        std::fill(g.begin(), g.end(), 0.0);
        for (int i = 0; i < N; ++i) {
            for (int a = 0; a < ny; ++a) {
                double res = feta[i * ny + a] - fz[a];
                for (int b = 0; b < nz; ++b) res -= theta[b * ny + a] * eta[i * nz + b];
                for (int b = 0; b < nz; ++b) g[b * ny + a] += -2.0 * res * eta[i * nz + b];
            }
        }
        double m = 0.0; for (double v : g) m = std::max(m, std::fabs(v)); return m;
    };
    double res = grad();
    int iter = 0;
    while (res > 1.0e-8 && iter < 100) {
        std::fill(H.begin(), H.end(), 0.0);
        for (int i = 0; i < N; ++i)
            for (int b = 0; b < nz; ++b) for (int c = 0; c < nz; ++c) {
                double w = 2.0 * eta[i * nz + b] * eta[i * nz + c];
                if (w != 0.0) for (int a = 0; a < ny; ++a) H[(b * ny + a) * nt + (c * ny + a)] += w;
            }
        if (!lu_factor(H.data(), piv.data(), nt)) return 2;
        d = g;
        lu_solve(H.data(), piv.data(), nt, d.data());
        for (int k = 0; k < nt; ++k) theta[k] -= d[k];
        res = grad();
        iter++;
    }
    return 0;
}

}  // namespace

extern "C" {

// tau_rule, apply_reg, mu_mode, soc_tau_cap — see ip.hpp Options.  All zero = the oracle's definition of parity.
void od_oracle_set_variant(int tau_rule, int apply_reg, int mu_mode, int soc_tau_cap) {
    g_variant[0] = tau_rule; g_variant[1] = apply_reg; g_variant[2] = mu_mode; g_variant[3] = soc_tau_cap;
}

int od_oracle_dims(int model, int* nq, int* nu, int* nz, int* nth) {
    Dims d; if (!dims_of(model, &d)) return 1; *nq = d.nq; *nu = d.nu; *nz = d.nz; *nth = d.nth; return 0;
}

// Residual and dense Jacobians at a point (row-major rz NZ×NZ, rθ NZ×NTH) — for model-level parity tests.
int od_oracle_residual(int model, const double* z, const double* th, double kappa, double* r, double* rz, double* rth) {
#define OD_EVAL(MODEL, INIT, NZ, NTH) { MODEL m; INIT; static const ConeIndex ci = idx_none(NZ); Options o; InteriorPoint<MODEL, NZ, NTH> ip(m, ci, o); \
        std::memcpy(ip.th, th, sizeof(double) * NTH); if (r) ip.eval_r(z, kappa, r); if (rz) ip.eval_rz(z, rz); if (rth) ip.eval_rth(z, rth); return 0; }
    switch (model) {
        case ACROBOT_IMPACT: OD_EVAL(Acrobot, m.impact = true, 6, 6)
        case ACROBOT_NOMINAL: OD_EVAL(Acrobot, m.impact = false, 2, 6)
        case CARTPOLE_FRICTION: OD_EVAL(Cartpole, m.friction = true, 10, 8)
        case CARTPOLE_FRICTIONLESS: OD_EVAL(Cartpole, m.friction = false, 2, 6)
        case PLANAR_PUSH: OD_EVAL(PlanarPush, (void)0, 35, 13)
        case HOPPER: OD_EVAL(Hopper, (void)0, 20, 13)
        case ROCKET: OD_EVAL(Rocket, (void)0, 12, 16)
        case ROCKET_PROJ: OD_EVAL(RocketProjection, (void)0, 10, 4)
    }
#undef OD_EVAL
    return 1;
}

// Batched step (+ optional IFT).  diff=0: eval solve only (f).  diff=1: grad solve (fx/fu).  Any output pointer may be NULL.
// q1,q2: B×nq, u: B×nu (row per problem).  Jacobians column-major per problem.  info: B×4 = [iterations, status, ls_steps, _],
// vio: B×5 = [r_vio, κ_vio, margin, ift_spread, q_uncertainty].  dz_full: B×nz×nθ row-major (tests only).  z_out: B×nz.
int od_oracle_step_batch(int model, int B, const double* q1, const double* q2, const double* u, const double* fric, double h,
                         double r_tol, double kappa_tol, int diff,
                         double* q3, double* dq1, double* dq2, double* du, double* dz_full, double* z_out, int* info, double* vio, int nthreads) {
    Dims d; if (!dims_of(model, &d) || model >= ROCKET) return 1;
    Options o = contact_opts(r_tol, kappa_tol, diff != 0);
    o.diagnostics = (vio != nullptr);
    parallel_for(B, nthreads, 16, [&](int i) {
        SolveInfo s = step_model(model, o, q1 + (size_t)i * d.nq, q2 + (size_t)i * d.nq, u + (size_t)i * d.nu, fric, h,
                                 q3 ? q3 + (size_t)i * d.nq : nullptr, dq1 ? dq1 + (size_t)i * d.nq * d.nq : nullptr,
                                 dq2 ? dq2 + (size_t)i * d.nq * d.nq : nullptr, du ? du + (size_t)i * d.nq * d.nu : nullptr,
                                 dz_full ? dz_full + (size_t)i * d.nz * d.nth : nullptr, z_out ? z_out + (size_t)i * d.nz : nullptr);
        if (info) { info[4 * i] = s.iterations; info[4 * i + 1] = s.status; info[4 * i + 2] = s.ls_steps; info[4 * i + 3] = 0; }
        if (vio) { vio[5 * i] = s.r_vio; vio[5 * i + 1] = s.k_vio; vio[5 * i + 2] = s.margin; vio[5 * i + 3] = s.ift_spread; vio[5 * i + 4] = s.q_uncertainty; }
    });
    return 0;
}

// Rocket: x B×12, u B×3 → y B×12, dx B×(12×12), du B×(12×3) column-major.  proj: apply the SOC thrust projection first.
int od_oracle_rocket_batch(int B, const double* x, const double* u, double h, double u_max, int proj, int diff,
                           double* y, double* dx, double* du, double* uproj, int* info, double* vio, int nthreads) {
    parallel_for(B, nthreads, 16, [&](int i) {
        SolveInfo pi;
        SolveInfo s = rocket_step(x + (size_t)i * 12, u + (size_t)i * 3, h, u_max, proj != 0, diff != 0, y ? y + (size_t)i * 12 : nullptr,
                                  dx ? dx + (size_t)i * 144 : nullptr, du ? du + (size_t)i * 36 : nullptr, &pi);
        if (uproj) { if (proj) { double dp[9]; rocket_projection(u + (size_t)i * 3, u_max, false, uproj + (size_t)i * 3, dp); } else for (int k = 0; k < 3; ++k) uproj[3 * i + k] = u[3 * i + k]; }
        if (info) { info[4 * i] = s.iterations; info[4 * i + 1] = s.status | (proj ? pi.status : 0); info[4 * i + 2] = s.ls_steps; info[4 * i + 3] = proj ? pi.iterations : 0; }
        if (vio) { vio[3 * i] = s.r_vio; vio[3 * i + 1] = s.k_vio; vio[3 * i + 2] = proj ? std::min(s.margin, pi.margin) : s.margin; }
    });
    return 0;
}

// soc_projection(+gradient) alone: u B×3 → up B×3, dproj B×(3×3) column-major
int od_oracle_rocket_projection_batch(int B, const double* u, double u_max, int diff, double* up, double* dproj, int* info, double* vio) {
    for (int i = 0; i < B; ++i) {
        SolveInfo s = rocket_projection(u + 3 * i, u_max, diff != 0, up + 3 * i, dproj ? dproj + 9 * i : nullptr);
        if (info) { info[4 * i] = s.iterations; info[4 * i + 1] = s.status; info[4 * i + 2] = s.ls_steps; info[4 * i + 3] = 0; }
        if (vio) { vio[3 * i] = s.r_vio; vio[3 * i + 1] = s.k_vio; vio[3 * i + 2] = s.margin; }
    }
    return 0;
}

// gradient! (src/gradient_bundle.jl:87-104): nominal + N perturbed eval-sim steps, then the LeastSquares Newton fit.
// eta: N×(2nq+nu) fixed perturbations shared by the batch.  dz: B × (nq × (2nq+nu)) column-major.
int od_oracle_bundle_batch(int model, int B, int N, const double* eta, const double* q1, const double* q2, const double* u, const double* fric,
                           double h, double r_tol, double kappa_tol, double* dz, int* status, int nthreads) {
    Dims d; if (!dims_of(model, &d) || model >= ROCKET) return 1;
    Options o = contact_opts(r_tol, kappa_tol, false);
    const int ny = d.nq, nzz = 2 * d.nq + d.nu;
    parallel_for(B, nthreads, 1, [&](int i) {
        std::vector<double> fz(ny), fe((size_t)N * ny), a(d.nq), b(d.nq), c(d.nu), theta((size_t)ny * nzz, 0.0);
        int st = step_model(model, o, q1 + (size_t)i * d.nq, q2 + (size_t)i * d.nq, u + (size_t)i * d.nu, fric, h, fz.data(), 0, 0, 0, 0, 0).status;
        for (int k = 0; k < N; ++k) {
            const double* e = eta + (size_t)k * nzz;
            for (int j = 0; j < d.nq; ++j) { a[j] = q1[(size_t)i * d.nq + j] + e[j]; b[j] = q2[(size_t)i * d.nq + j] + e[d.nq + j]; }
            for (int j = 0; j < d.nu; ++j) c[j] = u[(size_t)i * d.nu + j] + e[2 * d.nq + j];
            st |= step_model(model, o, a.data(), b.data(), c.data(), fric, h, fe.data() + (size_t)k * ny, 0, 0, 0, 0, 0).status;
        }
        st |= least_squares_fit(N, ny, nzz, fz.data(), fe.data(), eta, theta.data());
        std::memcpy(dz + (size_t)i * ny * nzz, theta.data(), sizeof(double) * ny * nzz);
        if (status) status[i] = st;
    });
    return 0;
}

// LeastSquares alone — known-answer test of reference src/ls.jl:62-144
int od_oracle_least_squares(int N, int ny, int nz, const double* fz, const double* feta, const double* eta, double* theta) {
    return least_squares_fit(N, ny, nz, fz, feta, eta, theta);
}

int od_oracle_num_threads() {
    return (int)std::thread::hardware_concurrency();
}

}  // extern "C"
