// ORACLE — TEST INFRASTRUCTURE ONLY (see dual.hpp header).  PARITY UNPINNED (see oracle/README.md).
//
// CPU restatement of the primal-dual interior-point solver + implicit-function-theorem sensitivities that the reference
// reaches through RoboDojo.jl (NOT IN TREE; compat "0.1.2", reference Project.toml:17,31):
//   RoboDojo.interior_point_solve!  — called at reference src/models/rocket/dynamics.jl:109,142,157,178,201,223,247,262
//                                     and, through RoboDojo.step!, at src/dynamics.jl:88,103,123
//   options                         — reference src/dynamics.jl:25-33, src/models/rocket/dynamics.jl:21-27,77-86
//   IndicesOptimization             — reference src/models/acrobot/simulator_impact.jl:20-31,
//                                     src/models/cartpole/simulator_friction.jl:22-33, src/models/planar_push/simulator.jl:19-49,
//                                     src/models/rocket/dynamics.jl:52-63
// The algorithm is the published RoboDojo v0.1.x Mehrotra predictor–corrector as written down in SURVEY.md Appendix A.3.
// Decisions the oracle fixes where the tree is silent (also in DESIGN.md §Oracle):
//   * linear solver: dense LU with partial (row) pivoting — `lu_solver` (reference src/gradient_bundle.jl:76);
//     the `reg` keyword that RoboDojo passes to `linear_solve!` is ignored by its LU solver, so κ_reg/γ_reg have no effect;
//   * central-path measure μ = Σ⟨primal,dual⟩ / (n_orthant + n_soc)  (cone degree, CVXOPT §5.1.3), σ = clamp(μaff/μ,0,1)³;
//   * fraction to the boundary τ = max(0.95, 1 − max(r_vio, κ_vio)²) for orthant and second-order cones alike.  The other
//     reading of SURVEY A.1, τ = 1 − min(ϵ_min, vio²), is ruled out by the reference itself: its rocket projection sets
//     ϵ_min = 0 (src/models/rocket/dynamics.jl:81) ⇒ τ = 1 ⇒ iterates land exactly on the cone boundary; restated that
//     way 41 % of a 1024-sample thrust batch fails to converge.  ϵ_min is therefore accepted and ignored;
//   * SOC step length: exact largest α with u − αΔ ∈ K via the CVXOPT §8.2 scaling formula, then α = min(1, τ·α_max).
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>
#include "dual.hpp"

namespace od_oracle {

struct ConeIndex {                       // 0-based restatement of RoboDojo's IndicesOptimization
    std::vector<int> ort_p, ort_d;       // ortz[1], ortz[2]
    std::vector<std::vector<int>> soc_p, soc_d;  // socz[i][1], socz[i][2]
    std::vector<int> equr;               // equality rows
    std::vector<int> ortr;               // orthant bilinear rows
    std::vector<std::vector<int>> socr;  // SOC bilinear rows per cone (socri)
};

struct Options {                         // RoboDojo InteriorPointOptions defaults
    double r_tol = 1e-5, kappa_tol = 1e-5, ls_scale = 0.5;
    int max_iter = 100, max_ls = 3;
    double eps_min = 0.05, kappa_reg = 1e-3, gamma_reg = 1e-1, undercut = 5.0;
    bool diff_sol = false;
    // ---- readings of the choices the reference tree does not pin (header above).  Defaults = the oracle's definition of parity;
    // the alternatives exist so that tests/test_unpinned_choices.py can MEASURE how far q3 and the sensitivities move under each.
    int tau_rule = 0;          // 0: τ = max(0.95, 1 − vio²);  1: τ = 1 − min(ϵ_min, vio²) (SURVEY A.1 first reading);  2: τ = 0.99 fixed
    int apply_reg = 0;         // 1: rz += reg·I with reg = κ_vio·γ_reg when κ_vio < κ_reg, and max(reg, κ_tol·γ_reg) in the IFT
    int mu_mode = 0;           // 0: μ = Σ⟨p,d⟩ / (#orthant pairs + #second-order cones);  1: / total cone dimension
    int soc_tau_cap = 0;       // 1: second-order-cone step length uses min(τ, 0.99)
    bool diagnostics = false;   // tests only: also report how far the returned iterate is from the exact root (SolveInfo::q_uncertainty)
};

struct SolveInfo {
    int iterations = 0;
    int status = 0;            // 0 converged, 1 max_iter, 2 non-finite / singular
    int ls_steps = 0;          // total backtracking halvings
    double r_vio = 0, k_vio = 0;
    double margin = std::numeric_limits<double>::infinity();  // min relative distance of any discrete decision from flipping
    double q_uncertainty = 0.0; // ‖(rz⁻¹ r(z*;θ,0))[output rows]‖∞: the next Newton step = distance of the returned q3 from the exact root.
                                // The solver stops at the FIRST iterate inside the tolerances, so q3 is only defined up to this amount.
    double ift_spread = 0.0;   // max |δz − δz'| where δz' re-solves with rz's rows rescaled by powers of two (different pivot order, same
                               // exact solution): large ⇒ rz(z*) is numerically singular and the sensitivities are not determined in fp64
};

// Dense LU, partial pivoting, row-major n×n.  Returns false on an exactly-zero pivot / non-finite entry.
inline bool lu_factor(double* A, int* piv, int n) {
    for (int k = 0; k < n; ++k) {
        int p = k; double best = std::fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i) { double a = std::fabs(A[i * n + k]); if (a > best) { best = a; p = i; } }
        piv[k] = p;
        if (!(best > 0.0) || !std::isfinite(best)) return false;
        if (p != k) for (int j = 0; j < n; ++j) std::swap(A[k * n + j], A[p * n + j]);
        double inv = 1.0 / A[k * n + k];
        for (int i = k + 1; i < n; ++i) {
            double l = A[i * n + k] * inv;
            A[i * n + k] = l;
            for (int j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
        }
    }
    return true;
}
inline void lu_solve(const double* LU, const int* piv, int n, double* x) {
    for (int k = 0; k < n; ++k) { if (piv[k] != k) std::swap(x[k], x[piv[k]]); }
    for (int i = 1; i < n; ++i) { double s = x[i]; for (int j = 0; j < i; ++j) s -= LU[i * n + j] * x[j]; x[i] = s; }
    for (int i = n - 1; i >= 0; --i) { double s = x[i]; for (int j = i + 1; j < n; ++j) s -= LU[i * n + j] * x[j]; x[i] = s / LU[i * n + i]; }
}

// Relative distance between the two sides of a comparison.  Two values that are both at rounding-noise level (< 1e-12, e.g. the
// residual of linear equality rows after a full Newton step) compare by noise: such a decision is reported as fragile (gap 0).
inline double rel_gap(double a, double b) {
    if (std::max(std::fabs(a), std::fabs(b)) < 1e-12) return 0.0;
    return std::fabs(a - b) / std::max(std::fabs(b), 1e-300);
}

template <class Model, int NZ, int NTH>
struct InteriorPoint {
    typedef Dual<double, NZ> DZ;
    typedef Dual<double, NTH> DT;
    const Model& model;
    const ConeIndex& idx;
    Options opts;
    double z[NZ], th[NTH], r[NZ], rz[NZ * NZ], rth[NZ * NTH], dz[NZ * NTH];
    int piv[NZ];
    int n_out_rows = NZ;   // rows of δz the caller consumes (the q block); ift_spread is measured on these rows only

    InteriorPoint(const Model& m, const ConeIndex& i, const Options& o) : model(m), idx(i), opts(o) {}

    void eval_r(const double* zz, double kappa, double* out) const { model.template residual<double>(zz, th, kappa, out); }
    void eval_rz(const double* zz, double* out) const {          // Symbolics.jacobian(r, z) — exact via duals
        DZ zd[NZ], td[NTH], rd[NZ];
        for (int i = 0; i < NZ; ++i) zd[i] = DZ::variable(zz[i], i);
        for (int i = 0; i < NTH; ++i) td[i] = DZ(th[i]);
        model.template residual<DZ>(zd, td, DZ(0.0), rd);
        for (int i = 0; i < NZ; ++i) for (int j = 0; j < NZ; ++j) out[i * NZ + j] = rd[i].d[j];
    }
    void eval_rth(const double* zz, double* out) const {         // Symbolics.jacobian(r, θ)
        DT zd[NZ], td[NTH], rd[NZ];
        for (int i = 0; i < NZ; ++i) zd[i] = DT(zz[i]);
        for (int i = 0; i < NTH; ++i) td[i] = DT::variable(th[i], i);
        model.template residual<DT>(zd, td, DT(0.0), rd);
        for (int i = 0; i < NZ; ++i) for (int j = 0; j < NTH; ++j) out[i * NTH + j] = rd[i].d[j];
    }
    double vio(const double* rr, const std::vector<int>& rows) const {
        double v = 0.0; for (int i : rows) v = std::max(v, std::fabs(rr[i])); return v;
    }
    double bil_vio(const double* rr) const {
        double v = vio(rr, idx.ortr);
        for (auto& c : idx.socr) v = std::max(v, vio(rr, c));
        return v;
    }
    bool has_cones() const { return !idx.ort_p.empty() || !idx.soc_p.empty(); }

    // largest α∈[0,1] with z − αΔ inside the cones, scaled by τ
    double step_length(const double* zz, const double* D, double tau) const {
        const double tau_soc = opts.soc_tau_cap ? std::min(tau, 0.99) : tau;
        double a = 1.0;
        for (size_t k = 0; k < idx.ort_p.size(); ++k) {
            int ip = idx.ort_p[k], id = idx.ort_d[k];
            if (D[ip] > 0.0) a = std::min(a, tau * zz[ip] / D[ip]);
            if (D[id] > 0.0) a = std::min(a, tau * zz[id] / D[id]);
        }
        for (size_t c = 0; c < idx.soc_p.size(); ++c) {
            a = std::min(a, soc_step(zz, D, idx.soc_p[c], tau_soc));
            a = std::min(a, soc_step(zz, D, idx.soc_d[c], tau_soc));
        }
        return a;
    }
    // CVXOPT §8.2: with λ in int(K), ρ = scaled(−Δ);  λ − αΔ ∈ K  ⇔  α ≤ 1 / max(0, ‖ρ_v‖ − ρ_s)
    static double soc_step(const double* zz, const double* D, const std::vector<int>& ix, double tau) {
        int n = (int)ix.size();
        double l0 = zz[ix[0]], ll = l0 * l0, lD = l0 * (-D[ix[0]]);
        for (int i = 1; i < n; ++i) { ll -= zz[ix[i]] * zz[ix[i]]; lD -= zz[ix[i]] * (-D[ix[i]]); }
        ll = std::max(ll, 1e-25);
        double sq = std::sqrt(ll);
        double rho_s = lD / ll;
        double coef = (lD / sq + (-D[ix[0]])) / (l0 / sq + 1.0);
        double nv = 0.0;
        for (int i = 1; i < n; ++i) { double rv = ((-D[ix[i]]) - coef * zz[ix[i]] / sq) / sq; nv += rv * rv; }
        nv = std::sqrt(nv);
        double a = 1.0;
        if (nv - rho_s > 0.0) a = std::min(a, tau / (nv - rho_s));
        return a;
    }
    double cone_dot(const double* zz, const double* D, double a) const {  // Σ⟨primal − aΔp, dual − aΔd⟩
        double s = 0.0;
        for (size_t k = 0; k < idx.ort_p.size(); ++k) { int ip = idx.ort_p[k], id = idx.ort_d[k]; s += (zz[ip] - a * D[ip]) * (zz[id] - a * D[id]); }
        for (size_t c = 0; c < idx.soc_p.size(); ++c)
            for (size_t k = 0; k < idx.soc_p[c].size(); ++k) { int ip = idx.soc_p[c][k], id = idx.soc_d[c][k]; s += (zz[ip] - a * D[ip]) * (zz[id] - a * D[id]); }
        return s;
    }
    void note_and(SolveInfo& info, bool c1, double m1, bool c2, double m2) const {   // decision = c1 && c2
        double m = (c1 && c2) ? std::min(m1, m2) : (!c1 && !c2) ? std::max(m1, m2) : (!c1 ? m1 : m2);
        info.margin = std::min(info.margin, m);
    }
    void note_or(SolveInfo& info, bool c1, double m1, bool c2, double m2) const {    // decision = c1 || c2
        double m = (c1 && c2) ? std::max(m1, m2) : (!c1 && !c2) ? std::min(m1, m2) : (c1 ? m1 : m2);
        info.margin = std::min(info.margin, m);
    }

    // z and th must be initialised by the caller (initialize_z! / θ pack).
    SolveInfo solve() {
        SolveInfo info;
        const bool cones = has_cones();
        int ncone = (int)idx.ort_p.size() + (int)idx.soc_p.size();
        if (opts.mu_mode == 1) { ncone = (int)idx.ort_p.size(); for (auto& c : idx.soc_p) ncone += (int)c.size(); }
        double reg = 0.0;
        double daff[NZ], dl[NZ], zc[NZ], rc[NZ];
        eval_r(z, 0.0, r);
        double r_vio = vio(r, idx.equr), k_vio = bil_vio(r);
        bool converged = false;
        for (int j = 0; j < opts.max_iter; ++j) {
            {
                bool c1 = r_vio < opts.r_tol, c2 = k_vio < opts.kappa_tol;
                note_and(info, c1, rel_gap(r_vio, opts.r_tol), c2, cones ? rel_gap(k_vio, opts.kappa_tol) : std::numeric_limits<double>::infinity());
                if (c1 && c2) { converged = true; break; }
            }
            info.iterations++;
            eval_rz(z, rz);
            if (opts.apply_reg) { reg = (k_vio < opts.kappa_reg) ? k_vio * opts.gamma_reg : 0.0; for (int i = 0; i < NZ; ++i) rz[i * NZ + i] += reg; }
            if (!lu_factor(rz, piv, NZ)) { info.status = 2; break; }
            for (int i = 0; i < NZ; ++i) daff[i] = r[i];
            lu_solve(rz, piv, NZ, daff);
            double kappa = 0.0;
            if (cones) {
                double a_aff = step_length(z, daff, 1.0);
                double mu = cone_dot(z, daff, 0.0) / ncone;
                double mu_aff = cone_dot(z, daff, a_aff) / ncone;
                double ratio = std::min(std::max(mu_aff / mu, 0.0), 1.0);
                double sigma = ratio * ratio * ratio;
                kappa = std::max(sigma * mu, opts.kappa_tol / opts.undercut);
            }
            eval_r(z, kappa, rc);
            // Mehrotra correction: r[bil] += Δaff_primal ∘ Δaff_dual
            for (size_t k = 0; k < idx.ort_p.size(); ++k) rc[idx.ortr[k]] += daff[idx.ort_p[k]] * daff[idx.ort_d[k]];
            for (size_t c = 0; c < idx.soc_p.size(); ++c) {
                const auto& ip = idx.soc_p[c]; const auto& id = idx.soc_d[c]; const auto& rr = idx.socr[c];
                double acc = 0.0;
                for (size_t k = 0; k < ip.size(); ++k) acc += daff[ip[k]] * daff[id[k]];
                rc[rr[0]] += acc;
                for (size_t k = 1; k < ip.size(); ++k) rc[rr[k]] += daff[ip[0]] * daff[id[k]] + daff[id[0]] * daff[ip[k]];
            }
            for (int i = 0; i < NZ; ++i) dl[i] = rc[i];
            lu_solve(rz, piv, NZ, dl);
            double viol = std::max(r_vio, k_vio);
            double tau = std::max(0.95, 1.0 - viol * viol);   // default reading: ϵ_min is carried in Options but unused (see header)
            if (opts.tau_rule == 1) tau = 1.0 - std::min(opts.eps_min, viol * viol);
            else if (opts.tau_rule == 2) tau = 0.99;
            double alpha = cones ? step_length(z, dl, tau) : 1.0;
            // residual line search
            double r_c = 0, k_c = 0;
            for (int i = 0; i < NZ; ++i) zc[i] = z[i] - alpha * dl[i];
            for (int ls = 1; ls <= opts.max_ls; ++ls) {
                eval_r(zc, 0.0, rc);
                r_c = vio(rc, idx.equr); k_c = bil_vio(rc);
                bool c1 = r_c <= r_vio, c2 = k_c <= k_vio;
                note_or(info, c1, rel_gap(r_c, r_vio), c2, cones ? rel_gap(k_c, k_vio) : std::numeric_limits<double>::infinity());
                if (c1 || c2) break;
                alpha *= opts.ls_scale;
                info.ls_steps++;
                for (int i = 0; i < NZ; ++i) zc[i] = z[i] - alpha * dl[i];
                if (ls == opts.max_ls) { eval_r(zc, 0.0, rc); r_c = vio(rc, idx.equr); k_c = bil_vio(rc); }
            }
            bool finite = true;
            for (int i = 0; i < NZ; ++i) { z[i] = zc[i]; r[i] = rc[i]; finite = finite && std::isfinite(zc[i]); }
            r_vio = r_c; k_vio = k_c;
            if (!finite || !std::isfinite(r_vio) || !std::isfinite(k_vio)) { info.status = 2; break; }
        }
        if (info.status == 0 && !converged) {
            converged = (r_vio < opts.r_tol) && (k_vio < opts.kappa_tol);
            if (!converged) info.status = 1;
        }
        info.r_vio = r_vio; info.k_vio = k_vio;
        if (opts.diagnostics && info.status != 2) {
            double rz2[NZ * NZ], st[NZ]; int pv[NZ];
            eval_rz(z, rz2);
            if (lu_factor(rz2, pv, NZ)) {
                for (int i = 0; i < NZ; ++i) st[i] = r[i];
                lu_solve(rz2, pv, NZ, st);
                double m = 0.0; for (int i = 0; i < n_out_rows; ++i) m = std::max(m, std::fabs(st[i]));
                info.q_uncertainty = m;
            } else info.q_uncertainty = std::numeric_limits<double>::infinity();
        }
        if (opts.diff_sol && info.status != 2) {
            if (!differentiate(&info.ift_spread, opts.apply_reg ? std::max(reg, opts.kappa_tol * opts.gamma_reg) : 0.0)) info.status = 2;
        }
        return info;
    }

    // δz = −rz(z*,θ)⁻¹ rθ(z*,θ)
    bool differentiate(double* spread = nullptr, double reg = 0.0) {
        eval_rz(z, rz);
        if (reg != 0.0) for (int i = 0; i < NZ; ++i) rz[i * NZ + i] += reg;
        eval_rth(z, rth);
        double alt[NZ * NZ], col[NZ];
        int piv2[NZ];
        for (int i = 0; i < NZ; ++i) { const double sc = std::ldexp(1.0, (7 * i) % 11 - 5); for (int j = 0; j < NZ; ++j) alt[i * NZ + j] = sc * rz[i * NZ + j]; }
        if (!lu_factor(rz, piv, NZ)) { if (spread) *spread = std::numeric_limits<double>::infinity(); return false; }
        const bool alt_ok = lu_factor(alt, piv2, NZ);
        double sp = alt_ok ? 0.0 : std::numeric_limits<double>::infinity();
        for (int j = 0; j < NTH; ++j) {
            for (int i = 0; i < NZ; ++i) col[i] = rth[i * NTH + j];
            lu_solve(rz, piv, NZ, col);
            for (int i = 0; i < NZ; ++i) dz[i * NTH + j] = -col[i];
            if (alt_ok) {
                for (int i = 0; i < NZ; ++i) col[i] = std::ldexp(1.0, (7 * i) % 11 - 5) * rth[i * NTH + j];
                lu_solve(alt, piv2, NZ, col);
                for (int i = 0; i < n_out_rows; ++i) { const double dd = std::fabs(-col[i] - dz[i * NTH + j]); if (!(dd <= sp)) sp = dd; }
            }
        }
        if (spread) *spread = sp;
        return true;
    }
};

}  // namespace od_oracle
