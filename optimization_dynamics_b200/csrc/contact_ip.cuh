// Condensed primal-dual interior-point step + IFT sensitivities for the contact-implicit models, one problem per thread,
// everything in registers.
//
// Replaces, for one (q1, q2, u) sample, the reference call chain
//     f / fx / fu            reference src/dynamics.jl:81-128
//       └ RoboDojo.step!  →  interior_point_solve!  →  differentiate (δz = −rz⁻¹ rθ)        [RoboDojo.jl, external]
// Algorithm = SURVEY.md Appendix A.3 / DESIGN.md §Algorithm (Mehrotra predictor–corrector, residual line search).
//
// Layout of one problem (RoboDojo IndicesZ: q, γ, sγ, ψ, b, sψ, sb — reference src/models/planar_push/simulator.jl:1-14):
//     z = [ q(NQ) | γ(NC) s(NC) | ψ(NP) b(NB) | sψ(NP) sb(NB) ]       NC orthant pairs (γ_i, s_i),
//                                                                     NP second-order cones (ψ_k; b_k) ∘ (sψ_k; sb_k)
//     rows: d(q,γ,b;θ)=0 | s − ϕ(q)=0 | ψ − ψ̂(γ;θ)=0 | vT(q;θ) − sb=0 | γ∘s = κ | (ψ;b)∘(sψ;sb) = (κ;0)
// The reference factors the dense nz×nz Jacobian with LU (`lu_solver`, reference src/gradient_bundle.jl:76).  Here the
// variables that enter their defining row with a unit coefficient — Δs = rs + NΔq, Δψ = rpsi + MψΔγ, Δsb = VΔq − rv — are
// substituted out exactly (no division, so no loss of accuracy however close the iterate is to the cone boundary), and the
// remaining NR = NQ+NC+NB+NP unknowns x = (Δq, Δγ, Δb, Δsψ) are solved with a partial-pivoting LU (hopper: 12×12 instead of
// 20×20 ⇒ ≈4.6× fewer flops).  Further condensation to NQ×NQ (dividing by s or sψ) is NOT used: with undercut = Inf the
// reference drives γ∘s to ~1e-26, where the normal-equation form D + Nᵀ(Γ/S)N loses all accuracy (measured: 22 % of the hopper
// batch diverged from the oracle).  The LU lives in a per-thread workspace (shared memory on the GPU), element e of lane l at
// ws[e·stride + l] — bank-conflict free.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

// Every solver routine is __host__ __device__ so that tests/host_check.cu can single-step the very same template code on the
// CPU against the oracle (a debugging aid for a GPU-less build container; liboptdyn_b200.so exports no host compute path).
#define OD_HD __host__ __device__ __forceinline__

namespace od {

OD_HD double rsqrt_d(double x) { return 1.0 / sqrt(x); }

struct SolverOpts {           // RoboDojo InteriorPointOptions as set at reference src/dynamics.jl:25-33
    double r_tol;             // ∞-norm tolerance on the equality rows
    double kappa_eval_tol;    // ∞-norm tolerance on the bilinear rows, eval simulator (f)
    double kappa_grad_tol;    // same, gradient simulator (fx, fu)
    double ls_scale;          // 0.5
    int max_iter;             // 100
    int max_ls;               // 25
};

// status nibble: 0 converged, 1 iteration cap, 2 non-finite iterate or singular system
enum : int { ST_OK = 0, ST_MAXIT = 1, ST_FAIL = 2 };

// CVXOPT §8.2 step to the boundary of the second-order cone for λ − αΔ, λ=(l0; l1..), returns min(1, τ α_max)
template <int DIM>
OD_HD double soc_step(double l0, const double* l1, double d0, const double* d1, double tau) {
    double ll = l0 * l0, lD = l0 * (-d0);
#pragma unroll
    for (int i = 0; i < DIM; ++i) { ll -= l1[i] * l1[i]; lD -= l1[i] * (-d1[i]); }
    ll = fmax(ll, 1e-25);
    const double sq = sqrt(ll);
    const double rho_s = lD / ll;
    const double coef = (lD / sq + (-d0)) / (l0 / sq + 1.0);
    double nv = 0.0;
#pragma unroll
    for (int i = 0; i < DIM; ++i) { const double rv = ((-d1[i]) - coef * l1[i] / sq) / sq; nv += rv * rv; }
    nv = sqrt(nv);
    double a = 1.0;
    if (nv - rho_s > 0.0) a = fmin(a, tau / (nv - rho_s));
    return a;
}

template <int N> struct cmax1 { static constexpr int v = N > 0 ? N : 1; };

template <class M>
struct ContactIP {
    static constexpr int NQ = M::NQ, NU = M::NU, NC = M::NC, NP = M::NP, NB = M::NB, NTH = M::NTH;
    static constexpr int NC1 = cmax1<NC>::v, NP1 = cmax1<NP>::v, NB1 = cmax1<NB>::v;
    static constexpr int NTP = 2 * NQ + NU;              // θ' = (q1, q2, u): the sensitivity columns that are returned
    static constexpr int NCONE = NC + NP;                // cone degree (orthant pairs + second-order cones)

    struct Z { double q[NQ], gam[NC1], s[NC1], psi[NP1], b[NB1], spsi[NP1], sb[NB1]; };
    // residual in block form; bilinear rows are stored at κ = 0 (r(z;κ) only shifts rgam and rc0 by −κ)
    struct R { double d[NQ], rs[NC1], rpsi[NP1], rv[NB1], rgam[NC1], rc0[NP1], rc1[NB1]; };
    static constexpr int NR = NQ + NC + NB + NP;         // reduced system size
    static constexpr int WS = NR * NR + 2 * NR;          // workspace doubles per problem: K (NR×NR) | x (NR) | piv (NR)
    struct Lin {
        double N[NC1 * NQ], V[NB1 * NQ], Mpsi[NP1 * NC1];
        double* ws; int stride;
        bool ok;
        OD_HD double& K(int i, int j) const { return ws[(size_t)(i * NR + j) * stride]; }
        OD_HD double& x(int i) const { return ws[(size_t)(NR * NR + i) * stride]; }
        OD_HD double& piv(int i) const { return ws[(size_t)(NR * NR + NR + i) * stride]; }
    };
    __host__ __device__ static constexpr int cone_of(int j) {   // friction cone that tangential component j belongs to
        int k = 0;
        for (int c = 0; c < NP; ++c) if (j >= M::cone_off(c) && j < M::cone_off(c) + M::cone_dim(c)) k = c;
        return k;
    }

    // ---- residual ------------------------------------------------------------------------------------------------
    OD_HD static void residual(const Z& z, const double* th, R& r, double& r_vio, double& k_vio) {
        double phi[NC1], psit[NP1], vT[NB1];
        M::eq(z.q, z.gam, z.b, th, r.d, phi, psit, vT);
        double rv = 0.0, kv = 0.0;
#pragma unroll
        for (int i = 0; i < NQ; ++i) rv = fmax(rv, fabs(r.d[i]));
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            r.rs[i] = z.s[i] - phi[i]; rv = fmax(rv, fabs(r.rs[i]));
            r.rgam[i] = z.gam[i] * z.s[i]; kv = fmax(kv, fabs(r.rgam[i]));
        }
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            r.rpsi[k] = z.psi[k] - psit[k]; rv = fmax(rv, fabs(r.rpsi[k]));
            double acc = z.psi[k] * z.spsi[k];
#pragma unroll
            for (int j = M::cone_off(k); j < M::cone_off(k) + M::cone_dim(k); ++j) {
                acc += z.b[j] * z.sb[j];
                r.rc1[j] = z.psi[k] * z.sb[j] + z.spsi[k] * z.b[j]; kv = fmax(kv, fabs(r.rc1[j]));
            }
            r.rc0[k] = acc; kv = fmax(kv, fabs(acc));
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) { r.rv[j] = vT[j] - z.sb[j]; rv = fmax(rv, fabs(r.rv[j])); }
        r_vio = rv; k_vio = kv;
    }

    // ---- linearise at z: model blocks → reduced matrix K → LU with partial pivoting (in the workspace) -----------------
    //   unknowns x = [Δq | Δγ | Δb | Δsψ];  rows:
    //   d_i  : D Δq + Eγ Δγ + Eb Δb                                              = rd
    //   γ_i  : γ_i N_i Δq + s_i Δγ_i                                             = rgam_i − γ_i rs_i
    //   c0_k : (Σ_j b_j V_j) Δq + sψ_k Mψ_k Δγ + Σ_j sb_j Δb_j + ψ_k Δsψ_k        = rc0_k − sψ_k rpsi_k + Σ_j b_j rv_j
    //   c1_j : ψ_k V_j Δq + sb_j Mψ_k Δγ + sψ_k Δb_j + b_j Δsψ_k                  = rc1_j − sb_j rpsi_k + ψ_k rv_j
    OD_HD static void linearize(const Z& z, const double* th, Lin& L) { assemble(z, th, L); factor(L); }
    OD_HD static void assemble(const Z& z, const double* th, Lin& L) {
        double D[NQ * NQ], Eg[NQ * NC1], Eb[NQ * NB1];
        M::jac(z.q, z.gam, z.b, th, D, Eg, Eb, L.N, L.V, L.Mpsi);
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
#pragma unroll
            for (int j = 0; j < NQ; ++j) L.K(i, j) = D[i * NQ + j];
#pragma unroll
            for (int j = 0; j < NC; ++j) L.K(i, NQ + j) = Eg[i * NC1 + j];
#pragma unroll
            for (int j = 0; j < NB; ++j) L.K(i, NQ + NC + j) = Eb[i * NB1 + j];
#pragma unroll
            for (int j = 0; j < NP; ++j) L.K(i, NQ + NC + NB + j) = 0.0;
        }
#pragma unroll
        for (int i = 0; i < NC; ++i) {
#pragma unroll
            for (int j = 0; j < NQ; ++j) L.K(NQ + i, j) = z.gam[i] * L.N[i * NQ + j];
#pragma unroll
            for (int j = NQ; j < NR; ++j) L.K(NQ + i, j) = (j == NQ + i) ? z.s[i] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const int row = NQ + NC + k;
#pragma unroll
            for (int c = 0; c < NQ; ++c) {
                double a = 0.0;
#pragma unroll
                for (int e = 0; e < M::cone_dim(k); ++e) a += z.b[M::cone_off(k) + e] * L.V[(M::cone_off(k) + e) * NQ + c];
                L.K(row, c) = a;
            }
#pragma unroll
            for (int i = 0; i < NC; ++i) L.K(row, NQ + i) = z.spsi[k] * L.Mpsi[k * NC1 + i];
#pragma unroll
            for (int j = 0; j < NB; ++j) L.K(row, NQ + NC + j) = (cone_of(j) == k) ? z.sb[j] : 0.0;
#pragma unroll
            for (int j = 0; j < NP; ++j) L.K(row, NQ + NC + NB + j) = (j == k) ? z.psi[k] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int k = cone_of(j), row = NQ + NC + NP + j;
#pragma unroll
            for (int c = 0; c < NQ; ++c) L.K(row, c) = z.psi[k] * L.V[j * NQ + c];
#pragma unroll
            for (int i = 0; i < NC; ++i) L.K(row, NQ + i) = z.sb[j] * L.Mpsi[k * NC1 + i];
#pragma unroll
            for (int jj = 0; jj < NB; ++jj) L.K(row, NQ + NC + jj) = (jj == j) ? z.spsi[k] : 0.0;
#pragma unroll
            for (int kk = 0; kk < NP; ++kk) L.K(row, NQ + NC + NB + kk) = (kk == k) ? z.b[j] : 0.0;
        }
    }
    // LU, partial (row) pivoting
    OD_HD static void factor(Lin& L) {
        bool ok = true;
        for (int k = 0; k < NR; ++k) {
            int p = k; double best = fabs(L.K(k, k));
            for (int i = k + 1; i < NR; ++i) { const double a = fabs(L.K(i, k)); if (a > best) { best = a; p = i; } }
            L.piv(k) = (double)p;
            ok = ok && (best > 0.0) && (best < INFINITY);
            if (p != k) for (int j = 0; j < NR; ++j) { const double t = L.K(k, j); L.K(k, j) = L.K(p, j); L.K(p, j) = t; }
            const double inv = 1.0 / L.K(k, k);
            L.K(k, k) = inv;                       // keep the reciprocal pivot
            for (int i = k + 1; i < NR; ++i) {
                const double l = L.K(i, k) * inv;
                L.K(i, k) = l;
                for (int j = k + 1; j < NR; ++j) L.K(i, j) -= l * L.K(k, j);
            }
        }
        L.ok = ok;
    }

    // x (in the workspace) ← K⁻¹ x
    OD_HD static void lu_solve(const Lin& L) {
        for (int k = 0; k < NR; ++k) { const int p = (int)L.piv(k); if (p != k) { const double t = L.x(k); L.x(k) = L.x(p); L.x(p) = t; } }
        for (int i = 1; i < NR; ++i) { double sacc = L.x(i); for (int j = 0; j < i; ++j) sacc -= L.K(i, j) * L.x(j); L.x(i) = sacc; }
        for (int i = NR - 1; i >= 0; --i) { double sacc = L.x(i); for (int j = i + 1; j < NR; ++j) sacc -= L.K(i, j) * L.x(j); L.x(i) = sacc * L.K(i, i); }
    }

    // reduced right-hand side of r into the workspace
    OD_HD static void load_rhs(const Lin& L, const Z& z, const R& r) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) L.x(i) = r.d[i];
#pragma unroll
        for (int i = 0; i < NC; ++i) L.x(NQ + i) = r.rgam[i] - z.gam[i] * r.rs[i];
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            double a = r.rc0[k] - z.spsi[k] * r.rpsi[k];
#pragma unroll
            for (int e = 0; e < M::cone_dim(k); ++e) a += z.b[M::cone_off(k) + e] * r.rv[M::cone_off(k) + e];
            L.x(NQ + NC + k) = a;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) L.x(NQ + NC + NP + j) = r.rc1[j] - z.sb[j] * r.rpsi[cone_of(j)] + z.psi[cone_of(j)] * r.rv[j];
    }

    // full Newton direction for right-hand side r:  rz Δ = r
    OD_HD static void solve(const Lin& L, const Z& z, const R& r, Z& D) {
        load_rhs(L, z, r);
        lu_solve(L);
#pragma unroll
        for (int i = 0; i < NQ; ++i) D.q[i] = L.x(i);
#pragma unroll
        for (int i = 0; i < NC; ++i) D.gam[i] = L.x(NQ + i);
#pragma unroll
        for (int j = 0; j < NB; ++j) D.b[j] = L.x(NQ + NC + j);
#pragma unroll
        for (int k = 0; k < NP; ++k) D.spsi[k] = L.x(NQ + NC + NB + k);
        // substituted variables
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            double ds = r.rs[i];
#pragma unroll
            for (int j = 0; j < NQ; ++j) ds += L.N[i * NQ + j] * D.q[j];
            D.s[i] = ds;
        }
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            double dpsi = r.rpsi[k];
#pragma unroll
            for (int i = 0; i < NC; ++i) dpsi += L.Mpsi[k * NC1 + i] * D.gam[i];
            D.psi[k] = dpsi;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            double v = -r.rv[j];
#pragma unroll
            for (int c = 0; c < NQ; ++c) v += L.V[j * NQ + c] * D.q[c];
            D.sb[j] = v;
        }
    }

    // ---- cone utilities --------------------------------------------------------------------------------------------
    OD_HD static double step_length(const Z& z, const Z& D, double tau) {
        double a = 1.0;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            if (D.gam[i] > 0.0) a = fmin(a, tau * z.gam[i] / D.gam[i]);
            if (D.s[i] > 0.0) a = fmin(a, tau * z.s[i] / D.s[i]);
        }
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            if (M::cone_dim(k) == 1) {
                a = fmin(a, soc_step<1>(z.psi[k], &z.b[M::cone_off(k)], D.psi[k], &D.b[M::cone_off(k)], tau));
                a = fmin(a, soc_step<1>(z.spsi[k], &z.sb[M::cone_off(k)], D.spsi[k], &D.sb[M::cone_off(k)], tau));
            } else {
                a = fmin(a, soc_step<2>(z.psi[k], &z.b[M::cone_off(k)], D.psi[k], &D.b[M::cone_off(k)], tau));
                a = fmin(a, soc_step<2>(z.spsi[k], &z.sb[M::cone_off(k)], D.spsi[k], &D.sb[M::cone_off(k)], tau));
            }
        }
        return a;
    }

    // Σ ⟨primal − aΔp, dual − aΔd⟩ over all cones
    OD_HD static double cone_dot(const Z& z, const Z& D, double a) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < NC; ++i) s += (z.gam[i] - a * D.gam[i]) * (z.s[i] - a * D.s[i]);
#pragma unroll
        for (int k = 0; k < NP; ++k) s += (z.psi[k] - a * D.psi[k]) * (z.spsi[k] - a * D.spsi[k]);
#pragma unroll
        for (int j = 0; j < NB; ++j) s += (z.b[j] - a * D.b[j]) * (z.sb[j] - a * D.sb[j]);
        return s;
    }

    OD_HD static void candidate(const Z& z, const Z& D, double a, Z& c) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) c.q[i] = z.q[i] - a * D.q[i];
#pragma unroll
        for (int i = 0; i < NC; ++i) { c.gam[i] = z.gam[i] - a * D.gam[i]; c.s[i] = z.s[i] - a * D.s[i]; }
#pragma unroll
        for (int k = 0; k < NP; ++k) { c.psi[k] = z.psi[k] - a * D.psi[k]; c.spsi[k] = z.spsi[k] - a * D.spsi[k]; }
#pragma unroll
        for (int j = 0; j < NB; ++j) { c.b[j] = z.b[j] - a * D.b[j]; c.sb[j] = z.sb[j] - a * D.sb[j]; }
    }

    // ---- one predictor–corrector iteration with residual line search (z, r, r_vio, k_vio updated in place) ------------
    OD_HD static void iterate(const Lin& L, const double* th, const SolverOpts& o, Z& z, R& r, double& r_vio, double& k_vio) {
        Z D;
        double kappa = 0.0;
        if (NCONE > 0) {
            solve(L, z, r, D);                                   // affine direction
            const double a_aff = step_length(z, D, 1.0);
            const double mu = cone_dot(z, D, 0.0) * (1.0 / (NCONE > 0 ? NCONE : 1));
            const double mu_aff = cone_dot(z, D, a_aff) * (1.0 / (NCONE > 0 ? NCONE : 1));
            const double ratio = fmin(fmax(mu_aff / mu, 0.0), 1.0);
            kappa = ratio * ratio * ratio * mu;                   // max(σμ, κ_tol/undercut) with undercut = Inf
            // corrector right-hand side: r(z;κ) + Δaff_primal ∘ Δaff_dual on the bilinear rows
            R rc = r;
#pragma unroll
            for (int i = 0; i < NC; ++i) rc.rgam[i] = (r.rgam[i] - kappa) + D.gam[i] * D.s[i];
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                double acc = D.psi[k] * D.spsi[k];
#pragma unroll
                for (int e = 0; e < M::cone_dim(k); ++e) {
                    const int j = M::cone_off(k) + e;
                    acc += D.b[j] * D.sb[j];
                    rc.rc1[j] = r.rc1[j] + (D.psi[k] * D.sb[j] + D.spsi[k] * D.b[j]);
                }
                rc.rc0[k] = (r.rc0[k] - kappa) + acc;
            }
            solve(L, z, rc, D);
        } else {
            solve(L, z, r, D);                                   // no cones: plain Newton direction
        }
        const double viol = fmax(r_vio, k_vio);
        const double tau = fmax(0.95, 1.0 - viol * viol);
        double alpha = (NCONE > 0) ? step_length(z, D, tau) : 1.0;
        Z zc; R rc2; double rv2, kv2;
        candidate(z, D, alpha, zc);
        for (int ls = 1; ls <= o.max_ls; ++ls) {
            residual(zc, th, rc2, rv2, kv2);
            if (rv2 <= r_vio || kv2 <= k_vio) break;
            alpha *= o.ls_scale;
            candidate(z, D, alpha, zc);
            if (ls == o.max_ls) residual(zc, th, rc2, rv2, kv2);
        }
        z = zc; r = rc2; r_vio = rv2; k_vio = kv2;
    }

    // ---- IFT: ∂q3/∂θ' = −(rz⁻¹ rθ')[q rows]; column c of the NQ×NTP column-major result goes to dq1 / dq2 / du ---------
    OD_HD static void sensitivities(const Lin& L, const Z& z, const double* th, double* dq1, double* dq2, double* du) {
        double Dth[NQ * NTP], Vth[NB1 * NTP];
        M::jacth(z.q, z.gam, z.b, th, Dth, Vth);
        R r;
#pragma unroll
        for (int i = 0; i < NC1; ++i) { r.rs[i] = 0.0; r.rgam[i] = 0.0; }
#pragma unroll
        for (int i = 0; i < NP1; ++i) { r.rpsi[i] = 0.0; r.rc0[i] = 0.0; }
#pragma unroll
        for (int i = 0; i < NB1; ++i) r.rc1[i] = 0.0;
#pragma unroll
        for (int c = 0; c < NTP; ++c) {
#pragma unroll
            for (int i = 0; i < NQ; ++i) r.d[i] = Dth[i * NTP + c];
#pragma unroll
            for (int j = 0; j < NB; ++j) r.rv[j] = Vth[j * NTP + c];
            load_rhs(L, z, r);
            lu_solve(L);
            double* dst = (c < NQ) ? (dq1 + c * NQ) : (c < 2 * NQ) ? (dq2 + (c - NQ) * NQ) : (du + (c - 2 * NQ) * NQ);
#pragma unroll
            for (int i = 0; i < NQ; ++i) dst[i] = -L.x(i);
        }
    }

    // initialize_z! (reference src/models/planar_push/simulator.jl:52-60 and the same pattern in the other models)
    OD_HD static void init_z(const double* q2, Z& z) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) z.q[i] = q2[i];
#pragma unroll
        for (int i = 0; i < NC1; ++i) { z.gam[i] = 1.0; z.s[i] = 1.0; }
#pragma unroll
        for (int i = 0; i < NP1; ++i) { z.psi[i] = 1.0; z.spsi[i] = 1.0; }
#pragma unroll
        for (int i = 0; i < NB1; ++i) { z.b[i] = 0.1; z.sb[i] = 0.1; }
    }

    OD_HD static bool finite(const Z& z, double r_vio, double k_vio) {
        double s = r_vio + k_vio;
#pragma unroll
        for (int i = 0; i < NQ; ++i) s += z.q[i];
#pragma unroll
        for (int i = 0; i < NC; ++i) s += z.gam[i] + z.s[i];
#pragma unroll
        for (int k = 0; k < NP; ++k) s += z.psi[k] + z.spsi[k];
#pragma unroll
        for (int j = 0; j < NB; ++j) s += z.b[j] + z.sb[j];
        return isfinite(s);
    }
};

// Per-launch arguments of the batched step kernel.  Arrays are row-per-problem with explicit strides (in doubles) so that the
// same kernel serves separate arrays (C ABI) and one packed row [q3 | ∂q3/∂q1 | ∂q3/∂q2 | ∂q3/∂u] (all-gather buffer).
struct StepArgs {
    int B;
    const double* q1; const double* q2; const double* u;
    int in_stride_q, in_stride_u;
    double* q3; double* dq1; double* dq2; double* du;          // any may be null
    int out_stride_q3, out_stride_dq, out_stride_du;
    int* status; int* iters;                                   // may be null
    double h;
    double fric[4];
    int want_eval, want_grad;
    // gradient bundle (reference src/gradient_bundle.jl:89-100): problem i = sample i/(n_eta+1) perturbed by eta row i%(n_eta+1) − 1
    // (row −1 = nominal); eta is n_eta × (2NQ+NU), null when unused.
    const double* eta; int n_eta;
    SolverOpts opts;
};

// One thread = one problem.  A single iterate sequence serves both simulators of ImplicitDynamics: eval_sim and grad_sim
// (reference src/dynamics.jl:60-64) start from the same initialisation and differ only in κ_tol, so the looser one is a prefix
// of the tighter one.  The IFT is taken at the first iterate meeting the gradient tolerance, q3 at the first meeting the eval one.
template <class M>
OD_HD void contact_step_one(const StepArgs& a, const int i, double* ws, const int ws_stride) {
    typedef ContactIP<M> IP;
    constexpr int NQ = M::NQ, NU = M::NU;
    double th[M::NTH];
    typename IP::Z z;
    {
        const int src = a.eta ? i / (a.n_eta + 1) : i;
        const int pert = a.eta ? i % (a.n_eta + 1) : 0;
        const double* p1 = a.q1 + (size_t)src * a.in_stride_q;
        const double* p2 = a.q2 + (size_t)src * a.in_stride_q;
        const double* pu = a.u + (size_t)src * a.in_stride_u;
        const double* pe = (pert > 0) ? a.eta + (size_t)(pert - 1) * (2 * NQ + NU) : nullptr;
        double q2v[NQ];
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            double x1 = p1[k], x2 = p2[k];
            if (pe) { x1 += pe[k]; x2 += pe[NQ + k]; }
            const double v1 = (x2 - x1) / a.h;                 // src/dynamics.jl:84-86
            th[k] = x2 - a.h * v1;                             // RoboDojo.step!: q1 = q2 − h v1
            th[NQ + k] = x2;
            q2v[k] = x2;
        }
#pragma unroll
        for (int k = 0; k < NU; ++k) th[2 * NQ + k] = pe ? pu[k] + pe[2 * NQ + k] : pu[k];
#pragma unroll
        for (int k = 0; k < M::NF; ++k) th[2 * NQ + NU + k] = a.fric[k];
        th[M::NTH - 1] = a.h;
        IP::init_z(q2v, z);
    }
    typename IP::R r;
    double r_vio, k_vio;
    IP::residual(z, th, r, r_vio, k_vio);
    bool eval_done = !a.want_eval, grad_done = !a.want_grad;
    int it = 0, it_e = 0, it_g = 0, st_e = 0, st_g = 0;
    for (;;) {
        const bool bad = !IP::finite(z, r_vio, k_vio);
        const bool capped = it >= a.opts.max_iter;
        const bool rok = r_vio < a.opts.r_tol;
        const bool conv_e = rok && (k_vio < a.opts.kappa_eval_tol);
        const bool conv_g = rok && (k_vio < a.opts.kappa_grad_tol);
        if (!eval_done && (conv_e || capped || bad)) {
            eval_done = true; it_e = it; st_e = bad ? ST_FAIL : (conv_e ? ST_OK : ST_MAXIT);
            if (a.q3) {
                double* o = a.q3 + (size_t)i * a.out_stride_q3;
#pragma unroll
                for (int k = 0; k < NQ; ++k) o[k] = z.q[k];
            }
        }
        const bool need_ift = !grad_done && (conv_g || capped || bad);
        const bool do_iter = !bad && !capped && (!eval_done || (!grad_done && !need_ift));
        if (!need_ift && !do_iter) break;
        typename IP::Lin L;
        L.ws = ws; L.stride = ws_stride;
        IP::linearize(z, th, L);
        if (need_ift) {
            grad_done = true; it_g = it; st_g = (bad || !L.ok) ? ST_FAIL : (conv_g ? ST_OK : ST_MAXIT);
            if (a.dq1) {
                IP::sensitivities(L, z, th, a.dq1 + (size_t)i * a.out_stride_dq, a.dq2 + (size_t)i * a.out_stride_dq, a.du + (size_t)i * a.out_stride_du);
            }
        }
        if (do_iter) {
            if (!L.ok) { st_e = eval_done ? st_e : ST_FAIL; st_g = grad_done ? st_g : ST_FAIL; it_e = eval_done ? it_e : it; it_g = grad_done ? it_g : it; break; }
            IP::iterate(L, th, a.opts, z, r, r_vio, k_vio);
            ++it;
        } else {
            break;
        }
    }
    if (a.status) a.status[i] = st_e | (st_g << 4);
    if (a.iters) a.iters[i] = it_e | (it_g << 16);
}

template <class M, int BLOCK>
__global__ void __launch_bounds__(BLOCK) contact_step_kernel(const StepArgs a) {
    extern __shared__ double od_smem[];            // BLOCK × ContactIP<M>::WS doubles, lane-interleaved
    const int i = blockIdx.x * BLOCK + threadIdx.x;
    if (i >= a.B) return;
    contact_step_one<M>(a, i, od_smem + threadIdx.x, BLOCK);
}

}  // namespace od
