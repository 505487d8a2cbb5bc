// Dense-Jacobian interior-point / Newton solve + IFT, one problem per thread — used for the rocket models whose residual
// Jacobians have no contact block structure to condense:
//   rocket dynamics (12×12, no cones)            reference src/models/rocket/codegen.jl:14-22, dynamics.jl:21-47,101-163
//   SOC thrust-limit projection (10×10, 2 orthant pairs + one 3-D second-order cone sharing u₃)
//                                                reference src/models/rocket/codegen.jl:45-64, dynamics.jl:49-86,168-210
// Same algorithm as contact_ip.cuh with a partial-pivoting LU on the full matrix.
#pragma once
#include "contact_ip.cuh"

namespace od {

template <class M>
struct DenseIP {
    static constexpr int NZ = M::NZ, NTH = M::NTH, NTHP = M::NTHP;
    static constexpr int NCONE = M::NORT + M::NSOC;

    OD_HD static void residual(const double* z, const double* th, double* r, double& r_vio, double& k_vio) {
        M::res(z, th, r);
        MaxAcc rv, kv;
#pragma unroll
        for (int i = 0; i < M::NEQ; ++i) rv.add(od_abs(r[i]));
#pragma unroll
        for (int i = M::NEQ; i < NZ; ++i) kv.add(od_abs(r[i]));
        r_vio = rv.v; k_vio = kv.v;
    }

    OD_HD static bool lu_factor(double* A, int* piv) {
        bool ok = true;
        for (int k = 0; k < NZ; ++k) {
            int p = k; double best = fabs(A[k * NZ + k]);
            for (int i = k + 1; i < NZ; ++i) { const double a = fabs(A[i * NZ + k]); if (a > best) { best = a; p = i; } }
            piv[k] = p;
            ok = ok && (best > 0.0) && (best < INFINITY);
            if (p != k) for (int j = 0; j < NZ; ++j) { const double t = A[k * NZ + j]; A[k * NZ + j] = A[p * NZ + j]; A[p * NZ + j] = t; }
            const double inv = 1.0 / A[k * NZ + k];
            for (int i = k + 1; i < NZ; ++i) {
                const double l = A[i * NZ + k] * inv;
                A[i * NZ + k] = l;
                for (int j = k + 1; j < NZ; ++j) A[i * NZ + j] -= l * A[k * NZ + j];
            }
        }
        return ok;
    }
    OD_HD static void lu_solve(const double* A, const int* piv, double* x) {
        for (int k = 0; k < NZ; ++k) { const int p = piv[k]; if (p != k) { const double t = x[k]; x[k] = x[p]; x[p] = t; } }
        for (int i = 1; i < NZ; ++i) { double s = x[i]; for (int j = 0; j < i; ++j) s -= A[i * NZ + j] * x[j]; x[i] = s; }
        for (int i = NZ - 1; i >= 0; --i) { double s = x[i]; for (int j = i + 1; j < NZ; ++j) s -= A[i * NZ + j] * x[j]; x[i] = s / A[i * NZ + i]; }
    }

    OD_HD static double step_length(const double* z, const double* D, double tau) {
        double bn = 1.0, bd = 1.0;
        for (int k = 0; k < M::NORT; ++k) {
            const int ip = M::ort_p(k), id = M::ort_d(k);
            if (D[ip] > 0.0) frac_min(bn, bd, tau * z[ip], D[ip]);
            if (D[id] > 0.0) frac_min(bn, bd, tau * z[id], D[id]);
        }
        for (int c = 0; c < M::NSOC; ++c) {
            double l1[2], d1[2];
            l1[0] = z[M::soc_p(c, 1)]; l1[1] = z[M::soc_p(c, 2)]; d1[0] = D[M::soc_p(c, 1)]; d1[1] = D[M::soc_p(c, 2)];
            soc_step<2>(z[M::soc_p(c, 0)], l1, D[M::soc_p(c, 0)], d1, tau, bn, bd);
            l1[0] = z[M::soc_d(c, 1)]; l1[1] = z[M::soc_d(c, 2)]; d1[0] = D[M::soc_d(c, 1)]; d1[1] = D[M::soc_d(c, 2)];
            soc_step<2>(z[M::soc_d(c, 0)], l1, D[M::soc_d(c, 0)], d1, tau, bn, bd);
        }
        return bn * pivot_rcp(bd);                 // bd > 0 by construction (starts at 1, replaced only by positive denominators)
    }
    OD_HD static double cone_dot(const double* z, const double* D, double a) {
        double s = 0.0;
        for (int k = 0; k < M::NORT; ++k) s += (z[M::ort_p(k)] - a * D[M::ort_p(k)]) * (z[M::ort_d(k)] - a * D[M::ort_d(k)]);
        for (int c = 0; c < M::NSOC; ++c)
            for (int e = 0; e < 3; ++e) s += (z[M::soc_p(c, e)] - a * D[M::soc_p(c, e)]) * (z[M::soc_d(c, e)] - a * D[M::soc_d(c, e)]);
        return s;
    }

    // Solve to (r_tol, κ_tol); returns status, iteration count in *iters.  z holds the initial point on entry.
    OD_HD static int solve(double* z, const double* th, double r_tol, double kappa_tol, int max_iter, int max_ls, double ls_scale, int* iters) {
        double r[NZ], A[NZ * NZ], da[NZ], dl[NZ], zc[NZ], rc[NZ];
        int piv[NZ];
        double r_vio, k_vio;
        residual(z, th, r, r_vio, k_vio);
        int it = 0, status = ST_MAXIT;
        for (;;) {
            bool fin = isfinite(r_vio + k_vio);
            for (int i = 0; i < NZ; ++i) fin = fin && isfinite(z[i]);
            if (!fin) { status = ST_FAIL; break; }
            if (r_vio < r_tol && k_vio < kappa_tol) { status = ST_OK; break; }
            if (it >= max_iter) { status = ST_MAXIT; break; }
            M::jac(z, th, A);
            if (!lu_factor(A, piv)) { status = ST_FAIL; break; }
            double kappa = 0.0;
            if (NCONE > 0) {
                for (int i = 0; i < NZ; ++i) da[i] = r[i];
                lu_solve(A, piv, da);
                const double a_aff = step_length(z, da, 1.0);
                const double mu = cone_dot(z, da, 0.0) / (NCONE > 0 ? NCONE : 1);
                const double mu_aff = cone_dot(z, da, a_aff) / (NCONE > 0 ? NCONE : 1);
                const double ratio = fmin(fmax(mu_aff / mu, 0.0), 1.0);
                kappa = ratio * ratio * ratio * mu;
                for (int i = 0; i < NZ; ++i) dl[i] = r[i];
                for (int k = 0; k < M::NORT; ++k) dl[M::ortr(k)] = (r[M::ortr(k)] - kappa) + da[M::ort_p(k)] * da[M::ort_d(k)];
                for (int c = 0; c < M::NSOC; ++c) {
                    double acc = 0.0;
                    for (int e = 0; e < 3; ++e) acc += da[M::soc_p(c, e)] * da[M::soc_d(c, e)];
                    dl[M::socr(c, 0)] = (r[M::socr(c, 0)] - kappa) + acc;
                    for (int e = 1; e < 3; ++e)
                        dl[M::socr(c, e)] = r[M::socr(c, e)] + (da[M::soc_p(c, 0)] * da[M::soc_d(c, e)] + da[M::soc_d(c, 0)] * da[M::soc_p(c, e)]);
                }
            } else {
                for (int i = 0; i < NZ; ++i) dl[i] = r[i];
            }
            lu_solve(A, piv, dl);
            const double viol = fmax(r_vio, k_vio);
            const double tau = fmax(0.95, 1.0 - viol * viol);
            double alpha = (NCONE > 0) ? step_length(z, dl, tau) : 1.0;
            double rv2 = 0.0, kv2 = 0.0;
            for (int i = 0; i < NZ; ++i) zc[i] = z[i] - alpha * dl[i];
            for (int ls = 1; ls <= max_ls; ++ls) {
                residual(zc, th, rc, rv2, kv2);
                if (rv2 <= r_vio || kv2 <= k_vio) break;
                alpha *= ls_scale;
                for (int i = 0; i < NZ; ++i) zc[i] = z[i] - alpha * dl[i];
                if (ls == max_ls) residual(zc, th, rc, rv2, kv2);
            }
            for (int i = 0; i < NZ; ++i) { z[i] = zc[i]; r[i] = rc[i]; }
            r_vio = rv2; k_vio = kv2;
            ++it;
        }
        *iters = it;
        return status;
    }

    // δz[rows 0..NROW) = −(rz⁻¹ rθ')[rows], written column-major NROW×NTHP into out.  Returns false if rz is singular.
    template <int NROW>
    OD_HD static bool sensitivities(const double* z, const double* th, double* out) {
        double A[NZ * NZ], rth[NZ * NTHP], col[NZ];
        int piv[NZ];
        M::jac(z, th, A);
        M::jacth(z, th, rth);
        if (!lu_factor(A, piv)) return false;
        for (int c = 0; c < NTHP; ++c) {
            for (int i = 0; i < NZ; ++i) col[i] = rth[i * NTHP + c];
            lu_solve(A, piv, col);
            for (int i = 0; i < NROW; ++i) out[c * NROW + i] = -col[i];
        }
        return true;
    }
};

}  // namespace od
