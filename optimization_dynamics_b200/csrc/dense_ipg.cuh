// Dense-Jacobian interior-point / Newton solve + IFT with G cooperating lanes per problem and the register-resident Gauss–Jordan of
// group_gj.cuh — the latency configuration of dense_ip.cuh for the rocket models (reference src/models/rocket/dynamics.jl:101-269;
// RoboDojo's interior_point_solve! and its dense `lu_solver` are external).  Same algorithm and the same warp-synchronous state
// machine as contact_step_one (contact_ip.cuh): all lanes of a warp iterate in lockstep until the slowest problem of the warp has
// finished; z, the residual and the direction are replicated in every lane, the Jacobian is evaluated redundantly, transposed
// through a small shared-memory staging area into per-lane rows and eliminated in registers with shuffles.
#pragma once
#include "dense_ip.cuh"
#include "group_gj.cuh"

namespace od {

template <class M, int G>
struct DenseIPG {
    static constexpr int NZ = M::NZ, NTH = M::NTH, NTHP = M::NTHP;
    static constexpr int NCONE = M::NORT + M::NSOC;
    // staging pitch: [rz | rθ' or the Newton right-hand side]; bank rule of ContactIP::PW (4 lanes: ≡ 12 mod 16 — dynamics 28 already, projection 14 → 28)
    static constexpr int PW0 = ((NZ + NTHP + 1) / 2) * 2;
    static constexpr int PW = (G <= 4) ? PW0 + ((12 - PW0 % 16) + 16) % 16 : ((PW0 / 2) % 2 == 1) ? PW0 : PW0 + 2;
    static constexpr int WS = NZ * PW;                           // staging doubles per problem
    typedef GroupGJ<NZ, NZ + 1, G> GJ;
    typedef GroupGJ<NZ, NZ + NTHP, G> GJS;
    static constexpr int RPL = GJ::RPL;

    struct Ctx { double* ws; int g; unsigned gm; };
    OD_HD static void sync() {
#ifdef __CUDA_ARCH__
        if (G > 1) __syncwarp();
#else
        if (G > 1) host_team_sync();
#endif
    }

    // rows of rz into the staging area (every lane writes the same values, 16-byte stores); the caller adds further columns
    OD_HD static void stage_jac(const Ctx& c, const double* z, const double* th) {
        double A[NZ * NZ];
        M::jac(z, th, A);
        sync();                                                  // nobody is still reading the staging area
#pragma unroll
        for (int i = 0; i < NZ; ++i) {
            double2* dst = reinterpret_cast<double2*>(c.ws + i * PW);
#pragma unroll
            for (int jj = 0; jj < NZ / 2; ++jj) dst[jj] = make_double2(A[i * NZ + 2 * jj], A[i * NZ + 2 * jj + 1]);
            if (NZ % 2) c.ws[i * PW + NZ - 1] = A[i * NZ + NZ - 1];
        }
    }
    template <int NCOLS>
    OD_HD static void fetch_rows(const Ctx& c, double (&a)[RPL][NCOLS]) {
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            const int r = s * G + c.g;
            const bool live = (G == 1) || ((s + 1) * G <= NZ) || (r < NZ);
            const double2* src = reinterpret_cast<const double2*>(c.ws + (live ? r : 0) * PW);
#pragma unroll
            for (int jj = 0; jj < NCOLS / 2; ++jj) { const double2 v = src[jj]; a[s][2 * jj] = live ? v.x : 0.0; a[s][2 * jj + 1] = live ? v.y : 0.0; }
            if (NCOLS % 2) { const double v = c.ws[(live ? r : 0) * PW + NCOLS - 1]; a[s][NCOLS - 1] = live ? v : 0.0; }
        }
    }

    // solution of the carried (affine) right-hand side, replicated in every lane
    OD_HD static void affine(const double (&a)[RPL][NZ + 1], const int (&piv)[NZ], double* sol, const Ctx& c) {
#if OD_EXTRACT_SMEM
        if constexpr (PW >= GJ::CINV + 2) { GJ::template extract_sm<PW>(a, piv, 0, sol, c.g, c.gm, c.ws); return; }
#endif
        GJ::extract(a, piv, 0, sol, c.gm);
    }

    // Solve to (r_tol, κ_tol); z holds the initial point on entry and the final iterate on return (replicated in every lane).
    OD_HD static int solve(const Ctx& c, double* z, const double* th, double r_tol, double kappa_tol, int max_iter, int max_ls, double ls_scale, int* iters) {
        typedef DenseIP<M> S;
        double r[NZ], D[NZ], zc[NZ], rc[NZ];
        double r_vio = 0.0, k_vio = 0.0, alpha = 0.0;
#pragma unroll
        for (int i = 0; i < NZ; ++i) { D[i] = 0.0; r[i] = 0.0; }
        bool first = true, active = true;
        int it = 0, ls = 0, status = ST_MAXIT;
        for (;;) {
            double rv2, kv2;
#pragma unroll
            for (int i = 0; i < NZ; ++i) zc[i] = z[i] - alpha * D[i];
            S::residual(zc, th, rc, rv2, kv2);
            const bool retry = active && !(first || rv2 <= r_vio || kv2 <= k_vio || ls >= max_ls);
            if (retry) { alpha *= ls_scale; ++ls; }
            if (warp_any(retry)) continue;
            if (active) {
#pragma unroll
                for (int i = 0; i < NZ; ++i) { z[i] = zc[i]; r[i] = rc[i]; }
                r_vio = rv2; k_vio = kv2;
                if (!first) ++it;
                first = false;
                double fs = r_vio + k_vio;
#pragma unroll
                for (int i = 0; i < NZ; ++i) fs += z[i];
                if (!isfinite(fs)) { status = ST_FAIL; active = false; }
                else if (r_vio < r_tol && k_vio < kappa_tol) { status = ST_OK; active = false; }
                else if (it >= max_iter) { status = ST_MAXIT; active = false; }
            }
            if (!warp_any(active)) break;
            // ---- linearise at z; the affine right-hand side (the residual) rides through the elimination as column NZ ---------
            stage_jac(c, z, th);
#pragma unroll
            for (int i = 0; i < NZ; ++i) c.ws[i * PW + NZ] = r[i];
            sync();
            double a[RPL][NZ + 1];
            int piv[NZ];
            fetch_rows<NZ + 1>(c, a);
#if OD_EXTRACT_SMEM
            bool ok;
            if constexpr (PW >= GJ::CINV + 2) ok = GJ::template factor_v2<PW>(a, piv, c.g, c.gm, c.ws);
            else ok = GJ::template factor_sm<PW>(a, piv, c.g, c.gm, c.ws);
#else
            const bool ok = GJ::template factor_sm<PW>(a, piv, c.g, c.gm, c.ws);     // pivot rows through the staging area (group_gj.cuh)
#endif
            if (active && !ok) { status = ST_FAIL; active = false; }
            double dl[NZ];
            if (NCONE > 0) {
                double da[NZ];
                affine(a, piv, da, c);
                const double a_aff = S::step_length(z, da, 1.0);
                const double mu = S::cone_dot(z, da, 0.0) * (1.0 / (NCONE > 0 ? NCONE : 1));
                const double mu_aff = S::cone_dot(z, da, a_aff) * (1.0 / (NCONE > 0 ? NCONE : 1));
                const double ratio = od_min(od_max(0.0, mu_aff * pivot_rcp(mu)), 1.0);
                const double kappa = ratio * ratio * ratio * mu;
#pragma unroll
                for (int i = 0; i < NZ; ++i) dl[i] = r[i];
#pragma unroll
                for (int k = 0; k < M::NORT; ++k) dl[M::ortr(k)] = (r[M::ortr(k)] - kappa) + da[M::ort_p(k)] * da[M::ort_d(k)];
#pragma unroll
                for (int q = 0; q < M::NSOC; ++q) {
                    double acc = 0.0;
#pragma unroll
                    for (int e = 0; e < 3; ++e) acc += da[M::soc_p(q, e)] * da[M::soc_d(q, e)];
                    dl[M::socr(q, 0)] = (r[M::socr(q, 0)] - kappa) + acc;
#pragma unroll
                    for (int e = 1; e < 3; ++e)
                        dl[M::socr(q, e)] = r[M::socr(q, e)] + (da[M::soc_p(q, 0)] * da[M::soc_d(q, e)] + da[M::soc_d(q, 0)] * da[M::soc_p(q, e)]);
                }
                double xm[RPL];
                GJ::mine(dl, xm, c.g);
#if OD_EXTRACT_SMEM
                if constexpr (PW >= GJ::CINV + 2) GJ::template solve_sm<PW>(a, piv, xm, dl, c.g, c.gm, c.ws);
                else GJ::solve(a, piv, xm, dl, c.g, c.gm);
#else
                GJ::solve(a, piv, xm, dl, c.g, c.gm);
#endif
                const double viol = od_max(r_vio, k_vio);
                alpha = S::step_length(z, dl, od_max(0.95, 1.0 - viol * viol));
            } else {
                affine(a, piv, dl, c);
                alpha = 1.0;
            }
#pragma unroll
            for (int i = 0; i < NZ; ++i) D[i] = dl[i];
            if (!active) alpha = 0.0;
            ls = 0;
        }
        *iters = it;
        return status;
    }

    // δz[rows 0..NROW) = −(rz⁻¹ rθ')[rows] into `out` (shared memory, column-major NROW×NTHP, visible to every lane after the
    // caller's sync).  Returns false if rz is singular.
    template <int NROW>
    OD_HD static bool sensitivities(const Ctx& c, const double* z, const double* th, double* out) {
        stage_jac(c, z, th);
        {
            double rth[NZ * NTHP];
            M::jacth(z, th, rth);
#pragma unroll
            for (int i = 0; i < NZ; ++i) {
#pragma unroll
                for (int q = 0; q < NTHP; ++q) c.ws[i * PW + NZ + q] = rth[i * NTHP + q];
            }
        }
        sync();
        double a[RPL][NZ + NTHP];
        int piv[NZ];
        fetch_rows<NZ + NTHP>(c, a);
        const bool ok = GJS::template factor_sm<PW>(a, piv, c.g, c.gm, c.ws);
#pragma unroll
        for (int i = 0; i < NROW; ++i) {
            const int wl = piv[i] & (G - 1), ws = piv[i] >> Grp<G>::LG;
            const double inv = GJS::pick(a, i, ws);
            if (c.g == wl) {
#pragma unroll
                for (int q = 0; q < NTHP; ++q) out[q * NROW + i] = -(GJS::pick(a, NZ + q, ws) * inv);
            }
        }
        return ok;
    }
};

}  // namespace od
