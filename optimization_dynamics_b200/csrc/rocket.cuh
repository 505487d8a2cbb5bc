// Rocket step kernel (see below).  Included by optdyn_b200.cu; host+device so tests/host_check.cu can run it on the CPU.
#pragma once
#include "models.cuh"
#include "dense_ip.cuh"

namespace od {

// ---------------------------------------------------------------------------------------------------------------------
// Rocket: [SOC thrust projection →] implicit-midpoint Newton solve → IFT, chain rule du = ∂y/∂u · ∂proj/∂u
// (reference src/models/rocket/dynamics.jl:101-163 without, :215-269 with projection).  One thread per problem.
// The reference re-solves identical problems in f, fx and fu (and the projection up to 4 times per (x,u)); here each solve
// runs once and its IFT is taken at the same final iterate.
// ---------------------------------------------------------------------------------------------------------------------
struct RocketArgs {
    int B;
    const double* x; const double* u;
    double* y; double* dx; double* du; double* uproj; double* duproj;     // any may be null
    int* status; int* iters;
    double h, u_max;
    int proj, want_grad;
    int proj_only;                                                       // soc_projection(+gradient) alone
    SolverOpts opts;
};

OD_HD void rocket_one(const RocketArgs& a, const int i) {
    double ue[3], dproj[9];
    int st_p = 0, it_p = 0, st_d = 0, it_d = 0;
    for (int k = 0; k < 3; ++k) ue[k] = a.u[(size_t)i * 3 + k];
    if (a.proj) {
        // soc_projection: z = 0.1, z[3] += 1, z[10] += 1, z[7] = 0 (1-based), θ = [u; u_max]   (rocket/dynamics.jl:168-176)
        double zp[10], thp[4];
        for (int k = 0; k < 10; ++k) zp[k] = 0.1;
        zp[2] += 1.0; zp[9] += 1.0; zp[6] = 0.0;
        thp[0] = ue[0]; thp[1] = ue[1]; thp[2] = ue[2]; thp[3] = a.u_max;
        // opts: r_tol 1e-8, κ_tol 1e-4, max_ls 25 (rocket/dynamics.jl:77-86)
        st_p = DenseIP<RocketProjModel>::solve(zp, thp, a.opts.r_tol, a.opts.kappa_eval_tol, a.opts.max_iter, a.opts.max_ls, a.opts.ls_scale, &it_p);
        for (int k = 0; k < 3; ++k) ue[k] = zp[k];
        if (a.want_grad && st_p != ST_FAIL) {
            if (!DenseIP<RocketProjModel>::template sensitivities<3>(zp, thp, dproj)) st_p = ST_FAIL;
        }
        if (a.uproj) for (int k = 0; k < 3; ++k) a.uproj[(size_t)i * 3 + k] = ue[k];
        if (a.duproj && a.want_grad) for (int k = 0; k < 9; ++k) a.duproj[(size_t)i * 9 + k] = dproj[k];
    }
    if (!a.proj_only) {
        double z[12], th[16];
        for (int k = 0; k < 12; ++k) { const double v = a.x[(size_t)i * 12 + k]; z[k] = v; th[k] = v; }   // warm start z = x (:103)
        th[12] = ue[0]; th[13] = ue[1]; th[14] = ue[2]; th[15] = a.h;
        // dynamics opts: r_tol 1e-8, κ_tol 1.0 (no cones), max_ls 25 (rocket/dynamics.jl:21-27)
        st_d = DenseIP<RocketDynModel>::solve(z, th, a.opts.r_tol, 1.0, a.opts.max_iter, a.opts.max_ls, a.opts.ls_scale, &it_d);
        if (a.y) for (int k = 0; k < 12; ++k) a.y[(size_t)i * 12 + k] = z[k];
        if (a.want_grad && st_d != ST_FAIL) {
            double dz[12 * 15];
            if (!DenseIP<RocketDynModel>::template sensitivities<12>(z, th, dz)) st_d = ST_FAIL;
            if (a.dx) for (int k = 0; k < 144; ++k) a.dx[(size_t)i * 144 + k] = dz[k];
            if (a.du) {
                for (int c = 0; c < 3; ++c)
                    for (int r = 0; r < 12; ++r) {
                        double v;
                        if (a.proj) { v = 0.0; for (int k = 0; k < 3; ++k) v += dz[(12 + k) * 12 + r] * dproj[c * 3 + k]; }   // mul!(du, du_dyn, du_proj) (:267)
                        else v = dz[(12 + c) * 12 + r];
                        a.du[(size_t)i * 36 + c * 12 + r] = v;
                    }
            }
        }
    }
    if (a.status) a.status[i] = a.proj_only ? st_p : (st_d | (st_p << 4));
    if (a.iters) a.iters[i] = a.proj_only ? it_p : (it_d | (it_p << 16));
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) rocket_kernel(const RocketArgs a) {
    const int i = blockIdx.x * BLOCK + threadIdx.x;
    if (i >= a.B) return;
    rocket_one(a, i);
}

}  // namespace od
