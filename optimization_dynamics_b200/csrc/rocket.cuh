// Rocket step kernel (see below).  Included by optdyn_b200.cu; host+device so tests/host_check.cu can run it on the CPU.
#pragma once
#include "models.cuh"
#include "dense_ip.cuh"
#include "dense_ipg.cuh"

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

// ---- cooperative-lane version (latency configuration): G lanes per problem, register Gauss–Jordan, warp-synchronous ------------
// Shared-memory workspace per problem: staging rows of the larger system, then the IFT results DZ (12×15) and DPROJ (3×3).
template <bool PHASED>
OD_HD void phase_barrier() {
#ifdef __CUDA_ARCH__
    if (PHASED) __syncthreads();          // (conditions around the call sites are uniform over the block: launch arguments only)
#endif
}
template <int G>
struct RocketG {
    typedef DenseIPG<RocketProjModel, G> P;
    typedef DenseIPG<RocketDynModel, G> Dy;
    static constexpr int STAGE = (P::WS > Dy::WS) ? P::WS : Dy::WS;
    static constexpr int O_DZ = STAGE, O_DP = O_DZ + 12 * 15, WS0 = ((O_DP + 9 + 1) / 2) * 2;
    static constexpr int WS = ((WS0 / 2) % 2 == 1) ? WS0 : WS0 + 2;

    // PHASED (blocks of several warps): a block barrier at each phase boundary — projection loop | its IFT | dynamics loop | its
    // IFT.  The kernel is 143 KB of mostly straight-line code that every warp walks once per problem (ncu: 2.2 stall cycles per
    // issue waiting for instructions, profiles/r02zj_*); warps that enter a phase together share its instruction fetches.
    template <bool PHASED = false>
    OD_HD static void run(const RocketArgs& a, const int i, double* ws, const int g) {
        double ue[3];
        int st_p = 0, it_p = 0, st_d = 0, it_d = 0;
        double* DZ = ws + O_DZ; double* DP = ws + O_DP;
#pragma unroll
        for (int k = 0; k < 3; ++k) ue[k] = a.u[(size_t)i * 3 + k];
        if (a.proj) {
            typename P::Ctx c{ws, g, 0xffffffffu};
            double zp[10], thp[4];
#pragma unroll
            for (int k = 0; k < 10; ++k) zp[k] = 0.1;
            zp[2] += 1.0; zp[9] += 1.0; zp[6] = 0.0;                              // rocket/dynamics.jl:168-176
            thp[0] = ue[0]; thp[1] = ue[1]; thp[2] = ue[2]; thp[3] = a.u_max;
            st_p = P::solve(c, zp, thp, a.opts.r_tol, a.opts.kappa_eval_tol, a.opts.max_iter, a.opts.max_ls, a.opts.ls_scale, &it_p);
#pragma unroll
            for (int k = 0; k < 3; ++k) ue[k] = zp[k];
            phase_barrier<PHASED>();
            if (a.want_grad) {
                if (!P::template sensitivities<3>(c, zp, thp, DP) && st_p != ST_FAIL) st_p = ST_FAIL;
                P::sync();
            }
            if (a.uproj && g == 0) for (int k = 0; k < 3; ++k) a.uproj[(size_t)i * 3 + k] = ue[k];
            if (a.duproj && a.want_grad) for (int k = g; k < 9; k += G) a.duproj[(size_t)i * 9 + k] = DP[k];
        }
        phase_barrier<PHASED>();
        if (!a.proj_only) {
            typename Dy::Ctx c{ws, g, 0xffffffffu};
            double z[12], th[16];
#pragma unroll
            for (int k = 0; k < 12; ++k) { const double v = a.x[(size_t)i * 12 + k]; z[k] = v; th[k] = v; }   // warm start z = x (:103)
            th[12] = ue[0]; th[13] = ue[1]; th[14] = ue[2]; th[15] = a.h;
            st_d = Dy::solve(c, z, th, a.opts.r_tol, 1.0, a.opts.max_iter, a.opts.max_ls, a.opts.ls_scale, &it_d);
            if (a.y) {
#pragma unroll
                for (int k = 0; k < 12; ++k) if (G == 1 || k % G == g) a.y[(size_t)i * 12 + k] = z[k];
            }
            phase_barrier<PHASED>();
            if (a.want_grad) {
                if (!Dy::template sensitivities<12>(c, z, th, DZ) && st_d != ST_FAIL) st_d = ST_FAIL;
                Dy::sync();
                if (a.dx) for (int k = g; k < 144; k += G) a.dx[(size_t)i * 144 + k] = DZ[k];
                if (a.du) {
                    for (int e = g; e < 36; e += G) {
                        const int q = e / 12, r = e % 12;
                        double v;
                        if (a.proj) { v = 0.0; for (int k = 0; k < 3; ++k) v += DZ[(12 + k) * 12 + r] * DP[q * 3 + k]; }   // mul!(du, du_dyn, du_proj) (:267)
                        else v = DZ[(12 + q) * 12 + r];
                        a.du[(size_t)i * 36 + e] = v;
                    }
                }
            }
        }
        if (g == 0) {
            if (a.status) a.status[i] = a.proj_only ? st_p : (st_d | (st_p << 4));
            if (a.iters) a.iters[i] = a.proj_only ? it_p : (it_d | (it_p << 16));
        }
    }
};

template <int G, int PPB, bool PHASED = false>
__global__ void __launch_bounds__(G * PPB) rocket_kernel_g(const RocketArgs a) {
    extern __shared__ __align__(16) double od_smem[];
    static_assert((G * PPB) % 32 == 0, "whole warps: the solve runs warp-synchronously");
    const int slot = threadIdx.x / G, g = threadIdx.x % G;
    int i = blockIdx.x * PPB + slot;
    if (i >= a.B) i = a.B - 1;               // padding lanes repeat the last problem (identical values, same addresses)
    RocketG<G>::template run<PHASED>(a, i, od_smem + slot * RocketG<G>::WS, g);
}

}  // namespace od
