// Batched Riccati backward pass of iLQR on the device — the consumer of the per-timestep Jacobians.
//
// In the reference the derivative sweep (fx, fu at every knot point; reference src/dynamics.jl:96-128) feeds IterativeLQR's
// backward pass (external package, driven by iLQR.solve!, reference examples/hopper.jl:292), which is sequential in t:
//     P_T = lxx_T,  p_T = lx_T
//     Qx = lx + fxᵀ p⁺,  Qu = lu + fuᵀ p⁺,  Qxx = lxx + fxᵀ P⁺ fx,  Quu = luu + fuᵀ P⁺ fu (+ reg·I),  Qux = lux + fuᵀ P⁺ fx
//     K = −Quu⁻¹ Qux,  k = −Quu⁻¹ Qu                                   (Cholesky; status 1 if Quu is not positive definite)
//     P = Qxx + Kᵀ Quu K + Kᵀ Qux + Quxᵀ K,   p = Qx + Kᵀ Quu k + Kᵀ Qu + Quxᵀ k
// Here one warp runs one trajectory; NT trajectories (samples × rollouts × shooting segments of a batch) run side by side, reading
// the packed rows [q3 | ∂q3/∂q1 | ∂q3/∂q2 | ∂q3/∂u1] exactly as contact_step_kernel (or the fused all-gather) left them in HBM:
//     fx = [0 I; ∂q3/∂q1 ∂q3/∂q2]  (2nq × 2nq),   fu = [0; ∂q3/∂u1]  (2nq × nu)          reference src/dynamics.jl:105-111,125
// and writing the gains in the layout od_rollout_batch consumes (K: [t][control][state]) — sweep → backward pass → line-search
// rollouts chain on one stream without the Jacobians ever visiting the host.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#ifndef OD_HD
#define OD_HD __host__ __device__ __forceinline__
#endif

namespace od {

struct RiccatiArgs {
    int NT, T, nq, nu;
    const double* jac;        // NT × (T−1) packed rows, width nq + nq(2nq+nu)
    const double* lx;         // NT × T × n          (n = 2nq)
    const double* lu;         // NT × (T−1) × m      (m = nu)
    const double* lxx;        // NT × T × n × n      (symmetric)
    const double* luu;        // NT × (T−1) × m × m  (symmetric)
    const double* lux;        // NT × (T−1) × m × n, or null (zero)
    double reg;               // added to the diagonal of Quu
    double* K;                // NT × (T−1) × m × n
    double* k;                // NT × (T−1) × m
    double* dV;               // NT × 2: Σ kᵀQu and ½ Σ kᵀQuu k (expected cost change of a unit step), or null
    int* status;              // NT: 0 ok, 1 = some Quu not positive definite (gains of that step are zero), or null
};

constexpr int RICCATI_MAX_N = 12, RICCATI_MAX_M = 3;
// workspace doubles per trajectory
__host__ __device__ constexpr int riccati_ws(int n, int m) { return 5 * n * n + 5 * n * m + 3 * m * m + 4 * n + 4 * m + 8; }

OD_HD void riccati_sync() {
#ifdef __CUDA_ARCH__
    __syncwarp();
#endif
}

// One trajectory; `lane`/`nl` = this thread's index and the number of cooperating threads (32 on the device, 1 on the host).
OD_HD void riccati_one(const RiccatiArgs& a, const int tr, double* ws, const int lane, const int nl) {
    const int nq = a.nq, m = a.nu, n = 2 * nq, T = a.T, S = T - 1;
    const int roww = nq + nq * (n + m);
    double* P = ws;                 // n×n   value Hessian at t+1
    double* A = P + n * n;          // n×n   fx
    double* PA = A + n * n;         // n×n   P fx
    double* Qxx = PA + n * n;       // n×n
    double* Pn = Qxx + n * n;       // n×n   value Hessian at t
    double* Bm = Pn + n * n;        // n×m   fu
    double* PB = Bm + n * m;        // n×m   P fu
    double* Qux = PB + n * m;       // m×n
    double* Kt = Qux + n * m;       // m×n
    double* QK = Kt + n * m;        // m×n   Quu K
    double* Quu = QK + n * m;       // m×m
    double* Lc = Quu + m * m;       // m×m   Cholesky factor
    double* sp = Lc + 2 * m * m;    // scalars
    double* p = sp + 8;             // n     value gradient at t+1
    double* pn = p + n;             // n
    double* Qx = pn + n;            // n
    double* Qu = Qx + n;            // m
    double* kt = Qu + m;            // m
    double* Qk = kt + m;            // m     Quu k
    const double* lxT = a.lx + ((size_t)tr * T + S) * n;
    const double* lxxT = a.lxx + ((size_t)tr * T + S) * n * n;
    for (int e = lane; e < n * n; e += nl) P[e] = lxxT[e];
    for (int e = lane; e < n; e += nl) p[e] = lxT[e];
    double dv1 = 0.0, dv2 = 0.0;
    int bad = 0;
    for (int t = S - 1; t >= 0; --t) {
        riccati_sync();
        const double* row = a.jac + ((size_t)tr * S + t) * roww;
        const double* d1 = row + nq; const double* d2 = d1 + nq * nq; const double* du = d2 + nq * nq;     // column-major blocks
        for (int e = lane; e < n * n; e += nl) {
            const int i = e / n, j = e % n;
            double v;
            if (i < nq) v = (j == i + nq) ? 1.0 : 0.0;
            else v = (j < nq) ? d1[j * nq + (i - nq)] : d2[(j - nq) * nq + (i - nq)];
            A[e] = v;
        }
        for (int e = lane; e < n * m; e += nl) { const int i = e / m, j = e % m; Bm[e] = (i < nq) ? 0.0 : du[j * nq + (i - nq)]; }
        riccati_sync();
        for (int e = lane; e < n * n; e += nl) { const int i = e / n, j = e % n; double s = 0.0; for (int l = 0; l < n; ++l) s += P[i * n + l] * A[l * n + j]; PA[e] = s; }
        for (int e = lane; e < n * m; e += nl) { const int i = e / m, j = e % m; double s = 0.0; for (int l = 0; l < n; ++l) s += P[i * n + l] * Bm[l * m + j]; PB[e] = s; }
        riccati_sync();
        const double* lxt = a.lx + ((size_t)tr * T + t) * n;
        const double* lut = a.lu + ((size_t)tr * S + t) * m;
        const double* lxxt = a.lxx + ((size_t)tr * T + t) * n * n;
        const double* luut = a.luu + ((size_t)tr * S + t) * m * m;
        const double* luxt = a.lux ? a.lux + ((size_t)tr * S + t) * m * n : nullptr;
        for (int e = lane; e < n * n; e += nl) { const int i = e / n, j = e % n; double s = lxxt[e]; for (int l = 0; l < n; ++l) s += A[l * n + i] * PA[l * n + j]; Qxx[e] = s; }
        for (int e = lane; e < m * n; e += nl) { const int i = e / n, j = e % n; double s = luxt ? luxt[e] : 0.0; for (int l = 0; l < n; ++l) s += Bm[l * m + i] * PA[l * n + j]; Qux[e] = s; }
        for (int e = lane; e < m * m; e += nl) { const int i = e / m, j = e % m; double s = luut[e] + (i == j ? a.reg : 0.0); for (int l = 0; l < n; ++l) s += Bm[l * m + i] * PB[l * m + j]; Quu[e] = s; }
        for (int e = lane; e < n; e += nl) { double s = lxt[e]; for (int l = 0; l < n; ++l) s += A[l * n + e] * p[l]; Qx[e] = s; }
        for (int e = lane; e < m; e += nl) { double s = lut[e]; for (int l = 0; l < n; ++l) s += Bm[l * m + e] * p[l]; Qu[e] = s; }
        riccati_sync();
        // Cholesky of Quu (m ≤ 3), redundantly in every thread; then thread j solves column j of −[Qux | Qu]
        double Lr[RICCATI_MAX_M * RICCATI_MAX_M];
        bool pd = true;
        for (int i = 0; i < m; ++i) {
            for (int j = 0; j <= i; ++j) {
                double s = Quu[i * m + j];
                for (int l = 0; l < j; ++l) s -= Lr[i * m + l] * Lr[j * m + l];
                if (i == j) { pd = pd && (s > 0.0); Lr[i * m + i] = sqrt(s); }
                else Lr[i * m + j] = s / Lr[j * m + j];
            }
        }
        if (!pd) bad = 1;
        for (int c = lane; c < n + 1; c += nl) {
            double y[RICCATI_MAX_M];
            for (int i = 0; i < m; ++i) {
                double s = -((c < n) ? Qux[i * n + c] : Qu[i]);
                for (int l = 0; l < i; ++l) s -= Lr[i * m + l] * y[l];
                y[i] = s / Lr[i * m + i];
            }
            for (int i = m - 1; i >= 0; --i) {
                double s = y[i];
                for (int l = i + 1; l < m; ++l) s -= Lr[l * m + i] * y[l];
                y[i] = s / Lr[i * m + i];
            }
            for (int i = 0; i < m; ++i) { const double v = pd ? y[i] : 0.0; if (c < n) Kt[i * n + c] = v; else kt[i] = v; }
        }
        riccati_sync();
        for (int e = lane; e < m * n; e += nl) { const int i = e / n, j = e % n; double s = 0.0; for (int l = 0; l < m; ++l) s += Quu[i * m + l] * Kt[l * n + j]; QK[e] = s; }
        for (int e = lane; e < m; e += nl) { double s = 0.0; for (int l = 0; l < m; ++l) s += Quu[e * m + l] * kt[l]; Qk[e] = s; }
        riccati_sync();
        for (int e = lane; e < n * n; e += nl) {
            const int i = e / n, j = e % n;
            double s = Qxx[e];
            for (int l = 0; l < m; ++l) s += Kt[l * n + i] * QK[l * n + j] + Kt[l * n + i] * Qux[l * n + j] + Qux[l * n + i] * Kt[l * n + j];
            Pn[e] = s;
        }
        for (int e = lane; e < n; e += nl) {
            double s = Qx[e];
            for (int l = 0; l < m; ++l) s += Kt[l * n + e] * Qk[l] + Kt[l * n + e] * Qu[l] + Qux[l * n + e] * kt[l];
            pn[e] = s;
        }
        double* Ko = a.K + ((size_t)tr * S + t) * m * n;
        double* ko = a.k + ((size_t)tr * S + t) * m;
        for (int e = lane; e < m * n; e += nl) Ko[e] = Kt[e];
        for (int e = lane; e < m; e += nl) ko[e] = kt[e];
        for (int l = 0; l < m; ++l) { dv1 += kt[l] * Qu[l]; dv2 += 0.5 * kt[l] * Qk[l]; }
        riccati_sync();
        for (int e = lane; e < n * n; e += nl) P[e] = Pn[e];
        for (int e = lane; e < n; e += nl) p[e] = pn[e];
    }
    if (lane == 0) {
        if (a.dV) { a.dV[2 * (size_t)tr] = dv1; a.dV[2 * (size_t)tr + 1] = dv2; }
        if (a.status) a.status[tr] = bad;
    }
}

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS) riccati_kernel(const RiccatiArgs a) {
    extern __shared__ __align__(16) double od_smem[];
    const int w = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int tr = blockIdx.x * WARPS + w;
    if (tr >= a.NT) return;                                   // whole warps only: no partial-warp divergence around __syncwarp
    riccati_one(a, tr, od_smem + (size_t)w * riccati_ws(2 * a.nq, a.nu), lane, 32);
}

}  // namespace od
