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
//
// The recursion is latency-bound (T−1 dependent steps of 8×8 products): dimensions are template parameters (constant index
// arithmetic, unrolled dot products), the matrices live in shared memory with each lane owning fixed elements, the value function
// ping-pongs between two buffers, and the global loads of step t−1 (Jacobian row, cost blocks) are issued into registers at the top
// of step t so that their HBM/L2 latency overlaps step t's arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#ifndef OD_HD
#define OD_HD __host__ __device__ __forceinline__
#endif

namespace od {

struct RiccatiArgs {
    int NT, T;
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

OD_HD void riccati_sync() {
#ifdef __CUDA_ARCH__
    __syncwarp();
#endif
}

// NL cooperating threads (32 on the device, 1 in the host build of tests/host_check.cu).
template <int NQ, int NU, int NL>
struct Riccati {
    static constexpr int N = 2 * NQ, M = NU, NN = N * N, NM = N * M, MM = M * M;
    static constexpr int ROWW = NQ + NQ * (N + M);
    static constexpr int E_NN = (NN + NL - 1) / NL, E_NM = (NM + NL - 1) / NL, E_N = (N + NL - 1) / NL, E_MM = (MM + NL - 1) / NL, E_M = (M + NL - 1) / NL;
    // shared-memory workspace (doubles)
    static constexpr int O_P0 = 0, O_P1 = O_P0 + NN, O_A = O_P1 + NN, O_PA = O_A + NN, O_QXX = O_PA + NN;
    static constexpr int O_B = O_QXX + NN, O_PB = O_B + NM, O_QUX = O_PB + NM, O_K = O_QUX + NM, O_QK = O_K + NM;
    static constexpr int O_QUU = O_QK + NM, O_p0 = O_QUU + MM, O_p1 = O_p0 + N, O_QX = O_p1 + N, O_QU = O_QX + N, O_k = O_QU + M, O_Qk = O_k + M;
    static constexpr int WS = O_Qk + M;

    struct Fetch { double a[E_NN], lxx[E_NN], b[E_NM], lux[E_NM], luu[E_MM], lx[E_N], lu[E_M]; };

    // this lane's elements of fx, fu and of the cost expansion at knot point t (global loads only; nothing is consumed here)
    OD_HD static void fetch(const RiccatiArgs& a, const int tr, const int t, const int lane, Fetch& f) {
        const int S = a.T - 1;
        const double* row = a.jac + ((size_t)tr * S + t) * ROWW;
        const double* d1 = row + NQ; const double* du = d1 + 2 * NQ * NQ;     // column-major blocks ∂q3/∂q1 | ∂q3/∂q2 | ∂q3/∂u1
        const double* lxxt = a.lxx + ((size_t)tr * a.T + t) * NN;
        const double* luxt = a.lux ? a.lux + ((size_t)tr * S + t) * NM : nullptr;
        const double* luut = a.luu + ((size_t)tr * S + t) * MM;
        const double* lxt = a.lx + ((size_t)tr * a.T + t) * N;
        const double* lut = a.lu + ((size_t)tr * S + t) * M;
#pragma unroll
        for (int q = 0; q < E_NN; ++q) {
            const int e = lane + q * NL;
            if (NN % NL == 0 || e < NN) {
                const int i = e / N, j = e % N;
                // rows 0..NQ−1 of fx are [0 I]; the load below always uses a valid offset of the packed row (top rows read element 0)
                const int off = (i < NQ) ? 0 : ((j < NQ) ? j * NQ + (i - NQ) : NQ * NQ + (j - NQ) * NQ + (i - NQ));
                const double v = d1[off];
                f.a[q] = (i < NQ) ? ((j == i + NQ) ? 1.0 : 0.0) : v;
                f.lxx[q] = lxxt[e];
            }
        }
#pragma unroll
        for (int q = 0; q < E_NM; ++q) {
            const int e = lane + q * NL;
            if (NM % NL == 0 || e < NM) {
                const int i = e / M, j = e % M;                     // fu is n×m row-major; lux is m×n row-major (same count)
                const double v = du[(i < NQ) ? 0 : j * NQ + (i - NQ)];
                f.b[q] = (i < NQ) ? 0.0 : v;
                f.lux[q] = luxt ? luxt[e] : 0.0;
            }
        }
#pragma unroll
        for (int q = 0; q < E_MM; ++q) { const int e = lane + q * NL; if (MM % NL == 0 || e < MM) f.luu[q] = luut[e]; }
#pragma unroll
        for (int q = 0; q < E_N; ++q) { const int e = lane + q * NL; if (N % NL == 0 || e < N) f.lx[q] = lxt[e]; }
#pragma unroll
        for (int q = 0; q < E_M; ++q) { const int e = lane + q * NL; if (M % NL == 0 || e < M) f.lu[q] = lut[e]; }
    }

    OD_HD static void run(const RiccatiArgs& a, const int tr, double* ws, const int lane) {
        const int T = a.T, S = T - 1;
        double* P = ws + O_P0; double* Pn = ws + O_P1; double* p = ws + O_p0; double* pn = ws + O_p1;
        double* A = ws + O_A; double* PA = ws + O_PA; double* Qxx = ws + O_QXX; double* Bm = ws + O_B; double* PB = ws + O_PB;
        double* Qux = ws + O_QUX; double* Kt = ws + O_K; double* QK = ws + O_QK; double* Quu = ws + O_QUU;
        double* Qx = ws + O_QX; double* Qu = ws + O_QU; double* kt = ws + O_k; double* Qk = ws + O_Qk;
        {
            const double* lxT = a.lx + ((size_t)tr * T + S) * N;
            const double* lxxT = a.lxx + ((size_t)tr * T + S) * NN;
            for (int e = lane; e < NN; e += NL) P[e] = lxxT[e];
            for (int e = lane; e < N; e += NL) p[e] = lxT[e];
        }
        double dv1 = 0.0, dv2 = 0.0;
        int bad = 0;
        Fetch cur, nxt;
        fetch(a, tr, S - 1, lane, cur);
        for (int t = S - 1; t >= 0; --t) {
            if (t > 0) fetch(a, tr, t - 1, lane, nxt);               // in flight during this step
#pragma unroll
            for (int q = 0; q < E_NN; ++q) { const int e = lane + q * NL; if (NN % NL == 0 || e < NN) A[e] = cur.a[q]; }
#pragma unroll
            for (int q = 0; q < E_NM; ++q) { const int e = lane + q * NL; if (NM % NL == 0 || e < NM) Bm[e] = cur.b[q]; }
            riccati_sync();                                          // A, B, and the P / p of the previous step are visible
#pragma unroll
            for (int q = 0; q < E_NN; ++q) {
                const int e = lane + q * NL;
                if (NN % NL == 0 || e < NN) { const int i = e / N, j = e % N; double s = 0.0;
#pragma unroll
                    for (int l = 0; l < N; ++l) s += P[i * N + l] * A[l * N + j];
                    PA[e] = s; }
            }
#pragma unroll
            for (int q = 0; q < E_NM; ++q) {
                const int e = lane + q * NL;
                if (NM % NL == 0 || e < NM) { const int i = e / M, j = e % M; double s = 0.0;
#pragma unroll
                    for (int l = 0; l < N; ++l) s += P[i * N + l] * Bm[l * M + j];
                    PB[e] = s; }
            }
            riccati_sync();
#pragma unroll
            for (int q = 0; q < E_NN; ++q) {
                const int e = lane + q * NL;
                if (NN % NL == 0 || e < NN) { const int i = e / N, j = e % N; double s = cur.lxx[q];
#pragma unroll
                    for (int l = 0; l < N; ++l) s += A[l * N + i] * PA[l * N + j];
                    Qxx[e] = s; }
            }
#pragma unroll
            for (int q = 0; q < E_NM; ++q) {
                const int e = lane + q * NL;
                if (NM % NL == 0 || e < NM) { const int i = e / N, j = e % N; double s = cur.lux[q];
#pragma unroll
                    for (int l = 0; l < N; ++l) s += Bm[l * M + i] * PA[l * N + j];
                    Qux[e] = s; }
            }
#pragma unroll
            for (int q = 0; q < E_MM; ++q) {
                const int e = lane + q * NL;
                if (MM % NL == 0 || e < MM) { const int i = e / M, j = e % M; double s = cur.luu[q] + (i == j ? a.reg : 0.0);
#pragma unroll
                    for (int l = 0; l < N; ++l) s += Bm[l * M + i] * PB[l * M + j];
                    Quu[e] = s; }
            }
#pragma unroll
            for (int q = 0; q < E_N; ++q) {
                const int e = lane + q * NL;
                if (N % NL == 0 || e < N) { double s = cur.lx[q];
#pragma unroll
                    for (int l = 0; l < N; ++l) s += A[l * N + e] * p[l];
                    Qx[e] = s; }
            }
#pragma unroll
            for (int q = 0; q < E_M; ++q) {
                const int e = lane + q * NL;
                if (M % NL == 0 || e < M) { double s = cur.lu[q];
#pragma unroll
                    for (int l = 0; l < N; ++l) s += Bm[l * M + e] * p[l];
                    Qu[e] = s; }
            }
            riccati_sync();
            // Cholesky of Quu redundantly in every thread; thread c then solves column c of −[Qux | Qu] and forms Quu·(that column)
            double Lr[MM], Qm[MM];
            bool pd = true;
#pragma unroll
            for (int e = 0; e < MM; ++e) Qm[e] = Quu[e];
#pragma unroll
            for (int i = 0; i < M; ++i) {
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    double s = Qm[i * M + j];
#pragma unroll
                    for (int l = 0; l < j; ++l) s -= Lr[i * M + l] * Lr[j * M + l];
                    if (i == j) { pd = pd && (s > 0.0); Lr[i * M + i] = 1.0 / sqrt(s); }      // reciprocal of the diagonal
                    else Lr[i * M + j] = s * Lr[j * M + j];
                }
            }
            if (!pd) bad = 1;
            for (int c = lane; c < N + 1; c += NL) {
                double y[M];
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    double s = -((c < N) ? Qux[i * N + c] : Qu[i]);
#pragma unroll
                    for (int l = 0; l < i; ++l) s -= Lr[i * M + l] * y[l];
                    y[i] = s * Lr[i * M + i];
                }
#pragma unroll
                for (int i = M - 1; i >= 0; --i) {
                    double s = y[i];
#pragma unroll
                    for (int l = i + 1; l < M; ++l) s -= Lr[l * M + i] * y[l];
                    y[i] = s * Lr[i * M + i];
                }
#pragma unroll
                for (int i = 0; i < M; ++i) y[i] = pd ? y[i] : 0.0;
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    double s = 0.0;
#pragma unroll
                    for (int l = 0; l < M; ++l) s += Qm[i * M + l] * y[l];
                    if (c < N) { Kt[i * N + c] = y[i]; QK[i * N + c] = s; } else { kt[i] = y[i]; Qk[i] = s; }
                }
            }
            riccati_sync();
#pragma unroll
            for (int q = 0; q < E_NN; ++q) {
                const int e = lane + q * NL;
                if (NN % NL == 0 || e < NN) { const int i = e / N, j = e % N; double s = Qxx[e];
#pragma unroll
                    for (int l = 0; l < M; ++l) s += Kt[l * N + i] * QK[l * N + j] + Kt[l * N + i] * Qux[l * N + j] + Qux[l * N + i] * Kt[l * N + j];
                    Pn[e] = s; }
            }
#pragma unroll
            for (int q = 0; q < E_N; ++q) {
                const int e = lane + q * NL;
                if (N % NL == 0 || e < N) { double s = Qx[e];
#pragma unroll
                    for (int l = 0; l < M; ++l) s += Kt[l * N + e] * Qk[l] + Kt[l * N + e] * Qu[l] + Qux[l * N + e] * kt[l];
                    pn[e] = s; }
            }
            double* Ko = a.K + ((size_t)tr * S + t) * NM;
            double* ko = a.k + ((size_t)tr * S + t) * M;
            for (int e = lane; e < NM; e += NL) Ko[e] = Kt[e];
            for (int e = lane; e < M; e += NL) ko[e] = kt[e];
#pragma unroll
            for (int l = 0; l < M; ++l) { dv1 += kt[l] * Qu[l]; dv2 += 0.5 * kt[l] * Qk[l]; }
            { double* tmp = P; P = Pn; Pn = tmp; tmp = p; p = pn; pn = tmp; }      // (the sync at the top of the next step publishes them)
            cur = nxt;
        }
        if (lane == 0) {
            if (a.dV) { a.dV[2 * (size_t)tr] = dv1; a.dV[2 * (size_t)tr + 1] = dv2; }
            if (a.status) a.status[tr] = bad;
        }
    }
};

template <int NQ, int NU, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) riccati_kernel(const RiccatiArgs a) {
    extern __shared__ __align__(16) double od_smem[];
    const int w = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int tr = blockIdx.x * WARPS + w;
    if (tr >= a.NT) return;                                   // whole warps leave together: no divergence around __syncwarp
    Riccati<NQ, NU, 32>::run(a, tr, od_smem + (size_t)w * Riccati<NQ, NU, 32>::WS, lane);
}

}  // namespace od
