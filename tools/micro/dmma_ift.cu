// Experiment for the north_star's tensor-core clause: does the fp64 tensor path (mma.sync.m8n8k4.f64, "DMMA" — the only
// tensor-core shape with fp64 inputs; tcgen05 has none) help the one dense contraction of the path, the trailing update of the planar
// push's 20×(20+12) IFT elimination?  One warp owns one augmented system, padded to 24×32 = 3×4 tiles of 8×8 held in the DMMA
// accumulator layout (24 doubles per lane).  Five dependent rank-4 panel updates C ← C − L·U (a blocked elimination's data flow; the
// panel factorisation itself is left out of BOTH variants), L = 24×4 panel of the current C, U = 4×32 panel of the current C:
//   variant 0 (SIMT):  every lane updates its 24 elements with 4 FMAs each per panel; L / U values come by shuffle from their owners
//   variant 1 (DMMA):  L / U are moved into the A / B fragment layouts by shuffle, then 12 mma.sync.m8n8k4 per panel
// Both compute the same numbers (checked).  Build + run (B200):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/dmma_ift tools/micro/dmma_ift.cu && tools/micro/dmma_ift
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

constexpr int MT = 3, NT = 4, PANELS = 5;

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// element (r, c) of the 24×32 matrix lives in lane (r%8)*4 + (c%8)/2, register c[r/8][c/8][c%2]
template <int VARIANT>
__global__ void __launch_bounds__(32) update_kernel(const double* __restrict__ in, double* __restrict__ out, int nprob, int reps) {
    const int lane = threadIdx.x, row = lane >> 2, quad = lane & 3;
    const int prob = blockIdx.x;
    if (prob >= nprob) return;
    double c[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) c[i][j][e] = in[((size_t)prob * 24 + i * 8 + row) * 32 + j * 8 + quad * 2 + e];
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
        for (int p = 0; p < PANELS; ++p) {
            // panel p = matrix columns / rows 4p … 4p+3: tile column jp = p/2 (columns (p%2)*4 … +3 of it), tile row ip = p/2
            const int jp = p / 2, ip = p / 2, off = (p % 2) * 4;
            if (VARIANT == 0) {
                // L (24×4) and U (4×32) are read BEFORE the update, as in a blocked elimination (the panel is final when the
                // trailing update starts): both variants compute C − L·U with the same operands
                double l[4][MT], u[4][NT][2];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int col = off + kk;                    // column inside tile jp: owner quad = col/2, element col%2
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        const double v0 = __shfl_sync(0xffffffffu, c[i][jp][0], (lane & ~3) | (col >> 1)), v1 = __shfl_sync(0xffffffffu, c[i][jp][1], (lane & ~3) | (col >> 1));
                        l[kk][i] = 0.015625 * ((col & 1) ? v1 : v0);
                    }
                    const int urow = off + kk;                   // row inside tile row ip: owner lanes urow*4 + quad
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        u[kk][j][0] = __shfl_sync(0xffffffffu, c[ip][j][0], urow * 4 + quad); u[kk][j][1] = __shfl_sync(0xffffffffu, c[ip][j][1], urow * 4 + quad);
                    }
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int j = 0; j < NT; ++j)
#pragma unroll
                        for (int i = 0; i < MT; ++i) { c[i][j][0] = fma(-l[kk][i], u[kk][j][0], c[i][j][0]); c[i][j][1] = fma(-l[kk][i], u[kk][j][1], c[i][j][1]); }
            } else {
                // A fragment of tile row i: A[row][k = quad] = −2⁻⁶·C[8i + row][4p + quad]  → owner quad (off+quad)/2, element (off+quad)%2
                double a[MT], b[NT];
                const int ac = off + quad;
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    const double v0 = __shfl_sync(0xffffffffu, c[i][jp][0], (lane & ~3) | (ac >> 1)), v1 = __shfl_sync(0xffffffffu, c[i][jp][1], (lane & ~3) | (ac >> 1));
                    a[i] = -0.015625 * ((ac & 1) ? v1 : v0);
                }
                // B fragment of tile column j: B[k = quad][n = row] = C[4p + quad][8j + row] → owner lane (off+quad)*4 + row/2, element row%2
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const double v0 = __shfl_sync(0xffffffffu, c[ip][j][0], (off + quad) * 4 + (row >> 1)), v1 = __shfl_sync(0xffffffffu, c[ip][j][1], (off + quad) * 4 + (row >> 1));
                    b[j] = (row & 1) ? v1 : v0;
                }
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) out[((size_t)prob * 24 + i * 8 + row) * 32 + j * 8 + quad * 2 + e] = c[i][j][e];
}

template <int V>
static float time_kernel(const double* in, double* out, int nprob, int reps, int launches) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    update_kernel<V><<<nprob, 32>>>(in, out, nprob, reps);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int k = 0; k < launches; ++k) update_kernel<V><<<nprob, 32>>>(in, out, nprob, reps);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    return ms / launches;
}

int main(int argc, char** argv) {
    const int reps = 20;
    for (int nprob : {592, 4736, 37888}) {      // one warp per SM sub-partition / 8 per sub-partition / saturating
        const size_t n = (size_t)nprob * 24 * 32;
        double* h = (double*)malloc(n * sizeof(double));
        srand(1);
        for (size_t i = 0; i < n; ++i) h[i] = (rand() / (double)RAND_MAX - 0.5) * 0.5;
        double *in, *o0, *o1;
        cudaMalloc(&in, n * sizeof(double)); cudaMalloc(&o0, n * sizeof(double)); cudaMalloc(&o1, n * sizeof(double));
        cudaMemcpy(in, h, n * sizeof(double), cudaMemcpyHostToDevice);
        const float t0 = time_kernel<0>(in, o0, nprob, reps, 20), t1 = time_kernel<1>(in, o1, nprob, reps, 20);
        double* r0 = (double*)malloc(n * sizeof(double)); double* r1 = (double*)malloc(n * sizeof(double));
        cudaMemcpy(r0, o0, n * sizeof(double), cudaMemcpyDeviceToHost); cudaMemcpy(r1, o1, n * sizeof(double), cudaMemcpyDeviceToHost);
        double md = 0, mx = 0;
        for (size_t i = 0; i < n; ++i) { md = fmax(md, fabs(r0[i] - r1[i])); mx = fmax(mx, fabs(r0[i])); }
        const double flop = 2.0 * 24 * 32 * 4 * PANELS * reps * nprob;
        printf("problems %6d (warps/SM %.1f): SIMT %.4f ms (%.2f TFLOP/s, %.3f us per panel chain)   DMMA %.4f ms (%.2f TFLOP/s, %.3f us per panel chain)   DMMA/SIMT time %.2f   max|diff| %.2e (max|x| %.2e)\n",
               nprob, nprob / 148.0, t0, flop / t0 * 1e-9, t0 * 1e3 / reps, t1, flop / t1 * 1e-9, t1 * 1e3 / reps, t1 / t0, md, mx);
        cudaFree(in); cudaFree(o0); cudaFree(o1); free(h); free(r0); free(r1);
    }
    return 0;
}
