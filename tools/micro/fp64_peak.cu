// fp64 FMA throughput of the GPU ("fp64 SIMT peak of B200 is not in MEASURED_PEAKS — measure with an FMA
// microbenchmark and commit").  tools/micro: measurement helper, not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/fp64_peak tools/micro/fp64_peak.cu && tools/micro/fp64_peak
// Every thread runs 8 independent DFMA chains (enough ILP to cover the pipe latency with 4 warps per scheduler); the result is
// written so the loop cannot be removed.  Prints TFLOP/s (2 flops per FMA) for a few occupancies and the best of them.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void fma_chain(double* out, double a, double b, int iters) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], b, a);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int iters = 1 << 14;
    double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 2048);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int threads : {128, 256, 512, 1024}) {
        for (int bps : {1, 2}) {
            if (threads * bps > 2048) continue;
            const int grid = p.multiProcessorCount * bps;
            fma_chain<8><<<grid, threads>>>(out, 1.0000001, 0.9999999, iters);   // warm-up
            cudaEventRecord(e0);
            for (int r = 0; r < 5; ++r) fma_chain<8><<<grid, threads>>>(out, 1.0000001, 0.9999999, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double flops = 2.0 * 8 * (double)iters * threads * grid * 5;
            const double tf = flops / (ms * 1e-3) / 1e12;
            printf("%4d threads x %d blocks/SM: %.2f TFLOP/s fp64\n", threads, bps, tf);
            if (tf > best) best = tf;
        }
    }
    printf("{\"fp64_fma_tflops\": %.2f, \"gpu\": \"%s\", \"sms\": %d}\n", best, p.name, p.multiProcessorCount);
    return 0;
}
