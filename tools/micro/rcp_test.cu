// accuracy of rcp.approx.ftz.f64 + Newton steps (tools/micro: measurement helpers, not part of the library)
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(const double* x, double* o0, double* o1, double* o2, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[i]));
    o0[i] = r; r = fma(r, fma(-x[i], r, 1.0), r); o1[i] = r; r = fma(r, fma(-x[i], r, 1.0), r); o2[i] = r;
}
#include "../../optimization_dynamics_b200/csrc/fastmath.cuh"
__global__ void k2(const double* x, double* o0, double* o1, double* o2, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    o0[i] = od::od_rsqrt(x[i]); o1[i] = od::od_sqrt(x[i]);
    double s, c; od::od_sincos(fmod(x[i], 1.0e4) - 5.0e3, &s, &c); o2[i] = s;
}
int main() {
    const int n = 1 << 20; double *x, *a, *b, *c;
    cudaMallocManaged(&x, n * 8); cudaMallocManaged(&a, n * 8); cudaMallocManaged(&b, n * 8); cudaMallocManaged(&c, n * 8);
    unsigned long long s = 88172645463325252ull;
    for (int i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; double m = 1.0 + (double)(s >> 11) / 9007199254740992.0; int e = (int)(s % 600) - 300; x[i] = ldexp(m, e) * ((s & 1) ? 1 : -1); }
    k<<<n / 256, 256>>>(x, a, b, c, n); cudaDeviceSynchronize();
    double e0 = 0, e1 = 0, e2 = 0;
    for (int i = 0; i < n; ++i) { long double t = 1.0L / (long double)x[i]; e0 = fmax(e0, fabs((double)((a[i] - t) / t))); e1 = fmax(e1, fabs((double)((b[i] - t) / t))); e2 = fmax(e2, fabs((double)((c[i] - t) / t))); }
    printf("rcp.approx.ftz.f64 max rel err: seed %.3e, +1 Newton %.3e, +2 Newton %.3e (eps = %.3e)\n", e0, e1, e2, ldexp(1.0, -53));
    // od_rsqrt / od_sqrt / od_sincos of csrc/fastmath.cuh on the device (positive arguments for the roots)
    for (int i = 0; i < n; ++i) x[i] = fabs(x[i]);
    k2<<<n / 256, 256>>>(x, a, b, c, n); cudaDeviceSynchronize();
    double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    for (int i = 0; i < n; ++i) {
        const long double t = 1.0L / sqrtl((long double)x[i]), u = sqrtl((long double)x[i]);
        r0 = fmax(r0, fabs((double)((a[i] - t) / t))); r1 = fmax(r1, fabs((double)((b[i] - u) / u)));
        const double ang = fmod(x[i], 1.0e4) - 5.0e3;       // the angle k2 used
        r2 = fmax(r2, fabs((double)(c[i] - sinl((long double)ang))));
    }
    printf("fastmath.cuh on the device: od_rsqrt max rel err %.3e, od_sqrt %.3e, od_sincos(sin) max abs err %.3e\n", r0, r1, r2);
    (void)r3;
    return 0;
}
