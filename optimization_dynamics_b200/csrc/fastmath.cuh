// Short fp64 replacements for the three libm-class operations that sit on the critical path of every interior-point iteration:
// the pivot / cone reciprocals, the two divisions of the centering and step-length rules, and the sin/cos pairs of the generated
// model code (reference: Symbolics-generated `sin`/`cos` calls in the residuals, e.g. src/models/cartpole/model.jl:28-63).
//
// Why: the kernel is bound by the dependent-instruction latency of the slowest problem of a batch.  CUDA's sincos() is ≈ 170 SASS
// instructions per call (two calls per hopper iteration = 10 % of the loop), an IEEE-rounded fp64 division ≈ 30.  The versions
// here are 35 and 5 instructions; they are accurate to ≈ 1 ulp (tests/test_host_logic.py::test_fast_math_accuracy), far inside
// the 1e-8 / 1e-6 parity tolerances, and they are the same code on the host tier (tests/host_check.cu) and on the device.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#ifndef OD_HD
#define OD_HD __host__ __device__ __forceinline__
#endif

namespace od {

// 1/x: hardware seed (MUFU.RCP64H, ≥ 20 bits) + two Newton steps = full double precision to within an ulp, in 5 dependent
// instructions instead of the ~20 of the IEEE-rounded division sequence.  (x = 0 gives NaN, not ±inf: callers never rely on it.)
OD_HD double pivot_rcp(double x) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
#else
    return 1.0 / x;
#endif
}

// 1/sqrt(x) and sqrt(x), x > 0: hardware seed (MUFU.RSQ64H) + two Newton steps (+ one correction for sqrt) ≈ 1 ulp, in 9 / 13
// instructions instead of the ≈ 30 of the IEEE-rounded sqrt() followed by a division.  Used by the 3-D second-order-cone step
// length (planar push, rocket thrust projection), eight times per iteration.  tools/micro/rcp_test.cu measures both on the device.
OD_HD double od_rsqrt(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(0.5 * y, fma(-x * y, y, 1.0), y);
    y = fma(0.5 * y, fma(-x * y, y, 1.0), y);
    return y;
#else
    return 1.0 / sqrt(x);
#endif
}
OD_HD double od_sqrt(double x) {
#ifdef __CUDA_ARCH__
    const double y = od_rsqrt(x);
    double s = x * y;
    s = fma(fma(-s, s, x), 0.5 * y, s);
    return (x == 0.0) ? 0.0 : s;
#else
    return sqrt(x);
#endif
}

// x^N for a compile-time N by square-and-multiply (the generated planar-push code needs x^8 … x^10: 4 multiplications instead of 9)
template <int N> OD_HD double ipow(double x) {
    static_assert(N >= 1, "positive integer power");
    if (N == 1) return x;
    const double h = ipow<(N > 1 ? N / 2 : 1)>(x);
    return (N % 2) ? h * h * x : h * h;
}

// max / min as one compare + select (3 instructions).  fmax()/fmin() expand to ≈ 8 (DSETP.MAX, selects, NaN quieting, moves), and
// the step-length rule and the residual norms take ≈ 45 of them per iteration.  A NaN in the SECOND argument is ignored (as fmax
// does); the first argument is the running value and is never NaN where these are used.
OD_HD double od_max(double a, double b) { return (b > a) ? b : a; }
OD_HD double od_min(double a, double b) { return (b < a) ? b : a; }
// |x| by clearing the sign bit (one integer instruction; fabs() in front of a compare + select costs an fp64-pipe DADD)
OD_HD double od_abs(double x) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x));
#else
    return fabs(x);
#endif
}
// Running maximum that starts from its first candidate (a `max(0, x)` start would be pattern-matched back into fmax).  Empty = 0.
struct MaxAcc {
    double v = 0.0; bool have = false;
    OD_HD void add(double c) { v = have ? od_max(v, c) : c; have = true; }
};

OD_HD int od_lo32(double t) {
#ifdef __CUDA_ARCH__
    return __double2loint(t);
#else
    long long b; memcpy(&b, &t, 8); return (int)(unsigned)(b & 0xffffffffll);
#endif
}
OD_HD double od_flip_sign(double v, int flip) {          // flip ∈ {0, 1}
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(v) ^ (flip << 31), __double2loint(v));
#else
    return flip ? -v : v;
#endif
}

// sin and cos of one argument.  Cody–Waite reduction x = k·π/2 + r with the two-term split of π/2 (exact inside the FMAs for
// |k| < 2^20, i.e. |x| < 1.6e6; beyond that the accuracy degrades gradually, NaN/inf give NaN), then the classical minimax
// polynomials on [−π/4, π/4] (the fdlibm kernel coefficients) and a quadrant rotation.  Max error measured against libm over
// [−1e6, 1e6]: 1.8e-16 absolute.
OD_HD void od_sincos(double x, double* sp, double* cp) {
    const double t = fma(x, 0.63661977236758138, 6755399441055744.0);      // 1.5·2^52: the integer lands in the low mantissa bits
    const int k = od_lo32(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(-kd, 1.5707963267948966, x);
    r = fma(-kd, 6.123233995736766e-17, r);
    const double r2 = r * r;
    double ps = fma(r2, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(r2, ps, 2.75573137070700676789e-06);
    ps = fma(r2, ps, -1.98412698298579493134e-04);
    ps = fma(r2, ps, 8.33333333332248946124e-03);
    ps = fma(r2, ps, -1.66666666666666324348e-01);
    const double S = fma(r * r2, ps, r);
    double pc = fma(r2, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(r2, pc, -2.75573143513906633035e-07);
    pc = fma(r2, pc, 2.48015872894767294178e-05);
    pc = fma(r2, pc, -1.38888888888741095749e-03);
    pc = fma(r2, pc, 4.16666666666666019037e-02);
    const double C = fma(r2 * r2, pc, fma(r2, -0.5, 1.0));
    const bool swap = (k & 1) != 0;
    const double s0 = swap ? C : S, c0 = swap ? S : C;
    *sp = od_flip_sign(s0, (k >> 1) & 1);
    *cp = od_flip_sign(c0, ((k + 1) >> 1) & 1);
}

}  // namespace od
