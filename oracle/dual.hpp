// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or called from the product path
// (optimization_dynamics_b200/csrc).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this code.
//
// Forward-mode dual numbers.  The reference obtains its Jacobians with `Symbolics.jacobian(r, z)` /
// `Symbolics.jacobian(r, θ)` at Pkg.build time (reference: src/models/acrobot/codegen.jl:22-31,
// src/models/cartpole/codegen.jl:20-34, src/models/planar_push/codegen.jl:15-17, src/models/rocket/codegen.jl:24-27,66-72).
// Exact differentiation of the same residual expression is what a dual number computes, so the oracle evaluates
// rz / rθ by seeding z (resp. θ) with unit tangents.  Duals nest (Dual<Dual<double,N>,M>) for the planar-push
// model, whose residual itself contains `Symbolics.jacobian(ϕ, q)` (reference: src/models/planar_push/model.jl:82-85,104).
#pragma once
#include <cmath>
#include <type_traits>

namespace od_oracle {

template <class S, int N>
struct Dual {
    S v;
    S d[N];
    Dual() : v(S(0.0)) { for (int i = 0; i < N; ++i) d[i] = S(0.0); }
    Dual(double c) : v(S(c)) { for (int i = 0; i < N; ++i) d[i] = S(0.0); }
    template <class Q = S, class = std::enable_if_t<!std::is_same<Q, double>::value>>
    Dual(const S& c) : v(c) { for (int i = 0; i < N; ++i) d[i] = S(0.0); }
    static Dual variable(const S& value, int k) { Dual r; r.v = value; r.d[k] = S(1.0); return r; }
};

template <class T> struct is_dual : std::false_type {};
template <class S, int N> struct is_dual<Dual<S, N>> : std::true_type {};

// ---- arithmetic -------------------------------------------------------------------------------------------------
template <class S, int N> inline Dual<S, N> operator+(const Dual<S, N>& a, const Dual<S, N>& b) {
    Dual<S, N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <class S, int N> inline Dual<S, N> operator-(const Dual<S, N>& a, const Dual<S, N>& b) {
    Dual<S, N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <class S, int N> inline Dual<S, N> operator-(const Dual<S, N>& a) {
    Dual<S, N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <class S, int N> inline Dual<S, N> operator*(const Dual<S, N>& a, const Dual<S, N>& b) {
    Dual<S, N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <class S, int N> inline Dual<S, N> operator/(const Dual<S, N>& a, const Dual<S, N>& b) {
    Dual<S, N> r; S inv = S(1.0) / b.v; r.v = a.v * inv;
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv; return r; }

// mixed with double
template <class S, int N> inline Dual<S, N> operator+(const Dual<S, N>& a, double b) { Dual<S, N> r = a; r.v = a.v + S(b); return r; }
template <class S, int N> inline Dual<S, N> operator+(double b, const Dual<S, N>& a) { return a + b; }
template <class S, int N> inline Dual<S, N> operator-(const Dual<S, N>& a, double b) { Dual<S, N> r = a; r.v = a.v - S(b); return r; }
template <class S, int N> inline Dual<S, N> operator-(double b, const Dual<S, N>& a) { return (-a) + b; }
template <class S, int N> inline Dual<S, N> operator*(const Dual<S, N>& a, double b) {
    Dual<S, N> r; r.v = a.v * S(b); for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * S(b); return r; }
template <class S, int N> inline Dual<S, N> operator*(double b, const Dual<S, N>& a) { return a * b; }
template <class S, int N> inline Dual<S, N> operator/(const Dual<S, N>& a, double b) { return a * (1.0 / b); }
template <class S, int N> inline Dual<S, N> operator/(double b, const Dual<S, N>& a) { return Dual<S, N>(b) / a; }

// mixed with the (non-double) inner scalar, needed for nested duals
template <class S, int N, class = std::enable_if_t<is_dual<S>::value>>
inline Dual<S, N> operator*(const Dual<S, N>& a, const S& b) {
    Dual<S, N> r; r.v = a.v * b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b; return r; }
template <class S, int N, class = std::enable_if_t<is_dual<S>::value>>
inline Dual<S, N> operator*(const S& b, const Dual<S, N>& a) { return a * b; }
template <class S, int N, class = std::enable_if_t<is_dual<S>::value>>
inline Dual<S, N> operator+(const Dual<S, N>& a, const S& b) { Dual<S, N> r = a; r.v = a.v + b; return r; }
template <class S, int N, class = std::enable_if_t<is_dual<S>::value>>
inline Dual<S, N> operator-(const Dual<S, N>& a, const S& b) { Dual<S, N> r = a; r.v = a.v - b; return r; }

template <class S, int N> inline Dual<S, N>& operator+=(Dual<S, N>& a, const Dual<S, N>& b) { a = a + b; return a; }
template <class S, int N> inline Dual<S, N>& operator-=(Dual<S, N>& a, const Dual<S, N>& b) { a = a - b; return a; }

// ---- elementary functions ---------------------------------------------------------------------------------------
inline double od_sin(double x) { return std::sin(x); }
inline double od_cos(double x) { return std::cos(x); }
inline double od_sqrt(double x) { return std::sqrt(x); }
inline double od_pow(double x, double p) { return std::pow(x, p); }

template <class S, int N> inline Dual<S, N> od_sin(const Dual<S, N>& a) {
    Dual<S, N> r; r.v = od_sin(a.v); S c = od_cos(a.v); for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i]; return r; }
template <class S, int N> inline Dual<S, N> od_cos(const Dual<S, N>& a) {
    Dual<S, N> r; r.v = od_cos(a.v); S s = -od_sin(a.v); for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }
template <class S, int N> inline Dual<S, N> od_sqrt(const Dual<S, N>& a) {
    Dual<S, N> r; r.v = od_sqrt(a.v); S g = S(0.5) / r.v; for (int i = 0; i < N; ++i) r.d[i] = g * a.d[i]; return r; }
// real power with constant exponent
template <class S, int N> inline Dual<S, N> od_pow(const Dual<S, N>& a, double p) {
    Dual<S, N> r; r.v = od_pow(a.v, p); S g = od_pow(a.v, p - 1.0) * p; for (int i = 0; i < N; ++i) r.d[i] = g * a.d[i]; return r; }

// integer power by repeated squaring (Julia's x^10 for a literal integer exponent)
template <class S> inline S od_ipow(const S& x, int n) {
    S result = S(1.0); S base = x; bool first = true;
    while (n > 0) {
        if (n & 1) { result = first ? base : result * base; first = false; }
        n >>= 1; if (n) base = base * base;
    }
    return result;
}

inline double value_of(double x) { return x; }
template <class S, int N> inline double value_of(const Dual<S, N>& a) { return value_of(a.v); }

}  // namespace od_oracle
