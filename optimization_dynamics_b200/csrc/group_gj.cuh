// Register-resident Gauss–Jordan elimination for ONE small dense system shared by a group of G lanes of a warp.
//
// Replaces RoboDojo's `lu_solver` / `linear_solve!` (dense partial-pivoting LU; used in-tree at reference src/gradient_bundle.jl:76,
// src/ls.jl:52 and inside interior_point_solve!) for the latency configuration of the step kernel, where a 4096-problem batch leaves
// one warp per SM sub-partition and the kernel time is the dependent-instruction latency of the slowest problem.
//
// Layout: row r of the NR×NCOL augmented matrix [K | right-hand sides] lives in lane g = r mod G, slot s = r div G, entirely in
// registers (a[s][j], every j a compile-time index):
//   * pivot search  = per-lane max over its slots + log2(G) shuffle rounds on a 32-bit key (high word of |a|, row index in the low
//                     5 bits — partial pivoting to ~15 significant bits of the candidates, ties broken by row index);
//   * row exchange  = none: pivoting is implicit (the pivot row of step k stays where it is, `piv[k]` remembers it);
//   * elimination   = the pivot row reaches every lane — through the shared-memory staging area the rows came from (factor_sm,
//                     the shipped path: 16-byte stores and broadcast loads) or by shuffles (factor) — and every lane updates its
//                     own rows (all rows, above and below: Gauss–Jordan costs the same as LU when rows are spread over lanes, and
//                     it needs no back-substitution).
// After factor(): column k (< NR) of row r holds the multiplier that step k applied to row r (1/pivot for the pivot row itself),
// so further right-hand sides can be pushed through the same elimination (solve()); the carried columns hold the reduced
// right-hand sides, x_k = a[piv[k]][NR + c] / pivot_k (extract()).
// Accuracy: measured on the hopper's reduced KKT matrices at converged iterates (cond ~1e19 from the complementarity scaling), the
// q-block of the solution agrees with extended-precision LU to the same 1e-15 relative as partial-pivoting LU does.
//
// G = 1 degenerates to a plain single-thread Gauss–Jordan (all shuffles are identities): that is what tests/host_check.cu runs on
// the CPU to check this algebra against the oracle without a GPU.
#pragma once
#include <cuda_runtime.h>
#include <string.h>
#include "fastmath.cuh"

#ifndef OD_HD
#define OD_HD __host__ __device__ __forceinline__
#endif

// 1 (prepared experiment, not measured yet): the inverse pivots are kept in the shared-memory mirror and extract() / the tail of
// solve() publish each lane's scaled rows there and gather the solution with plain loads (extract_sm / solve_sm) instead of two
// slot selects, a multiply and two shuffles per unknown.  Static count on the hopper kernel: see DESIGN.md §9.
#ifndef OD_EXTRACT_SMEM
#define OD_EXTRACT_SMEM 0
#endif

namespace od {

// Host tier only (tests/host_check.cu): a team of lock-stepped host threads can stand in for the lanes of a warp.  When one is
// installed for the calling thread, the G > 1 collectives below exchange real values between the threads, so the multi-lane code
// paths (row distribution, pivot search, shared-memory mirror, warp votes) run on the CPU exactly as written.  Without a team
// (the default, and always in liboptdyn_b200.so, which never runs these templates on the host) they are the G = 1 identities.
struct HostLaneTeam {
    int lane = 0;                                        // this thread's lane index within the emulated warp
    virtual ~HostLaneTeam() {}
    virtual double shfl_f64(double v, int src_lane) = 0;
    virtual unsigned shfl_u32(unsigned v, int src_lane) = 0;
    virtual bool any(bool p) = 0;
    virtual void sync() = 0;
};
inline HostLaneTeam*& host_lane_team() { static thread_local HostLaneTeam* t = nullptr; return t; }
OD_HD void host_team_sync() {
#ifndef __CUDA_ARCH__
    if (HostLaneTeam* t = host_lane_team()) t->sync();
#endif
}

template <int G> struct Grp {
    static constexpr int LG = (G == 1 ? 0 : G == 2 ? 1 : G == 4 ? 2 : G == 8 ? 3 : G == 16 ? 4 : 5);
    static_assert((1 << LG) == G, "lanes per problem must be a power of two");
    // value of lane `src` (index within the group) in every lane of the group
    OD_HD static double bcast(double v, int src, unsigned m) {
#ifdef __CUDA_ARCH__
        if (G > 1) return __shfl_sync(m, v, src, G);
#else
        if (G > 1) if (HostLaneTeam* t = host_lane_team()) return t->shfl_f64(v, (t->lane & ~(G - 1)) + src);
#endif
        return v;
    }
    // (largest value, smallest index among equals) over the lanes of the group, replicated in every lane
    OD_HD static void argmax_all(double& best, int& idx, unsigned m) {
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            const double ob = __shfl_xor_sync(m, best, d, G); const int oi = __shfl_xor_sync(m, idx, d, G);
            if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
        }
#else
        if (HostLaneTeam* t = host_lane_team())
            for (int d = 1; d < G; d <<= 1) {
                const double ob = t->shfl_f64(best, t->lane ^ d); const int oi = (int)t->shfl_u32((unsigned)idx, t->lane ^ d);
                if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
            }
#endif
    }
    // largest value over the lanes of the group (values are never NaN where this is used), replicated in every lane
    OD_HD static double dmax_all(double v, unsigned m) {
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int d = 1; d < G; d <<= 1) { const double o = __shfl_xor_sync(m, v, d, G); v = o > v ? o : v; }
#else
        if (HostLaneTeam* t = host_lane_team())
            for (int d = 1; d < G; d <<= 1) { const double o = t->shfl_f64(v, t->lane ^ d); v = o > v ? o : v; }
#endif
        return v;
    }
    OD_HD static unsigned umax_all(unsigned v, unsigned m) {
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int d = 1; d < G; d <<= 1) { const unsigned o = __shfl_xor_sync(m, v, d, G); v = o > v ? o : v; }
#else
        if (HostLaneTeam* t = host_lane_team())
            for (int d = 1; d < G; d <<= 1) { const unsigned o = t->shfl_u32(v, t->lane ^ d); v = o > v ? o : v; }
#endif
        return v;
    }
};

OD_HD unsigned abs_hi32(double v) {
#ifdef __CUDA_ARCH__
    return (unsigned)__double2hiint(v) & 0x7fffffffu;
#else
    unsigned long long b; memcpy(&b, &v, 8); return (unsigned)(b >> 32) & 0x7fffffffu;
#endif
}

template <int NR, int NCOL, int G>
struct GroupGJ {
    static constexpr int RPL = (NR + G - 1) / G;      // rows (slots) per lane
    static constexpr int LG = Grp<G>::LG;
    static constexpr int CINV = NCOL + (NCOL % 2), CSOL = CINV + 1;   // OD_EXTRACT_SMEM: spare mirror columns (inverse pivot, scaled solution)
    static_assert(RPL * G <= 32, "row index must fit the 5 low key bits");
    static_assert(NCOL >= NR, "augmented matrix");

    // a[s][j] for a run-time slot s and a compile-time column j (registers cannot be indexed dynamically: select chain)
    OD_HD static double pick(const double (&a)[RPL][NCOL], const int j, const int s) {
        double v = a[0][j];
#pragma unroll
        for (int t = 1; t < RPL; ++t) v = (s == t) ? a[t][j] : v;
        return v;
    }
    OD_HD static double pickv(const double (&x)[RPL], const int s) {
        double v = x[0];
#pragma unroll
        for (int t = 1; t < RPL; ++t) v = (s == t) ? x[t] : v;
        return v;
    }
    // entry r = s·G + g of a vector that every lane holds in full (the lane's own rows of a replicated right-hand side)
    OD_HD static void mine(const double* full, double (&x)[RPL], const int g) {
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            double v = (s * G < NR) ? full[s * G] : 0.0;
#pragma unroll
            for (int t = 1; t < G; ++t) if (s * G + t < NR) v = (g == t) ? full[s * G + t] : v;
            if (G > 1 && (s + 1) * G > NR) v = (s * G + g < NR) ? v : 0.0;
            x[s] = v;
        }
    }

    // Gauss–Jordan with implicit partial pivoting on the lane-distributed rows; returns false on a zero / non-finite pivot.
    OD_HD static bool factor(double (&a)[RPL][NCOL], int (&piv)[NR], const int g, const unsigned gm) {
        bool ok = true;
        unsigned done = 0;                                       // slots of this lane that are padding or have been pivot rows
#pragma unroll
        for (int s = 0; s < RPL; ++s) if (s * G + g >= NR) done |= 1u << s;
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            unsigned key = 0;
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                const unsigned ks = (abs_hi32(a[s][k]) & ~31u) | (unsigned)(s * G + g);
                const unsigned kk = ((done >> s) & 1u) ? 0u : ks;
                key = kk > key ? kk : key;
            }
            // the reciprocal of this lane's best candidate is started before the group reduction: if the lane wins, it is the pivot's
            const double myinv = pivot_rcp(pick(a, k, (int)(key & 31u) >> LG));
            key = Grp<G>::umax_all(key, gm);
            ok = ok && (key >= 32u) && (key < 0x7ff00000u);
            const int pr = (int)(key & 31u), wl = pr & (G - 1), ws = pr >> LG;
            piv[k] = pr;
            const double inv = Grp<G>::bcast(myinv, wl, gm);
            const bool own = (g == wl);
            if (own) done |= 1u << ws;
            double prow[NCOL];
#pragma unroll
            for (int j = k + 1; j < NCOL; ++j) prow[j] = Grp<G>::bcast(pick(a, j, ws), wl, gm);
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                const bool isp = own && (s == ws);
                const double m = a[s][k] * inv;
                a[s][k] = isp ? inv : m;
                const double me = isp ? 0.0 : m;
#pragma unroll
                for (int j = k + 1; j < NCOL; ++j) a[s][j] -= me * prow[j];
            }
        }
        return ok;
    }

    // factor() with the pivot row passed through shared memory instead of shuffles.  `S` is the staging area the rows were fetched
    // from (row r at S + r·PITCH, PITCH even, 16-byte aligned, at least NCOL + NCOL%2 columns): on entry it still holds the matrix,
    // so step 0 reads its pivot row straight from it; before every later step each lane re-publishes the columns ≥ k+1 of its own
    // rows (16-byte stores, issued while the pivot search of the step is in flight) and the pivot row comes back as 16-byte
    // broadcast loads.  Per matrix element that is 1.5 instructions (2 STS.128 + 1 LDS.128 per pair of columns) instead of the
    // 4 of the register path (a 2-way select + 2 SHFL per double), and one scoreboard wait per pair instead of per word:
    // 12×13 system: 129 memory instructions instead of 312 shuffles + selects.
    // Only rows that can still become a pivot row are published (not the padding rows, not former pivot rows): a lane only ever
    // writes its own rows, never the row of the previous step that another lane may still be reading, so one warp barrier per
    // step — between publishing and reading — is enough.  (Measured: publishing every row costs 7 % at 262 144 problems.)
    // Same arithmetic, in the same order, as factor().
    template <int PITCH>
    OD_HD static bool factor_sm(double (&a)[RPL][NCOL], int (&piv)[NR], const int g, const unsigned gm, double* S) {
        static_assert(PITCH % 2 == 0 && PITCH >= NCOL + (NCOL % 2), "pairs of columns are moved as 16-byte words");
        bool ok = true;
        unsigned done = 0;
#pragma unroll
        for (int s = 0; s < RPL; ++s) if (s * G + g >= NR) done |= 1u << s;
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            const int j0 = (k + 1) & ~1;                       // first column of the first 16-byte pair that holds a column > k
            if (k > 0) {
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    const int r = s * G + g;
                    if (!((done >> s) & 1u)) {                 // padding rows and former pivot rows are never read again
                        double2* dst = reinterpret_cast<double2*>(S + r * PITCH + j0);
#pragma unroll
                        for (int j = j0; j < NCOL; j += 2) dst[(j - j0) / 2] = make_double2(a[s][j], (j + 1 < NCOL) ? a[s][j + 1] : 0.0);
                    }
                }
            }
            unsigned key = 0;
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                const unsigned ks = (abs_hi32(a[s][k]) & ~31u) | (unsigned)(s * G + g);
                const unsigned kk = ((done >> s) & 1u) ? 0u : ks;
                key = kk > key ? kk : key;
            }
            const double myinv = pivot_rcp(pick(a, k, (int)(key & 31u) >> LG));
            key = Grp<G>::umax_all(key, gm);
            ok = ok && (key >= 32u) && (key < 0x7ff00000u);
            const int pr = (int)(key & 31u), wl = pr & (G - 1), ws = pr >> LG;
            piv[k] = pr;
            const double inv = Grp<G>::bcast(myinv, wl, gm);
            const bool own = (g == wl);
            if (own) done |= 1u << ws;
#ifdef __CUDA_ARCH__
            if (G > 1 && k > 0) __syncwarp(gm);               // (the reduction above is not a memory barrier)
#else
            if (G > 1 && k > 0) host_team_sync();
#endif
            double prow[NCOL + 1];
            {
                const double2* src = reinterpret_cast<const double2*>(S + pr * PITCH + j0);
#pragma unroll
                for (int j = j0; j < NCOL; j += 2) { const double2 v = src[(j - j0) / 2]; prow[j] = v.x; prow[j + 1] = v.y; }
            }
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                const bool isp = own && (s == ws);
                const double m = a[s][k] * inv;
                a[s][k] = isp ? inv : m;
                const double me = isp ? 0.0 : m;
#pragma unroll
                for (int j = k + 1; j < NCOL; ++j) a[s][j] -= me * prow[j];
            }
        }
        return ok;
    }

    // Prepared variant of factor_sm() (-DOD_EXTRACT_SMEM=1; needs two spare mirror columns, PITCH ≥ CINV + 2):
    //   * the inverse pivot of step k goes to S[pivot row][CINV] and column k of the pivot row keeps a ZERO multiplier, so that
    //     solve_sm() needs no pivot-row predicate and extract_sm() / solve_sm() scale and gather through the mirror;
    //   * an odd last column is moved as 8 bytes (no zero pad to materialise);
    //   * a lane recognises its pivot row by comparing the winning row index with its own row numbers, and keeps one "still a
    //     candidate" flag per slot — no lane / slot decoding of the winner beyond the source lane of the reciprocal.
    // Same arithmetic, in the same order, as factor_sm().
    template <int PITCH>
    OD_HD static bool factor_v2(double (&a)[RPL][NCOL], int (&piv)[NR], const int g, const unsigned gm, double* S) {
        static_assert(PITCH % 2 == 0 && PITCH >= CINV + 2, "two spare columns behind the (padded) matrix");
        bool ok = true;
        bool cand[RPL];                                            // slot can still become a pivot row (not padding, not used yet)
#pragma unroll
        for (int s = 0; s < RPL; ++s) cand[s] = (s * G + g < NR);
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            const int j0 = (k + 1) & ~1;
            if (k > 0) {
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    if (cand[s]) {
                        double* row = S + (s * G + g) * PITCH;
#pragma unroll
                        for (int j = j0; j < NCOL; j += 2) {
                            if (j + 1 >= NCOL) row[j] = a[s][j];
                            else *reinterpret_cast<double2*>(row + j) = make_double2(a[s][j], a[s][j + 1]);
                        }
                    }
                }
            }
            unsigned key = 0;
            double best = a[0][k];                                 // this lane's best candidate (value of the slot that carries `key`)
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                const unsigned ks = cand[s] ? ((abs_hi32(a[s][k]) & ~31u) | (unsigned)(s * G + g)) : 0u;
                if (s > 0) best = (ks > key) ? a[s][k] : best;
                key = ks > key ? ks : key;
            }
            const double myinv = pivot_rcp(best);
            key = Grp<G>::umax_all(key, gm);
            ok = ok && (key >= 32u) && (key < 0x7ff00000u);
            const int pr = (int)(key & 31u);
            piv[k] = pr;
            const double inv = Grp<G>::bcast(myinv, pr & (G - 1), gm);
#ifdef __CUDA_ARCH__
            if (G > 1 && k > 0) __syncwarp(gm);
#else
            if (G > 1 && k > 0) host_team_sync();
#endif
            const double* prow_s = S + pr * PITCH;
            S[pr * PITCH + CINV] = inv;                            // every lane stores the same value: its owner reads it back without a barrier
            double prow[NCOL + 1];
#pragma unroll
            for (int j = j0; j < NCOL; j += 2) {
                if (j + 1 >= NCOL) { prow[j] = prow_s[j]; prow[j + 1] = 0.0; }
                else { const double2 v = *reinterpret_cast<const double2*>(prow_s + j); prow[j] = v.x; prow[j + 1] = v.y; }
            }
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                const bool isp = (pr == s * G + g);
                cand[s] = cand[s] && !isp;
                const double m = a[s][k] * inv;
                const double me = isp ? 0.0 : m;
                a[s][k] = me;
#pragma unroll
                for (int j = k + 1; j < NCOL; ++j) a[s][j] -= me * prow[j];
            }
        }
        return ok;
    }

    // Solution of the carried right-hand side in column NR + c, replicated in every lane.
    OD_HD static void extract(const double (&a)[RPL][NCOL], const int (&piv)[NR], const int c, double* sol, const unsigned gm) {
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            const int wl = piv[k] & (G - 1), ws = piv[k] >> LG;
            sol[k] = Grp<G>::bcast(pick(a, NR + c, ws) * pick(a, k, ws), wl, gm);
        }
    }

    // OD_EXTRACT_SMEM variants.  Row r's inverse pivot sits in S[r·PITCH + CINV] (factor_v2); every lane scales its own rows,
    // publishes them in column CSOL and reads unknown k from the row that was the pivot of step k.
    template <int PITCH>
    OD_HD static void gather_sm(const double (&y)[RPL], const int (&piv)[NR], double* sol, const int g, const unsigned gm, double* S) {
#ifdef __CUDA_ARCH__
        if (G > 1) __syncwarp(gm);                                // earlier readers of column CSOL are done
#else
        if (G > 1) host_team_sync();
#endif
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            const int r = s * G + g;
            if (G == 1 || (s + 1) * G <= NR || r < NR) S[r * PITCH + CSOL] = y[s] * S[r * PITCH + CINV];
        }
#ifdef __CUDA_ARCH__
        if (G > 1) __syncwarp(gm);
#else
        if (G > 1) host_team_sync();
#endif
#pragma unroll
        for (int k = 0; k < NR; ++k) sol[k] = S[piv[k] * PITCH + CSOL];
    }
    template <int PITCH>
    OD_HD static void extract_sm(const double (&a)[RPL][NCOL], const int (&piv)[NR], const int c, double* sol, const int g, const unsigned gm, double* S) {
        double y[RPL];
#pragma unroll
        for (int s = 0; s < RPL; ++s) y[s] = a[s][NR + c];
        gather_sm<PITCH>(y, piv, sol, g, gm, S);
    }
    template <int PITCH>
    OD_HD static void solve_sm(const double (&a)[RPL][NCOL], const int (&piv)[NR], double (&x)[RPL], double* sol, const int g, const unsigned gm, double* S) {
        static_assert(OD_EXTRACT_SMEM && PITCH >= CINV + 2, "needs the factorisation of factor_v2 (zero multiplier in pivot rows, inverse pivots in the mirror)");
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            const int wl = piv[k] & (G - 1), ws = piv[k] >> LG;
            const double xp = Grp<G>::bcast(pickv(x, ws), wl, gm);
#pragma unroll
            for (int s = 0; s < RPL; ++s) x[s] -= a[s][k] * xp;        // the pivot row of step k holds a zero in column k
        }
        gather_sm<PITCH>(x, piv, sol, g, gm, S);
    }

    // A further right-hand side (x = this lane's rows of it) through the stored elimination; solution replicated in every lane.
    OD_HD static void solve(const double (&a)[RPL][NCOL], const int (&piv)[NR], double (&x)[RPL], double* sol, const int g, const unsigned gm) {
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            const int wl = piv[k] & (G - 1), ws = piv[k] >> LG;
            const double xp = Grp<G>::bcast(pickv(x, ws), wl, gm);
            const bool own = (g == wl);
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                const bool isp = own && (s == ws);
                const double me = isp ? 0.0 : a[s][k];
                x[s] -= me * xp;
            }
        }
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            const int wl = piv[k] & (G - 1), ws = piv[k] >> LG;
            sol[k] = Grp<G>::bcast(pickv(x, ws) * pick(a, k, ws), wl, gm);
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// The same elimination with the matrix RESIDENT in shared memory and every loop over elimination steps rolled.
//
// Why: the register-resident GroupGJ needs compile-time column indices, so its NR steps are NR copies of the code — 1.9 k
// instructions for the planar push's 20×21 system, 0.9 k for the hopper's 12×13 — and the interior-point loop around it no longer
// fits the 32 KB instruction cache: ncu shows 6.0 (planar push), 3.3 (rocket) and 0.8 (hopper) stall cycles per issued instruction
// waiting for instructions.  Here row r still belongs to lane r mod G (only its owner writes it; everybody reads the pivot row), the
// step loop is one copy of the code with run-time addresses, pairs of columns move as 16-byte words, and the per-lane register
// image of the matrix (RPL·NCOL doubles) is gone.  Same arithmetic in the same order as GroupGJ::factor_v2 / solve_sm / extract_sm
// (bit-identical results; tests/test_host_logic.py holds the two against each other).
// S: row r at S + r·PITCH (PITCH even, ≥ CINV + 2: inverse pivots in column CINV, scaled solution in CSOL); PV: NR ints in shared memory.
template <int NR, int NCOL, int G, int PITCH>
struct GroupGJS {
    static constexpr int RPL = (NR + G - 1) / G;
    static constexpr int LG = Grp<G>::LG;
    static constexpr int CINV = NCOL + (NCOL % 2), CSOL = CINV + 1;
    static_assert(RPL * G <= 32 && NCOL >= NR && PITCH % 2 == 0 && PITCH >= CINV + 2, "layout");
    OD_HD static void sync(const unsigned gm) {
#ifdef __CUDA_ARCH__
        if (G > 1) __syncwarp(gm);
#else
        if (G > 1) host_team_sync();
#endif
    }
    OD_HD static bool factor(double* S, int* PV, const int g, const unsigned gm) {
        bool ok = true;
        bool cand[RPL];
#pragma unroll
        for (int s = 0; s < RPL; ++s) cand[s] = (s * G + g < NR);
#pragma unroll 1
        for (int k = 0; k < NR; ++k) {
            if (k > 0) sync(gm);                                   // the rows as updated by step k − 1
            unsigned key = 0;
            double best = (G <= NR || g < NR) ? S[g * PITCH + k] : 0.0;
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                const int r = s * G + g;
                const bool in = (G == 1) || ((s + 1) * G <= NR) || (r < NR);
                const double v = in ? S[r * PITCH + k] : 0.0;
                const unsigned ks = cand[s] ? ((abs_hi32(v) & ~31u) | (unsigned)r) : 0u;
                if (s > 0) best = (ks > key) ? v : best;
                key = ks > key ? ks : key;
            }
            const double myinv = pivot_rcp(best);
            key = Grp<G>::umax_all(key, gm);
            ok = ok && (key >= 32u) && (key < 0x7ff00000u);
            const int pr = (int)(key & 31u);
            PV[k] = pr;                                            // every lane stores the same value
            const double inv = Grp<G>::bcast(myinv, pr & (G - 1), gm);
            const double* prow = S + pr * PITCH;
            S[pr * PITCH + CINV] = inv;
            double me[RPL]; bool upd[RPL];
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                const int r = s * G + g;
                const bool in = (G == 1) || ((s + 1) * G <= NR) || (r < NR);
                const bool isp = (pr == r);
                cand[s] = cand[s] && !isp;
                const double m = (in ? S[r * PITCH + k] : 0.0) * inv;
                me[s] = isp ? 0.0 : m;
                upd[s] = in && !isp;                               // the pivot row itself stays as it is (a − 0·a)
                if (in) S[r * PITCH + k] = me[s];
            }
            int j = k + 1;
            if (j & 1) {
                if (j < NCOL) {
                    const double pv = prow[j];
#pragma unroll
                    for (int s = 0; s < RPL; ++s) if (upd[s]) S[(s * G + g) * PITCH + j] -= me[s] * pv;
                }
                ++j;
            }
#pragma unroll 2
            for (; j < NCOL; j += 2) {                             // (an odd NCOL drags its pad column along: never read back)
                const double2 pv = *reinterpret_cast<const double2*>(prow + j);
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    if (upd[s]) {
                        double2* p = reinterpret_cast<double2*>(S + (s * G + g) * PITCH + j);
                        double2 o = *p;
                        o.x -= me[s] * pv.x; o.y -= me[s] * pv.y;
                        *p = o;
                    }
                }
            }
        }
        sync(gm);
        return ok;
    }
    // solution of the carried right-hand side in column NR + c, replicated in every lane (sol[k]: compile-time indices)
    OD_HD static void extract(const double* S, const int* PV, const int c, double* sol) {
#pragma unroll
        for (int k = 0; k < NR; ++k) { const double* row = S + PV[k] * PITCH; sol[k] = row[NR + c] * row[CINV]; }
    }
    // a further right-hand side (x = this lane's rows of it) through the stored multipliers; solution replicated in every lane
    OD_HD static void solve(double* S, const int* PV, double (&x)[RPL], double* sol, const int g, const unsigned gm) {
#pragma unroll 1
        for (int k = 0; k < NR; ++k) {
            const int pr = PV[k], wl = pr & (G - 1), ws = pr >> LG;
            double xs = x[0];
#pragma unroll
            for (int t = 1; t < RPL; ++t) xs = (ws == t) ? x[t] : xs;
            const double xp = Grp<G>::bcast(xs, wl, gm);
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                const int r = s * G + g;
                const bool in = (G == 1) || ((s + 1) * G <= NR) || (r < NR);
                x[s] -= (in ? S[r * PITCH + k] : 0.0) * xp;        // the pivot row of step k holds a zero in column k
            }
        }
        sync(gm);                                                  // earlier readers of column CSOL are done
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            const int r = s * G + g;
            if (G == 1 || (s + 1) * G <= NR || r < NR) S[r * PITCH + CSOL] = x[s] * S[r * PITCH + CINV];
        }
        sync(gm);
#pragma unroll
        for (int k = 0; k < NR; ++k) sol[k] = S[PV[k] * PITCH + CSOL];
    }
};

}  // namespace od
