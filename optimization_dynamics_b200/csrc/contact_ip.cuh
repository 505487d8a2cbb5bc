// Condensed primal-dual interior-point step + IFT sensitivities for the contact-implicit models, one problem per thread,
// everything in registers.
//
// Replaces, for one (q1, q2, u) sample, the reference call chain
//     f / fx / fu            reference src/dynamics.jl:81-128
//       └ RoboDojo.step!  →  interior_point_solve!  →  differentiate (δz = −rz⁻¹ rθ)        [RoboDojo.jl, external]
// Algorithm = SURVEY.md Appendix A.3 / DESIGN.md §Algorithm (Mehrotra predictor–corrector, residual line search).
//
// Layout of one problem (RoboDojo IndicesZ: q, γ, sγ, ψ, b, sψ, sb — reference src/models/planar_push/simulator.jl:1-14):
//     z = [ q(NQ) | γ(NC) s(NC) | ψ(NP) b(NB) | sψ(NP) sb(NB) ]       NC orthant pairs (γ_i, s_i),
//                                                                     NP second-order cones (ψ_k; b_k) ∘ (sψ_k; sb_k)
//     rows: d(q,γ,b;θ)=0 | s − ϕ(q)=0 | ψ − ψ̂(γ;θ)=0 | vT(q;θ) − sb=0 | γ∘s = κ | (ψ;b)∘(sψ;sb) = (κ;0)
// The reference factors the dense nz×nz Jacobian with LU (`lu_solver`, reference src/gradient_bundle.jl:76).  Here the
// variables that enter their defining row with a unit coefficient — Δs = rs + NΔq, Δψ = rpsi + MψΔγ, Δsb = VΔq − rv — are
// substituted out exactly (no division, so no loss of accuracy however close the iterate is to the cone boundary), and the
// remaining NR = NQ+NC+NB+NP unknowns x = (Δq, Δγ, Δb, Δsψ) are solved with a partial-pivoting LU (hopper: 12×12 instead of
// 20×20 ⇒ ≈4.6× fewer flops).  Further condensation to NQ×NQ (dividing by s or sψ) is NOT used: with undercut = Inf the
// reference drives γ∘s to ~1e-26, where the normal-equation form D + Nᵀ(Γ/S)N loses all accuracy (measured: 22 % of the hopper
// batch diverged from the oracle).  The LU lives in a per-thread workspace (shared memory on the GPU), element e of lane l at
// ws[e·stride + l] — bank-conflict free.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "group_gj.cuh"

// 1: the pivot row of every elimination step travels through the shared-memory staging area (GroupGJ::factor_sm), 0: by shuffles.
#ifndef OD_GJ_SMEM
#define OD_GJ_SMEM 1
#endif
// 1 (experiment, off): the block's input rows are staged by one cp.async.bulk (TMA) copy, see contact_step_kernel
#ifndef OD_TMA_INPUT
#define OD_TMA_INPUT 0
#endif
// Reduced systems of this size and larger are factorised in shared memory with rolled loops (ContactIP::LASM, GroupGJS).  OFF in the
// shipped build (99): measured on B200 (profiles/r02h_*), the rolled shared-memory elimination removes the instruction-cache stalls
// and the spills but pays for it in shared-memory latency on the dependent LDS → FMA → STS chain (short-scoreboard stalls 1.0 → 2.5
// per issue, half the wavefronts bank-conflicted): hopper 4096 0.064 → 0.107 ms, planar push 25 600 6.5 → 10.3 ms.
// OD_SPLIT_SOC (default 1): the 3-D second-order-cone step lengths of a group are evaluated one per lane and combined by a shuffle
// maximum instead of all of them in every lane (see ContactIP::step_length); 0 restores the replicated evaluation (A/B).
// OD_ZSMEM_MIN_NZ (default 30: planar push only): see ContactIP::ZSM.  99 = off (A/B).
#ifndef OD_ZSMEM_MIN_NZ
#define OD_ZSMEM_MIN_NZ 30
#endif
// OD_PITCH_G4 (default 1): row pitch rule of the 4-lane configurations (see ContactIP::PW); 0 = the PW/2-odd rule for every G (A/B).
#ifndef OD_PITCH_G4
#define OD_PITCH_G4 1
#endif
#ifndef OD_SPLIT_SOC
#define OD_SPLIT_SOC 1
#endif
#ifndef OD_LA_SMEM_MIN_NR
#define OD_LA_SMEM_MIN_NR 99
#endif
// 1 (prepared, off): iterate advanced in place, see contact_step_one
#ifndef OD_INPLACE_Z
#define OD_INPLACE_Z 0
#endif

// Every solver routine is __host__ __device__ so that tests/host_check.cu can single-step the very same template code on the
// CPU against the oracle (a debugging aid for a GPU-less build container; liboptdyn_b200.so exports no host compute path).
#ifndef OD_HD
#define OD_HD __host__ __device__ __forceinline__
#endif

namespace od {

OD_HD double rsqrt_d(double x) { return 1.0 / sqrt(x); }

struct SolverOpts {           // RoboDojo InteriorPointOptions as set at reference src/dynamics.jl:25-33
    double r_tol;             // ∞-norm tolerance on the equality rows
    double kappa_eval_tol;    // ∞-norm tolerance on the bilinear rows, eval simulator (f)
    double kappa_grad_tol;    // same, gradient simulator (fx, fu)
    double ls_scale;          // 0.5
    int max_iter;             // 100
    int max_ls;               // 25
};

// status nibble: 0 converged, 1 iteration cap, 2 non-finite iterate or singular system
enum : int { ST_OK = 0, ST_MAXIT = 1, ST_FAIL = 2, ST_PEND = 3 };   // ST_PEND: device-internal (parked problem of the persistent sweep), never returned

// Step lengths are tracked as fractions num/den (den > 0) and compared by cross-multiplication, so that a whole step-length
// computation costs one division instead of one per cone variable (fp64 division is ~20 instructions on the GPU).
OD_HD void frac_min(double& bn, double& bd, double n, double d) { if (n * bd < bn * d) { bn = n; bd = d; } }

// CVXOPT §8.2 step to the boundary of the second-order cone for λ − αΔ, λ=(l0; l1..): candidate α = τ / (‖ρ_v‖ − ρ_s) when that
// denominator is positive (same formula, including the 1e-25 clamp, as the oracle's soc_step; reciprocals are shared).
template <int DIM>
OD_HD void soc_step(double l0, const double* l1, double d0, const double* d1, double tau, double& bn, double& bd) {
    double ll = l0 * l0, lD = l0 * (-d0);
#pragma unroll
    for (int i = 0; i < DIM; ++i) { ll -= l1[i] * l1[i]; lD -= l1[i] * (-d1[i]); }
    ll = fmax(ll, 1e-25);
    const double inv_sq = od_rsqrt(ll);                       // short reciprocal square root / reciprocal / square root (fastmath.cuh)
    const double inv_ll = inv_sq * inv_sq;
    const double rho_s = lD * inv_ll;
    const double coef = (lD * inv_sq + (-d0)) * pivot_rcp(l0 * inv_sq + 1.0);
    double nv = 0.0;
    if (DIM == 1) {
        nv = fabs(((-d1[0]) - coef * l1[0] * inv_sq) * inv_sq);
    } else {
#pragma unroll
        for (int i = 0; i < DIM; ++i) { const double rv = ((-d1[i]) - coef * l1[i] * inv_sq) * inv_sq; nv += rv * rv; }
        nv = od_sqrt(nv);
    }
    const double den = nv - rho_s;
    if (den > 0.0) frac_min(bn, bd, tau, den);
}

// Two-dimensional second-order cone (ψ; b), b scalar — every friction cone of the hopper and the cartpole: ψ ≥ |b| is the pair of
// half-planes ψ − b ≥ 0, ψ + b ≥ 0, so the step to the boundary is the smaller of two orthant-type ratios.  In exact arithmetic
// this equals the CVXOPT expression above (the eigenvalues of the scaled direction are the two ratios); it needs no division
// and no square root.
OD_HD void soc2_step(double l0, double l1, double d0, double d1, double tau, double& bn, double& bd) {
    const double dm = d0 - d1, dp = d0 + d1;
    if (dm > 0.0) frac_min(bn, bd, tau * (l0 - l1), dm);
    if (dp > 0.0) frac_min(bn, bd, tau * (l0 + l1), dp);
}

template <int N> struct cmax1 { static constexpr int v = N > 0 ? N : 1; };

// G lanes cooperate on one problem (G = 1: one thread per problem — the throughput configuration; G = 4/8: the latency
// configuration for small batches, where 4096 problems would otherwise occupy one warp per SM).  All lanes of a group hold
// identical replicated state (z, r, direction) and evaluate the model redundantly; only the O(n³) factorisation is split by rows.
// PPB = problems per block; the per-problem workspace is lane-interleaved in shared memory: element e of problem p at ws[e·PPB + p].
//
// REG = true selects the register-resident linear algebra (group_gj.cuh): the reduced matrix is assembled through a small
// shared-memory staging area (written redundantly by the lanes, read back one row set per lane), then factorised, and both Newton
// right-hand sides solved, entirely in registers with warp shuffles.  Its workspace is contiguous per problem (ws[e]).
template <class M, int G = 1, int PPB = 1, bool REG = false>
struct ContactIP {
    static constexpr int NQ = M::NQ, NU = M::NU, NC = M::NC, NP = M::NP, NB = M::NB, NTH = M::NTH;
    static constexpr int NC1 = cmax1<NC>::v, NP1 = cmax1<NP>::v, NB1 = cmax1<NB>::v;
    static constexpr int NTC1 = cmax1<M::NTC>::v, NTV1 = cmax1<M::NTV>::v;   // sin/cos tables (θ-only / q-dependent arguments)
    static constexpr int NTP = 2 * NQ + NU;              // θ' = (q1, q2, u): the sensitivity columns that are returned
    static constexpr int NCONE = NC + NP;                // cone degree (orthant pairs + second-order cones)
    static constexpr int NR = NQ + NC + NB + NP;         // reduced system size
    static constexpr int NRP = (NR % 2 == 0) ? NR + 1 : NR;   // odd row pitch: the G row-owners of a step hit disjoint banks
    static constexpr int OFF_X = NR * NRP;               // NTP right-hand-side / solution vectors, pitch NRP (odd: lane-private vectors on disjoint banks)
    static constexpr int OFF_PIV = OFF_X + NTP * NRP;    // row interchanges of the LU (LAPACK convention), stored as doubles
    static constexpr int OFF_CP = OFF_PIV + NR;          // column permutation of the rank-revealing factorisation (robust IFT only)
    static constexpr int NZ = NQ + 2 * NC + 2 * NP + 2 * NB;
    static constexpr int OFF_ZS = OFF_CP + (M::ROBUST_IFT ? NR : 0);   // snapshot of the iterate at which the IFT is taken
    // REG workspace (contiguous per problem): staging rows [NR][PW] (K | carried right-hand sides; reused for the packed output
    // row), iterate snapshot, q3.  Even pitch and offsets: rows are moved as 16-byte pairs; WS/2 odd spreads problems over banks.
    // Models with the rank-revealing IFT run it (once per problem) in the shared-memory-LU layout at the start of the same
    // workspace, so the snapshot / q3 slots sit behind both areas.
    // Row pitch: even (16-byte pairs) with PW/2 odd — the lanes of a group move their own rows r, r+1, … as 16-byte accesses, and a
    // pitch of 4 (mod 8) words puts 8 consecutive rows on disjoint banks.  (Planar push had 32 doubles = 256 bytes: every row on the
    // same banks, 59 % of its shared-memory wavefronts were bank-conflict replays, profiles/r02x_planar_push_sweep_and_resume.json.)
    // With 4 lanes per problem a quarter-warp (one 16-byte wavefront) holds TWO groups: rows 24 (mod 32) words apart take the even
    // 4-bank units and the neighbouring problem (workspaces 4 mod 8 words apart) the odd ones — PW ≡ 12 (mod 16): hopper 28 (22 gave
    // a 2-way conflict on every row move: 1.6 M of the hopper kernel's 5.9 M shared-memory wavefronts), cartpole 12, acrobot 12.
    static constexpr int PW0 = ((NR + NTP + 1) / 2) * 2;
    static constexpr int PW = (G <= 4 && OD_PITCH_G4) ? PW0 + ((12 - PW0 % 16) + 16) % 16 : ((PW0 / 2) % 2 == 1) ? PW0 : PW0 + 2;
    static constexpr int ROBUST_END = M::ROBUST_IFT ? OFF_CP + NR : 0;
    static constexpr int ROFF_ZS = ((((NR * PW > ROBUST_END) ? NR * PW : ROBUST_END) + 1) / 2) * 2;
    static constexpr int ROFF_Q3 = ROFF_ZS + NZ;
    // LASM (A/B switch, off by default — see OD_LA_SMEM_MIN_NR): the Newton systems are factorised in place in the staging area
    // with rolled loops (GroupGJS) instead of in registers; -DOD_LA_SMEM_MIN_NR=1 builds every model that way.
    static constexpr bool LASM = REG && (NR >= OD_LA_SMEM_MIN_NR) && (PW >= NR + 1 + ((NR + 1) % 2) + 2);
    static constexpr int ROFF_PV = ((ROFF_Q3 + NQ + 1) / 2) * 2;          // pivot rows of the elimination steps (ints)
    // ZSM (models with NZ ≥ OD_ZSMEM_MIN_NZ: planar push): the iterate objects of the solver loop — iterate, direction, candidate and
    // the candidate's residual, identical in all lanes of a group — live ONCE per group in the workspace instead of G times in
    // registers / local memory (every lane stores the same value with the same warp instruction; reads are broadcasts).  Measured on
    // B200 (profiles/r02w_*): planar push sweep 1696 → 800 bytes of stack per thread, 25 600 problems 5.52 → 5.02 ms.
    static constexpr bool ZSM = REG && (NZ >= OD_ZSMEM_MIN_NZ);
    static constexpr int ROFF_ZX = ((ROFF_PV + (LASM ? ((NR + 3) / 4) * 2 : 0) + 1) / 2) * 2;
    static constexpr int RWS0 = ROFF_ZX + (ZSM ? 4 * NZ + (4 * NZ) % 2 : 0);
    static constexpr int RWS = ((RWS0 / 2) % 2 == 1) ? RWS0 : RWS0 + 2;
    static constexpr int NOUT = NQ + NQ * NTP;           // packed output row [q3 | ∂q3/∂q1 | ∂q3/∂q2 | ∂q3/∂u1]
    static_assert(!REG || NOUT <= NR * PW, "output row is staged in the matrix area");
    static constexpr int WS = REG ? RWS : OFF_ZS + NZ;   // workspace doubles per problem
    static constexpr int WS_STRIDE = REG ? 1 : PPB;      // distance between consecutive elements of one problem
    static constexpr int WS_SLOT = REG ? RWS : 1;        // distance between the workspaces of consecutive problems of a block
    typedef GroupGJ<NR, NR + 1, G> GJ;                   // Newton systems: one carried right-hand side (the affine one)
    typedef GroupGJS<NR, NR + 1, G, PW> GJSM;            // the same, matrix resident in shared memory (LASM)
    typedef GroupGJ<NR, NR + NTP, G> GJS;                // sensitivity system: all NTP right-hand sides carried
    static constexpr int RPL = GJ::RPL;

    struct Z { double q[NQ], gam[NC1], s[NC1], psi[NP1], b[NB1], spsi[NP1], sb[NB1]; };
    // ZSM: slot `k` (0..3, NZ doubles each) of the group's shared area as an object of type T; otherwise — and always on the host,
    // where the emulated lanes of tests/host_check.cu are free-running threads between collectives — the caller's own object
    template <class T>
    OD_HD static T& shared_obj(double* ws, int k, T& local) {
#ifdef __CUDA_ARCH__
        if constexpr (ZSM) { static_assert(sizeof(T) <= NZ * sizeof(double), "slot size"); return *reinterpret_cast<T*>(ws + ROFF_ZX + k * NZ); }
#endif
        return local;
    }
    // residual in block form; bilinear rows are stored at κ = 0 (r(z;κ) only shifts rgam and rc0 by −κ)
    struct R { double d[NQ], rs[NC1], rpsi[NP1], rv[NB1], rgam[NC1], rc0[NP1], rc1[NB1]; };
    struct Lin {
        double N[NC1 * NQ], V[NB1 * NQ], Mpsi[NP1 * NC1];
        double* ws;                  // workspace of this problem (already offset by the problem's slot)
        int g;                       // lane within the group
        unsigned gmask;              // shuffle / __syncwarp mask (the whole warp: execution is warp-synchronous)
        bool ok;
        double a[(REG && !LASM) ? RPL : 1][(REG && !LASM) ? NR + 1 : 1];   // REG: this lane's rows of the eliminated [K | affine rhs]
        int piv[(REG && !LASM) ? NR : 1];                                  // REG: pivot row of every elimination step
        OD_HD int* PV() const { return reinterpret_cast<int*>(ws + ROFF_PV); }   // LASM: the same, in the workspace
        OD_HD double& S(int r, int j) const { return ws[r * PW + j]; }     // REG staging area
        OD_HD double& K(int i, int j) const { return ws[(i * NRP + j) * WS_STRIDE]; }
        OD_HD double& X(int v, int i) const { return ws[(OFF_X + v * NRP + i) * WS_STRIDE]; }
        OD_HD double& PIV(int i) const { return ws[(OFF_PIV + i) * WS_STRIDE]; }
        OD_HD double& CP(int i) const { return ws[(OFF_CP + i) * WS_STRIDE]; }
        OD_HD void sync() const {
#ifdef __CUDA_ARCH__
            if (G > 1) __syncwarp(gmask);
#else
            if (G > 1) host_team_sync();
#endif
        }
    };
    __host__ __device__ static constexpr int cone_of(int j) {   // friction cone that tangential component j belongs to
        int k = 0;
        for (int c = 0; c < NP; ++c) if (j >= M::cone_off(c) && j < M::cone_off(c) + M::cone_dim(c)) k = c;
        return k;
    }

    // ---- residual ------------------------------------------------------------------------------------------------
    OD_HD static void residual(const Z& z, const double* th, const double* trc, const double* trv, R& r, double& r_vio, double& k_vio) {
        double phi[NC1], psit[NP1], vT[NB1];
        M::eq(z.q, z.gam, z.b, th, trc, trv, r.d, phi, psit, vT);
        MaxAcc rv, kv;                                         // ∞-norms of the equality rows / the bilinear rows
#pragma unroll
        for (int i = 0; i < NQ; ++i) rv.add(od_abs(r.d[i]));
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            r.rs[i] = z.s[i] - phi[i]; rv.add(od_abs(r.rs[i]));
            r.rgam[i] = z.gam[i] * z.s[i]; kv.add(od_abs(r.rgam[i]));
        }
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            r.rpsi[k] = z.psi[k] - psit[k]; rv.add(od_abs(r.rpsi[k]));
            double acc = z.psi[k] * z.spsi[k];
#pragma unroll
            for (int j = M::cone_off(k); j < M::cone_off(k) + M::cone_dim(k); ++j) {
                acc += z.b[j] * z.sb[j];
                r.rc1[j] = z.psi[k] * z.sb[j] + z.spsi[k] * z.b[j]; kv.add(od_abs(r.rc1[j]));
            }
            r.rc0[k] = acc; kv.add(od_abs(acc));
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) { r.rv[j] = vT[j] - z.sb[j]; rv.add(od_abs(r.rv[j])); }
        r_vio = rv.v; k_vio = kv.v;
    }

    // ---- linearise at z: model blocks → reduced matrix K → LU with partial pivoting (in the workspace) -----------------
    //   unknowns x = [Δq | Δγ | Δb | Δsψ];  rows:
    //   d_i  : D Δq + Eγ Δγ + Eb Δb                                              = rd
    //   γ_i  : γ_i N_i Δq + s_i Δγ_i                                             = rgam_i − γ_i rs_i
    //   c0_k : (Σ_j b_j V_j) Δq + sψ_k Mψ_k Δγ + Σ_j sb_j Δb_j + ψ_k Δsψ_k        = rc0_k − sψ_k rpsi_k + Σ_j b_j rv_j
    //   c1_j : ψ_k V_j Δq + sb_j Mψ_k Δγ + sψ_k Δb_j + b_j Δsψ_k                  = rc1_j − sb_j rpsi_k + ψ_k rv_j
    // `r` is the residual at z: the REG path carries its reduced right-hand side (the affine Newton system) through the elimination.
    OD_HD static void linearize(const Z& z, const double* th, const double* trc, const double* trv, const R& r, Lin& L) {
        if constexpr (REG) {
            double x[NR];
            load_rhs(z, r, x);
            stage_matrix(z, th, trc, trv, L);
#pragma unroll
            for (int i = 0; i < NR; ++i) L.S(i, NR) = x[i];
            L.sync();
            if constexpr (LASM) {
                L.ok = GJSM::factor(&L.S(0, 0), L.PV(), L.g, L.gmask);
            } else {
                fetch_rows<NR + 1>(L, L.a);
#if OD_GJ_SMEM && OD_EXTRACT_SMEM
                if constexpr (PW >= GJ::CINV + 2) L.ok = GJ::template factor_v2<PW>(L.a, L.piv, L.g, L.gmask, &L.S(0, 0));
                else L.ok = GJ::template factor_sm<PW>(L.a, L.piv, L.g, L.gmask, &L.S(0, 0));
#elif OD_GJ_SMEM
                L.ok = GJ::template factor_sm<PW>(L.a, L.piv, L.g, L.gmask, &L.S(0, 0));
#else
                L.ok = GJ::factor(L.a, L.piv, L.g, L.gmask);
#endif
            }
        } else {
            assemble(z, th, trc, trv, L); factor(L);
        }
    }

    // Row i of the reduced matrix (i is a compile-time constant wherever this is called from an unrolled loop).
    OD_HD static void krow(const int i, const Z& z, const double* D, const double* Eg, const double* Eb, const Lin& L, double* row) {
        if (i < NQ) {
#pragma unroll
            for (int j = 0; j < NQ; ++j) row[j] = D[i * NQ + j];
#pragma unroll
            for (int j = 0; j < NC; ++j) row[NQ + j] = Eg[i * NC1 + j];
#pragma unroll
            for (int j = 0; j < NB; ++j) row[NQ + NC + j] = Eb[i * NB1 + j];
#pragma unroll
            for (int j = 0; j < NP; ++j) row[NQ + NC + NB + j] = 0.0;
        } else if (i < NQ + NC) {
            const int c = (i - NQ < NC1) ? i - NQ : 0;
#pragma unroll
            for (int j = 0; j < NQ; ++j) row[j] = z.gam[c] * L.N[c * NQ + j];
#pragma unroll
            for (int j = NQ; j < NR; ++j) row[j] = (j == i) ? z.s[c] : 0.0;
        } else if (i < NQ + NC + NP) {
            const int k = (i - NQ - NC < NP1) ? i - NQ - NC : 0;
#pragma unroll
            for (int c = 0; c < NQ; ++c) {
                double acc = 0.0;
#pragma unroll
                for (int e = 0; e < M::cone_dim(k); ++e) acc += z.b[M::cone_off(k) + e] * L.V[(M::cone_off(k) + e) * NQ + c];
                row[c] = acc;
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) row[NQ + c] = z.spsi[k] * L.Mpsi[k * NC1 + c];
#pragma unroll
            for (int j = 0; j < NB; ++j) row[NQ + NC + j] = (cone_of(j) == k) ? z.sb[j] : 0.0;
#pragma unroll
            for (int j = 0; j < NP; ++j) row[NQ + NC + NB + j] = (j == k) ? z.psi[k] : 0.0;
        } else {
            const int j = (i - NQ - NC - NP < NB1) ? i - NQ - NC - NP : 0;
            const int k = cone_of(j);
#pragma unroll
            for (int c = 0; c < NQ; ++c) row[c] = z.psi[k] * L.V[j * NQ + c];
#pragma unroll
            for (int c = 0; c < NC; ++c) row[NQ + c] = z.sb[j] * L.Mpsi[k * NC1 + c];
#pragma unroll
            for (int jj = 0; jj < NB; ++jj) row[NQ + NC + jj] = (jj == j) ? z.spsi[k] : 0.0;
#pragma unroll
            for (int kk = 0; kk < NP; ++kk) row[NQ + NC + NB + kk] = (kk == k) ? z.b[j] : 0.0;
        }
    }

    // REG: model blocks → K rows into the staging area.  Every lane writes the same values (16-byte stores); the caller adds the
    // right-hand-side columns and synchronises before fetch_rows.
    OD_HD static void stage_matrix(const Z& z, const double* th, const double* trc, const double* trv, Lin& L) {
        double D[NQ * NQ], Eg[NQ * NC1], Eb[NQ * NB1];
        M::jac(z.q, z.gam, z.b, th, trc, trv, D, Eg, Eb, L.N, L.V, L.Mpsi);
        L.sync();                                             // no lane may still be reading the staging area
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            double row[NR + 1];
            krow(i, z, D, Eg, Eb, L, row);
            double2* dst = reinterpret_cast<double2*>(&L.S(i, 0));
#pragma unroll
            for (int jj = 0; jj < NR / 2; ++jj) dst[jj] = make_double2(row[2 * jj], row[2 * jj + 1]);
            if (NR % 2) L.S(i, NR - 1) = row[NR - 1];
        }
    }

    // REG: this lane's rows r = s·G + g (first NCOLS columns) from the staging area into registers; padding rows are zero.
    template <int NCOLS, int RA>
    OD_HD static void fetch_rows(const Lin& L, double (&a)[RA][NCOLS]) {
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            const int r = s * G + L.g;
            const bool live = (G == 1) || ((s + 1) * G <= NR) || (r < NR);
            const double2* src = reinterpret_cast<const double2*>(&L.S(live ? r : 0, 0));
#pragma unroll
            for (int jj = 0; jj < NCOLS / 2; ++jj) { const double2 v = src[jj]; a[s][2 * jj] = live ? v.x : 0.0; a[s][2 * jj + 1] = live ? v.y : 0.0; }
            if (NCOLS % 2) { const double v = L.S(live ? r : 0, NCOLS - 1); a[s][NCOLS - 1] = live ? v : 0.0; }
        }
    }
    // every lane of the group writes the same values (benign); factor() synchronises before reading
    OD_HD static void assemble(const Z& z, const double* th, const double* trc, const double* trv, Lin& L) {
        static_assert(REG || NTP >= G, "each lane needs a private scratch vector");
        double D[NQ * NQ], Eg[NQ * NC1], Eb[NQ * NB1];
        M::jac(z.q, z.gam, z.b, th, trc, trv, D, Eg, Eb, L.N, L.V, L.Mpsi);
        L.sync();                                             // no lane may still be reading the previous factorisation
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
#pragma unroll
            for (int j = 0; j < NQ; ++j) L.K(i, j) = D[i * NQ + j];
#pragma unroll
            for (int j = 0; j < NC; ++j) L.K(i, NQ + j) = Eg[i * NC1 + j];
#pragma unroll
            for (int j = 0; j < NB; ++j) L.K(i, NQ + NC + j) = Eb[i * NB1 + j];
#pragma unroll
            for (int j = 0; j < NP; ++j) L.K(i, NQ + NC + NB + j) = 0.0;
        }
#pragma unroll
        for (int i = 0; i < NC; ++i) {
#pragma unroll
            for (int j = 0; j < NQ; ++j) L.K(NQ + i, j) = z.gam[i] * L.N[i * NQ + j];
#pragma unroll
            for (int j = NQ; j < NR; ++j) L.K(NQ + i, j) = (j == NQ + i) ? z.s[i] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const int row = NQ + NC + k;
#pragma unroll
            for (int c = 0; c < NQ; ++c) {
                double a = 0.0;
#pragma unroll
                for (int e = 0; e < M::cone_dim(k); ++e) a += z.b[M::cone_off(k) + e] * L.V[(M::cone_off(k) + e) * NQ + c];
                L.K(row, c) = a;
            }
#pragma unroll
            for (int i = 0; i < NC; ++i) L.K(row, NQ + i) = z.spsi[k] * L.Mpsi[k * NC1 + i];
#pragma unroll
            for (int j = 0; j < NB; ++j) L.K(row, NQ + NC + j) = (cone_of(j) == k) ? z.sb[j] : 0.0;
#pragma unroll
            for (int j = 0; j < NP; ++j) L.K(row, NQ + NC + NB + j) = (j == k) ? z.psi[k] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int k = cone_of(j), row = NQ + NC + NP + j;
#pragma unroll
            for (int c = 0; c < NQ; ++c) L.K(row, c) = z.psi[k] * L.V[j * NQ + c];
#pragma unroll
            for (int i = 0; i < NC; ++i) L.K(row, NQ + i) = z.sb[j] * L.Mpsi[k * NC1 + i];
#pragma unroll
            for (int jj = 0; jj < NB; ++jj) L.K(row, NQ + NC + jj) = (jj == j) ? z.spsi[k] : 0.0;
#pragma unroll
            for (int kk = 0; kk < NP; ++kk) L.K(row, NQ + NC + NB + kk) = (kk == k) ? z.b[j] : 0.0;
        }
    }

    // LU with partial (row) pivoting, fully unrolled: every workspace address is base + immediate except the pivot row.
    // Pivot search is done redundantly by every lane of the group (same result); the row interchange is split by columns and the
    // trailing update by rows across the G lanes.
    OD_HD static void factor(Lin& L) {
        bool ok = true;
        L.sync();
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            int p = k; double best = fabs(L.K(k, k));
#pragma unroll
            for (int i = k + 1; i < NR; ++i) { const double a = fabs(L.K(i, k)); if (a > best) { best = a; p = i; } }
            if (L.g == 0) L.PIV(k) = (double)p;
            ok = ok && (best > 0.0) && (best < INFINITY);
            if (G > 1) L.sync();                              // all lanes have read column k before rows move
            if (p != k) {
                double* rk = &L.K(k, 0); double* rp = &L.K(p, 0);
#pragma unroll
                for (int t = 0; t < (NR + G - 1) / G; ++t) {
                    const int j = L.g + G * t;
                    if (G == 1 || j < NR) { const double a = rk[j * PPB]; rk[j * PPB] = rp[j * PPB]; rp[j * PPB] = a; }
                }
            }
            if (G > 1) L.sync();
            const double inv = 1.0 / L.K(k, k);
            double prow[NR];
#pragma unroll
            for (int j = k + 1; j < NR; ++j) prow[j] = L.K(k, j);
#pragma unroll
            for (int t = 0; t < (NR - k - 1 + G - 1) / G; ++t) {
                const int i = k + 1 + L.g + G * t;
                if (G == 1 || i < NR) {
                    double* ri = &L.K(i, 0);
                    const double l = ri[k * PPB] * inv;
                    ri[k * PPB] = l;
#pragma unroll
                    for (int j = k + 1; j < NR; ++j) ri[j * PPB] -= l * prow[j];
                }
            }
            if (G > 1) L.sync();
            if (L.g == 0) L.K(k, k) = inv;                    // keep the reciprocal pivot (nobody reads (k,k) again before the solves)
        }
        L.sync();
        L.ok = ok;
    }

    // x ← K⁻¹ x for one right-hand side held in registers; run redundantly by every lane that needs the result.
    // `scratch` is a workspace vector private to the calling lane (used only for the dynamic-index row interchanges).
    OD_HD static void lu_solve(const Lin& L, double* x, double* scratch) {
        int piv[NR];
        bool any = false;
#pragma unroll
        for (int k = 0; k < NR; ++k) { piv[k] = (int)L.PIV(k); any = any || (piv[k] != k); }
        if (any) {                                            // interchanges need run-time indexing: do them in the workspace
#pragma unroll
            for (int i = 0; i < NR; ++i) scratch[i * PPB] = x[i];
#pragma unroll
            for (int k = 0; k < NR; ++k) {
                const int p = piv[k];
                if (p != k) { const double t = scratch[k * PPB]; scratch[k * PPB] = scratch[p * PPB]; scratch[p * PPB] = t; }
            }
#pragma unroll
            for (int i = 0; i < NR; ++i) x[i] = scratch[i * PPB];
        }
#pragma unroll
        for (int i = 1; i < NR; ++i) {
#pragma unroll
            for (int j = 0; j < i; ++j) x[i] -= L.K(i, j) * x[j];
        }
#pragma unroll
        for (int i = NR - 1; i >= 0; --i) {
#pragma unroll
            for (int j = i + 1; j < NR; ++j) x[i] -= L.K(i, j) * x[j];
            x[i] *= L.K(i, i);
        }
    }

    // reduced right-hand side of r
    OD_HD static void load_rhs(const Z& z, const R& r, double* x) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) x[i] = r.d[i];
#pragma unroll
        for (int i = 0; i < NC; ++i) x[NQ + i] = r.rgam[i] - z.gam[i] * r.rs[i];
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            double a = r.rc0[k] - z.spsi[k] * r.rpsi[k];
#pragma unroll
            for (int e = 0; e < M::cone_dim(k); ++e) a += z.b[M::cone_off(k) + e] * r.rv[M::cone_off(k) + e];
            x[NQ + NC + k] = a;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) x[NQ + NC + NP + j] = r.rc1[j] - z.sb[j] * r.rpsi[cone_of(j)] + z.psi[cone_of(j)] * r.rv[j];
    }

    // full Newton direction for right-hand side r:  rz Δ = r   (every lane computes the same D)
    OD_HD static void solve(const Lin& L, const Z& z, const R& r, Z& D) {
        double x[NR];
        load_rhs(z, r, x);
        if constexpr (LASM) {
            double xm[RPL];
            GJ::mine(x, xm, L.g);
            GJSM::solve(&L.S(0, 0), L.PV(), xm, x, L.g, L.gmask);
        } else if constexpr (REG) {
            double xm[RPL];
            GJ::mine(x, xm, L.g);
#if OD_EXTRACT_SMEM && OD_GJ_SMEM
            if constexpr (PW >= GJ::CINV + 2) GJ::template solve_sm<PW>(L.a, L.piv, xm, x, L.g, L.gmask, &L.S(0, 0));
            else GJ::solve(L.a, L.piv, xm, x, L.g, L.gmask);
#else
            GJ::solve(L.a, L.piv, xm, x, L.g, L.gmask);
#endif
        } else {
            lu_solve(L, x, &L.X(L.g, 0));
        }
        expand(L, r, x, D);
    }
    // REG: direction for the right-hand side that linearize() carried through the elimination (the residual at z itself)
    OD_HD static void solve_carried(const Lin& L, const Z& z, const R& r, Z& D) {
        if constexpr (LASM) {
            double x[NR];
            GJSM::extract(&L.S(0, 0), L.PV(), 0, x);
            expand(L, r, x, D);
        } else if constexpr (REG) {
            double x[NR];
#if OD_EXTRACT_SMEM && OD_GJ_SMEM
            if constexpr (PW >= GJ::CINV + 2) GJ::template extract_sm<PW>(L.a, L.piv, 0, x, L.g, L.gmask, &L.S(0, 0));
            else GJ::extract(L.a, L.piv, 0, x, L.gmask);
#else
            GJ::extract(L.a, L.piv, 0, x, L.gmask);
#endif
            expand(L, r, x, D);
        } else {
            solve(L, z, r, D);
        }
    }
    // reduced solution x = (Δq, Δγ, Δb, Δsψ) → full direction
    OD_HD static void expand(const Lin& L, const R& r, const double* x, Z& D) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) D.q[i] = x[i];
#pragma unroll
        for (int i = 0; i < NC; ++i) D.gam[i] = x[NQ + i];
#pragma unroll
        for (int j = 0; j < NB; ++j) D.b[j] = x[NQ + NC + j];
#pragma unroll
        for (int k = 0; k < NP; ++k) D.spsi[k] = x[NQ + NC + NB + k];
        // substituted variables
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            double ds = r.rs[i];
#pragma unroll
            for (int j = 0; j < NQ; ++j) ds += L.N[i * NQ + j] * D.q[j];
            D.s[i] = ds;
        }
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            double dpsi = r.rpsi[k];
#pragma unroll
            for (int i = 0; i < NC; ++i) dpsi += L.Mpsi[k * NC1 + i] * D.gam[i];
            D.psi[k] = dpsi;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            double v = -r.rv[j];
#pragma unroll
            for (int c = 0; c < NQ; ++c) v += L.V[j * NQ + c] * D.q[c];
            D.sb[j] = v;
        }
    }

    // ---- cone utilities --------------------------------------------------------------------------------------------
    // Step to the cone boundary as α = min(1, τ / ρ),  ρ = max over the cone variables of Δ_i / z_i (orthant pairs), of
    // (Δψ ∓ Δb)/(ψ ∓ b) (2-D second-order cones = two half-planes each: identical in exact arithmetic to the CVXOPT §8.2 expression
    // the oracle uses) and of the CVXOPT denominator ‖ρ_v‖ − ρ_s (3-D cones).  The reciprocals of the cone variables depend on the
    // iterate only, so the affine and the corrector step length of one iteration share them: one multiply + one max per candidate
    // instead of a cross-multiplied fraction comparison (9 instructions), 2 divisions per iteration instead of per cone.
    struct ConeRcp { double gam[NC1], s[NC1], pm[NP1], pp[NP1], dm[NP1], dp[NP1]; };
    OD_HD static void cone_rcp(const Z& z, ConeRcp& c) {
#pragma unroll
        for (int i = 0; i < NC; ++i) { c.gam[i] = pivot_rcp(z.gam[i]); c.s[i] = pivot_rcp(z.s[i]); }
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            if (M::cone_dim(k) == 1) {
                const int o = M::cone_off(k);
                c.pm[k] = pivot_rcp(z.psi[k] - z.b[o]); c.pp[k] = pivot_rcp(z.psi[k] + z.b[o]);
                c.dm[k] = pivot_rcp(z.spsi[k] - z.sb[o]); c.dp[k] = pivot_rcp(z.spsi[k] + z.sb[o]);
            }
        }
    }
    // Cones with two tangential components (3-D second-order cones; planar push: 4 contacts × primal / dual = 8 per step length):
    // their CVXOPT step lengths are the expensive part of step_length (reciprocal square root, reciprocal, square root each) and
    // identical in all lanes of a group.  With cooperating lanes (SPLIT_SOC) lane g evaluates candidate g, g+G, … and the group
    // combines the maxima by shuffle — a maximum of never-NaN, non-negative values: the same number in any order.
    __host__ __device__ static constexpr int nsoc3() { int n = 0; for (int k = 0; k < NP; ++k) if (M::cone_dim(k) == 2) ++n; return n; }
    __host__ __device__ static constexpr int soc3_cone(int j) {
        int n = 0, r = 0;
        for (int k = 0; k < NP; ++k) if (M::cone_dim(k) == 2) { if (n == j) r = k; ++n; }
        return r;
    }
    static constexpr bool SPLIT_SOC = OD_SPLIT_SOC && REG && G >= 4 && nsoc3() > 0 && NC > 0;

    OD_HD static double step_length(const Z& z, const ConeRcp& c, const Z& D, double tau, const int g = 0, const unsigned gmask = 0xffffffffu) {
        MaxAcc rho, rho2;                                      // two independent running maxima (shorter dependent chain)
#pragma unroll
        for (int i = 0; i < NC; ++i) { rho.add(D.gam[i] * c.gam[i]); rho2.add(D.s[i] * c.s[i]); }
        if constexpr (SPLIT_SOC) {
            constexpr int T = 2 * nsoc3();
            double m = 0.0;
#pragma unroll
            for (int r0 = 0; r0 < T; r0 += G) {
                // a lane without a candidate in this round keeps an interior point with a zero direction: contributes 0
                double l0 = 1.0, l1[2] = {0.0, 0.0}, d0 = 0.0, d1[2] = {0.0, 0.0};
#pragma unroll
                for (int t = r0; t < ((r0 + G < T) ? r0 + G : T); ++t) {
                    const int k = soc3_cone(t / 2), o = M::cone_off(k);
                    if (t - r0 == g) {
                        if (t % 2 == 0) { l0 = z.psi[k]; l1[0] = z.b[o]; l1[1] = z.b[o + 1]; d0 = D.psi[k]; d1[0] = D.b[o]; d1[1] = D.b[o + 1]; }
                        else { l0 = z.spsi[k]; l1[0] = z.sb[o]; l1[1] = z.sb[o + 1]; d0 = D.spsi[k]; d1[0] = D.sb[o]; d1[1] = D.sb[o + 1]; }
                    }
                }
                double bn = 1.0, bd = 0.0;                     // soc_step leaves (τ, den) when den > 0
                soc_step<2>(l0, l1, d0, d1, 1.0, bn, bd);
                m = od_max(m, bd);
            }
            rho2.add(Grp<G>::dmax_all(m, gmask));
        }
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const int o = M::cone_off(k);
            if (M::cone_dim(k) == 1) {
                rho.add((D.psi[k] - D.b[o]) * c.pm[k]); rho2.add((D.psi[k] + D.b[o]) * c.pp[k]);
                rho.add((D.spsi[k] - D.sb[o]) * c.dm[k]); rho2.add((D.spsi[k] + D.sb[o]) * c.dp[k]);
            } else if constexpr (!SPLIT_SOC) {
                double bn = 1.0, bd = 0.0;                     // soc_step leaves (τ, den) when den > 0
                soc_step<2>(z.psi[k], &z.b[o], D.psi[k], &D.b[o], 1.0, bn, bd);
                rho.add(bd);
                bn = 1.0; bd = 0.0;
                soc_step<2>(z.spsi[k], &z.sb[o], D.spsi[k], &D.sb[o], 1.0, bn, bd);
                rho2.add(bd);
            }
        }
        rho.add(rho2.v);                                       // only a ratio above τ > 0 shortens the step: no clamp at 0 needed
        return (rho.v > tau) ? tau * pivot_rcp(rho.v) : 1.0;        // rho > τ ≥ 0.95: the short reciprocal is safe (fastmath.cuh)
    }

    // Σ ⟨primal − aΔp, dual − aΔd⟩ over all cones
    OD_HD static double cone_dot(const Z& z, const Z& D, double a) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < NC; ++i) s += (z.gam[i] - a * D.gam[i]) * (z.s[i] - a * D.s[i]);
#pragma unroll
        for (int k = 0; k < NP; ++k) s += (z.psi[k] - a * D.psi[k]) * (z.spsi[k] - a * D.spsi[k]);
#pragma unroll
        for (int j = 0; j < NB; ++j) s += (z.b[j] - a * D.b[j]) * (z.sb[j] - a * D.sb[j]);
        return s;
    }

    OD_HD static void candidate(const Z& z, const Z& D, double a, Z& c) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) c.q[i] = z.q[i] - a * D.q[i];
#pragma unroll
        for (int i = 0; i < NC; ++i) { c.gam[i] = z.gam[i] - a * D.gam[i]; c.s[i] = z.s[i] - a * D.s[i]; }
#pragma unroll
        for (int k = 0; k < NP; ++k) { c.psi[k] = z.psi[k] - a * D.psi[k]; c.spsi[k] = z.spsi[k] - a * D.spsi[k]; }
#pragma unroll
        for (int j = 0; j < NB; ++j) { c.b[j] = z.b[j] - a * D.b[j]; c.sb[j] = z.sb[j] - a * D.sb[j]; }
    }


    // Newton direction at z (predictor, centering, Mehrotra corrector) and the step length along it
    OD_HD static void direction(const Lin& L, const Z& z, const R& r, double r_vio, double k_vio, Z& D, double& alpha) {
        if (NCONE > 0) {
            solve_carried(L, z, r, D);                           // affine direction
            ConeRcp rcp;
            cone_rcp(z, rcp);
            const double a_aff = step_length(z, rcp, D, 1.0, L.g, L.gmask);
            const double mu = cone_dot(z, D, 0.0) * (1.0 / (NCONE > 0 ? NCONE : 1));
            const double mu_aff = cone_dot(z, D, a_aff) * (1.0 / (NCONE > 0 ? NCONE : 1));
            const double ratio = od_min(od_max(0.0, mu_aff * pivot_rcp(mu)), 1.0);   // μ > 0 at any non-converged iterate
            const double kappa = ratio * ratio * ratio * mu;      // max(σμ, κ_tol/undercut) with undercut = Inf
            // corrector right-hand side: r(z;κ) + Δaff_primal ∘ Δaff_dual on the bilinear rows
            R rc = r;
#pragma unroll
            for (int i = 0; i < NC; ++i) rc.rgam[i] = (r.rgam[i] - kappa) + D.gam[i] * D.s[i];
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                double acc = D.psi[k] * D.spsi[k];
#pragma unroll
                for (int e = 0; e < M::cone_dim(k); ++e) {
                    const int j = M::cone_off(k) + e;
                    acc += D.b[j] * D.sb[j];
                    rc.rc1[j] = r.rc1[j] + (D.psi[k] * D.sb[j] + D.spsi[k] * D.b[j]);
                }
                rc.rc0[k] = (r.rc0[k] - kappa) + acc;
            }
            solve(L, z, rc, D);
            const double viol = od_max(r_vio, k_vio);
            alpha = step_length(z, rcp, D, od_max(0.95, 1.0 - viol * viol), L.g, L.gmask);
        } else {
            solve_carried(L, z, r, D);                           // no cones: plain Newton direction, full step
            alpha = 1.0;
        }
    }

    // ---- IFT: ∂q3/∂θ' = −(rz⁻¹ rθ')[q rows]; column c of the NQ×NTP column-major result goes to dq1 / dq2 / du ---------
    // The NTP reduced right-hand sides are written to the workspace (redundantly by all lanes), then lane g solves columns
    // g, g+G, … with the factorisation of rz at the final iterate.
    OD_HD static void sensitivities(const Lin& L, const Z& z, const double* th, const double* trc, const double* trv, double* dq1, double* dq2, double* du) {
        {
            double Dth[NQ * NTP], Vth[NB1 * NTP];
            M::jacth(z.q, z.gam, z.b, th, trc, trv, Dth, Vth);
            R r;
#pragma unroll
            for (int i = 0; i < NC1; ++i) { r.rs[i] = 0.0; r.rgam[i] = 0.0; }
#pragma unroll
            for (int i = 0; i < NP1; ++i) { r.rpsi[i] = 0.0; r.rc0[i] = 0.0; }
#pragma unroll
            for (int i = 0; i < NB1; ++i) r.rc1[i] = 0.0;
#pragma unroll
            for (int c = 0; c < NTP; ++c) {
#pragma unroll
                for (int i = 0; i < NQ; ++i) r.d[i] = Dth[i * NTP + c];
#pragma unroll
                for (int j = 0; j < NB; ++j) r.rv[j] = Vth[j * NTP + c];
                double x[NR];
                load_rhs(z, r, x);
#pragma unroll
                for (int i = 0; i < NR; ++i) L.X(c, i) = x[i];
            }
        }
        L.sync();
#pragma unroll 1
        for (int c = L.g; c < NTP; c += G) {
            double* col = &L.X(c, 0);
            double x[NR];
#pragma unroll
            for (int i = 0; i < NR; ++i) x[i] = col[i * PPB];
            lu_solve(L, x, col);
            double* dst = (c < NQ) ? (dq1 + c * NQ) : (c < 2 * NQ) ? (dq2 + (c - NQ) * NQ) : (du + (c - 2 * NQ) * NQ);
#pragma unroll
            for (int i = 0; i < NQ; ++i) dst[i] = -x[i];
        }
        L.sync();
    }

    // REG variant: [K | NTP reduced right-hand sides] is staged once, every lane takes its rows (NR + NTP columns) into registers and
    // one Gauss–Jordan pass leaves, in the pivot row of unknown i, the solution component i of every right-hand side.  The q rows
    // are written (negated) into the packed output row in the staging area: OUT[NQ + c·NQ + i] = ∂q3_i/∂θ'_c.  Returns factor's ok.
    OD_HD static bool sensitivities_reg(Lin& L, const Z& z, const double* th, const double* trc, const double* trv) {
        stage_matrix(z, th, trc, trv, L);
        {
            double Dth[NQ * NTP], Vth[NB1 * NTP];
            M::jacth(z.q, z.gam, z.b, th, trc, trv, Dth, Vth);
            R r;
#pragma unroll
            for (int i = 0; i < NC1; ++i) { r.rs[i] = 0.0; r.rgam[i] = 0.0; }
#pragma unroll
            for (int i = 0; i < NP1; ++i) { r.rpsi[i] = 0.0; r.rc0[i] = 0.0; }
#pragma unroll
            for (int i = 0; i < NB1; ++i) r.rc1[i] = 0.0;
#pragma unroll
            for (int c = 0; c < NTP; ++c) {
#pragma unroll
                for (int i = 0; i < NQ; ++i) r.d[i] = Dth[i * NTP + c];
#pragma unroll
                for (int j = 0; j < NB; ++j) r.rv[j] = Vth[j * NTP + c];
                double x[NR];
                load_rhs(z, r, x);
#pragma unroll
                for (int i = 0; i < NR; ++i) L.S(i, NR + c) = x[i];
            }
        }
        L.sync();
        double a[RPL][NR + NTP];
        int piv[NR];
        fetch_rows<NR + NTP>(L, a);
#if OD_GJ_SMEM
        const bool ok = GJS::template factor_sm<PW>(a, piv, L.g, L.gmask, &L.S(0, 0));
#else
        const bool ok = GJS::factor(a, piv, L.g, L.gmask);
#endif
        L.sync();                                             // every lane has its rows: the staging area becomes the output row
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            const int wl = piv[i] & (G - 1), ws = piv[i] >> Grp<G>::LG;
            const double inv = GJS::pick(a, i, ws);
            if (L.g == wl) {
#pragma unroll
                for (int c = 0; c < NTP; ++c) L.ws[NQ + c * NQ + i] = -(GJS::pick(a, NR + c, ws) * inv);
            }
        }
        return ok;
    }

    // ---- robust IFT for models with redundant contact constraints (planar push: 4 corner contacts on a 3-DoF block) ---------
    // At a sticking iterate with κ → 0 the friction multipliers are not unique and rz is numerically rank deficient, while the
    // q rows of δz stay well defined.  Partial pivoting then divides by rounding-level pivots (or hits an exact zero); here the
    // factorisation at the final iterate uses COMPLETE pivoting, stops at the numerical rank, and solves the consistent system
    // with the free multipliers set to zero.  Runs once per problem, on all lanes of the group (dynamic loops: small code).
    OD_HD static bool sensitivities_robust(Lin& L, const Z& z, const double* th, const double* trc, const double* trv, double* dq1, double* dq2, double* du) {
        {
            double Dth[NQ * NTP], Vth[NB1 * NTP];
            M::jacth(z.q, z.gam, z.b, th, trc, trv, Dth, Vth);
            R r;
#pragma unroll
            for (int i = 0; i < NC1; ++i) { r.rs[i] = 0.0; r.rgam[i] = 0.0; }
#pragma unroll
            for (int i = 0; i < NP1; ++i) { r.rpsi[i] = 0.0; r.rc0[i] = 0.0; }
#pragma unroll
            for (int i = 0; i < NB1; ++i) r.rc1[i] = 0.0;
#pragma unroll
            for (int c = 0; c < NTP; ++c) {
#pragma unroll
                for (int i = 0; i < NQ; ++i) r.d[i] = Dth[i * NTP + c];
#pragma unroll
                for (int j = 0; j < NB; ++j) r.rv[j] = Vth[j * NTP + c];
                double x[NR];
                load_rhs(z, r, x);
#pragma unroll
                for (int i = 0; i < NR; ++i) L.X(c, i) = x[i];
            }
        }
        L.sync();
        // All G lanes of the group work on the shared-memory system: the pivot search scans rows g, g+G, … per lane and combines by
        // shuffle (largest magnitude, first in row-major order among equals — the choice a sequential scan makes), interchanges are
        // split by column / row, the elimination by row, the back-substitution by right-hand side.  Every element sees the same
        // operations in the same order as with one lane, so the result does not depend on G.  The groups of a warp run in lockstep
        // (full-mask shuffles): a group that has reached its numerical rank idles through the remaining steps.
        double amax = 0.0;
        for (int i = L.g; i < NR; i += G) for (int j = 0; j < NR; ++j) amax = fmax(amax, fabs(L.K(i, j)));
        { int dummy = 0; Grp<G>::argmax_all(amax, dummy, L.gmask); }
        const double tol = 1e-13 * amax;      // exact redundancy leaves pivots at rounding level (≲1e-16·amax); κ-level pivots (≳1e-11) are genuine
        int rank = NR;
        for (int k = 0; k < NR; ++k) {
            const bool live = rank == NR;     // false once the numerical rank has been found
            double best = -1.0; int at = k * NR + k;
            if (live)
                for (int i = k + L.g; i < NR; i += G) for (int j = k; j < NR; ++j) { const double a = fabs(L.K(i, j)); if (a > best) { best = a; at = i * NR + j; } }
            Grp<G>::argmax_all(best, at, L.gmask);
            if (live && !(best > tol)) rank = k;
            const bool go = rank == NR;
            const int p = at / NR, q = at % NR;
            if (go) {
                if (L.g == 0) L.CP(k) = (double)q;
                if (p != k) {
                    for (int j = L.g; j < NR; j += G) { const double t = L.K(k, j); L.K(k, j) = L.K(p, j); L.K(p, j) = t; }
                    for (int c = L.g; c < NTP; c += G) { const double t = L.X(c, k); L.X(c, k) = L.X(c, p); L.X(c, p) = t; }     // row swap applied to every rhs
                }
            }
            L.sync();
            if (go && q != k) for (int i = L.g; i < NR; i += G) { const double t = L.K(i, k); L.K(i, k) = L.K(i, q); L.K(i, q) = t; }
            L.sync();
            if (go) {
                const double inv = 1.0 / L.K(k, k);
                for (int i = k + 1 + L.g; i < NR; i += G) {
                    const double l = L.K(i, k) * inv;
                    if (l != 0.0) {
                        for (int j = k + 1; j < NR; ++j) L.K(i, j) -= l * L.K(k, j);
                        for (int c = 0; c < NTP; ++c) L.X(c, i) -= l * L.X(c, k);                                        // forward elimination of every rhs
                    }
                }
            }
            L.sync();
        }
        for (int c = L.g; c < NTP; c += G) {
            for (int i = rank; i < NR; ++i) L.X(c, i) = 0.0;                                  // free (non-unique) unknowns
            for (int i = rank - 1; i >= 0; --i) {
                double sacc = L.X(c, i);
                for (int j = i + 1; j < rank; ++j) sacc -= L.K(i, j) * L.X(c, j);
                L.X(c, i) = sacc / L.K(i, i);
            }
            for (int k = rank - 1; k >= 0; --k) { const int q = (int)L.CP(k); if (q != k) { const double t = L.X(c, k); L.X(c, k) = L.X(c, q); L.X(c, q) = t; } }
            double* dst = (c < NQ) ? (dq1 + c * NQ) : (c < 2 * NQ) ? (dq2 + (c - NQ) * NQ) : (du + (c - 2 * NQ) * NQ);
            for (int i = 0; i < NQ; ++i) dst[i] = -L.X(c, i);
        }
        L.sync();
        return rank > 0;
    }

    // iterate ↔ flat array of NZ doubles, element e at p[e·ST] (the order of the snapshots handed to contact_ift_kernel)
    template <int ST>
    OD_HD static void pack_z(const Z& z, double* p) {
        int e = 0;
#pragma unroll
        for (int i = 0; i < NQ; ++i) p[(e++) * ST] = z.q[i];
#pragma unroll
        for (int i = 0; i < NC; ++i) { p[(e++) * ST] = z.gam[i]; p[(e++) * ST] = z.s[i]; }
#pragma unroll
        for (int i = 0; i < NP; ++i) { p[(e++) * ST] = z.psi[i]; p[(e++) * ST] = z.spsi[i]; }
#pragma unroll
        for (int i = 0; i < NB; ++i) { p[(e++) * ST] = z.b[i]; p[(e++) * ST] = z.sb[i]; }
    }
    template <int ST>
    OD_HD static void unpack_z(const double* p, Z& z) {
        int e = 0;
#pragma unroll
        for (int i = 0; i < NQ; ++i) z.q[i] = p[(e++) * ST];
#pragma unroll
        for (int i = 0; i < NC; ++i) { z.gam[i] = p[(e++) * ST]; z.s[i] = p[(e++) * ST]; }
#pragma unroll
        for (int i = 0; i < NP; ++i) { z.psi[i] = p[(e++) * ST]; z.spsi[i] = p[(e++) * ST]; }
#pragma unroll
        for (int i = 0; i < NB; ++i) { z.b[i] = p[(e++) * ST]; z.sb[i] = p[(e++) * ST]; }
    }
    // iterate ↔ workspace snapshot (all lanes of a group write identical values)
    OD_HD static void store_z(const Lin& L, const Z& z) { pack_z<WS_STRIDE>(z, &L.ws[(REG ? ROFF_ZS : OFF_ZS) * WS_STRIDE]); }
    OD_HD static void load_z(const Lin& L, Z& z) { unpack_z<WS_STRIDE>(&L.ws[(REG ? ROFF_ZS : OFF_ZS) * WS_STRIDE], z); }

    // initialize_z! (reference src/models/planar_push/simulator.jl:52-60 and the same pattern in the other models)
    OD_HD static void init_z(const double* q2, Z& z) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) z.q[i] = q2[i];
#pragma unroll
        for (int i = 0; i < NC1; ++i) { z.gam[i] = 1.0; z.s[i] = 1.0; }
#pragma unroll
        for (int i = 0; i < NP1; ++i) { z.psi[i] = 1.0; z.spsi[i] = 1.0; }
#pragma unroll
        for (int i = 0; i < NB1; ++i) { z.b[i] = 0.1; z.sb[i] = 0.1; }
    }

    OD_HD static bool finite(const Z& z, double r_vio, double k_vio) {
        double s = r_vio + k_vio;
#pragma unroll
        for (int i = 0; i < NQ; ++i) s += z.q[i];
#pragma unroll
        for (int i = 0; i < NC; ++i) s += z.gam[i] + z.s[i];
#pragma unroll
        for (int k = 0; k < NP; ++k) s += z.psi[k] + z.spsi[k];
#pragma unroll
        for (int j = 0; j < NB; ++j) s += z.b[j] + z.sb[j];
        return isfinite(s);
    }
};

// Per-launch arguments of the batched step kernel.  Arrays are row-per-problem with explicit strides (in doubles) so that the
// same kernel serves separate arrays (C ABI) and one packed row [q3 | ∂q3/∂q1 | ∂q3/∂q2 | ∂q3/∂u] (all-gather buffer).
struct StepArgs {
    int B;
    const double* q1; const double* q2; const double* u;
    int in_stride_q, in_stride_u;
    double* q3; double* dq1; double* dq2; double* du;          // any may be null
    int out_stride_q3, out_stride_dq, out_stride_du;
    int* status; int* iters;                                   // may be null
    double h;
    double fric[4];
    int want_eval, want_grad;
    // gradient bundle (reference src/gradient_bundle.jl:89-100): problem i = sample i/(n_eta+1) perturbed by eta row i%(n_eta+1) − 1
    // (row −1 = nominal); eta is n_eta × (2NQ+NU), null when unused.
    const double* eta; int n_eta;
    long long eta_i0;          // index of this launch's first problem on the flattened (sample × perturbation) axis (multi-GPU slices)
    // Fused all-gather (multi-GPU): when n_peers > 1 the packed output row of problem i is also stored into the gather buffers of
    // the other ranks (peer-mapped device pointers, NVLink P2P) as soon as the problem has finished — the transfer overlaps the
    // remaining problems' compute and replaces the separate ncclAllGather.  `gather_row0` = first row of this rank's shard,
    // `gather_width` = doubles per row; the local outputs (q3/dq1/dq2/du) must already point into this rank's own gather buffer.
    int n_peers, self_rank;
    long long gather_row0; int gather_width;
    double* peer_out[8];
    // Packed output (q3/dq1/dq2/du point into one row of width NOUT, 16-byte aligned): the register path stores it as 16-byte pairs.
    int packed_out;
    // Packed input (q2 = q1 + NQ, u = q1 + 2NQ, one stride): the register path loads the row cooperatively.
    int in_packed;
    // RoboDojo.step!(sim, q, v, u, t) call shape (reference examples/hopper.jl:63,89,112,133,157): the `q1` array holds the velocity
    // v1 and the data vector takes q1 = q2 − h·v1 directly (no (q2 − q1)/h round trip).
    int in_vel;
    // Cross-GPU barrier fused into the kernel (replaces the separate barrier launch after the fused all-gather): every thread
    // fences its peer stores, the last block of the grid to finish publishes `sync_epoch` into slot `self_rank` of every peer's
    // flag array and waits until every peer has published it here.  sync_flags[r] = rank r's flag array (world × u64) as mapped
    // in this process; sync_counter = this rank's block counter (device memory, zero before the first launch).  Null = off.
    unsigned long long* sync_flags[8];
    unsigned int* sync_counter;
    unsigned long long sync_epoch;
    // sync_epoch == 0: the epoch lives on the device — sync_epoch_dev (this rank's u64, zero before the first launch) is advanced by
    // one per launch by the last block, so a CUDA graph holding such launches can be replayed (every rank replays the same launches).
    unsigned long long* sync_epoch_dev;
    // Persistent sweep (contact_sweep_kernel / contact_ift_kernel): work_queue = a zeroed counter in device memory from which the
    // groups claim problem indices; z_snapshots = B × NZ doubles, the iterate of every problem at its gradient tolerance, handed from
    // the sweep kernel to the IFT kernel.
    unsigned int* work_queue;
    double* z_snapshots;
    // Park and resume (two launches of contact_sweep_kernel): with park_iter > 0 a problem that is still unfinished after park_iter
    // iterations leaves the sweep — iterate to z_park (B × NZ), progress words to park_info (2 ints per problem), its index
    // appended to park_list (counter: work_queue[1]) — so that the few problems that run to max_iter do not start their 100
    // dependent iterations at whatever time the queue happens to reach them.  The second launch (resume = 1; queue: work_queue[2])
    // continues all of them at once, spread over the SMs with one warp each and more lanes per problem.
    int park_iter, resume;
    double* z_park; int* park_list; int* park_info;
    // HOST side only (never read by a kernel): a second stream and two events of the handle.  When present, the resume launch and the IFT
    // of the problems that finished in the sweep run side by side (fork after the sweep, join before the IFT of the parked problems) —
    // ≈ 1 % of a batch is parked, but it has 84 dependent iterations left: a millisecond during which the SMs would otherwise idle.
    void* side_stream; void* ev_fork; void* ev_join;
    // NVLink multicast alias of the gather buffers (torch symmetric memory multicast_ptr; null = per-peer stores): ONE multimem.st
    // per 16 bytes reaches the gather buffer of every rank — this one included — and NVSwitch does the replication.
    double* mc_out;
    unsigned long long* mc_flags;                              // multicast alias of the flag arrays (or null): one store publishes the epoch everywhere
    SolverOpts opts;
};

// One problem (G cooperating lanes).  A single iterate sequence serves both simulators of ImplicitDynamics: eval_sim and grad_sim
// (reference src/dynamics.jl:60-64) start from the same initialisation and differ only in κ_tol, so the looser one is a prefix
// of the tighter one.  The IFT is taken at the first iterate meeting the gradient tolerance, q3 at the first meeting the eval one.
// The loop is a small state machine with ONE call site each for the residual, the factorisation and the direction, so that the
// instruction footprint stays small and lanes of a warp that are in different phases (line search / new iteration) share code.
//
// Warp-synchronous execution: every lane of a warp walks the state machine in lockstep until the slowest problem of the warp has
// finished (finished problems keep executing with their state frozen; the padding lanes of a partial last warp repeat the last problem).
// That costs nothing — a warp lasts as long as its slowest problem anyway — and makes every shuffle / __syncwarp a full-warp
// operation on converged lanes: a partial-mask shuffle is bracketed by WARPSYNC/ENDCOLLECTIVE plus register moves (≈10 SASS
// instructions per shuffle, measured), a full-mask one is a single SHFL.
OD_HD bool warp_any(bool p) {
#ifdef __CUDA_ARCH__
    return __any_sync(0xffffffffu, p);
#else
    if (HostLaneTeam* t = host_lane_team()) return t->any(p);
    return p;
#endif
}

// BSYNC (block-phased execution; large models): the warps of a block take the two votes of the state machine block-wide, through
// __syncthreads_or — a barrier, so all warps of the block walk the loop in phase and every instruction-cache line that one warp
// fetches serves the others (the planar push's loop is 108 KB of straight-line code: with 8 independent warps per SM every warp
// streamed it from L2 on its own, 6 stall cycles per issued instruction).  A warp whose problems have all finished skips the body
// and only keeps the barriers.
OD_HD bool block_or(bool p) {
#ifdef __CUDA_ARCH__
    return __syncthreads_or(p) != 0;
#else
    return warp_any(p);
#endif
}

// θ.q1 = q2 − h·v1 with v1 = (q2 − q1)/h (src/dynamics.jl:84-86, RoboDojo.step!): rounded operation by operation, never contracted
// into a fused multiply-add — the reference (Julia) and the oracle (gcc -ffp-contract=off) round the product first, and every
// instantiation of the solver (per-warp kernel, persistent sweep, IFT kernel, rollouts) must see the same data vector bit for bit.
OD_HD void theta_q1(const double x1, const double x2, const double h, const bool in_vel, double& q1p, double& v1out) {
#ifdef __CUDA_ARCH__
    const double v1 = in_vel ? x1 : __ddiv_rn(__dsub_rn(x2, x1), h);
    q1p = __dsub_rn(x2, __dmul_rn(h, v1));
#else
    const double v1 = in_vel ? x1 : (x2 - x1) / h;
    volatile double hv = h * v1;
    q1p = x2 - hv;
#endif
    v1out = v1;
}

// Data vector θ and the initial configuration of problem i, every lane loading every value (no collectives: callable from
// divergent code).  Same arithmetic as the prologue of contact_step_one.
template <class M>
OD_HD void load_problem_plain(const StepArgs& a, const int i, double* th, double* q2v) {
    constexpr int NQ = M::NQ, NU = M::NU;
    const long long ig = a.eta ? a.eta_i0 + i : i;
    const int src = a.eta ? (int)(ig / (a.n_eta + 1)) : i;
    const int pert = a.eta ? (int)(ig % (a.n_eta + 1)) : 0;
    const double* p1 = a.q1 + (size_t)src * a.in_stride_q;
    const double* p2 = a.q2 + (size_t)src * a.in_stride_q;
    const double* pu = a.u + (size_t)src * a.in_stride_u;
    const double* pe = (pert > 0) ? a.eta + (size_t)(pert - 1) * (2 * NQ + NU) : nullptr;
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
        double x1 = p1[k], x2 = p2[k];
        if (pe) { x1 += pe[k]; x2 += pe[NQ + k]; }
        double v1;
        theta_q1(x1, x2, a.h, a.in_vel != 0, th[k], v1);
        th[NQ + k] = x2;
        q2v[k] = x2;
    }
#pragma unroll
    for (int k = 0; k < NU; ++k) th[2 * NQ + k] = pe ? pu[k] + pe[2 * NQ + k] : pu[k];
#pragma unroll
    for (int k = 0; k < M::NF; ++k) th[2 * NQ + NU + k] = a.fric[k];
    th[M::NTH - 1] = a.h;
}

template <class M, int G, int PPB, bool REG = false, bool BSYNC = false>
OD_HD void contact_step_one(const StepArgs& a, const int i, double* ws, const int g, const unsigned gmask) {
    typedef ContactIP<M, G, PPB, REG> IP;
    constexpr int NQ = M::NQ, NU = M::NU;
    double th[M::NTH];
    typename IP::Z z_r, D_r, zc_r;
    typename IP::R rc_r;
    typename IP::Z& zc = IP::shared_obj(ws, 0, zc_r);        // ZSM models: one copy per group in the workspace (see ContactIP::ZSM)
    typename IP::R& rc = IP::shared_obj(ws, 1, rc_r);
    typename IP::Z& D = IP::shared_obj(ws, 2, D_r);
    typename IP::Z& z = IP::shared_obj(ws, 3, z_r);
    {
        const long long ig = a.eta ? a.eta_i0 + i : i;
        const int src = a.eta ? (int)(ig / (a.n_eta + 1)) : i;
        const int pert = a.eta ? (int)(ig % (a.n_eta + 1)) : 0;
        const double* p1 = a.q1 + (size_t)src * a.in_stride_q;
        const double* p2 = a.q2 + (size_t)src * a.in_stride_q;
        const double* pu = a.u + (size_t)src * a.in_stride_u;
        const double* pe = (pert > 0) ? a.eta + (size_t)(pert - 1) * (2 * NQ + NU) : nullptr;
        // One packed input row [q1 | q2 | u] per problem: the G lanes load it once between them (consecutive lanes, consecutive
        // doubles — whole sectors even when the row lives in pinned host memory) and exchange the values by shuffle; otherwise
        // every lane loads every value (same addresses within a group: a broadcast).
        constexpr int NIN = 2 * NQ + NU;
        double xin[NIN];
        bool coop = false;
        if constexpr (REG && G > 1) {
            coop = a.in_packed != 0;
            if (coop) {
                double v[(NIN + G - 1) / G];
#pragma unroll
                for (int t = 0; t < (NIN + G - 1) / G; ++t) { const int e = g + G * t; v[t] = (e < NIN) ? p1[e] : 0.0; }
#pragma unroll
                for (int e = 0; e < NIN; ++e) xin[e] = Grp<G>::bcast(v[e / G], e % G, gmask);
            }
        }
        if (!coop) {
#pragma unroll
            for (int k = 0; k < NQ; ++k) { xin[k] = p1[k]; xin[NQ + k] = p2[k]; }
#pragma unroll
            for (int k = 0; k < NU; ++k) xin[2 * NQ + k] = pu[k];
        }
        double q2v[NQ];
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            double x1 = xin[k], x2 = xin[NQ + k];
            if (pe) { x1 += pe[k]; x2 += pe[NQ + k]; }
            double v1;                                         // src/dynamics.jl:84-86 (or the caller's v1: step!(sim, q2, v1, u1, t))
            theta_q1(x1, x2, a.h, a.in_vel != 0, th[k], v1);   // RoboDojo.step!: q1 = q2 − h v1
            th[NQ + k] = x2;
            q2v[k] = x2;
        }
#pragma unroll
        for (int k = 0; k < NU; ++k) th[2 * NQ + k] = pe ? xin[2 * NQ + k] + pe[2 * NQ + k] : xin[2 * NQ + k];
#pragma unroll
        for (int k = 0; k < M::NF; ++k) th[2 * NQ + NU + k] = a.fric[k];
        th[M::NTH - 1] = a.h;
        IP::init_z(q2v, z);
    }
    typename IP::Lin L;
    L.ws = ws; L.g = g; L.gmask = gmask; L.ok = true;
    double trc[IP::NTC1], trv[IP::NTV1];            // sin/cos of the θ-only arguments (once) and of the q-dependent ones (per candidate)
    M::trig_const(th, trc);
    double r_vio = 0.0, k_vio = 0.0, alpha = 0.0;
#if OD_INPLACE_Z
    double step = 0.0;
#endif
    D = z;                                  // any finite values: the first candidate uses alpha = 0
    bool first = true, eval_done = !a.want_eval, grad_done = !a.want_grad;
    bool active = true;                    // this problem is still iterating
    int it = 0, ls = 0, it_e = 0, it_g = 0, st_e = 0, st_g = 0;
    for (;;) {
        // ---- candidate z − αΔ and its residual (the only residual call site) ----------------------------------------------
        double rv2, kv2;
#if OD_INPLACE_Z
        // Prepared variant: the iterate is advanced in place, z ← z − step·Δ, and a rejected step is taken back by the
        // difference of the step lengths — no second copy of the iterate, no copy on acceptance.  Identical arithmetic on the
        // accepted path; a retried candidate (rare: no hopper / cartpole / acrobot problem of the test batches ever retries) is
        // z − αΔ + (α − α')Δ instead of z − α'Δ, a rounding-level difference.
        IP::candidate(z, D, step, z);
        M::trig_var(z.q, th, trv);
        IP::residual(z, th, trc, trv, rc, rv2, kv2);
        const bool retry = active && !(first || rv2 <= r_vio || kv2 <= k_vio || ls >= a.opts.max_ls);
        step = 0.0;                                                 // a problem that does not retry re-evaluates its unchanged candidate
        if (retry) { const double a2 = alpha * a.opts.ls_scale; step = a2 - alpha; alpha = a2; ++ls; }
        if (warp_any(retry)) continue;
        if (active) {
            r_vio = rv2; k_vio = kv2;
#else
        bool retry = false;
        if (!BSYNC || warp_any(active)) {                            // (BSYNC: a warp without live problems only keeps the barriers)
            IP::candidate(z, D, alpha, zc);
            M::trig_var(zc.q, th, trv);
            IP::residual(zc, th, trc, trv, rc, rv2, kv2);
            retry = active && !(first || rv2 <= r_vio || kv2 <= k_vio || ls >= a.opts.max_ls);
            if (retry) { alpha *= a.opts.ls_scale; ++ls; }          // residual line search: halve and retry
        }
        if (BSYNC ? block_or(retry) : warp_any(retry)) continue;    // (the other problems of the warp / block re-evaluate their unchanged candidate)
        if (active) {
            z = zc; r_vio = rv2; k_vio = kv2;
#endif
            if (!first) ++it;
            first = false;
            // ---- accepted iterate: termination tests ------------------------------------------------------------------------
            const bool bad = !IP::finite(z, r_vio, k_vio);
            const bool capped = it >= a.opts.max_iter;
            const bool rok = r_vio < a.opts.r_tol;
            const bool conv_e = rok && (k_vio < a.opts.kappa_eval_tol);
            const bool conv_g = rok && (k_vio < a.opts.kappa_grad_tol);
            if (!eval_done && (conv_e || capped || bad)) {
                eval_done = true; it_e = it; st_e = bad ? ST_FAIL : (conv_e ? ST_OK : ST_MAXIT);
                if constexpr (REG) {                             // kept in the workspace; written out with the rest of the row
#pragma unroll
                    for (int k = 0; k < NQ; ++k) ws[IP::ROFF_Q3 + k] = z.q[k];
                } else if (a.q3 && g == 0) {
                    double* o = a.q3 + (size_t)i * a.out_stride_q3;
#pragma unroll
                    for (int k = 0; k < NQ; ++k) o[k] = z.q[k];
                }
            }
            if (!grad_done && (conv_g || capped || bad)) {       // the IFT itself is deferred until the loop has finished
                grad_done = true; it_g = it; st_g = bad ? ST_FAIL : (conv_g ? ST_OK : ST_MAXIT);
                IP::store_z(L, z);
            }
            if (bad || capped || (eval_done && grad_done)) active = false;
        }
        if (BSYNC ? !block_or(active) : !warp_any(active)) break;
        if constexpr (BSYNC) { if (!warp_any(active)) continue; }
        IP::linearize(z, th, trc, trv, rc, L);                   // trv still belongs to z (the candidate that was just accepted)
        if (active && !L.ok) {
            if (!eval_done) { st_e = ST_FAIL; it_e = it; eval_done = true; }
            if (!grad_done) { st_g = ST_FAIL; it_g = it; IP::store_z(L, z); grad_done = true; }
            active = false;
        }
        IP::direction(L, z, rc, r_vio, k_vio, D, alpha);
        if (!active) alpha = 0.0;                                // a finished problem stays where it is
#if OD_INPLACE_Z
        step = alpha;
#endif
        ls = 0;
    }
    // ---- IFT at the snapshot.  It sits after the loop on purpose: the problems of a warp converge at different iterations, and a
    // sensitivity pass inside the loop would be executed once per distinct convergence iteration (up to 8× per warp).
    if constexpr (REG) {
        bool grad = a.want_grad && a.dq1;
        if (grad) {
            IP::load_z(L, z);
            M::trig_var(z.q, th, trv);
            if constexpr (M::ROBUST_IFT) {
                // rank-revealing factorisation in the shared-memory-LU layout; it writes the Jacobian blocks to their destination
                // itself, only q3 goes through the staged row below
                IP::assemble(z, th, trc, trv, L);
                if (!IP::sensitivities_robust(L, z, th, trc, trv, a.dq1 + (size_t)i * a.out_stride_dq, a.dq2 + (size_t)i * a.out_stride_dq,
                                              a.du + (size_t)i * a.out_stride_du)) st_g = ST_FAIL;
                grad = false;
            } else {
                if (!IP::sensitivities_reg(L, z, th, trc, trv)) st_g = ST_FAIL;
            }
        } else {
            L.sync();
        }
#pragma unroll
        for (int k = 0; k < NQ; ++k) ws[k] = ws[IP::ROFF_Q3 + k];          // (garbage when the eval solve failed before converging: status says so)
        L.sync();
        // ---- the packed row [q3 | ∂q3/∂q1 | ∂q3/∂q2 | ∂q3/∂u1] leaves the workspace: consecutive lanes take consecutive elements
        constexpr int NOUT = IP::NOUT, NQQ = NQ * NQ;
        double* oq = a.q3 ? a.q3 + (size_t)i * a.out_stride_q3 : nullptr;
        double* o1 = grad ? a.dq1 + (size_t)i * a.out_stride_dq - NQ : nullptr;
        double* o2 = grad ? a.dq2 + (size_t)i * a.out_stride_dq - (NQ + NQQ) : nullptr;
        double* o3 = grad ? a.du + (size_t)i * a.out_stride_du - (NQ + 2 * NQQ) : nullptr;
        if (a.packed_out && grad && oq && NOUT % 2 == 0) {
            // whole row contiguous and 16-byte aligned: consecutive lanes store consecutive 16-byte pairs (128 B per 8 lanes)
            const double2* src = reinterpret_cast<const double2*>(ws);
            double2* dst = reinterpret_cast<double2*>(oq);
#pragma unroll
            for (int t = 0; t < (NOUT / 2 + G - 1) / G; ++t) {
                const int e = g + G * t;
                if (G == 1 || e < NOUT / 2) {
                    const double2 v = src[e];
#ifdef __CUDA_ARCH__
                    if (a.n_peers > 1 && a.mc_out) {
                        // one multicast store: every rank's gather buffer (this rank's too) receives the 16 bytes
                        const size_t off = ((size_t)(a.gather_row0 + i) * a.gather_width) / 2 + e;
                        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<double2*>(a.mc_out) + off),
                                     "f"(__int_as_float(__double2loint(v.x))), "f"(__int_as_float(__double2hiint(v.x))),
                                     "f"(__int_as_float(__double2loint(v.y))), "f"(__int_as_float(__double2hiint(v.y))) : "memory");
                    } else {
                        dst[e] = v;
                        if (a.n_peers > 1) {
                            const size_t off = ((size_t)(a.gather_row0 + i) * a.gather_width) / 2 + e;
                            for (int p = 0; p < a.n_peers; ++p) if (p != a.self_rank) reinterpret_cast<double2*>(a.peer_out[p])[off] = v;
                        }
                    }
#else
                    dst[e] = v;
#endif
                }
            }
        } else {
#pragma unroll
        for (int t = 0; t < (NOUT + G - 1) / G; ++t) {
            const int e = g + G * t;
            if (G == 1 || e < NOUT) {
                double* dst = (e < NQ) ? oq : (e < NQ + NQQ) ? o1 : (e < NQ + 2 * NQQ) ? o2 : o3;
                if (dst) {
                    const double v = ws[e];
                    dst[e] = v;
#ifdef __CUDA_ARCH__
                    if (a.n_peers > 1) {
                        const size_t off = (size_t)(a.gather_row0 + i) * a.gather_width + e;
                        for (int p = 0; p < a.n_peers; ++p) if (p != a.self_rank) a.peer_out[p][off] = v;
                    }
#endif
                }
            }
        }
        }
#ifdef __CUDA_ARCH__
        if constexpr (M::ROBUST_IFT) {
            // fused all-gather: the rank-revealing IFT wrote the Jacobian blocks of the LOCAL row itself (lane 0, straight to global
            // memory) and only q3 went through the staged row above — forward the rest of the row to the peers from this rank's own
            // gather buffer, as the shared-memory-LU path does
            if (a.n_peers > 1 && a.want_grad && a.dq1) {
                L.sync();                                          // lane 0's stores to the local row are visible to the group
                const size_t off = (size_t)(a.gather_row0 + i) * a.gather_width;
                const double* row = a.peer_out[a.self_rank] + off;
                for (int k = NQ + g; k < a.gather_width; k += G) {
                    const double v = __ldcg(row + k);
                    for (int p = 0; p < a.n_peers; ++p) if (p != a.self_rank) a.peer_out[p][off + k] = v;
                }
            }
        }
#endif
    } else {
        if (a.want_grad && a.dq1) {
            IP::load_z(L, z);
            M::trig_var(z.q, th, trv);
            double* o1 = a.dq1 + (size_t)i * a.out_stride_dq; double* o2 = a.dq2 + (size_t)i * a.out_stride_dq; double* o3 = a.du + (size_t)i * a.out_stride_du;
            if (M::ROBUST_IFT) {
                IP::assemble(z, th, trc, trv, L);
                if (!IP::sensitivities_robust(L, z, th, trc, trv, o1, o2, o3)) st_g = ST_FAIL;
            } else {
                IP::assemble(z, th, trc, trv, L); IP::factor(L);
                if (!L.ok) st_g = ST_FAIL;
                IP::sensitivities(L, z, th, trc, trv, o1, o2, o3);
            }
        }
#ifdef __CUDA_ARCH__
        if (a.n_peers > 1) {
            L.sync();                                              // every lane's share of the row has been written
            const size_t off = (size_t)(a.gather_row0 + i) * a.gather_width;
            const double* row = a.peer_out[a.self_rank] + off;
            for (int k = g; k < a.gather_width; k += G) {
                const double v = __ldcg(row + k);
                for (int p = 0; p < a.n_peers; ++p) if (p != a.self_rank) a.peer_out[p][off + k] = v;
            }
        }
#endif
    }
    if (g == 0) {
        if (a.status) a.status[i] = st_e | (st_g << 4);
        if (a.iters) a.iters[i] = it_e | (it_g << 16);
    }
}

// BLOCK = G·PPB threads; dynamic shared memory = PPB × ContactIP::WS doubles.
// OD_MIN_BLOCKS (A/B switch): minimum resident blocks per SM asked of ptxas — caps the registers per thread (16 → 128) to trade
// spills for occupancy at saturating batches.
#ifndef OD_MIN_BLOCKS
#define OD_MIN_BLOCKS 1
#endif
template <class M, int G, int PPB, bool REG, bool BSYNC = false>
__global__ void __launch_bounds__(G * PPB, OD_MIN_BLOCKS) contact_step_kernel(const StepArgs a) {
    extern __shared__ __align__(16) double od_smem[];
    const int slot = threadIdx.x / G, g = threadIdx.x % G;
    static_assert((G * PPB) % 32 == 0, "whole warps: the step runs warp-synchronously");
    int i = blockIdx.x * PPB + slot;
    if (i >= a.B) i = a.B - 1;               // padding lanes of the last warp repeat the last problem (identical values, same addresses)
    static_assert(!BSYNC || !OD_INPLACE_Z, "block-phased execution is written for the two-copy iterate update");
#if OD_TMA_INPUT
    // Experiment for the north_star's "TMA staging" clause (A/B switch, off in the shipped build): the block's packed
    // input rows (PPB × 80 B for the hopper) arrive in shared memory as ONE bulk asynchronous copy (cp.async.bulk, the 1-D TMA path,
    // completion on an mbarrier) instead of per-lane global loads; the solver then reads its row from shared memory.
    if (a.in_packed && !a.eta && (reinterpret_cast<uintptr_t>(a.q1) & 15) == 0) {
        constexpr int NIN = 2 * M::NQ + M::NU;
        static_assert((NIN * 8 * PPB) % 16 == 0, "bulk copies move multiples of 16 bytes");
        __shared__ __align__(16) double tile[PPB * NIN];
        __shared__ __align__(8) unsigned long long mbar;
        const int row0 = blockIdx.x * PPB;
        const int rows = (a.B - row0 < PPB) ? a.B - row0 : PPB;
        const unsigned bytes = (unsigned)(rows * NIN * 8);
        const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar), dst = (unsigned)__cvta_generic_to_shared(tile);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(a.q1 + (size_t)row0 * NIN), "r"(bytes), "r"(mb) : "memory");
        }
        unsigned done = 0;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb) : "memory");
        StepArgs b = a;
        b.q1 = tile - (size_t)row0 * NIN; b.q2 = b.q1 + M::NQ; b.u = b.q1 + 2 * M::NQ;       // row i of the batch = row i − row0 of the tile
        contact_step_one<M, G, PPB, REG, BSYNC>(b, i, od_smem + slot * ContactIP<M, G, PPB, REG>::WS_SLOT, g, 0xffffffffu);
    } else
#endif
    contact_step_one<M, G, PPB, REG, BSYNC>(a, i, od_smem + slot * ContactIP<M, G, PPB, REG>::WS_SLOT, g, 0xffffffffu);
    if (a.sync_counter) {                                   // fused cross-GPU barrier (see StepArgs)
        // The block barrier orders every thread's peer / multicast stores before thread 0's system-scope fence, which is cumulative
        // (the pattern of cooperative-groups grid sync: bar.sync, then ONE thread fences and signals) — one fence per block, not per thread.
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned prev = atomicAdd(a.sync_counter, 1u);
            if (prev == gridDim.x - 1) {                    // last block of this rank: every row of the shard is on its way / there
                *a.sync_counter = 0u;                       // ready for the next launch (stream-ordered)
                const unsigned long long epoch = a.sync_epoch ? a.sync_epoch : *a.sync_epoch_dev + 1ull;
                // ONE system-scope fence, then the flags go out as relaxed stores that travel concurrently (fence + relaxed store =
                // release; a st.release per peer is a fence per peer: each waits for the previous flag's round trip — at 8 ranks the
                // seven sequential releases cost more than the rest of the barrier) — or as a single multicast store.
                __threadfence_system();
                if (a.mc_flags) {
                    asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a.mc_flags + a.self_rank), "l"(epoch) : "memory");
                } else {
                    for (int p = 0; p < a.n_peers; ++p)
                        if (p != a.self_rank) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a.sync_flags[p] + a.self_rank), "l"(epoch) : "memory");
                }
                for (int p = 0; p < a.n_peers; ++p) {
                    if (p == a.self_rank) continue;
                    unsigned long long v;
                    do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.sync_flags[a.self_rank] + p) : "memory"); } while (v < epoch);
                }
                if (!a.sync_epoch) *a.sync_epoch_dev = epoch;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent, block-phased sweep for the models whose loop does not fit the instruction cache (planar push: 108 KB).
//
// Measured on B200 (profiles/r02e_*, r02g_*): with one-warp blocks, 8 unsynchronised warps per SM each stream the loop from L2 —
// 6 stall cycles per issued instruction waiting for instructions; phasing the warps of a 256-thread block with barriers removes
// those stalls, but a block that ends with its slowest problem drains its SM (and 0.5 % of the planar-push batch runs 100
// iterations).  Here both are kept: one resident block per SM, all its warps walk the state machine in phase (the two votes go
// through __syncthreads_or), and a GROUP whose problem has finished claims the next problem index from an atomic queue at the top
// of the next pass — a straggler then holds one group slot, not a block.  The IFT is not run inside the loop (it would serialise
// every pass behind whichever warp happens to differentiate): the sweep kernel leaves q3, the status words and the iterate at the
// gradient tolerance (z_snapshots), and contact_ift_kernel — naturally in phase, every problem needs exactly one pass —
// differentiates all problems afterwards.  Per problem the arithmetic is that of contact_step_one, operation for operation: the
// results are bit-identical (tests/test_gpu_parity.py).
template <class M, int G, int PPB>
__device__ __forceinline__ void contact_sweep_body(const StepArgs& a) {
#ifdef __CUDA_ARCH__
    typedef ContactIP<M, G, PPB, true> IP;
    constexpr int NQ = M::NQ;
    static_assert((G * PPB) % 32 == 0 && !OD_INPLACE_Z, "whole warps; two-copy iterate update");
    extern __shared__ __align__(16) double od_smem[];
    const int slot = threadIdx.x / G, g = threadIdx.x % G, lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    double* ws = od_smem + slot * IP::WS_SLOT;
    double th[M::NTH];
    typename IP::Z z_r, D_r, zc_r;
    typename IP::R rc_r;
    typename IP::Z& zc = IP::shared_obj(ws, 0, zc_r);
    typename IP::R& rc = IP::shared_obj(ws, 1, rc_r);
    typename IP::Z& D = IP::shared_obj(ws, 2, D_r);
    typename IP::Z& z = IP::shared_obj(ws, 3, z_r);
    typename IP::Lin L;
    L.ws = ws; L.g = g; L.gmask = full; L.ok = true;
    double trc[IP::NTC1], trv[IP::NTV1];
    double r_vio = 0.0, k_vio = 0.0, alpha = 0.0;
    bool first = true, eval_done = true, grad_done = true, active = false, have = false;
    int it = 0, ls = 0, it_e = 0, it_g = 0, st_e = 0, st_g = 0, i = 0;
    {
        double q0[NQ];
#pragma unroll
        for (int k = 0; k < NQ; ++k) q0[k] = 0.0;
        IP::init_z(q0, z);
#pragma unroll
        for (int k = 0; k < M::NTH; ++k) th[k] = 1.0;
        M::trig_const(th, trc);
        D = z; zc = z;
    }
    // the finished problem's status words; its q3 and gradient snapshot have been written when they were taken
    auto finish = [&]() {
        if (g == 0) {
            if (a.status) a.status[i] = st_e | (st_g << 4);
            if (a.iters) a.iters[i] = it_e | (it_g << 16);
        }
        active = false; have = false;
    };
    unsigned int* const queue = a.resume ? a.work_queue + 2 : a.work_queue;
    // resume: the parked problems of the previous launch (their count is final)
    const int bound = a.resume ? (int)*reinterpret_cast<volatile unsigned int*>(a.work_queue + 1) : a.B;
    for (;;) {
        // ---- groups without a problem claim the next index
        int claim = -1;
        if (!have && g == 0) claim = (int)atomicAdd(queue, 1u);
        claim = __shfl_sync(full, claim, lane & ~(G - 1));
        if (!have && claim >= 0 && claim < bound) {
            i = a.resume ? a.park_list[claim] : claim;
            double q2v[NQ];
            load_problem_plain<M>(a, i, th, q2v);
            IP::init_z(q2v, z);
            M::trig_const(th, trc);
            alpha = 0.0; r_vio = 0.0; k_vio = 0.0;
            first = true; eval_done = !a.want_eval; grad_done = !a.want_grad; active = true; have = true;
            it = 0; ls = 0; it_e = 0; it_g = 0; st_e = 0; st_g = 0;
            if (a.resume) {
                // the parked iterate re-enters as a "first" candidate (z − 0·Δ: its residual is evaluated, it is accepted without a
                // line-search test and without counting an iteration) with the progress words of the first launch
                IP::template unpack_z<1>(a.z_park + (size_t)i * IP::NZ, z);
                const int sw = a.park_info[2 * i], iw = a.park_info[2 * i + 1];
                st_e = sw & 15; st_g = (sw >> 4) & 15; it_e = iw & 0xffff; it_g = (iw >> 16) & 0xffff;
                eval_done = st_e != ST_PEND; grad_done = st_g != ST_PEND;
                it = eval_done ? it_g : it_e;
            }
            D = z;
        }
        // ---- candidate z − αΔ and its residual
        double rv2 = 0.0, kv2 = 0.0;
        bool retry = false;
        if (warp_any(active)) {
            IP::candidate(z, D, alpha, zc);
            M::trig_var(zc.q, th, trv);
            IP::residual(zc, th, trc, trv, rc, rv2, kv2);
            retry = active && !(first || rv2 <= r_vio || kv2 <= k_vio || ls >= a.opts.max_ls);
            if (retry) { alpha *= a.opts.ls_scale; ++ls; }
        }
        if (block_or(retry)) continue;
        // ---- accepted iterate: termination tests (a group's lanes agree on every one of these; the warp's groups may not: the
        // snapshot's group barriers are warp barriers, so it is taken by whole warps with the other groups masked)
        bool take_snapshot = false, done_now = false, park = false;
        if (active) {
            z = zc; r_vio = rv2; k_vio = kv2;
            if (!first) ++it;
            first = false;
            const bool bad = !IP::finite(z, r_vio, k_vio);
            const bool capped = it >= a.opts.max_iter;
            const bool rok = r_vio < a.opts.r_tol;
            const bool conv_e = rok && (k_vio < a.opts.kappa_eval_tol);
            const bool conv_g = rok && (k_vio < a.opts.kappa_grad_tol);
            if (!eval_done && (conv_e || capped || bad)) {
                eval_done = true; it_e = it; st_e = bad ? ST_FAIL : (conv_e ? ST_OK : ST_MAXIT);
                if (a.q3 && g == 0) {
                    double* o = a.q3 + (size_t)i * a.out_stride_q3;
#pragma unroll
                    for (int k = 0; k < NQ; ++k) o[k] = z.q[k];
                }
            }
            if (!grad_done && (conv_g || capped || bad)) {
                grad_done = true; it_g = it; st_g = bad ? ST_FAIL : (conv_g ? ST_OK : ST_MAXIT);
                take_snapshot = true;
            }
            done_now = bad || capped || (eval_done && grad_done);
            park = !done_now && a.park_iter > 0 && it >= a.park_iter;
        }
        if (warp_any(take_snapshot || park)) {
            IP::store_z(L, z);                                  // (groups that are not taking one rewrite their own workspace: harmless)
            L.sync();
            if (take_snapshot) {
                double* dst = a.z_snapshots + (size_t)i * IP::NZ;
                for (int e = g; e < IP::NZ; e += G) dst[e] = ws[IP::ROFF_ZS + e];
            }
            if (park) {
                double* dst = a.z_park + (size_t)i * IP::NZ;
                for (int e = g; e < IP::NZ; e += G) dst[e] = ws[IP::ROFF_ZS + e];
                if (g == 0) {
                    a.park_info[2 * i] = (eval_done ? st_e : ST_PEND) | ((grad_done ? st_g : ST_PEND) << 4);
                    a.park_info[2 * i + 1] = (eval_done ? it_e : it) | ((grad_done ? it_g : it) << 16);
                    a.park_list[atomicAdd(a.work_queue + 1, 1u)] = i;
                }
            }
            L.sync();
        }
        if (done_now) finish();
        if (park) { active = false; have = false; }
        // ---- anything left for this block?  (a group without a problem will claim one at the top if the queue still has any)
        bool more = false;
        if (!have && g == 0) more = *reinterpret_cast<volatile unsigned int*>(queue) < (unsigned)bound;
        if (!block_or(active || more)) break;
        if (!warp_any(active)) continue;
        // ---- Newton system at z, direction, step length
        IP::linearize(z, th, trc, trv, rc, L);
        bool failed = false;
        if (active && !L.ok) {
            if (!eval_done) { st_e = ST_FAIL; it_e = it; eval_done = true; }
            if (!grad_done) { st_g = ST_FAIL; it_g = it; grad_done = true; failed = true; }
            active = false;
        }
        IP::direction(L, z, rc, r_vio, k_vio, D, alpha);
        if (!active) alpha = 0.0;
        ls = 0;
        if (warp_any(failed)) {                                  // singular Newton system: the IFT is still attempted at this iterate
            IP::store_z(L, z);
            L.sync();
            if (failed) {
                double* dst = a.z_snapshots + (size_t)i * IP::NZ;
                for (int e = g; e < IP::NZ; e += G) dst[e] = ws[IP::ROFF_ZS + e];
            }
            L.sync();
        }
        if (have && !active) { L.ok = true; finish(); }
    }
#endif
}

template <class M, int G, int PPB>
__global__ void __launch_bounds__(G * PPB) contact_sweep_kernel(const StepArgs a) { contact_sweep_body<M, G, PPB>(a); }

// IFT at the snapshots of contact_sweep_kernel: one group per problem.  Writes the Jacobian blocks and folds a failed factorisation
// into the gradient nibble of the status word.  MODE 0: problems block·PPB + slot of an ordinary grid.  MODE 1: the same, but a
// PARKED problem (park_info word non-zero; its snapshot and status are still being produced by the resume launch running beside this
// one) is only walked through — the group must keep its warp's barriers — at the initial iterate, and its status is left alone; its
// rows are written again, from the real snapshot, by the MODE 2 launch that follows: the parked list, any grid, whole blocks striding.
template <class M, int G, int PPB, int MODE>
__device__ __forceinline__ void contact_ift_body(const StepArgs& a, const int block, const int nblocks) {
#ifdef __CUDA_ARCH__
    typedef ContactIP<M, G, PPB, true> IP;
    constexpr int NQ = M::NQ;
    static_assert(M::ROBUST_IFT, "the persistent sweep is instantiated for the models with the rank-revealing IFT");
    extern __shared__ __align__(16) double od_smem[];
    const int slot = threadIdx.x / G, g = threadIdx.x % G;
    double* ws = od_smem + slot * IP::WS_SLOT;
    typename IP::Lin L;
    L.ws = ws; L.g = g; L.gmask = 0xffffffffu; L.ok = true;
    const int count = (MODE == 2) ? (int)*reinterpret_cast<volatile unsigned int*>(a.work_queue + 1) : a.B;
    for (int c0 = block * PPB; c0 < count; c0 += nblocks * PPB) {          // (MODE 0 / 1: one trip)
        int c = c0 + slot;
        if (c >= count) c = count - 1;
        const int i = (MODE == 2) ? a.park_list[c] : c;
        const bool dry = (MODE == 1) && a.park_info[2 * i] != 0;
        double th[M::NTH], q2v[NQ], trc[IP::NTC1], trv[IP::NTV1];
        load_problem_plain<M>(a, i, th, q2v);
        M::trig_const(th, trc);
        const double* src = a.z_snapshots + (size_t)i * IP::NZ;
        typename IP::Z z;
        if (dry) {
            IP::init_z(q2v, z);
            IP::store_z(L, z);
        } else {
            for (int e = g; e < IP::NZ; e += G) ws[IP::ROFF_ZS + e] = src[e];
        }
        L.sync();
        IP::load_z(L, z);
        M::trig_var(z.q, th, trv);
        IP::assemble(z, th, trc, trv, L);
        const bool ok = IP::sensitivities_robust(L, z, th, trc, trv, a.dq1 + (size_t)i * a.out_stride_dq, a.dq2 + (size_t)i * a.out_stride_dq,
                                                 a.du + (size_t)i * a.out_stride_du);
        if (!ok && !dry && g == 0 && a.status) a.status[i] = (a.status[i] & 15) | (ST_FAIL << 4);
        L.sync();
    }
#endif
}
template <class M, int G, int PPB, int MODE>
__global__ void __launch_bounds__(G * PPB) contact_ift_kernel(const StepArgs a) { contact_ift_body<M, G, PPB, MODE>(a, blockIdx.x, gridDim.x); }

// A rank whose shard is empty still has to take part in the fused cross-GPU barrier: publish the epoch, wait for the peers.
static __global__ void gather_sync_only_kernel(const StepArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned long long epoch = a.sync_epoch ? a.sync_epoch : *a.sync_epoch_dev + 1ull;
    __threadfence_system();
    for (int p = 0; p < a.n_peers; ++p)
        if (p != a.self_rank) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a.sync_flags[p] + a.self_rank), "l"(epoch) : "memory");
    for (int p = 0; p < a.n_peers; ++p) {
        if (p == a.self_rank) continue;
        unsigned long long v;
        do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.sync_flags[a.self_rank] + p) : "memory"); } while (v < epoch);
    }
    if (!a.sync_epoch) *a.sync_epoch_dev = epoch;
}

// ---------------------------------------------------------------------------------------------------------------------
// Batched closed-loop rollouts: the caller of f in the reference's outer solver — iLQR.rollout(model, x1, ū) and the forward
// pass / Armijo line search of IterativeLQR (reference examples/cartpole.jl:79,86; the package itself is external).  Time is
// sequential, rollouts (line-search candidates α_r, or independent rollouts) are parallel: ONE launch instead of T−1, the state
// x_t = [q1; q2] never leaves the device.  Step t of rollout r:
//     u_t = ū_t + α_r k_t + K_t (x_t − x̄_t)          (k, K, x̄ may be null: open-loop rollout)
//     x_{t+1} = f(x_t, u_t) = [q2; q3],  q3 = step!(eval_sim, q2, (q2 − q1)/h, u_t)        reference src/dynamics.jl:81-94
// X (R × T × 2NQ) and U (R × (T−1) × NU) are outputs and also the working storage: every step is the same contact_step_one as
// the batched entry points, reading row t of X/U and writing q3 into row t+1.
struct RolloutArgs {
    int R, T;
    const double* x1;                      // R × 2NQ
    const double* ubar; long long ubar_stride;   // (T−1) × NU nominal controls; stride between rollouts in doubles (0 = shared)
    const double* xbar;                    // T × 2NQ nominal states, or null
    const double* K;                       // (T−1) × NU × 2NQ feedback gains, row-major [t][u component][x component], or null
    const double* kff;                     // (T−1) × NU feed-forward terms, or null
    const double* alpha;                   // R step sizes, or null (α = 1)
    double* X; double* U;
    int* status; int* iters;               // R × (T−1), may be null
    double h; double fric[4];
    SolverOpts opts;
};

template <class M, int G, int PPB, bool REG>
OD_HD void contact_rollout_one(const RolloutArgs& ra, const int r, double* ws, const int g, const unsigned gmask) {
    constexpr int NQ = M::NQ, NU = M::NU, NX = 2 * M::NQ;
    const int T = ra.T;
    double* Xr = ra.X + (size_t)r * T * NX;
    double* Ur = ra.U + (size_t)r * (T - 1) * NU;
    StepArgs a;
    a.B = T - 1; a.q1 = Xr; a.q2 = Xr + NQ; a.u = Ur; a.in_stride_q = NX; a.in_stride_u = NU;
    a.q3 = Xr + NX + NQ; a.dq1 = nullptr; a.dq2 = nullptr; a.du = nullptr;
    a.out_stride_q3 = NX; a.out_stride_dq = 0; a.out_stride_du = 0;
    a.status = ra.status ? ra.status + (size_t)r * (T - 1) : nullptr;
    a.iters = ra.iters ? ra.iters + (size_t)r * (T - 1) : nullptr;
    a.h = ra.h;
#pragma unroll
    for (int k = 0; k < 4; ++k) a.fric[k] = ra.fric[k];
    a.want_eval = 1; a.want_grad = 0; a.eta = nullptr; a.n_eta = 0; a.eta_i0 = 0;
    a.n_peers = 0; a.self_rank = 0; a.gather_row0 = 0; a.gather_width = 0;
    a.packed_out = 0; a.in_packed = 0; a.in_vel = 0; a.sync_counter = nullptr; a.sync_epoch = 0; a.sync_epoch_dev = nullptr; a.mc_out = nullptr; a.mc_flags = nullptr; a.work_queue = nullptr; a.z_snapshots = nullptr; a.park_iter = 0; a.resume = 0; a.z_park = nullptr; a.park_list = nullptr; a.park_info = nullptr; a.side_stream = nullptr; a.ev_fork = nullptr; a.ev_join = nullptr;
    a.opts = ra.opts;
    const double alpha = ra.alpha ? ra.alpha[r] : 1.0;
    const double* ub = ra.ubar + (size_t)r * ra.ubar_stride;
    for (int e = g; e < NX; e += G) Xr[e] = ra.x1[(size_t)r * NX + e];
    for (int t = 0; t < T - 1; ++t) {
#ifdef __CUDA_ARCH__
        __syncwarp(gmask);                                     // x_t (written by other lanes) is visible
#else
        host_team_sync();
#endif
        for (int e = g; e < NU; e += G) {
            double u = ub[(size_t)t * NU + e];
            if (ra.kff) u += alpha * ra.kff[(size_t)t * NU + e];
            if (ra.K) {
                const double* Kr = ra.K + ((size_t)t * NU + e) * NX;
                double acc = 0.0;
                for (int j = 0; j < NX; ++j) acc += Kr[j] * (Xr[(size_t)t * NX + j] - (ra.xbar ? ra.xbar[(size_t)t * NX + j] : 0.0));
                u += acc;
            }
            Ur[(size_t)t * NU + e] = u;
        }
        for (int e = g; e < NQ; e += G) Xr[(size_t)(t + 1) * NX + e] = Xr[(size_t)t * NX + NQ + e];
#ifdef __CUDA_ARCH__
        __syncwarp(gmask);
#else
        host_team_sync();
#endif
        contact_step_one<M, G, PPB, REG>(a, t, ws, g, gmask);
    }
}

template <class M, int G, int PPB, bool REG>
__global__ void __launch_bounds__(G * PPB) contact_rollout_kernel(const RolloutArgs ra) {
    extern __shared__ __align__(16) double od_smem[];
    const int slot = threadIdx.x / G, g = threadIdx.x % G;
    int r = blockIdx.x * PPB + slot;
    if (r >= ra.R) r = ra.R - 1;             // padding lanes repeat the last rollout (identical values, same addresses)
    contact_rollout_one<M, G, PPB, REG>(ra, r, od_smem + slot * ContactIP<M, G, PPB, REG>::WS_SLOT, g, 0xffffffffu);
}

}  // namespace od
