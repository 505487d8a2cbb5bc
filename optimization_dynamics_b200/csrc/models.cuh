// Model traits binding the generated device code (csrc/gen/, produced by tools/codegen/gen_models.py) to the solvers.
// Index sets restate the reference's IndicesOptimization constructions:
//   acrobot   reference src/models/acrobot/simulator_impact.jl:16-32       cartpole  reference src/models/cartpole/simulator_friction.jl:18-34
//   planar push  reference src/models/planar_push/simulator.jl:16-50       rocket projection  reference src/models/rocket/dynamics.jl:52-63
#pragma once
#include "contact_ip.cuh"
#include "gen/model_hopper.cuh"
#include "gen/model_acrobot_impact.cuh"
#include "gen/model_acrobot_nominal.cuh"
#include "gen/model_cartpole_friction.cuh"
#include "gen/model_cartpole_frictionless.cuh"
#include "gen/model_planar_push.cuh"
#include "gen/model_rocket.cuh"
#include "gen/model_rocket_proj.cuh"

namespace od {

#define OD_CONTACT_MODEL(NAME, NS, NFRIC, OFF_EXPR, DIM_EXPR, ROBUST)                                                                     \
    struct NAME {                                                                                                                \
        static constexpr int NQ = NS::NQ, NU = NS::NU, NC = NS::NC, NP = NS::NP, NB = NS::NB, NTH = NS::NTH, NF = NFRIC;           \
        static constexpr int NTC = NS::NTC, NTV = NS::NTV;   /* sin/cos table sizes: θ-only and q-dependent arguments */          \
        static constexpr bool ROBUST_IFT = ROBUST;   /* redundant contact constraints: rank-revealing IFT (contact_ip.cuh) */  \
        __host__ __device__ static constexpr int cone_off(int k) { return OFF_EXPR; }                                             \
        __host__ __device__ static constexpr int cone_dim(int k) { return DIM_EXPR; }                                             \
        OD_HD static void trig_const(const double* th, double* trc) { NS::trig_const(th, trc); }                                  \
        OD_HD static void trig_var(const double* q, const double* th, double* trv) { NS::trig_var(q, th, trv); }                  \
        OD_HD static void eq(const double* q, const double* g, const double* b, const double* th, const double* trc,              \
                             const double* trv, double* d, double* phi, double* psit, double* vT) {                             \
            NS::eq(q, g, b, th, trc, trv, d, phi, psit, vT); }                                                                    \
        OD_HD static void jac(const double* q, const double* g, const double* b, const double* th, const double* trc,             \
                              const double* trv, double* D, double* Eg, double* Eb, double* N, double* V, double* Mpsi) {       \
            NS::jac(q, g, b, th, trc, trv, D, Eg, Eb, N, V, Mpsi); }                                                              \
        OD_HD static void jacth(const double* q, const double* g, const double* b, const double* th, const double* trc,           \
                                const double* trv, double* Dth, double* Vth) { NS::jacth(q, g, b, th, trc, trv, Dth, Vth); }      \
    };

// cone k of the friction block: offset into b / sb and number of tangential components
OD_CONTACT_MODEL(HopperModel, gen_hopper, 2, k, 1, false)
OD_CONTACT_MODEL(AcrobotImpactModel, gen_acrobot_impact, 0, 0, 0, false)
OD_CONTACT_MODEL(AcrobotNominalModel, gen_acrobot_nominal, 0, 0, 0, false)
OD_CONTACT_MODEL(CartpoleFrictionModel, gen_cartpole_friction, 2, k, 1, false)
OD_CONTACT_MODEL(CartpoleFrictionlessModel, gen_cartpole_frictionless, 0, 0, 0, false)
OD_CONTACT_MODEL(PlanarPushModel, gen_planar_push, 0, 2 * k, (k < 4 ? 2 : 1), true)
#undef OD_CONTACT_MODEL

struct RocketDynModel {
    static constexpr int NZ = 12, NTH = 16, NTHP = 15, NEQ = 12, NORT = 0, NSOC = 0;
    OD_HD static constexpr int ort_p(int) { return 0; }
    OD_HD static constexpr int ort_d(int) { return 0; }
    OD_HD static constexpr int ortr(int) { return 0; }
    OD_HD static constexpr int soc_p(int, int) { return 0; }
    OD_HD static constexpr int soc_d(int, int) { return 0; }
    OD_HD static constexpr int socr(int, int) { return 0; }
    OD_HD static void res(const double* z, const double* th, double* r) { gen_rocket::res(z, th, r); }
    OD_HD static void jac(const double* z, const double* th, double* A) { gen_rocket::jac(z, th, A); }
    OD_HD static void jacth(const double* z, const double* th, double* A) { gen_rocket::jacth(z, th, A); }
};

// z = [u(3), p, s, w, y, v(3)];  ortz = [[5,3],[6,4]], socz = [[3,1,2],[10,8,9]] (1-based) — rocket/dynamics.jl:52-63
struct RocketProjModel {
    static constexpr int NZ = 10, NTH = 4, NTHP = 3, NEQ = 5, NORT = 2, NSOC = 1;
    OD_HD static constexpr int ort_p(int k) { return k == 0 ? 4 : 2; }
    OD_HD static constexpr int ort_d(int k) { return k == 0 ? 5 : 3; }
    OD_HD static constexpr int ortr(int k) { return 5 + k; }
    OD_HD static constexpr int soc_p(int, int e) { return e == 0 ? 2 : e - 1; }
    OD_HD static constexpr int soc_d(int, int e) { return e == 0 ? 9 : 6 + e; }
    OD_HD static constexpr int socr(int, int e) { return 7 + e; }
    OD_HD static void res(const double* z, const double* th, double* r) { gen_rocket_proj::res(z, th, r); }
    OD_HD static void jac(const double* z, const double* th, double* A) { gen_rocket_proj::jac(z, th, A); }
    OD_HD static void jacth(const double* z, const double* th, double* A) { gen_rocket_proj::jacth(z, th, A); }
};

}  // namespace od
