/* optdyn_b200 — C ABI of the B200 optimization-based-dynamics hot path.
 *
 * Drop-in boundary for the inner loop the reference (thowell/optimization_dynamics) exposes to IterativeLQR:
 *     f / fx / fu                      reference src/dynamics.jl:81-128
 *     fx_gb / fu_gb (gradient bundle)  reference src/gradient_bundle.jl:87-147, src/ls.jl:44-60
 *     f/fx/fu_rocket[_proj]            reference src/models/rocket/dynamics.jl:101-269
 * Every entry point is batched: B independent (timestep × sample × rollout) problems per call.  Plain C types only; the
 * Julia `ccall` / Python `ctypes` bindings are shown in INTEGRATION.md.  All real data is IEEE fp64.
 *
 * Array conventions
 *   q1, q2 : B×nq, u : B×nu, row per problem (x = [q1; q2] of the reference, src/dynamics.jl:82-83).
 *   q3     : B×nq.
 *   dq3dq1, dq3dq2 : B×(nq×nq), dq3du1 : B×(nq×nu) — each block COLUMN-major (Julia layout), i.e. exactly
 *            grad_sim.grad.∂q3∂q1[1], ∂q3∂q2[1], ∂q3∂u1[1] of src/dynamics.jl:110-111,125.
 *   packed : in  row = [q1 | q2 | u]                         (2nq+nu doubles)
 *            out row = [q3 | ∂q3/∂q1 | ∂q3/∂q2 | ∂q3/∂u1]    (nq + nq(2nq+nu) doubles; hopper: 44 = 352 B)
 *   status : int32 per problem, 0 = both solves converged.  Low nibble = eval solve (f), next nibble = gradient solve
 *            (fx/fu): 1 = iteration cap (max_iter), 2 = non-finite iterate or singular Jacobian.  The reference ignores the
 *            solver status (src/dynamics.jl:88); callers of this ABI should not.
 *   iters  : int32 per problem, eval iterations | gradient iterations << 16 (may be NULL).
 * "_device" entry points take device pointers, enqueue on the handle's stream and return without synchronising.
 * The others take host pointers (pinned or pageable), copy H2D/D2H on the handle's stream and synchronise before returning.
 * Return value: 0 on success, non-zero on error (message via od_last_error()).  There is no CPU fallback.
 */
#ifndef OPTDYN_B200_H
#define OPTDYN_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum od_model {
    OD_ACROBOT_IMPACT = 0,        /* reference src/models/acrobot/model.jl:121-142  */
    OD_ACROBOT_NOMINAL = 1,       /* reference src/models/acrobot/model.jl:144-157  */
    OD_CARTPOLE_FRICTION = 2,     /* reference src/models/cartpole/model.jl:81-114  */
    OD_CARTPOLE_FRICTIONLESS = 3, /* reference src/models/cartpole/model.jl:116-129 */
    OD_PLANAR_PUSH = 4,           /* reference src/models/planar_push/model.jl:121-187 */
    OD_HOPPER = 5,                /* RoboDojo.hopper, reference examples/hopper.jl:14,38-50 */
    OD_ROCKET = 6                 /* reference src/models/rocket/codegen.jl:14-22,45-64 */
};

/* InteriorPointOptions as the reference sets them (src/dynamics.jl:25-33; rocket: src/models/rocket/dynamics.jl:21-27,77-86). */
typedef struct od_options {
    double r_tol;          /* 1e-8 */
    double kappa_eval_tol; /* κ_eval_tol, e.g. 1e-4 (examples/hopper.jl:42) */
    double kappa_grad_tol; /* κ_grad_tol, e.g. 1e-3 */
    double ls_scale;       /* 0.5 */
    int32_t max_iter;      /* 100 */
    int32_t max_ls;        /* 25 */
} od_options;

typedef struct od_handle od_handle;

/* Fills *opts with the settings the reference's EXAMPLES pass for `model` (κ_eval_tol 1e-4, κ_grad_tol 1e-3 / 1e-2;
 * examples/hopper.jl:42, planar_push.jl:22).  od_create(opts = NULL) instead uses the reference CONSTRUCTOR's defaults
 * (κ_eval_tol = κ_grad_tol = 1e-6, src/dynamics.jl:51-53). */
int od_default_options(int model, od_options* opts);
/* nq, nu, nz (decision variables), ntheta (data vector) of a model; any pointer may be NULL. */
int od_model_dims(int model, int* nq, int* nu, int* nz, int* ntheta);

/* ImplicitDynamics(model, h, …; r_tol, κ_eval_tol, κ_grad_tol) — reference src/dynamics.jl:51-79.
 * params: friction coefficients for OD_CARTPOLE_FRICTION [μ_slider, μ_angle] and OD_HOPPER [μ_body, μ_foot] (the reference
 * mutates model.friction after construction, examples/cartpole.jl:21), u_max for OD_ROCKET; NULL/0 = model defaults.
 * device: CUDA device ordinal.  Returns NULL on failure. */
od_handle* od_create(int model, double h, const od_options* opts, const double* params, int nparams, int device);
void od_destroy(od_handle* hd);
/* Use an existing cudaStream_t (e.g. the caller's framework stream) instead of the handle's own stream. */
int od_set_stream(od_handle* hd, void* cuda_stream);
int od_synchronize(od_handle* hd);

/* f for a batch: q3 = step!(eval_sim, q2, (q2−q1)/h, u)  — reference src/dynamics.jl:81-94.  Host pointers. */
int od_step_batch(od_handle* hd, int B, const double* q1, const double* q2, const double* u, double* q3, int32_t* status);
/* f + fx + fu for a batch with the duplicate gradient solve removed — reference src/dynamics.jl:81-128.  Host pointers;
 * q3 may be NULL (fx/fu only). */
int od_step_grad_batch(od_handle* hd, int B, const double* q1, const double* q2, const double* u,
                       double* q3, double* dq3dq1, double* dq3dq2, double* dq3du1, int32_t* status);
/* RoboDojo.step!(sim, q, v, u, t) for a batch — the call shape the reference's hopper example uses directly
 * (reference examples/hopper.jl:63,89,112,133,157; inside f / fx / fu at src/dynamics.jl:88,103,123): q: B×nq current
 * configuration, v: B×nq velocity (the data vector takes q1 = q − h·v), u: B×nu.
 * grad_sim = 0: the eval simulator (κ_eval_tol, diff_sol = false) → q3.
 * grad_sim = 1: the gradient simulator (κ_grad_tol, diff_sol = true) → q3 (may be NULL) at κ_grad_tol and
 *               grad.∂q3∂q1[1], ∂q3∂q2[1], ∂q3∂u1[1] (column-major blocks, all three required).  Host pointers. */
int od_sim_step_batch(od_handle* hd, int B, int grad_sim, const double* q, const double* v, const double* u,
                      double* q3, double* dq3dq1, double* dq3dq2, double* dq3du1, int32_t* status);
/* Same, one packed host row per problem in and out (fewest transfers: 1 H2D, 2 D2H). */
int od_step_grad_packed(od_handle* hd, int B, const double* in, double* out, int32_t* status);

/* Device-pointer variants (asynchronous).  want_eval / want_grad select f and/or fx+fu. */
int od_step_grad_batch_device(od_handle* hd, int B, const double* q1, const double* q2, const double* u,
                              double* q3, double* dq3dq1, double* dq3dq2, double* dq3du1, int32_t* status, int32_t* iters,
                              int want_eval, int want_grad);
int od_step_grad_packed_device(od_handle* hd, int B, const double* in, double* out, int32_t* status, int32_t* iters,
                               int want_eval, int want_grad);

/* Multi-GPU derivative sweep with the all-gather fused into the kernel (single node, NVLink/NVSwitch peer access).
 * Every rank solves its contiguous shard [row0, row0+B) of the global batch and stores each finished packed output row into the
 * gather buffer of EVERY rank: gather_buffers[r] is rank r's [B_total × (nq + nq(2nq+nu))] buffer as mapped in THIS process
 * (peer-mapped device pointers, e.g. from torch symmetric memory / cudaIpcOpenMemHandle); world ≤ 8.  Asynchronous; the caller
 * must run a cross-rank barrier after the kernel before reading rows produced by other ranks.  Replaces kernel + ncclAllGather
 * (the Jacobians the sequential Riccati pass needs). */
int od_step_grad_packed_gather_device(od_handle* hd, int B, const double* in, long long row0, int world, int rank,
                                      const uint64_t* gather_buffers, int32_t* status, int32_t* iters);
/* Same, with the cross-rank barrier fused into the kernel as well: flag_buffers[r] = rank r's flag array (world × uint64, zeroed
 * once, peer-mapped like the gather buffers), block_counter = a zeroed uint32 in this rank's device memory, epoch = 1, 2, 3, …
 * (the same value on every rank for the same step).  The last block of each rank to finish publishes `epoch` to every peer and
 * waits for every peer's, so when the kernel has completed on this rank's stream all rows of all ranks are in this rank's buffer:
 * no separate barrier launch.  Alternate two gather buffers between consecutive steps if a consumer of step k may still be
 * reading when step k+1 starts on a faster rank.  Register-path models only (hopper, cartpole, acrobot). */
int od_step_grad_packed_gather_sync_device(od_handle* hd, int B, const double* in, long long row0, int world, int rank,
                                           const uint64_t* gather_buffers, const uint64_t* flag_buffers, uint32_t* block_counter,
                                           uint64_t epoch, int32_t* status, int32_t* iters);

/* Every variant of the fused gather through one descriptor (the two entry points above are special cases):
 *   multicast_buffer : NVLink multicast alias of the SAME gather buffers (e.g. torch symmetric memory's multicast_ptr), or 0.
 *                      When set, each finished 16 bytes leave as ONE multimem.st that NVSwitch replicates into every rank's buffer
 *                      (this rank's included) instead of world−1 peer stores.  Register-path models only.
 *   flag_buffers     : NULL = no fused barrier (caller synchronises the ranks after the kernel).
 *   epoch / epoch_dev: epoch >= 1 = the host supplies the step's epoch;  epoch == 0 = the epoch lives in *epoch_dev (a zeroed uint64
 *                      in this rank's device memory), advanced by one per launch on the device — such launches can sit in a CUDA
 *                      graph and be replayed (every rank must replay the same launches).
 *   An empty shard (B == 0) with the fused barrier still publishes / waits, so ragged splits and B_total < world do not hang.
 *   Planar push (rank-revealing IFT; peer stores without the fused barrier only): shards of at least OD_PERSIST (default 3072)
 *   problems run the persistent sweep into this rank's own rows, then ONE forwarding kernel stores those rows into every peer's
 *   buffer; smaller shards forward each row from the step kernel.  Either way the caller's barrier follows. */
typedef struct od_gather_desc {
    int32_t world, rank;
    int64_t row0;
    const uint64_t* gather_buffers;
    uint64_t multicast_buffer;
    const uint64_t* flag_buffers;
    uint32_t* block_counter;
    uint64_t* epoch_dev;
    uint64_t epoch;
    uint64_t multicast_flags;   /* multicast alias of the flag arrays, or 0: the epoch is then published by one multimem.st */
} od_gather_desc;
int od_step_grad_packed_gather_ex_device(od_handle* hd, int B, const double* in, const od_gather_desc* gather, int32_t* status, int32_t* iters);

/* Batched closed-loop rollouts — the caller of f in the outer solver: iLQR.rollout(model, x1, ū) (reference examples/cartpole.jl:79,
 * acrobot.jl:92, planar_push.jl:113, hopper.jl:272) and the forward pass / Armijo line search of IterativeLQR (step sizes down to
 * 1e-5, examples/cartpole.jl:86).  R rollouts of T knot points in ONE launch; time is sequential inside the kernel:
 *     u_t = ū_t + α_r k_t + K_t (x_t − x̄_t),     x_{t+1} = f(x_t, u_t) = [q2; q3]        (src/dynamics.jl:81-94)
 * x1: R×2nq initial states [q1; q2].  ubar: (T−1)×nu shared by all rollouts (ubar_per_rollout = 0) or R×(T−1)×nu.
 * xbar: T×2nq, K: (T−1)×nu×2nq row-major [t][control][state], kff: (T−1)×nu, alpha: R — each may be NULL (K needs xbar;
 * alpha NULL = 1): all NULL is the open-loop rollout.  X: R×T×2nq, U: R×(T−1)×nu (U may be NULL in the host variant),
 * status: R×(T−1) per-step solver status (0 = converged; may be NULL).  T = 1 (no steps) returns X = x1. */
int od_rollout_batch(od_handle* hd, int R, int T, const double* x1, const double* ubar, int ubar_per_rollout, const double* xbar,
                     const double* K, const double* kff, const double* alpha, double* X, double* U, int32_t* status);
/* Device pointers, asynchronous; ubar_stride = doubles between the controls of consecutive rollouts (0 = shared). */
int od_rollout_batch_device(od_handle* hd, int R, int T, const double* x1, const double* ubar, long long ubar_stride,
                            const double* xbar, const double* K, const double* kff, const double* alpha, double* X, double* U,
                            int32_t* status, int32_t* iters);

/* Batched Riccati backward pass — the consumer of fx / fu in the outer solver (IterativeLQR's backward pass inside iLQR.solve!,
 * reference examples/hopper.jl:292; sequential in t, one warp per trajectory, NT trajectories per launch).
 * jac: NT×(T−1) packed output rows of od_step_grad_packed (fx = [0 I; ∂q3/∂q1 ∂q3/∂q2], fu = [0; ∂q3/∂u1], src/dynamics.jl:105-125);
 * cost expansion along the trajectory, n = 2nq, m = nu: lx NT×T×n, lu NT×(T−1)×m, lxx NT×T×n×n, luu NT×(T−1)×m×m,
 * lux NT×(T−1)×m×n (NULL = 0); reg is added to the diagonal of Quu.
 * Out: K NT×(T−1)×m×n row-major [t][control][state] and k NT×(T−1)×m — the layout od_rollout_batch takes — dV NT×2
 * (Σ kᵀQu, ½ Σ kᵀQuu k; may be NULL), status NT (1 = a Quu was not positive definite; that step's gains are zero; may be NULL). */
int od_riccati_batch(od_handle* hd, int NT, int T, const double* jac, const double* lx, const double* lu, const double* lxx,
                     const double* luu, const double* lux, double reg, double* K, double* k, double* dV, int32_t* status);
int od_riccati_batch_device(od_handle* hd, int NT, int T, const double* jac, const double* lx, const double* lu, const double* lxx,
                            const double* luu, const double* lux, double reg, double* K, double* k, double* dV, int32_t* status);

/* Gradient bundle — gradient!(eval_sim, gb, q1, q2, u1), reference src/gradient_bundle.jl:87-104: one nominal and N perturbed
 * eval-sim steps per problem ((N+1)·B solves in one launch), then the least-squares fit of src/ls.jl:44-60 in closed form
 * (normal equations).  eta: N×(2nq+nu) perturbations shared by the batch (host).  dz: B×(nq×(2nq+nu)) column-major (host).
 * status: OR of the statuses of the N+1 solves; 8 = singular normal equations (a coordinate never perturbed). */
int od_bundle_batch(od_handle* hd, int B, int N, const double* eta, const double* q1, const double* q2, const double* u,
                    double* dz, int32_t* status);

/* The same in pieces, device pointers, asynchronous — for device-resident callers and for sharding the sample axis over GPUs
 * (SURVEY.md §8e): od_bundle_prepare inverts Σηηᵀ on the host (Hinv: ncol×ncol, ncol = 2nq+nu ≤ 16; non-zero = singular);
 * od_bundle_solve_device runs the eval-sim steps p ∈ [p0, p0+P) of the flattened B×(N+1) axis (p = b·(N+1) + k, k = 0 nominal,
 * k ≥ 1 perturbation k−1) and writes feta[p] (nq doubles) and st_work[p]; od_bundle_fit_device fits all B problems from the
 * complete feta / st_work (e.g. after an all-gather of the slices).  eta: N×ncol and Hinv in device memory; stride_q / stride_u =
 * doubles between consecutive rows of q1, q2 / u (0 = dense; 2nq+nu for packed rows [q1 | q2 | u]). */
int od_bundle_prepare(int ncol, int N, const double* eta_host, double* Hinv_host);
int od_bundle_solve_device(od_handle* hd, int B, int N, const double* eta, const double* q1, const double* q2, const double* u,
                           int stride_q, int stride_u, long long p0, long long P, double* feta, int32_t* st_work);
int od_bundle_fit_device(od_handle* hd, int B, int N, const double* eta, const double* Hinv, const double* feta, const int32_t* st_work,
                         double* dz, int32_t* status);

/* Rocket: f/fx/fu_rocket (proj = 0) and f/fx/fu_rocket_proj (proj = 1) — reference src/models/rocket/dynamics.jl:101-269.
 * x: B×12, u: B×3 → y: B×12, dx: B×(12×12), du: B×(12×3) column-major; dx/du may be NULL (f only).  Host pointers.
 * status low nibble = dynamics solve, next nibble = projection solve. */
int od_rocket_batch(od_handle* hd, int B, const double* x, const double* u, int proj, double* y, double* dx, double* du, int32_t* status);
int od_rocket_batch_device(od_handle* hd, int B, const double* x, const double* u, int proj, double* y, double* dx, double* du,
                           int32_t* status, int32_t* iters);
/* soc_projection / soc_projection_gradient alone — reference src/models/rocket/dynamics.jl:168-210.  Host pointers. */
int od_rocket_projection_batch(od_handle* hd, int B, const double* u, double* u_proj, double* du_proj, int32_t* status);

/* Number of kernels launched through this handle so far (bench.py's gpu_launches). */
int64_t od_launch_count(const od_handle* hd);
const char* od_last_error(void);
const char* od_version(void);

#ifdef __cplusplus
}
#endif
#endif
