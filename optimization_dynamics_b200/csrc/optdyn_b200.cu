// liboptdyn_b200.so — kernels + C ABI (include/optdyn_b200.h).  sm_100a only; there is no CPU fallback.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <new>
#include "../../include/optdyn_b200.h"
#include "launch.cuh"
#include "dense_ip.cuh"
#include "rocket.cuh"
#include "riccati.cuh"

namespace od {

// ---------------------------------------------------------------------------------------------------------------------
// Gradient-bundle fit: M = argmin Σ_k ‖fη_k − fz − M η_k‖²  (reference src/gradient_bundle.jl:33-47,101-102; src/ls.jl:44-60).
// The cost is quadratic, so the reference's Newton loop converges in one step to the normal-equation solution
// M = (Σ (fη_k − fz) η_kᵀ)(Σ η_k η_kᵀ)⁻¹; the second factor depends only on the shared perturbations and is inverted once.
// ---------------------------------------------------------------------------------------------------------------------
// One block per problem, one thread per (output component i, coordinate l): G[l][i] = Σ_k (fη_k − fz)_i η_k[l] in sample order, then
// M[i][j] = Σ_l G[l][i] Hinv[l][j] (the first version ran the whole fit of a problem on one thread: 81 µs for the 50-problem sweep of
// BASELINE configs[1], five times the solves it follows).
__global__ void bundle_fit_kernel(int B, int N, int nq, int ncol, const double* __restrict__ feta /*B×(N+1)×nq*/, const double* __restrict__ eta /*N×ncol*/,
                                  const double* __restrict__ Hinv /*ncol×ncol*/, const int* __restrict__ st_in /*B×(N+1)*/, double* __restrict__ dz, int* __restrict__ st_out) {
    __shared__ double G[16 * 16];
    __shared__ int st_sh;
    const int b = blockIdx.x, t = threadIdx.x;
    if (b >= B) return;
    if (t == 0) st_sh = 0;
    __syncthreads();
    const double* f = feta + (size_t)b * (N + 1) * nq;
    int st = 0;
    for (int k = t; k <= N; k += blockDim.x) st |= st_in[(size_t)b * (N + 1) + k];
    if (st) atomicOr(&st_sh, st);
    const int i = t % nq, l = t / nq;
    if (l < ncol) {
        double g = 0.0;
        const double fz = f[i];
        for (int k = 0; k < N; ++k) g += (f[(size_t)(k + 1) * nq + i] - fz) * eta[(size_t)k * ncol + l];
        G[l * nq + i] = g;
    }
    __syncthreads();
    if (l < ncol) {
        const int j = l;
        double m = 0.0;
        for (int ll = 0; ll < ncol; ++ll) m += G[ll * nq + i] * Hinv[ll * ncol + j];
        dz[(size_t)b * nq * ncol + (size_t)j * nq + i] = m;
    }
    if (t == 0 && st_out) st_out[b] = st_sh;
}

}  // namespace od

using namespace od;

// =====================================================================================================================
// host side
// =====================================================================================================================
static thread_local char g_err[512] = "";
static int fail(const char* what, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess) snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    else snprintf(g_err, sizeof(g_err), "%s", what);
    return 1;
}
#define OD_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(#call, e_); } while (0)

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {               // page-locked, device-mapped host staging (cudaHostAlloc)
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocMapped);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct od_handle {
    int model, device;
    double h;
    od_options opts;
    double params[4];
    cudaStream_t stream; bool own_stream;
    cudaStream_t side; cudaEvent_t ev_fork, ev_join;   // planar push: second stream of the persistent sweep's tail (launch.cuh), else null
    DevBuf in, out, st, aux, aux2, sweep;      // sweep: work queue + iterate snapshots of the persistent sweep (planar push)
    PinBuf hin, hout;
    int64_t launches;
};

struct Dims { int nq, nu, nz, nth, nf; };
static bool dims_of(int model, Dims* d) {
    switch (model) {
        case OD_ACROBOT_IMPACT: *d = {2, 1, 6, 6, 0}; return true;
        case OD_ACROBOT_NOMINAL: *d = {2, 1, 2, 6, 0}; return true;
        case OD_CARTPOLE_FRICTION: *d = {2, 1, 10, 8, 2}; return true;
        case OD_CARTPOLE_FRICTIONLESS: *d = {2, 1, 2, 6, 0}; return true;
        case OD_PLANAR_PUSH: *d = {5, 2, 35, 13, 0}; return true;
        case OD_HOPPER: *d = {4, 2, 20, 13, 2}; return true;
        case OD_ROCKET: *d = {12, 3, 12, 16, 0}; return true;
    }
    return false;
}

extern "C" {

const char* od_last_error(void) { return g_err; }
const char* od_version(void) { return "optdyn_b200 0.1 (sm_100a)"; }

int od_default_options(int model, od_options* o) {
    Dims d; if (!o || !dims_of(model, &d)) return fail("od_default_options: bad model");
    o->r_tol = 1e-8; o->ls_scale = 0.5; o->max_iter = 100; o->max_ls = 25;
    o->kappa_eval_tol = 1e-4; o->kappa_grad_tol = 1e-3;                       // examples/{acrobot,cartpole,hopper}.jl
    if (model == OD_PLANAR_PUSH) o->kappa_grad_tol = 1e-2;                    // examples/planar_push.jl:21-22
    if (model == OD_ROCKET) o->kappa_grad_tol = 1e-4;                         // projection κ_tol (rocket/dynamics.jl:79)
    return 0;
}

int od_model_dims(int model, int* nq, int* nu, int* nz, int* nth) {
    Dims d; if (!dims_of(model, &d)) return fail("od_model_dims: bad model");
    if (nq) *nq = d.nq; if (nu) *nu = d.nu; if (nz) *nz = d.nz; if (nth) *nth = d.nth;
    return 0;
}

od_handle* od_create(int model, double h, const od_options* opts, const double* params, int nparams, int device) {
    Dims d;
    if (!dims_of(model, &d)) { fail("od_create: bad model"); return nullptr; }
    if (!(h > 0.0)) { fail("od_create: h must be positive"); return nullptr; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { fail("od_create: no CUDA device (this library has no CPU fallback)", e); return nullptr; }
    if (device < 0 || device >= ndev) { fail("od_create: bad device ordinal"); return nullptr; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { fail("cudaSetDevice", e); return nullptr; }
    od_handle* hd = new (std::nothrow) od_handle();
    if (!hd) { fail("od_create: out of memory"); return nullptr; }
    hd->model = model; hd->device = device; hd->h = h; hd->launches = 0;
    if (opts) hd->opts = *opts;
    else {
        od_default_options(model, &hd->opts);
        // NULL = the reference CONSTRUCTOR's defaults, ImplicitDynamics(...; r_tol=1e-8, κ_eval_tol=1e-6, κ_grad_tol=1e-6) (src/dynamics.jl:51-53);
        // od_default_options() holds the looser values the reference's EXAMPLES pass.  The rocket keeps RocketInfo's own settings.
        if (model != OD_ROCKET) { hd->opts.kappa_eval_tol = 1e-6; hd->opts.kappa_grad_tol = 1e-6; }
    }
    for (int k = 0; k < 4; ++k) hd->params[k] = 0.0;
    if (model == OD_CARTPOLE_FRICTION) { hd->params[0] = 0.1; hd->params[1] = 0.1; }      // cartpole/model.jl:132
    if (model == OD_HOPPER) { hd->params[0] = 0.5; hd->params[1] = 0.5; }                 // DESIGN.md §Hopper constants
    if (model == OD_ROCKET) hd->params[0] = 12.5;                                         // examples/rocket.jl:16
    for (int k = 0; k < nparams && k < 4 && params; ++k) hd->params[k] = params[k];
    if ((e = cudaStreamCreateWithFlags(&hd->stream, cudaStreamNonBlocking)) != cudaSuccess) { fail("cudaStreamCreate", e); delete hd; return nullptr; }
    hd->own_stream = true;
    hd->side = nullptr; hd->ev_fork = nullptr; hd->ev_join = nullptr;
    if (model == OD_PLANAR_PUSH) {                            // optional: without them the tail of the sweep runs on one stream
        if (cudaStreamCreateWithFlags(&hd->side, cudaStreamNonBlocking) != cudaSuccess) hd->side = nullptr;
        if (hd->side && (cudaEventCreateWithFlags(&hd->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                         cudaEventCreateWithFlags(&hd->ev_join, cudaEventDisableTiming) != cudaSuccess)) {
            if (hd->ev_fork) cudaEventDestroy(hd->ev_fork);
            cudaStreamDestroy(hd->side); hd->side = nullptr; hd->ev_fork = nullptr; hd->ev_join = nullptr;
            cudaGetLastError();
        }
    }
    return hd;
}

void od_destroy(od_handle* hd) {
    if (!hd) return;
    cudaSetDevice(hd->device);
    cudaStreamSynchronize(hd->stream);
    hd->in.release(); hd->out.release(); hd->st.release(); hd->aux.release(); hd->aux2.release(); hd->sweep.release();
    hd->hin.release(); hd->hout.release();
    if (hd->side) { cudaStreamSynchronize(hd->side); cudaEventDestroy(hd->ev_fork); cudaEventDestroy(hd->ev_join); cudaStreamDestroy(hd->side); }
    if (hd->own_stream) cudaStreamDestroy(hd->stream);
    delete hd;
}

int od_set_stream(od_handle* hd, void* s) {
    if (!hd) return fail("od_set_stream: null handle");
    if (hd->own_stream) { cudaStreamSynchronize(hd->stream); cudaStreamDestroy(hd->stream); hd->own_stream = false; }
    hd->stream = (cudaStream_t)s;
    return 0;
}

int od_synchronize(od_handle* hd) {
    if (!hd) return fail("od_synchronize: null handle");
    OD_CUDA(cudaStreamSynchronize(hd->stream));
    return 0;
}

int64_t od_launch_count(const od_handle* hd) { return hd ? hd->launches : 0; }

}  // extern "C"

// The step / rollout kernels of each contact model are instantiated in their own translation unit (csrc/inst/contact_<model>.cu,
// through launch.cuh) so that the library builds in parallel; this file holds the host side of the C ABI.
// Zero-copy output needs whole rows leaving the kernel as contiguous stores (the register path stages the packed row in shared
// memory); the shared-memory-LU path scatters 8-byte stores, which PCIe handles badly (measured: 2.4x slower end to end).
static bool rows_leave_coalesced(int model, int B) {
    const bool regok = model == OD_HOPPER || model == OD_CARTPOLE_FRICTION || model == OD_ACROBOT_IMPACT;
    return regok && reg_path() && lanes_for(B) >= 4;
}

// grad_sim_q3: the eval tolerance is replaced by the gradient tolerance, so that q3 is the gradient simulator's own result
// (step!(grad_sim, …) of the reference returns q3 at κ_grad_tol)
static int launch_step(od_handle* hd, StepArgs& a, bool grad_sim_q3 = false) {
    if (a.B <= 0) return 0;
    a.h = hd->h;
    for (int k = 0; k < 4; ++k) a.fric[k] = hd->params[k];
    a.opts.r_tol = hd->opts.r_tol; a.opts.kappa_eval_tol = hd->opts.kappa_eval_tol; a.opts.kappa_grad_tol = hd->opts.kappa_grad_tol;
    a.opts.ls_scale = hd->opts.ls_scale; a.opts.max_iter = hd->opts.max_iter; a.opts.max_ls = hd->opts.max_ls;
    if (grad_sim_q3) a.opts.kappa_eval_tol = hd->opts.kappa_grad_tol;
    if (hd->model == OD_PLANAR_PUSH && a.n_peers <= 1) {         // scratch of the persistent sweep (launch.cuh decides whether it is used)
        Dims d; dims_of(hd->model, &d);
        const size_t zb = sizeof(double) * (size_t)d.nz * a.B;   // [counters | snapshots | parked iterates | parked progress words | parked list]
        if (hd->sweep.reserve(256 + 2 * zb + 3 * sizeof(int) * (size_t)a.B) == cudaSuccess) {
            char* base = (char*)hd->sweep.p;
            a.work_queue = (unsigned int*)base;
            a.z_snapshots = (double*)(base + 256);
            a.z_park = (double*)(base + 256 + zb);
            a.park_info = (int*)(base + 256 + 2 * zb);
            a.park_list = a.park_info + 2 * (size_t)a.B;
        }
        a.side_stream = hd->side; a.ev_fork = hd->ev_fork; a.ev_join = hd->ev_join;
    }
    cudaError_t e;
    switch (hd->model) {
        case OD_ACROBOT_IMPACT: e = od_launch_step_acrobot_impact(a, hd->stream); break;
        case OD_ACROBOT_NOMINAL: e = od_launch_step_acrobot_nominal(a, hd->stream); break;
        case OD_CARTPOLE_FRICTION: e = od_launch_step_cartpole_friction(a, hd->stream); break;
        case OD_CARTPOLE_FRICTIONLESS: e = od_launch_step_cartpole_frictionless(a, hd->stream); break;
        case OD_PLANAR_PUSH: e = od_launch_step_planar_push(a, hd->stream); break;
        case OD_HOPPER: e = od_launch_step_hopper(a, hd->stream); break;
        default: return fail("this entry point needs a contact model handle (not OD_ROCKET)");
    }
    if (e != cudaSuccess) return fail("contact_step_kernel launch", e);
    hd->launches += last_launch_kernels();
    return 0;
}

extern "C" {

int od_step_grad_batch_device(od_handle* hd, int B, const double* q1, const double* q2, const double* u,
                              double* q3, double* dq1, double* dq2, double* du, int32_t* status, int32_t* iters, int want_eval, int want_grad) {
    if (!hd) return fail("null handle");
    Dims d; dims_of(hd->model, &d);
    if (want_grad && !(dq1 && dq2 && du)) return fail("od_step_grad_batch_device: gradient outputs must all be non-null");
    OD_CUDA(cudaSetDevice(hd->device));
    StepArgs a; memset(&a, 0, sizeof(a));
    a.B = B; a.q1 = q1; a.q2 = q2; a.u = u; a.in_stride_q = d.nq; a.in_stride_u = d.nu;
    a.q3 = q3; a.dq1 = want_grad ? dq1 : nullptr; a.dq2 = dq2; a.du = du;
    a.out_stride_q3 = d.nq; a.out_stride_dq = d.nq * d.nq; a.out_stride_du = d.nq * d.nu;
    a.status = status; a.iters = iters; a.want_eval = want_eval; a.want_grad = want_grad;
    return launch_step(hd, a);
}

int od_step_grad_packed_device(od_handle* hd, int B, const double* in, double* out, int32_t* status, int32_t* iters, int want_eval, int want_grad) {
    if (!hd) return fail("null handle");
    Dims d; dims_of(hd->model, &d);
    OD_CUDA(cudaSetDevice(hd->device));
    const int inw = 2 * d.nq + d.nu, outw = d.nq + d.nq * inw;
    StepArgs a; memset(&a, 0, sizeof(a));
    a.B = B; a.q1 = in; a.q2 = in + d.nq; a.u = in + 2 * d.nq; a.in_stride_q = inw; a.in_stride_u = inw; a.in_packed = 1;
    a.q3 = out; a.dq1 = want_grad ? out + d.nq : nullptr; a.dq2 = out + d.nq + d.nq * d.nq; a.du = out + d.nq + 2 * d.nq * d.nq;
    a.out_stride_q3 = outw; a.out_stride_dq = outw; a.out_stride_du = outw;
    a.status = status; a.iters = iters; a.want_eval = want_eval; a.want_grad = want_grad;
    a.packed_out = ((reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
    return launch_step(hd, a);
}

}  // extern "C"

// Rows [row0, row0 + B) of this rank's gather buffer → the same rows of every peer's buffer (peer-mapped stores over NVLink).  Used
// where the kernels do not forward their rows themselves: the persistent sweep of the planar push.
static __global__ void gather_forward_kernel(const StepArgs a) {
    const size_t n = (size_t)a.B * a.gather_width, off = (size_t)a.gather_row0 * a.gather_width;
    const double* src = a.peer_out[a.self_rank] + off;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const double v = src[k];
        for (int p = 0; p < a.n_peers; ++p)
            if (p != a.self_rank) a.peer_out[p][off + k] = v;
    }
}

extern "C" {

// Multi-GPU derivative sweep, every variant (see od_gather_desc in the header)
int od_step_grad_packed_gather_ex_device(od_handle* hd, int B, const double* in, const od_gather_desc* g, int32_t* status, int32_t* iters) {
    if (!hd) return fail("null handle");
    if (!g || g->world < 1 || g->world > 8 || g->rank < 0 || g->rank >= g->world || !g->gather_buffers)
        return fail("od_step_grad_packed_gather: need 1 <= world <= 8 and the peer buffer table");
    const bool sync = g->flag_buffers != nullptr;
    if (sync && (!g->block_counter || (g->epoch == 0 && !g->epoch_dev)))
        return fail("od_step_grad_packed_gather: the fused barrier needs the flag table, a block counter and an epoch (value >= 1 or device counter)");
    if (hd->model == OD_ROCKET) return fail("od_step_grad_packed_gather: contact models only");
    if ((sync || g->multicast_buffer) && !rows_leave_coalesced(hd->model, B > 0 ? B : 1))
        return fail("od_step_grad_packed_gather: the fused barrier / multicast stores need the cooperative-lane register path (hopper, cartpole, acrobot)");
    Dims d; dims_of(hd->model, &d);
    OD_CUDA(cudaSetDevice(hd->device));
    const int inw = 2 * d.nq + d.nu, outw = d.nq + d.nq * inw;
    double* out = (double*)g->gather_buffers[g->rank] + (size_t)g->row0 * outw;
    StepArgs a; memset(&a, 0, sizeof(a));
    a.B = B; a.q1 = in; a.q2 = in + d.nq; a.u = in + 2 * d.nq; a.in_stride_q = inw; a.in_stride_u = inw; a.in_packed = 1;
    a.q3 = out; a.dq1 = out + d.nq; a.dq2 = out + d.nq + d.nq * d.nq; a.du = out + d.nq + 2 * d.nq * d.nq;
    a.out_stride_q3 = outw; a.out_stride_dq = outw; a.out_stride_du = outw;
    a.status = status; a.iters = iters; a.want_eval = 1; a.want_grad = 1;
    a.n_peers = g->world; a.self_rank = g->rank; a.gather_row0 = g->row0; a.gather_width = outw;
    bool aligned = (g->multicast_buffer & 15) == 0;
    for (int r = 0; r < g->world; ++r) {
        a.peer_out[r] = (double*)g->gather_buffers[r];
        aligned = aligned && ((g->gather_buffers[r] & 15) == 0);
        if (sync) a.sync_flags[r] = (unsigned long long*)g->flag_buffers[r];
    }
    a.packed_out = aligned ? 1 : 0;
    a.mc_out = (aligned && outw % 2 == 0) ? (double*)g->multicast_buffer : nullptr;
    if (sync) {
        a.sync_counter = g->block_counter; a.sync_epoch = g->epoch; a.sync_epoch_dev = (unsigned long long*)g->epoch_dev;
        a.mc_flags = (unsigned long long*)g->multicast_flags;
    }
    if (B <= 0) {
        if (!sync) return 0;
        // empty shard (ragged split, B_total < world): the peers still wait for this rank's flag
        gather_sync_only_kernel<<<1, 32, 0, hd->stream>>>(a);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail("gather_sync_only_kernel launch", e);
        hd->launches++;
        return 0;
    }
    if (hd->model == OD_PLANAR_PUSH && !sync && !a.mc_out && g->world > 1 && reg_path() && persist_min_batch() > 0 && B >= persist_min_batch()) {
        // large shard of the planar push: the persistent sweep (launch.cuh) writes this rank's rows into its own gather buffer, one
        // copy kernel forwards them to the peers; the caller's barrier follows as after the per-warp kernel
        StepArgs local = a;
        local.n_peers = 1;
        const int rc = launch_step(hd, local);
        if (rc != 0) return rc;
        int blocks = (int)(((size_t)B * outw + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        gather_forward_kernel<<<blocks, 256, 0, hd->stream>>>(a);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail("gather_forward_kernel launch", e);
        hd->launches++;
        return 0;
    }
    return launch_step(hd, a);
}

int od_step_grad_packed_gather_device(od_handle* hd, int B, const double* in, long long row0, int world, int rank,
                                      const uint64_t* gather_buffers, int32_t* status, int32_t* iters) {
    od_gather_desc g; memset(&g, 0, sizeof(g));
    g.world = world; g.rank = rank; g.row0 = row0; g.gather_buffers = gather_buffers;
    return od_step_grad_packed_gather_ex_device(hd, B, in, &g, status, iters);
}

int od_step_grad_packed_gather_sync_device(od_handle* hd, int B, const double* in, long long row0, int world, int rank,
                                           const uint64_t* gather_buffers, const uint64_t* flag_buffers, uint32_t* block_counter,
                                           uint64_t epoch, int32_t* status, int32_t* iters) {
    if (!flag_buffers || !block_counter || epoch == 0)
        return fail("od_step_grad_packed_gather_sync_device: need 1 <= world <= 8, the peer buffer and flag tables, a block counter and epoch >= 1");
    od_gather_desc g; memset(&g, 0, sizeof(g));
    g.world = world; g.rank = rank; g.row0 = row0; g.gather_buffers = gather_buffers; g.flag_buffers = flag_buffers;
    g.block_counter = block_counter; g.epoch = epoch;
    return od_step_grad_packed_gather_ex_device(hd, B, in, &g, status, iters);
}

// Device-visible alias of a pinned (page-locked, mapped) host buffer, or null for pageable memory (queried on every call: a
// cached answer would go stale if the caller unpinned or freed the buffer).
static void* mapped_alias(const void* host) {
    if (!host) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged) return at.devicePointer;
    return nullptr;
}
// OD_ZEROCOPY: 0 = always stage through device buffers, 1 = outputs written in place, 2 = inputs read in place as well (default).
// Measured on B200 (hopper, 4096 problems, pinned buffers), µs per call: round 1 (kernel 0.098 ms, every lane loading every input
// value) 152 / 120 / 127; round 2 (kernel 0.061 ms, the lanes of a group load their packed 80-B row once between them, whole
// sectors) 120.0 / 98.3 / 92.6 — the cooperative row load made reading the inputs over PCIe cheaper than a cudaMemcpyAsync ahead
// of the launch (profiles/r02v_e2e_zero_copy_modes.txt).
static int zero_copy_mode() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("OD_ZEROCOPY"); v = e ? atoi(e) : 2; }
    return v;
}
// Host-buffer entry point.  When the caller's buffers are pinned host memory (cudaHostAlloc / cudaHostRegister — what a Julia
// or Python host uses for its trajectory arrays), the kernel reads the 80-B input rows and writes each finished 352-B output row
// straight over PCIe: the device→host transfer of a row overlaps the problems that are still iterating, and the separate
// cudaMemcpy calls (and their launch gaps) disappear.  Pageable buffers are staged through device memory as before.
int od_step_grad_packed(od_handle* hd, int B, const double* in, double* out, int32_t* status) {
    if (!hd) return fail("null handle");
    Dims d; dims_of(hd->model, &d);
    if (hd->model == OD_ROCKET) return fail("od_step_grad_packed: use od_rocket_batch for OD_ROCKET");
    if (B <= 0) return 0;
    if (!in || !out) return fail("od_step_grad_packed: null buffer");
    OD_CUDA(cudaSetDevice(hd->device));
    const size_t inw = 2 * d.nq + d.nu, outw = d.nq + d.nq * inw;
    const int zc = rows_leave_coalesced(hd->model, B) ? zero_copy_mode() : 0;
    const double* din = zc >= 2 ? (const double*)mapped_alias(in) : nullptr;
    double* dout = zc >= 1 ? (double*)mapped_alias(out) : nullptr;
    int32_t* dst = (zc >= 1 && status) ? (int32_t*)mapped_alias(status) : nullptr;
    if (!din) {
        OD_CUDA(hd->in.reserve(sizeof(double) * inw * B));
        OD_CUDA(cudaMemcpyAsync(hd->in.p, in, sizeof(double) * inw * B, cudaMemcpyHostToDevice, hd->stream));
        din = (const double*)hd->in.p;
    }
    const bool stage_out = !dout, stage_st = !dst;
    if (stage_out) { OD_CUDA(hd->out.reserve(sizeof(double) * outw * B)); dout = (double*)hd->out.p; }
    if (stage_st) { OD_CUDA(hd->st.reserve(sizeof(int32_t) * B)); dst = (int32_t*)hd->st.p; }
    if (od_step_grad_packed_device(hd, B, din, dout, dst, nullptr, 1, 1)) return 1;
    if (stage_out) OD_CUDA(cudaMemcpyAsync(out, dout, sizeof(double) * outw * B, cudaMemcpyDeviceToHost, hd->stream));
    if (stage_st && status) OD_CUDA(cudaMemcpyAsync(status, dst, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, hd->stream));
    OD_CUDA(cudaStreamSynchronize(hd->stream));
    return 0;
}

// Separate host arrays → ONE packed, pinned staging row per problem → one H2D copy, one launch; the kernel writes the finished
// packed rows (and, behind them, the status words) straight into the pinned output staging (or, for the models whose rows do not
// leave coalesced, one D2H copy brings rows + status back) → unpack on the host.  Replaces 3 H2D + 5 D2H copies per call; for a
// single problem (the reference's per-timestep f / fx / fu call shape) the call is launch + synchronise.
//   in_vel: the first array is v1, not q1 (RoboDojo.step! call shape);  sim: -1 = f + fx/fu in one pass (q3 at κ_eval_tol, IFT at
//   κ_grad_tol), 0 = eval simulator only, 1 = gradient simulator only (q3 AND the IFT at κ_grad_tol, like step!(grad_sim, …)).
static int step_host_rows(od_handle* hd, int B, const double* a0, const double* q2, const double* u, int in_vel, int sim,
                          double* q3, double* dq1, double* dq2, double* du, int32_t* status) {
    Dims d; dims_of(hd->model, &d);
    const size_t nq = d.nq, nu = d.nu, inw = 2 * nq + nu;
    const bool want_grad = sim != 0 && dq1;
    const size_t outw = want_grad ? nq + nq * inw : nq;          // eval only: compact q3 rows
    OD_CUDA(cudaSetDevice(hd->device));
    const size_t in_bytes = sizeof(double) * inw * B, out_bytes = sizeof(double) * outw * B, st_bytes = sizeof(int32_t) * B;
    OD_CUDA(hd->hin.reserve(in_bytes));
    OD_CUDA(hd->hout.reserve(out_bytes + st_bytes));
    OD_CUDA(hd->in.reserve(in_bytes));
    double* hin = (double*)hd->hin.p; double* hout = (double*)hd->hout.p; int32_t* hst = (int32_t*)((char*)hd->hout.p + out_bytes);
    for (int i = 0; i < B; ++i) {
        double* r = hin + (size_t)i * inw;
        memcpy(r, a0 + (size_t)i * nq, sizeof(double) * nq);
        memcpy(r + nq, q2 + (size_t)i * nq, sizeof(double) * nq);
        memcpy(r + 2 * nq, u + (size_t)i * nu, sizeof(double) * nu);
    }
    const bool coal = rows_leave_coalesced(hd->model, B);
    const double* din = nullptr;
    if (coal && zero_copy_mode() >= 2) {                    // the kernel reads the pinned staging rows in place (cooperative row loads)
        OD_CUDA(cudaHostGetDevicePointer((void**)&din, hin, 0));
    } else {
        OD_CUDA(cudaMemcpyAsync(hd->in.p, hin, in_bytes, cudaMemcpyHostToDevice, hd->stream));
        din = (const double*)hd->in.p;
    }
    const bool zc = want_grad && coal && zero_copy_mode() >= 1;
    double* dout = nullptr; int32_t* dst = nullptr;
    if (zc) {
        OD_CUDA(cudaHostGetDevicePointer((void**)&dout, hout, 0));
        dst = (int32_t*)((char*)dout + out_bytes);
    } else {
        OD_CUDA(hd->out.reserve(out_bytes + st_bytes));
        dout = (double*)hd->out.p; dst = (int32_t*)((char*)hd->out.p + out_bytes);
    }
    StepArgs a; memset(&a, 0, sizeof(a));
    a.B = B; a.q1 = din; a.q2 = din + nq; a.u = din + 2 * nq; a.in_stride_q = (int)inw; a.in_stride_u = (int)inw; a.in_packed = 1; a.in_vel = in_vel;
    a.q3 = dout; a.dq1 = want_grad ? dout + nq : nullptr; a.dq2 = dout + nq + nq * nq; a.du = dout + nq + 2 * nq * nq;
    a.out_stride_q3 = (int)outw; a.out_stride_dq = (int)outw; a.out_stride_du = (int)outw;
    a.status = dst; a.want_eval = (sim != 1 || q3) ? 1 : 0; a.want_grad = want_grad ? 1 : 0;
    a.packed_out = ((reinterpret_cast<uintptr_t>(dout) & 15) == 0) ? 1 : 0;
    if (launch_step(hd, a, sim == 1)) return 1;
    if (!zc) OD_CUDA(cudaMemcpyAsync(hout, dout, out_bytes + st_bytes, cudaMemcpyDeviceToHost, hd->stream));
    OD_CUDA(cudaStreamSynchronize(hd->stream));
    for (int i = 0; i < B; ++i) {
        const double* r = hout + (size_t)i * outw;
        if (q3) memcpy(q3 + (size_t)i * nq, r, sizeof(double) * nq);
        if (want_grad) {
            memcpy(dq1 + (size_t)i * nq * nq, r + nq, sizeof(double) * nq * nq);
            memcpy(dq2 + (size_t)i * nq * nq, r + nq + nq * nq, sizeof(double) * nq * nq);
            memcpy(du + (size_t)i * nq * nu, r + nq + 2 * nq * nq, sizeof(double) * nq * nu);
        }
    }
    if (status) memcpy(status, hst, st_bytes);
    return 0;
}

int od_step_grad_batch(od_handle* hd, int B, const double* q1, const double* q2, const double* u,
                       double* q3, double* dq1, double* dq2, double* du, int32_t* status) {
    if (!hd) return fail("null handle");
    if (hd->model == OD_ROCKET) return fail("od_step_grad_batch: use od_rocket_batch for OD_ROCKET");
    if (B <= 0) return 0;
    if (!q1 || !q2 || !u) return fail("od_step_grad_batch: null input");
    const bool want_grad = dq1 || dq2 || du;
    if (want_grad && !(dq1 && dq2 && du)) return fail("od_step_grad_batch: gradient outputs must be all null or all non-null");
    if (!want_grad && !q3) return fail("od_step_grad_batch: nothing to compute");
    return step_host_rows(hd, B, q1, q2, u, 0, want_grad ? (q3 ? -1 : 1) : 0, q3, dq1, dq2, du, status);
}

// RoboDojo.step!(sim, q, v, u, t) for a batch — the call the reference's hopper example makes directly
// (reference examples/hopper.jl:63,89,112,133,157; inside f / fx / fu at src/dynamics.jl:88,103,123).
int od_sim_step_batch(od_handle* hd, int B, int grad_sim, const double* q, const double* v, const double* u,
                      double* q3, double* dq3dq1, double* dq3dq2, double* dq3du1, int32_t* status) {
    if (!hd) return fail("null handle");
    if (hd->model == OD_ROCKET) return fail("od_sim_step_batch: contact models only");
    if (B <= 0) return 0;
    if (!q || !v || !u) return fail("od_sim_step_batch: null input");
    if (grad_sim) {
        if (!(dq3dq1 && dq3dq2 && dq3du1)) return fail("od_sim_step_batch: the gradient simulator needs all three Jacobian outputs");
        return step_host_rows(hd, B, v, q, u, 1, 1, q3, dq3dq1, dq3dq2, dq3du1, status);
    }
    if (!q3) return fail("od_sim_step_batch: q3 is null");
    return step_host_rows(hd, B, v, q, u, 1, 0, q3, nullptr, nullptr, nullptr, status);
}

int od_step_batch(od_handle* hd, int B, const double* q1, const double* q2, const double* u, double* q3, int32_t* status) {
    if (!q3) return fail("od_step_batch: q3 is null");
    return od_step_grad_batch(hd, B, q1, q2, u, q3, nullptr, nullptr, nullptr, status);
}

int od_rollout_batch_device(od_handle* hd, int R, int T, const double* x1, const double* ubar, long long ubar_stride, const double* xbar,
                            const double* K, const double* kff, const double* alpha, double* X, double* U, int32_t* status, int32_t* iters) {
    if (!hd) return fail("null handle");
    if (hd->model == OD_ROCKET) return fail("od_rollout_batch: contact models only");
    if (R <= 0 || T <= 0) return 0;
    if (T == 1) {                              // zero steps: the rollout is [x1], like iLQR.rollout with an empty control sequence
        if (!x1 || !X) return fail("od_rollout_batch: x1 and X are required");
        Dims d1; dims_of(hd->model, &d1);
        OD_CUDA(cudaSetDevice(hd->device));
        OD_CUDA(cudaMemcpyAsync(X, x1, sizeof(double) * 2 * d1.nq * (size_t)R, cudaMemcpyDeviceToDevice, hd->stream));
        return 0;
    }
    if (!x1 || !ubar || !X || !U) return fail("od_rollout_batch: x1, ubar, X and U are required");
    if (K && !xbar) return fail("od_rollout_batch: feedback gains K need the nominal states xbar");
    OD_CUDA(cudaSetDevice(hd->device));
    RolloutArgs a; memset(&a, 0, sizeof(a));
    a.R = R; a.T = T; a.x1 = x1; a.ubar = ubar; a.ubar_stride = ubar_stride; a.xbar = xbar; a.K = K; a.kff = kff; a.alpha = alpha;
    a.X = X; a.U = U; a.status = status; a.iters = iters; a.h = hd->h;
    for (int k = 0; k < 4; ++k) a.fric[k] = hd->params[k];
    a.opts.r_tol = hd->opts.r_tol; a.opts.kappa_eval_tol = hd->opts.kappa_eval_tol; a.opts.kappa_grad_tol = hd->opts.kappa_grad_tol;
    a.opts.ls_scale = hd->opts.ls_scale; a.opts.max_iter = hd->opts.max_iter; a.opts.max_ls = hd->opts.max_ls;
    cudaError_t e;
    switch (hd->model) {
        case OD_ACROBOT_IMPACT: e = od_launch_rollout_acrobot_impact(a, hd->stream); break;
        case OD_ACROBOT_NOMINAL: e = od_launch_rollout_acrobot_nominal(a, hd->stream); break;
        case OD_CARTPOLE_FRICTION: e = od_launch_rollout_cartpole_friction(a, hd->stream); break;
        case OD_CARTPOLE_FRICTIONLESS: e = od_launch_rollout_cartpole_frictionless(a, hd->stream); break;
        case OD_PLANAR_PUSH: e = od_launch_rollout_planar_push(a, hd->stream); break;
        case OD_HOPPER: e = od_launch_rollout_hopper(a, hd->stream); break;
        default: return fail("od_rollout_batch: bad model");
    }
    if (e != cudaSuccess) return fail("contact_rollout_kernel launch", e);
    hd->launches++;
    return 0;
}

int od_rollout_batch(od_handle* hd, int R, int T, const double* x1, const double* ubar, int ubar_per_rollout, const double* xbar,
                     const double* K, const double* kff, const double* alpha, double* X, double* U, int32_t* status) {
    if (!hd) return fail("null handle");
    Dims d; dims_of(hd->model, &d);
    if (hd->model == OD_ROCKET) return fail("od_rollout_batch: contact models only");
    if (R <= 0 || T <= 0) return 0;
    if (T == 1) {
        if (!x1 || !X) return fail("od_rollout_batch: x1 and X are required");
        memcpy(X, x1, sizeof(double) * 2 * d.nq * (size_t)R);
        return 0;
    }
    if (!x1 || !ubar || !X) return fail("od_rollout_batch: x1, ubar and X are required");
    if (K && !xbar) return fail("od_rollout_batch: feedback gains K need the nominal states xbar");
    OD_CUDA(cudaSetDevice(hd->device));
    const size_t nx = 2 * d.nq, nu = d.nu, S = T - 1;
    const size_t n_x1 = (size_t)R * nx, n_ub = (ubar_per_rollout ? (size_t)R : 1) * S * nu, n_xb = xbar ? (size_t)T * nx : 0;
    const size_t n_K = K ? S * nu * nx : 0, n_k = kff ? S * nu : 0, n_al = alpha ? (size_t)R : 0;
    const size_t n_in = n_x1 + n_ub + n_xb + n_K + n_k + n_al;
    const size_t n_X = (size_t)R * T * nx, n_U = (size_t)R * S * nu;
    OD_CUDA(hd->in.reserve(sizeof(double) * n_in));
    OD_CUDA(hd->out.reserve(sizeof(double) * (n_X + n_U)));
    OD_CUDA(hd->st.reserve(sizeof(int32_t) * (size_t)R * S));
    double* p = (double*)hd->in.p;
    double* d_x1 = p; p += n_x1; double* d_ub = p; p += n_ub;
    double* d_xb = xbar ? p : nullptr; p += n_xb; double* d_K = K ? p : nullptr; p += n_K;
    double* d_k = kff ? p : nullptr; p += n_k; double* d_al = alpha ? p : nullptr;
    OD_CUDA(cudaMemcpyAsync(d_x1, x1, sizeof(double) * n_x1, cudaMemcpyHostToDevice, hd->stream));
    OD_CUDA(cudaMemcpyAsync(d_ub, ubar, sizeof(double) * n_ub, cudaMemcpyHostToDevice, hd->stream));
    if (xbar) OD_CUDA(cudaMemcpyAsync(d_xb, xbar, sizeof(double) * n_xb, cudaMemcpyHostToDevice, hd->stream));
    if (K) OD_CUDA(cudaMemcpyAsync(d_K, K, sizeof(double) * n_K, cudaMemcpyHostToDevice, hd->stream));
    if (kff) OD_CUDA(cudaMemcpyAsync(d_k, kff, sizeof(double) * n_k, cudaMemcpyHostToDevice, hd->stream));
    if (alpha) OD_CUDA(cudaMemcpyAsync(d_al, alpha, sizeof(double) * n_al, cudaMemcpyHostToDevice, hd->stream));
    double* d_X = (double*)hd->out.p; double* d_U = d_X + n_X;
    if (od_rollout_batch_device(hd, R, T, d_x1, d_ub, ubar_per_rollout ? (long long)(S * nu) : 0, d_xb, d_K, d_k, d_al, d_X, d_U, (int32_t*)hd->st.p, nullptr)) return 1;
    OD_CUDA(cudaMemcpyAsync(X, d_X, sizeof(double) * n_X, cudaMemcpyDeviceToHost, hd->stream));
    if (U) OD_CUDA(cudaMemcpyAsync(U, d_U, sizeof(double) * n_U, cudaMemcpyDeviceToHost, hd->stream));
    if (status) OD_CUDA(cudaMemcpyAsync(status, hd->st.p, sizeof(int32_t) * (size_t)R * S, cudaMemcpyDeviceToHost, hd->stream));
    OD_CUDA(cudaStreamSynchronize(hd->stream));
    return 0;
}

int od_riccati_batch_device(od_handle* hd, int NT, int T, const double* jac, const double* lx, const double* lu, const double* lxx,
                            const double* luu, const double* lux, double reg, double* K, double* k, double* dV, int32_t* status) {
    if (!hd) return fail("null handle");
    if (hd->model == OD_ROCKET) return fail("od_riccati_batch: contact models only (state x = [q1; q2])");
    Dims d; dims_of(hd->model, &d);
    if (NT <= 0 || T <= 1) return 0;
    if (!jac || !lx || !lu || !lxx || !luu || !K || !k) return fail("od_riccati_batch: jac, lx, lu, lxx, luu, K and k are required");
    OD_CUDA(cudaSetDevice(hd->device));
    RiccatiArgs a; memset(&a, 0, sizeof(a));
    a.NT = NT; a.T = T; a.jac = jac; a.lx = lx; a.lu = lu; a.lxx = lxx; a.luu = luu; a.lux = lux; a.reg = reg;
    a.K = K; a.k = k; a.dV = dV; a.status = status;
    constexpr int WARPS = 2;
    const int grid = (NT + WARPS - 1) / WARPS;
    if (d.nq == 2 && d.nu == 1) riccati_kernel<2, 1, WARPS><<<grid, 32 * WARPS, sizeof(double) * WARPS * Riccati<2, 1, 32>::WS, hd->stream>>>(a);
    else if (d.nq == 4 && d.nu == 2) riccati_kernel<4, 2, WARPS><<<grid, 32 * WARPS, sizeof(double) * WARPS * Riccati<4, 2, 32>::WS, hd->stream>>>(a);
    else if (d.nq == 5 && d.nu == 2) riccati_kernel<5, 2, WARPS><<<grid, 32 * WARPS, sizeof(double) * WARPS * Riccati<5, 2, 32>::WS, hd->stream>>>(a);
    else return fail("od_riccati_batch: no instantiation for this model's dimensions");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("riccati_kernel launch", e);
    hd->launches++;
    return 0;
}

int od_riccati_batch(od_handle* hd, int NT, int T, const double* jac, const double* lx, const double* lu, const double* lxx,
                     const double* luu, const double* lux, double reg, double* K, double* k, double* dV, int32_t* status) {
    if (!hd) return fail("null handle");
    if (hd->model == OD_ROCKET) return fail("od_riccati_batch: contact models only (state x = [q1; q2])");
    Dims d; dims_of(hd->model, &d);
    if (NT <= 0 || T <= 1) return 0;
    if (!jac || !lx || !lu || !lxx || !luu || !K || !k) return fail("od_riccati_batch: jac, lx, lu, lxx, luu, K and k are required");
    OD_CUDA(cudaSetDevice(hd->device));
    const size_t n = 2 * d.nq, m = d.nu, S = T - 1, roww = d.nq + d.nq * (n + m);
    const size_t c_jac = (size_t)NT * S * roww, c_lx = (size_t)NT * T * n, c_lu = (size_t)NT * S * m, c_lxx = (size_t)NT * T * n * n;
    const size_t c_luu = (size_t)NT * S * m * m, c_lux = lux ? (size_t)NT * S * m * n : 0;
    const size_t c_K = (size_t)NT * S * m * n, c_k = (size_t)NT * S * m, c_dV = (size_t)NT * 2;
    OD_CUDA(hd->in.reserve(sizeof(double) * (c_jac + c_lx + c_lu + c_lxx + c_luu + c_lux)));
    OD_CUDA(hd->out.reserve(sizeof(double) * (c_K + c_k + c_dV)));
    OD_CUDA(hd->st.reserve(sizeof(int32_t) * NT));
    double* p = (double*)hd->in.p;
    double* d_jac = p; p += c_jac; double* d_lx = p; p += c_lx; double* d_lu = p; p += c_lu; double* d_lxx = p; p += c_lxx;
    double* d_luu = p; p += c_luu; double* d_lux = lux ? p : nullptr;
    OD_CUDA(cudaMemcpyAsync(d_jac, jac, sizeof(double) * c_jac, cudaMemcpyHostToDevice, hd->stream));
    OD_CUDA(cudaMemcpyAsync(d_lx, lx, sizeof(double) * c_lx, cudaMemcpyHostToDevice, hd->stream));
    OD_CUDA(cudaMemcpyAsync(d_lu, lu, sizeof(double) * c_lu, cudaMemcpyHostToDevice, hd->stream));
    OD_CUDA(cudaMemcpyAsync(d_lxx, lxx, sizeof(double) * c_lxx, cudaMemcpyHostToDevice, hd->stream));
    OD_CUDA(cudaMemcpyAsync(d_luu, luu, sizeof(double) * c_luu, cudaMemcpyHostToDevice, hd->stream));
    if (lux) OD_CUDA(cudaMemcpyAsync(d_lux, lux, sizeof(double) * c_lux, cudaMemcpyHostToDevice, hd->stream));
    double* d_K = (double*)hd->out.p; double* d_k = d_K + c_K; double* d_dV = d_k + c_k;
    if (od_riccati_batch_device(hd, NT, T, d_jac, d_lx, d_lu, d_lxx, d_luu, d_lux, reg, d_K, d_k, d_dV, (int32_t*)hd->st.p)) return 1;
    OD_CUDA(cudaMemcpyAsync(K, d_K, sizeof(double) * c_K, cudaMemcpyDeviceToHost, hd->stream));
    OD_CUDA(cudaMemcpyAsync(k, d_k, sizeof(double) * c_k, cudaMemcpyDeviceToHost, hd->stream));
    if (dV) OD_CUDA(cudaMemcpyAsync(dV, d_dV, sizeof(double) * c_dV, cudaMemcpyDeviceToHost, hd->stream));
    if (status) OD_CUDA(cudaMemcpyAsync(status, hd->st.p, sizeof(int32_t) * NT, cudaMemcpyDeviceToHost, hd->stream));
    OD_CUDA(cudaStreamSynchronize(hd->stream));
    return 0;
}

// H = Σ η ηᵀ inverted on the host by Gauss–Jordan with partial pivoting (ncol ≤ 16); returns non-zero when singular
int od_bundle_prepare(int ncol, int N, const double* eta, double* Hinv) {
    if (ncol <= 0 || ncol > 16 || N <= 0 || !eta || !Hinv) return fail("od_bundle_prepare: need 0 < 2nq+nu <= 16, N > 0, eta and Hinv");
    double H[16 * 16]; double* Hi = Hinv;
    for (int i = 0; i < ncol * ncol; ++i) { H[i] = 0.0; Hi[i] = 0.0; }
    for (int k = 0; k < N; ++k) for (int i = 0; i < ncol; ++i) for (int j = 0; j < ncol; ++j) H[i * ncol + j] += eta[(size_t)k * ncol + i] * eta[(size_t)k * ncol + j];
    for (int i = 0; i < ncol; ++i) Hi[i * ncol + i] = 1.0;
    for (int k = 0; k < ncol; ++k) {
        int p = k; double best = fabs(H[k * ncol + k]);
        for (int i = k + 1; i < ncol; ++i) if (fabs(H[i * ncol + k]) > best) { best = fabs(H[i * ncol + k]); p = i; }
        // the reference's LU would divide by zero here (a coordinate that no perturbation touches)
        if (!(best > 0.0)) return fail("od_bundle: Σηηᵀ is singular (some coordinate is never perturbed)");
        if (p != k) for (int j = 0; j < ncol; ++j) { double t = H[k * ncol + j]; H[k * ncol + j] = H[p * ncol + j]; H[p * ncol + j] = t; t = Hi[k * ncol + j]; Hi[k * ncol + j] = Hi[p * ncol + j]; Hi[p * ncol + j] = t; }
        const double inv = 1.0 / H[k * ncol + k];
        for (int j = 0; j < ncol; ++j) { H[k * ncol + j] *= inv; Hi[k * ncol + j] *= inv; }
        for (int i = 0; i < ncol; ++i) if (i != k) {
            const double l = H[i * ncol + k];
            if (l != 0.0) for (int j = 0; j < ncol; ++j) { H[i * ncol + j] -= l * H[k * ncol + j]; Hi[i * ncol + j] -= l * Hi[k * ncol + j]; }
        }
    }
    return 0;
}

int od_bundle_solve_device(od_handle* hd, int B, int N, const double* eta, const double* q1, const double* q2, const double* u,
                           int stride_q, int stride_u, long long p0, long long P, double* feta, int32_t* st_work) {
    if (!hd) return fail("null handle");
    Dims d; dims_of(hd->model, &d);
    if (hd->model == OD_ROCKET) return fail("od_bundle: contact models only");
    if (P <= 0) return 0;
    if (N <= 0 || !eta || !feta || !st_work || p0 < 0 || p0 + P > (long long)B * (N + 1)) return fail("od_bundle_solve_device: bad arguments");
    OD_CUDA(cudaSetDevice(hd->device));
    StepArgs a; memset(&a, 0, sizeof(a));
    a.B = (int)P; a.q1 = q1; a.q2 = q2; a.u = u; a.in_stride_q = stride_q > 0 ? stride_q : d.nq; a.in_stride_u = stride_u > 0 ? stride_u : d.nu;
    a.q3 = feta + (size_t)p0 * d.nq; a.out_stride_q3 = d.nq; a.status = st_work + p0; a.want_eval = 1; a.want_grad = 0; a.eta = eta; a.n_eta = N; a.eta_i0 = p0;
    return launch_step(hd, a);
}

int od_bundle_fit_device(od_handle* hd, int B, int N, const double* eta, const double* Hinv, const double* feta, const int32_t* st_work,
                         double* dz, int32_t* status) {
    if (!hd) return fail("null handle");
    Dims d; dims_of(hd->model, &d);
    if (B <= 0) return 0;
    if (N <= 0 || !eta || !Hinv || !feta || !st_work || !dz) return fail("od_bundle_fit_device: bad arguments");
    OD_CUDA(cudaSetDevice(hd->device));
    bundle_fit_kernel<<<B, ((d.nq * (2 * d.nq + d.nu) + 31) / 32) * 32, 0, hd->stream>>>(B, N, d.nq, 2 * d.nq + d.nu, feta, eta, Hinv, st_work, dz, status);
    OD_CUDA(cudaGetLastError());
    hd->launches++;
    return 0;
}

int od_bundle_batch(od_handle* hd, int B, int N, const double* eta, const double* q1, const double* q2, const double* u, double* dz, int32_t* status) {
    if (!hd) return fail("null handle");
    Dims d; dims_of(hd->model, &d);
    if (hd->model == OD_ROCKET) return fail("od_bundle_batch: contact models only");
    if (B <= 0) return 0;
    if (N <= 0 || !eta || !dz) return fail("od_bundle_batch: need N > 0, eta and dz");
    const int nq = d.nq, nu = d.nu, ncol = 2 * nq + nu;
    if (ncol > 16) return fail("od_bundle_batch: 2nq+nu > 16 unsupported");
    double Hi[16 * 16];
    if (od_bundle_prepare(ncol, N, eta, Hi)) {
        if (status) for (int b = 0; b < B; ++b) status[b] = 8;
        return 1;
    }
    OD_CUDA(cudaSetDevice(hd->device));
    const size_t P = (size_t)B * (N + 1);
    const size_t inw = ncol;
    OD_CUDA(hd->in.reserve(sizeof(double) * inw * B));
    OD_CUDA(hd->aux.reserve(sizeof(double) * ((size_t)N * ncol + ncol * ncol)));
    OD_CUDA(hd->aux2.reserve(sizeof(double) * P * nq + sizeof(int32_t) * P));
    OD_CUDA(hd->out.reserve(sizeof(double) * (size_t)B * nq * ncol));
    OD_CUDA(hd->st.reserve(sizeof(int32_t) * B));
    double* din = (double*)hd->in.p;
    double *d_q1 = din, *d_q2 = din + (size_t)nq * B, *d_u = din + (size_t)2 * nq * B;
    double* d_eta = (double*)hd->aux.p; double* d_Hi = d_eta + (size_t)N * ncol;
    double* d_f = (double*)hd->aux2.p; int32_t* d_st = (int32_t*)(d_f + P * nq);
    OD_CUDA(cudaMemcpyAsync(d_q1, q1, sizeof(double) * nq * B, cudaMemcpyHostToDevice, hd->stream));
    OD_CUDA(cudaMemcpyAsync(d_q2, q2, sizeof(double) * nq * B, cudaMemcpyHostToDevice, hd->stream));
    OD_CUDA(cudaMemcpyAsync(d_u, u, sizeof(double) * nu * B, cudaMemcpyHostToDevice, hd->stream));
    OD_CUDA(cudaMemcpyAsync(d_eta, eta, sizeof(double) * N * ncol, cudaMemcpyHostToDevice, hd->stream));
    OD_CUDA(cudaMemcpyAsync(d_Hi, Hi, sizeof(double) * ncol * ncol, cudaMemcpyHostToDevice, hd->stream));
    if (od_bundle_solve_device(hd, B, N, d_eta, d_q1, d_q2, d_u, 0, 0, 0, (long long)P, d_f, d_st)) return 1;
    if (od_bundle_fit_device(hd, B, N, d_eta, d_Hi, d_f, d_st, (double*)hd->out.p, (int32_t*)hd->st.p)) return 1;
    OD_CUDA(cudaMemcpyAsync(dz, hd->out.p, sizeof(double) * (size_t)B * nq * ncol, cudaMemcpyDeviceToHost, hd->stream));
    if (status) OD_CUDA(cudaMemcpyAsync(status, hd->st.p, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, hd->stream));
    OD_CUDA(cudaStreamSynchronize(hd->stream));
    return 0;
}

// OD_ROCKET_PHASED (default 4): 4-lane rocket kernel with the projection in phased blocks of 2 / 3 / 4 warps; 1 = one-warp blocks, no barriers
static int rocket_phased_warps() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("OD_ROCKET_PHASED"); v = e ? atoi(e) : 4; }
    return v;
}

}  // extern "C"

template <int PPB>
static cudaError_t launch_rocket_phased(const RocketArgs& a, cudaStream_t s) {
    constexpr int G = 4;
    constexpr size_t smem = sizeof(double) * PPB * RocketG<G>::WS;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(rocket_kernel_g<G, PPB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    rocket_kernel_g<G, PPB, true><<<(a.B + PPB - 1) / PPB, G * PPB, smem, s>>>(a);
    return cudaGetLastError();
}

extern "C" {

static int launch_rocket(od_handle* hd, RocketArgs& a) {
    if (hd->model != OD_ROCKET) return fail("rocket entry point needs an OD_ROCKET handle");
    if (a.B <= 0) return 0;
    a.h = hd->h; a.u_max = hd->params[0];
    a.opts.r_tol = hd->opts.r_tol; a.opts.kappa_eval_tol = hd->opts.kappa_eval_tol; a.opts.kappa_grad_tol = hd->opts.kappa_grad_tol;
    a.opts.ls_scale = hd->opts.ls_scale; a.opts.max_iter = hd->opts.max_iter; a.opts.max_ls = hd->opts.max_ls;
    // cooperative lanes (register Gauss–Jordan) while the batch cannot fill the machine with one thread per problem; OD_REG=0 or
    // OD_LANES=1 force the thread-per-problem kernel
    const int lanes = lanes_for(a.B);
    if (reg_path() && lanes >= 8) {
        constexpr int G = 8, PPB = 4;
        rocket_kernel_g<G, PPB><<<(a.B + PPB - 1) / PPB, G * PPB, sizeof(double) * PPB * RocketG<G>::WS, hd->stream>>>(a);
    } else if (reg_path() && lanes >= 4 && a.proj && rocket_phased_warps() > 1) {
        // with the projection: 128-thread blocks with a barrier at each phase boundary (rocket.cuh: PHASED) — measured on B200, 8192
        // problems (profiles/r02zl_*): 0.166 -> 0.132 ms; without the projection the one-warp blocks stay (0.0437 vs 0.0478 ms)
        const int w = rocket_phased_warps();
        cudaError_t e = (w == 2) ? launch_rocket_phased<16>(a, hd->stream) : (w == 3) ? launch_rocket_phased<24>(a, hd->stream) : launch_rocket_phased<32>(a, hd->stream);
        if (e != cudaSuccess) return fail("rocket_kernel_g (phased) launch", e);
    } else if (reg_path() && lanes >= 4) {
        constexpr int G = 4, PPB = 8;
        rocket_kernel_g<G, PPB><<<(a.B + PPB - 1) / PPB, G * PPB, sizeof(double) * PPB * RocketG<G>::WS, hd->stream>>>(a);
    } else {
        constexpr int BLOCK = 32;
        rocket_kernel<BLOCK><<<(a.B + BLOCK - 1) / BLOCK, BLOCK, 0, hd->stream>>>(a);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("rocket_kernel launch", e);
    hd->launches++;
    return 0;
}

int od_rocket_batch_device(od_handle* hd, int B, const double* x, const double* u, int proj, double* y, double* dx, double* du, int32_t* status, int32_t* iters) {
    if (!hd) return fail("null handle");
    OD_CUDA(cudaSetDevice(hd->device));
    RocketArgs a; memset(&a, 0, sizeof(a));
    a.B = B; a.x = x; a.u = u; a.y = y; a.dx = dx; a.du = du; a.status = status; a.iters = iters; a.proj = proj; a.want_grad = (dx || du) ? 1 : 0;
    return launch_rocket(hd, a);
}

int od_rocket_batch(od_handle* hd, int B, const double* x, const double* u, int proj, double* y, double* dx, double* du, int32_t* status) {
    if (!hd) return fail("null handle");
    if (hd->model != OD_ROCKET) return fail("od_rocket_batch needs an OD_ROCKET handle");
    if (B <= 0) return 0;
    OD_CUDA(cudaSetDevice(hd->device));
    OD_CUDA(hd->in.reserve(sizeof(double) * 15 * (size_t)B));
    OD_CUDA(hd->out.reserve(sizeof(double) * (12 + 144 + 36) * (size_t)B));
    OD_CUDA(hd->st.reserve(sizeof(int32_t) * B));
    double* d_x = (double*)hd->in.p; double* d_u = d_x + (size_t)12 * B;
    double* d_y = (double*)hd->out.p; double* d_dx = d_y + (size_t)12 * B; double* d_du = d_dx + (size_t)144 * B;
    OD_CUDA(cudaMemcpyAsync(d_x, x, sizeof(double) * 12 * B, cudaMemcpyHostToDevice, hd->stream));
    OD_CUDA(cudaMemcpyAsync(d_u, u, sizeof(double) * 3 * B, cudaMemcpyHostToDevice, hd->stream));
    const bool grad = dx || du;
    if (od_rocket_batch_device(hd, B, d_x, d_u, proj, d_y, grad ? d_dx : nullptr, grad ? d_du : nullptr, (int32_t*)hd->st.p, nullptr)) return 1;
    if (y) OD_CUDA(cudaMemcpyAsync(y, d_y, sizeof(double) * 12 * B, cudaMemcpyDeviceToHost, hd->stream));
    if (dx) OD_CUDA(cudaMemcpyAsync(dx, d_dx, sizeof(double) * 144 * B, cudaMemcpyDeviceToHost, hd->stream));
    if (du) OD_CUDA(cudaMemcpyAsync(du, d_du, sizeof(double) * 36 * B, cudaMemcpyDeviceToHost, hd->stream));
    if (status) OD_CUDA(cudaMemcpyAsync(status, hd->st.p, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, hd->stream));
    OD_CUDA(cudaStreamSynchronize(hd->stream));
    return 0;
}

int od_rocket_projection_batch(od_handle* hd, int B, const double* u, double* up, double* dup, int32_t* status) {
    if (!hd) return fail("null handle");
    if (hd->model != OD_ROCKET) return fail("od_rocket_projection_batch needs an OD_ROCKET handle");
    if (B <= 0) return 0;
    OD_CUDA(cudaSetDevice(hd->device));
    OD_CUDA(hd->in.reserve(sizeof(double) * 3 * (size_t)B));
    OD_CUDA(hd->out.reserve(sizeof(double) * 12 * (size_t)B));
    OD_CUDA(hd->st.reserve(sizeof(int32_t) * B));
    double* d_u = (double*)hd->in.p; double* d_up = (double*)hd->out.p; double* d_dup = d_up + (size_t)3 * B;
    OD_CUDA(cudaMemcpyAsync(d_u, u, sizeof(double) * 3 * B, cudaMemcpyHostToDevice, hd->stream));
    RocketArgs a; memset(&a, 0, sizeof(a));
    a.B = B; a.u = d_u; a.uproj = d_up; a.duproj = d_dup; a.status = (int32_t*)hd->st.p; a.proj = 1; a.proj_only = 1; a.want_grad = dup ? 1 : 0;
    if (launch_rocket(hd, a)) return 1;
    if (up) OD_CUDA(cudaMemcpyAsync(up, d_up, sizeof(double) * 3 * B, cudaMemcpyDeviceToHost, hd->stream));
    if (dup) OD_CUDA(cudaMemcpyAsync(dup, d_dup, sizeof(double) * 9 * B, cudaMemcpyDeviceToHost, hd->stream));
    if (status) OD_CUDA(cudaMemcpyAsync(status, hd->st.p, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, hd->stream));
    OD_CUDA(cudaStreamSynchronize(hd->stream));
    return 0;
}

}  // extern "C"
