// One user-specified contact model as its own shared library (the successor of the reference's per-model
// codegen.jl + deps/build.jl).  `tools/codegen/gen_models.py --spec my_model.py` writes model_<name>.cuh (device code + traits struct);
// this unit instantiates the SAME solver templates / launch heuristics the shipped models use (launch.cuh) for that struct:
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared \
//          -DOD_USER_MODEL_HEADER='"…/model_<name>.cuh"' -DOD_USER_MODEL=<Name>Model -o libodmodel_<name>.so csrc/user_model_unit.cu
// (optimization_dynamics_b200/user_model.py: build_user_model() does exactly that and UserModelDynamics binds the result).
// No CPU fallback: the entry points fail without a CUDA device.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/optdyn_b200.h"
#include "launch.cuh"
#include OD_USER_MODEL_HEADER

using namespace od;
typedef OD_USER_MODEL UM;
typedef ContactIP<UM, 1, 1, false> UIP;
// cooperative register path for cone models whose packed row fits the staging area; 8 lanes only for systems of 8+ unknowns
static constexpr bool U_REGOK = (UIP::NCONE > 0) && !UM::ROBUST_IFT && (UIP::NOUT <= UIP::NR * (((UIP::NR + UIP::NTP + 1) / 2) * 2)) && (UIP::NR >= 4);
static constexpr bool U_WIDE = U_REGOK && UIP::NR >= 8 && UIP::NTP >= 8;

static thread_local char u_err[256] = "";
static int ufail(const char* what, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess) snprintf(u_err, sizeof(u_err), "%s: %s", what, cudaGetErrorString(e)); else snprintf(u_err, sizeof(u_err), "%s", what);
    return 1;
}

extern "C" {
const char* odu_last_error(void) { return u_err; }
int odu_dims(int* nq, int* nu, int* nfric) { if (nq) *nq = UM::NQ; if (nu) *nu = UM::NU; if (nfric) *nfric = UM::NF; return 0; }

// Packed rows in [q1 | q2 | u] → out [q3 | ∂q3/∂q1 | ∂q3/∂q2 | ∂q3/∂u1] (blocks column-major), device pointers, asynchronous on `stream`.
int odu_step_grad_packed_device(int B, const double* in, double* out, int32_t* status, int32_t* iters, double h, const double* fric, int nfric,
                                const od_options* opts, int want_eval, int want_grad, void* stream) {
    if (B <= 0) return 0;
    if (!in || !out || !opts || !(h > 0.0)) return ufail("odu_step_grad_packed_device: bad arguments");
    constexpr int NQ = UM::NQ, NU = UM::NU, inw = 2 * NQ + NU, outw = NQ + NQ * inw;
    StepArgs a; memset(&a, 0, sizeof(a));
    a.B = B; a.q1 = in; a.q2 = in + NQ; a.u = in + 2 * NQ; a.in_stride_q = inw; a.in_stride_u = inw; a.in_packed = 1;
    a.q3 = out; a.dq1 = want_grad ? out + NQ : nullptr; a.dq2 = out + NQ + NQ * NQ; a.du = out + NQ + 2 * NQ * NQ;
    a.out_stride_q3 = outw; a.out_stride_dq = outw; a.out_stride_du = outw;
    a.status = status; a.iters = iters; a.want_eval = want_eval; a.want_grad = want_grad;
    a.packed_out = ((reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
    a.h = h;
    for (int k = 0; k < 4; ++k) a.fric[k] = (fric && k < nfric) ? fric[k] : 0.0;
    a.opts.r_tol = opts->r_tol; a.opts.kappa_eval_tol = opts->kappa_eval_tol; a.opts.kappa_grad_tol = opts->kappa_grad_tol;
    a.opts.ls_scale = opts->ls_scale; a.opts.max_iter = opts->max_iter; a.opts.max_ls = opts->max_ls;
    cudaError_t e = launch_contact<UM, U_WIDE, U_REGOK>(a, (cudaStream_t)stream);
    if (e != cudaSuccess) return ufail("contact_step_kernel<user model> launch", e);
    return 0;
}

// Host pointers: H2D, launch, D2H, synchronise.
int odu_step_grad_packed(int B, const double* in, double* out, int32_t* status, double h, const double* fric, int nfric, const od_options* opts, int device) {
    if (B <= 0) return 0;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return ufail("no CUDA device (user-model libraries have no CPU fallback)", e);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return ufail("cudaSetDevice", e);
    constexpr size_t inw = 2 * UM::NQ + UM::NU, outw = UM::NQ + UM::NQ * inw;
    double *din = nullptr, *dout = nullptr; int32_t* dst = nullptr;
    int rc = 1;
    if ((e = cudaMalloc(&din, sizeof(double) * inw * B)) == cudaSuccess && (e = cudaMalloc(&dout, sizeof(double) * outw * B)) == cudaSuccess &&
        (e = cudaMalloc(&dst, sizeof(int32_t) * B)) == cudaSuccess && (e = cudaMemcpy(din, in, sizeof(double) * inw * B, cudaMemcpyHostToDevice)) == cudaSuccess) {
        if (!odu_step_grad_packed_device(B, din, dout, dst, nullptr, h, fric, nfric, opts, 1, 1, nullptr) &&
            (e = cudaMemcpy(out, dout, sizeof(double) * outw * B, cudaMemcpyDeviceToHost)) == cudaSuccess &&
            (!status || (e = cudaMemcpy(status, dst, sizeof(int32_t) * B, cudaMemcpyDeviceToHost)) == cudaSuccess)) rc = 0;
    }
    if (rc && e != cudaSuccess) ufail("odu_step_grad_packed", e);
    cudaFree(din); cudaFree(dout); cudaFree(dst);
    return rc;
}
}  // extern "C"
