// TEST INFRASTRUCTURE — not part of liboptdyn_b200.so.
// Host-tier harness for a model produced by `tools/codegen/gen_models.py --spec …`: the generated header is compiled against the
// product's solver templates (csrc/contact_ip.cuh) and stepped on the CPU — shared-memory-LU templates (reg = 0) and the register
// path (reg = 1).  Built by tests/test_host_logic.py with
//   nvcc -O2 -std=c++17 -Xcompiler -fPIC -shared -DUSER_MODEL_HEADER='"…/model_<name>.cuh"' -DUSER_MODEL=<Name>Model …
#include <string.h>
#include "../optimization_dynamics_b200/csrc/contact_ip.cuh"
#include USER_MODEL_HEADER

using namespace od;

extern "C" int um_dims(int* nq, int* nu) { *nq = USER_MODEL::NQ; *nu = USER_MODEL::NU; return 0; }

extern "C" int um_step(int B, const double* q1, const double* q2, const double* u, double h, const double* fric, double k_eval, double k_grad,
                       double* q3, double* dq1, double* dq2, double* du, int* status, int reg) {
    typedef USER_MODEL M;
    StepArgs a; memset(&a, 0, sizeof(a));
    a.B = B; a.q1 = q1; a.q2 = q2; a.u = u; a.in_stride_q = M::NQ; a.in_stride_u = M::NU;
    a.q3 = q3; a.dq1 = dq1; a.dq2 = dq2; a.du = du;
    a.out_stride_q3 = M::NQ; a.out_stride_dq = M::NQ * M::NQ; a.out_stride_du = M::NQ * M::NU;
    a.status = status; a.h = h; a.want_eval = 1; a.want_grad = 1;
    for (int k = 0; k < 4; ++k) a.fric[k] = fric ? fric[k] : 0.0;
    a.opts.r_tol = 1e-8; a.opts.kappa_eval_tol = k_eval; a.opts.kappa_grad_tol = k_grad; a.opts.ls_scale = 0.5; a.opts.max_iter = 100; a.opts.max_ls = 25;
    if (reg) { alignas(16) double ws[ContactIP<M, 1, 1, true>::WS]; for (int i = 0; i < B; ++i) contact_step_one<M, 1, 1, true>(a, i, ws, 0, 0u); }
    else { alignas(16) double ws[ContactIP<M>::WS]; for (int i = 0; i < B; ++i) contact_step_one<M, 1, 1>(a, i, ws, 0, 0u); }
    return 0;
}
