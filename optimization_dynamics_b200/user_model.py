"""User-specified contact models on the GPU — the successor of the reference's per-model `codegen.jl` +
`deps/build.jl` (reference deps/build.jl:27-49): a model written as a specification file (format: tools/codegen/examples/particle_spec.py)
goes through the sympy generator and is compiled with the SAME solver templates and launch heuristics as the shipped models into
its own shared library, next to liboptdyn_b200.so — no edit of the package's sources.

    so = build_user_model("my_model_spec.py")                   # generator + nvcc (sm_100a), in-tree under user_models/
    dyn = UserModelDynamics(so, h=0.05, κ_eval_tol=1e-4, κ_grad_tol=1e-3, friction=[0.5])
    q3, dq1, dq2, du, status = dyn.step_grad_batch(q1, q2, u)   # one launch for the batch

No CPU fallback: the library's entry points fail without a CUDA device."""
import ctypes as C
import importlib.util
import os
import subprocess
import sys

import numpy as np

from . import _lib

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
USER_DIR = os.path.join(_PKG, "user_models")


def _spec_name(spec_path):
    gen_dir = os.path.join(_ROOT, "tools", "codegen")
    if gen_dir not in sys.path:
        sys.path.insert(0, gen_dir)
    spec = importlib.util.spec_from_file_location("od_user_spec", spec_path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.model()["name"]


def build_user_model(spec_path, out_dir=None, force=False):
    """Specification file → model_<name>.cuh (tools/codegen/gen_models.py --spec) → libodmodel_<name>.so (nvcc, sm_100a)."""
    out_dir = out_dir or USER_DIR
    os.makedirs(out_dir, exist_ok=True)
    name = _spec_name(spec_path)
    hdr = os.path.join(out_dir, "model_%s.cuh" % name)
    so = os.path.join(out_dir, "libodmodel_%s.so" % name)
    unit = os.path.join(_PKG, "csrc", "user_model_unit.cu")
    deps = [spec_path, unit, os.path.join(_PKG, "csrc", "contact_ip.cuh"), os.path.join(_PKG, "csrc", "group_gj.cuh"), os.path.join(_PKG, "csrc", "launch.cuh")]
    if not force and os.path.exists(so) and os.path.getmtime(so) >= max(os.path.getmtime(d) for d in deps):
        return so
    subprocess.check_call([sys.executable, os.path.join(_ROOT, "tools", "codegen", "gen_models.py"), "--spec", spec_path, "--out", out_dir],
                          stdout=subprocess.DEVNULL)
    traits = "".join(p.capitalize() for p in name.split("_")) + "Model"
    cmd = ["nvcc"] + _lib.NVCC_FLAGS + _lib.DEFAULT_DEFINES + ["-shared", "-DOD_USER_MODEL_HEADER=\"%s\"" % hdr, "-DOD_USER_MODEL=%s" % traits, "-o", so, unit]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed on the user model:\n" + res.stdout)
    return so


class UserModelDynamics:
    """The batched step + IFT-gradient of a user model (same outputs and layouts as ImplicitDynamics.step_grad_batch)."""

    def __init__(self, so_path, h, r_tol=1.0e-8, κ_eval_tol=1.0e-6, κ_grad_tol=1.0e-6, friction=(), device=0, max_iter=100, max_ls=25):
        if not os.path.exists(so_path):
            raise RuntimeError("user-model library %s is not built (build_user_model); there is no CPU fallback" % so_path)
        self.L = C.CDLL(so_path)
        self.L.odu_last_error.restype = C.c_char_p
        nq, nu, nf = C.c_int(), C.c_int(), C.c_int()
        self.L.odu_dims(C.byref(nq), C.byref(nu), C.byref(nf))
        self.nq, self.nu, self.nf = nq.value, nu.value, nf.value
        self.h, self.device = float(h), device
        self.friction = np.ascontiguousarray(friction, dtype=np.float64)
        self.opts = _lib.od_options(r_tol, κ_eval_tol, κ_grad_tol, 0.5, max_iter, max_ls)
        self.L.odu_step_grad_packed.argtypes = [C.c_int, _lib.c_double_p, _lib.c_double_p, _lib.c_int32_p, C.c_double, _lib.c_double_p, C.c_int,
                                                C.POINTER(_lib.od_options), C.c_int]
        self.L.odu_step_grad_packed_device.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, _lib.c_double_p, C.c_int,
                                                       C.POINTER(_lib.od_options), C.c_int, C.c_int, C.c_void_p]

    def _check(self, rc):
        if rc:
            raise RuntimeError("optdyn_b200 user model: " + self.L.odu_last_error().decode())

    def step_grad_batch(self, q1, q2, u):
        """q3 [B,nq], ∂q3/∂q1, ∂q3/∂q2 [B,nq,nq], ∂q3/∂u1 [B,nq,nu] (row = q3 component), status [B]."""
        nq, nu = self.nq, self.nu
        q1 = np.asarray(q1, dtype=np.float64).reshape(-1, nq); B = q1.shape[0]
        xin = np.ascontiguousarray(np.concatenate([q1, np.asarray(q2, dtype=np.float64).reshape(B, nq), np.asarray(u, dtype=np.float64).reshape(B, nu)], axis=1))
        outw = nq + nq * (2 * nq + nu)
        out = np.empty((B, outw)); st = np.empty(B, dtype=np.int32)
        fr = self.friction
        self._check(self.L.odu_step_grad_packed(B, xin.ctypes.data_as(_lib.c_double_p), out.ctypes.data_as(_lib.c_double_p), st.ctypes.data_as(_lib.c_int32_p),
                                                self.h, fr.ctypes.data_as(_lib.c_double_p) if fr.size else None, int(fr.size), C.byref(self.opts), self.device))
        from .device import unpack_outputs
        q3, d1, d2, du = unpack_outputs(out, nq, nu)
        return q3, d1, d2, du, st

    def step_grad_packed_device(self, xin, out, status, iters=None, stream=None, want_eval=True, want_grad=True):
        """torch CUDA tensors (packed rows), asynchronous on `stream` (a cudaStream_t value; None = the legacy default stream)."""
        fr = self.friction
        self._check(self.L.odu_step_grad_packed_device(xin.shape[0], C.c_void_p(xin.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(status.data_ptr()),
                                                       C.c_void_p(iters.data_ptr()) if iters is not None else None, self.h,
                                                       fr.ctypes.data_as(_lib.c_double_p) if fr.size else None, int(fr.size), C.byref(self.opts),
                                                       int(want_eval), int(want_grad), C.c_void_p(stream) if stream else None))
        return out, status
