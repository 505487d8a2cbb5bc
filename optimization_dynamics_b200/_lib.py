"""Loader / builder of liboptdyn_b200.so (the C ABI in include/optdyn_b200.h).  No CPU fallback: if the library cannot be
loaded or there is no CUDA device, every entry point raises."""
import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
SO_PATH = os.path.join(_PKG, "liboptdyn_b200.so")
_SRC = os.path.join(_PKG, "csrc", "optdyn_b200.cu")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]

_lib = None

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class od_options(C.Structure):
    _fields_ = [("r_tol", C.c_double), ("kappa_eval_tol", C.c_double), ("kappa_grad_tol", C.c_double), ("ls_scale", C.c_double),
                ("max_iter", C.c_int32), ("max_ls", C.c_int32)]


def _sources():
    out = []
    for dp, _, fs in os.walk(os.path.join(_PKG, "csrc")):
        out += [os.path.join(dp, f) for f in fs]
    out.append(os.path.join(_ROOT, "include", "optdyn_b200.h"))
    return out


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a … → optimization_dynamics_b200/liboptdyn_b200.so (in-tree)."""
    if not force and os.path.exists(SO_PATH) and os.path.getmtime(SO_PATH) >= max(os.path.getmtime(s) for s in _sources()):
        return SO_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH, _SRC]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return SO_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("OD_B200_LIB", SO_PATH)      # A/B timing of differently built libraries (tools/micro/ab_time.sh)
    if not os.path.exists(path):
        raise RuntimeError("liboptdyn_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "optimization_dynamics_b200 has no CPU fallback")
    L = C.CDLL(path)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    dp, ip = c_double_p, c_int32_p
    sig = {
        "od_default_options": (i, [i, C.POINTER(od_options)]),
        "od_model_dims": (i, [i, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
        "od_create": (vp, [i, d, C.POINTER(od_options), dp, i, i]),
        "od_destroy": (None, [vp]),
        "od_set_stream": (i, [vp, vp]),
        "od_synchronize": (i, [vp]),
        "od_step_batch": (i, [vp, i, dp, dp, dp, dp, ip]),
        "od_step_grad_batch": (i, [vp, i, dp, dp, dp, dp, dp, dp, dp, ip]),
        "od_step_grad_packed": (i, [vp, i, dp, dp, ip]),
        "od_step_grad_batch_device": (i, [vp, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i]),
        "od_step_grad_packed_device": (i, [vp, i, vp, vp, vp, vp, i, i]),
        "od_step_grad_packed_gather_device": (i, [vp, i, vp, C.c_longlong, i, i, C.POINTER(C.c_uint64), vp, vp]),
        "od_step_grad_packed_gather_sync_device": (i, [vp, i, vp, C.c_longlong, i, i, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), vp, C.c_uint64, vp, vp]),
        "od_bundle_batch": (i, [vp, i, i, dp, dp, dp, dp, dp, ip]),
        "od_riccati_batch": (i, [vp, i, i, dp, dp, dp, dp, dp, dp, d, dp, dp, dp, ip]),
        "od_riccati_batch_device": (i, [vp, i, i, vp, vp, vp, vp, vp, vp, d, vp, vp, vp, vp]),
        "od_rollout_batch": (i, [vp, i, i, dp, dp, i, dp, dp, dp, dp, dp, dp, ip]),
        "od_rollout_batch_device": (i, [vp, i, i, vp, vp, C.c_longlong, vp, vp, vp, vp, vp, vp, vp, vp]),
        "od_rocket_batch": (i, [vp, i, dp, dp, i, dp, dp, dp, ip]),
        "od_rocket_batch_device": (i, [vp, i, vp, vp, i, vp, vp, vp, vp, vp]),
        "od_rocket_projection_batch": (i, [vp, i, dp, dp, dp, ip]),
        "od_launch_count": (C.c_int64, [vp]),
        "od_last_error": (C.c_char_p, []),
        "od_version": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)      # AttributeError if the library does not export a symbol the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = ["od_default_options", "od_model_dims", "od_create", "od_destroy", "od_set_stream", "od_synchronize", "od_step_batch",
                    "od_step_grad_batch", "od_step_grad_packed", "od_step_grad_batch_device", "od_step_grad_packed_device", "od_step_grad_packed_gather_device", "od_step_grad_packed_gather_sync_device", "od_bundle_batch",
                    "od_rollout_batch", "od_rollout_batch_device", "od_riccati_batch", "od_riccati_batch_device",
                    "od_rocket_batch", "od_rocket_batch_device", "od_rocket_projection_batch", "od_launch_count", "od_last_error", "od_version"]


def check(rc):
    if rc != 0:
        raise RuntimeError("optdyn_b200: " + lib().od_last_error().decode())
