"""Loader / builder of liboptdyn_b200.so (the C ABI in include/optdyn_b200.h).  No CPU fallback: if the library cannot be
loaded or there is no CUDA device, every entry point raises."""
import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
SO_PATH = os.path.join(_PKG, "liboptdyn_b200.so")
_CSRC = os.path.join(_PKG, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
# kernel switches of the shipped build (A/B variants: build(extra_flags=[...], out=...), tools/gpu_validate_variant.sh)
DEFAULT_DEFINES = ["-DOD_EXTRACT_SMEM=1"]

_lib = None

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class od_gather_desc(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("row0", C.c_int64), ("gather_buffers", C.POINTER(C.c_uint64)),
                ("multicast_buffer", C.c_uint64), ("flag_buffers", C.POINTER(C.c_uint64)), ("block_counter", C.c_void_p),
                ("epoch_dev", C.c_void_p), ("epoch", C.c_uint64), ("multicast_flags", C.c_uint64)]


class od_options(C.Structure):
    _fields_ = [("r_tol", C.c_double), ("kappa_eval_tol", C.c_double), ("kappa_grad_tol", C.c_double), ("ls_scale", C.c_double),
                ("max_iter", C.c_int32), ("max_ls", C.c_int32)]


def _sources():
    out = []
    for dp, _, fs in os.walk(_CSRC):
        out += [os.path.join(dp, f) for f in fs]
    out.append(os.path.join(_ROOT, "include", "optdyn_b200.h"))
    return out


def translation_units():
    """optdyn_b200.cu (C ABI, host side, small kernels) + one unit per contact model (csrc/inst/contact_<model>.cu)."""
    inst = os.path.join(_CSRC, "inst")
    return [os.path.join(_CSRC, "optdyn_b200.cu")] + sorted(os.path.join(inst, f) for f in os.listdir(inst) if f.endswith(".cu"))


def build(force=False, verbose=False, extra_flags=None, out=None, jobs=None):
    """nvcc -gencode arch=compute_100a,code=sm_100a … → optimization_dynamics_b200/liboptdyn_b200.so (in-tree).  The translation
    units are compiled side by side (one nvcc process each) and linked into one shared library."""
    from concurrent.futures import ThreadPoolExecutor
    so = out or SO_PATH
    defines = DEFAULT_DEFINES if extra_flags is None else list(extra_flags)
    if not force and os.path.exists(so) and os.path.getmtime(so) >= max(os.path.getmtime(s) for s in _sources() + [os.path.abspath(__file__)]):
        return so
    objdir = os.path.join(_PKG, "build", os.path.basename(so) + ".o")
    os.makedirs(objdir, exist_ok=True)
    units = translation_units()
    objs = [os.path.join(objdir, os.path.basename(u)[:-3] + ".o") for u in units]

    def compile_one(uo):
        u, o = uo
        cmd = ["nvcc"] + NVCC_FLAGS + defines + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", o, u]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return u, res.returncode, res.stdout

    with ThreadPoolExecutor(max_workers=jobs or min(len(units), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, zip(units, objs)))
    for u, rc, log in results:
        if rc != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (u, log))
        if verbose:
            print(log)
    res = subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so] + objs, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout)
    return so


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
        "od_sim_step_batch": (i, [vp, i, i, dp, dp, dp, dp, dp, dp, dp, ip]),
        "od_step_grad_batch_device": (i, [vp, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i]),
        "od_step_grad_packed_device": (i, [vp, i, vp, vp, vp, vp, i, i]),
        "od_step_grad_packed_gather_device": (i, [vp, i, vp, C.c_longlong, i, i, C.POINTER(C.c_uint64), vp, vp]),
        "od_step_grad_packed_gather_sync_device": (i, [vp, i, vp, C.c_longlong, i, i, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), vp, C.c_uint64, vp, vp]),
        "od_step_grad_packed_gather_ex_device": (i, [vp, i, vp, C.POINTER(od_gather_desc), vp, vp]),
        "od_bundle_batch": (i, [vp, i, i, dp, dp, dp, dp, dp, ip]),
        "od_bundle_prepare": (i, [i, i, dp, dp]),
        "od_bundle_solve_device": (i, [vp, i, i, vp, vp, vp, vp, i, i, C.c_longlong, C.c_longlong, vp, vp]),
        "od_bundle_fit_device": (i, [vp, i, i, vp, vp, vp, vp, vp, vp]),
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
                    "od_step_grad_batch", "od_step_grad_packed", "od_sim_step_batch", "od_step_grad_batch_device", "od_step_grad_packed_device", "od_step_grad_packed_gather_device", "od_step_grad_packed_gather_sync_device", "od_step_grad_packed_gather_ex_device", "od_bundle_batch", "od_bundle_prepare", "od_bundle_solve_device", "od_bundle_fit_device",
                    "od_rollout_batch", "od_rollout_batch_device", "od_riccati_batch", "od_riccati_batch_device",
                    "od_rocket_batch", "od_rocket_batch_device", "od_rocket_projection_batch", "od_launch_count", "od_last_error", "od_version"]


def check(rc):
    if rc != 0:
        raise RuntimeError("optdyn_b200: " + lib().od_last_error().decode())
