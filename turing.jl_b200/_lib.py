"""ctypes binding of libtnuts.so -- the same C ABI (include/tnuts.h) Julia binds with ccall.

There is no CPU fallback: if the CUDA library is missing or no sm_100 GPU is present, every
entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TNUTS_LIB: development knob, an alternative build of the same ABI (A/B timing of kernel variants)
LIB_PATH = os.environ.get("TNUTS_LIB") or os.path.join(_HERE, "libtnuts.so")

NSTAT = 9
STAT_NAMES = ("n_steps", "acceptance_rate", "log_density", "hamiltonian_energy",
              "hamiltonian_energy_error", "max_hamiltonian_energy_error", "tree_depth",
              "numerical_error", "step_size")

OK, E_ARG, E_CUDA, E_INIT, E_UNSUPPORTED, E_NCCL = 0, -1, -2, -3, -4, -5

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class TnutsModel(C.Structure):
    _fields_ = [("family", C.c_int32), ("D", C.c_int32), ("N", C.c_int64), ("G", C.c_int32),
                ("X", _dp), ("y", _dp), ("sigma", _dp), ("group", _ip), ("hyper", C.c_double * 8)]


class TnutsConfig(C.Structure):
    _fields_ = [("algorithm", C.c_int32), ("delta", C.c_double), ("max_depth", C.c_int32),
                ("delta_max", C.c_double), ("init_eps", C.c_double), ("n_adapts", C.c_int64),
                ("n_leapfrog", C.c_int32), ("lambda_", C.c_double), ("seed", C.c_uint64),
                ("n_chains", C.c_int32), ("chain_offset", C.c_int32), ("device", C.c_int32),
                ("row_shards", C.c_int32), ("row_rank", C.c_int32), ("strategy", C.c_int32),
                ("metric", C.c_int32), ("reserved", C.c_int32 * 4)]


class TnutsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("tnuts error %d: %s" % (code, msg))
        self.code = code
        self.message = msg


# every symbol include/tnuts.h declares: name -> (restype, argtypes)
_H = C.c_void_p
SYMBOLS = {
    "tnuts_abi_version": (C.c_int, []),
    "tnuts_last_error": (C.c_char_p, [_H]),
    "tnuts_create": (C.c_int, [C.POINTER(TnutsConfig), C.POINTER(_H)]),
    "tnuts_destroy": (C.c_int, [_H]),
    "tnuts_set_model": (C.c_int, [_H, C.POINTER(TnutsModel)]),
    "tnuts_set_model_device": (C.c_int, [_H, C.POINTER(TnutsModel)]),
    "tnuts_dimension": (C.c_int, [_H]),
    "tnuts_num_params": (C.c_int, [_H]),
    "tnuts_init": (C.c_int, [_H, _dp]),
    "tnuts_run": (C.c_int, [_H, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]),
    "tnuts_run_streamed": (C.c_int, [_H, C.c_int64, C.c_int64, C.c_int64, _dp, C.c_int64,
                                     C.POINTER(C.c_int64)]),
    "tnuts_run_streamed_strided": (C.c_int, [_H, C.c_int64, C.c_int64, C.c_int64, _dp, C.c_int64, C.c_int64,
                                             C.POINTER(C.c_int64)]),
    "tnuts_fetch": (C.c_int, [_H, _dp, _dp, _dp, _dp]),
    "tnuts_num_columns": (C.c_int, [_H]),
    "tnuts_fetch_draws": (C.c_int, [_H, _dp]),
    "tnuts_set_option": (C.c_int, [_H, C.c_char_p, C.c_double]),
    "tnuts_get_theta": (C.c_int, [_H, _dp]),
    "tnuts_constrain": (C.c_int, [_H, _dp, C.c_int64, _dp, _dp]),
    "tnuts_get_stepsize": (C.c_int, [_H, _dp]),
    "tnuts_get_inv_metric": (C.c_int, [_H, _dp]),
    "tnuts_logdensity_and_gradient": (C.c_int, [_H, _dp, C.c_int64, _dp, _dp]),
    "tnuts_logdensity": (C.c_int, [_H, _dp, C.c_int64, _dp]),
    "tnuts_leapfrog": (C.c_int, [_H, _dp, _dp, _dp, C.c_double, C.c_int32, C.c_int64, _dp, _dp, _dp]),
    "tnuts_get_state": (C.c_int, [_H, C.c_void_p, C.POINTER(C.c_size_t)]),
    "tnuts_set_state": (C.c_int, [_H, C.c_void_p, C.c_size_t]),
    "tnuts_diagnostics": (C.c_int, [_H, _dp, _dp]),
    "tnuts_num_grad_evals": (C.c_int64, [_H]),
    "tnuts_num_divergences": (C.c_int64, [_H]),
    "tnuts_num_kernel_launches": (C.c_int64, [_H]),
    "tnuts_device_seconds": (C.c_double, [_H]),
    "tnuts_init_device_seconds": (C.c_double, [_H]),
    "tnuts_init_tries_max": (C.c_int, [_H]),
    "tnuts_get_dense_inv_metric": (C.c_int, [_H, _dp]),
    "tnuts_kernel_times": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                     C.POINTER(C.c_int64)]),
    "tnuts_bench_gradient": (C.c_int, [_H, C.c_int32, C.POINTER(C.c_double)]),
    "tnuts_comm_unique_id": (C.c_int, [C.c_void_p]),
    "tnuts_comm_init": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32]),
    "tnuts_comm_allreduce_host": (C.c_int, [_H, _dp, C.c_int64]),
    "tnuts_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "tnuts_host_free": (C.c_int, [C.c_void_p]),
}

_lib = None


def load():
    """dlopen libtnuts.so and declare every prototype.  Raises if the library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TnutsError(E_CUDA, "libtnuts.so is not built (%s); run `python -c 'import "
                             "__graft_entry__ as g; g.build()'` -- there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(handle, rc):
    if rc != OK:
        msg = load().tnuts_last_error(handle)
        raise TnutsError(rc, (msg or b"").decode())
    return rc
