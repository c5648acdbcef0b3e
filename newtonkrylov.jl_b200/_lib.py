"""ctypes binding of libariadne_b200.so (the C ABI in include/ariadne_b200.h).

There is no CPU fallback: if the shared library is missing or does not load, every entry
point raises.  The library itself refuses to create a context without a CUDA device.
"""
import ctypes as C
import os

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libariadne_b200.so")

_lib = None


class AriadneError(RuntimeError):
    """Negative status from the C ABI (CUDA / NCCL / usage error)."""

    def __init__(self, code, msg):
        super().__init__(f"libariadne_b200 error {code}: {msg}")
        self.code = code


_dp = A.c_double_p
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/ariadne_b200.h declares
SIGNATURES = {
    "ak_problem_size": (C.c_int64, [C.POINTER(A.ak_problem)]),
    "ak_abi_version": (C.c_int, []),
    "ak_last_error": (C.c_char_p, []),
    "ak_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "ak_ctx_destroy": (C.c_int, [_vp]),
    "ak_ctx_sync": (C.c_int, [_vp]),
    "ak_ctx_stream": (C.c_uint64, [_vp]),
    "ak_ctx_launch_count": (C.c_int64, [_vp, C.c_int]),
    "ak_timer_start": (C.c_int, [_vp]),
    "ak_timer_stop": (C.c_int, [_vp, _dp]),
    "ak_profile_enable": (C.c_int, [_vp, C.c_int]),
    "ak_profile_read": (C.c_int, [_vp, C.c_int, A.c_int64_p, _dp]),
    "ak_malloc": (C.c_int, [_vp, C.c_int64, C.POINTER(_vp)]),
    "ak_free": (C.c_int, [_vp, _vp]),
    "ak_upload": (C.c_int, [_vp, _vp, _vp, C.c_int64]),
    "ak_download": (C.c_int, [_vp, _vp, _vp, C.c_int64]),
    "ak_host_alloc": (C.c_int, [C.c_int64, C.POINTER(_vp)]),
    "ak_host_free": (C.c_int, [_vp]),
    "ak_halo_pack": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int64]),
    "ak_halo_unpack": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int32]),
    "ak_comm_unique_id": (C.c_int, [C.c_char_p]),
    "ak_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, C.c_char_p]),
    "ak_comm_enable_p2p": (C.c_int, [_vp, C.c_int64]),
    "ak_comm_p2p_enabled": (C.c_int, [_vp]),
    "ak_comm_use_p2p": (C.c_int, [_vp, C.c_int]),
    "ak_comm_rank": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ak_comm_barrier": (C.c_int, [_vp]),
    "ak_residual": (C.c_int, [_vp, C.POINTER(A.ak_problem), _vp, _vp, _dp]),
    "ak_jvp": (C.c_int, [_vp, C.POINTER(A.ak_problem), _vp, _vp, _vp]),
    "ak_jvp_transpose": (C.c_int, [_vp, C.POINTER(A.ak_problem), _vp, _vp, _vp]),
    "ak_jvp_batched": (C.c_int, [_vp, C.POINTER(A.ak_problem), _vp, _vp, C.c_int64, _vp, C.c_int64, C.c_int32]),
    "ak_dot": (C.c_int, [_vp, C.c_int64, _vp, _vp, _dp]),
    "ak_nrm2": (C.c_int, [_vp, C.c_int64, _vp, _dp]),
    "ak_scal": (C.c_int, [_vp, C.c_int64, C.c_double, _vp]),
    "ak_axpy": (C.c_int, [_vp, C.c_int64, C.c_double, _vp, _vp]),
    "ak_axpby": (C.c_int, [_vp, C.c_int64, C.c_double, _vp, C.c_double, _vp]),
    "ak_copy": (C.c_int, [_vp, C.c_int64, _vp, _vp]),
    "ak_fill": (C.c_int, [_vp, C.c_int64, _vp, C.c_double]),
    "ak_ref": (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_double, C.c_double]),
    "ak_divcopy": (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_double]),
    "ak_krylov_default_opts": (None, [C.POINTER(A.ak_krylov_opts)]),
    "ak_krylov_create": (C.c_int, [_vp, C.c_int32, C.c_int64, C.c_int32, C.c_int64, C.POINTER(_vp)]),
    "ak_krylov_destroy": (C.c_int, [_vp]),
    "ak_krylov_solve": (C.c_int, [_vp, C.POINTER(A.ak_problem), _vp, _vp, C.POINTER(A.ak_krylov_opts),
                                  C.POINTER(A.ak_krylov_stats), _dp, C.c_int64]),
    "ak_krylov_x": (_vp, [_vp]),
    "ak_krylov_basis": (C.c_int, [_vp, C.c_int64, C.POINTER(_vp), _dp, A.c_int64_p]),
    "ak_precond_apply": (C.c_int, [_vp, C.POINTER(A.ak_problem), _vp, C.c_int32, C.c_int32, _vp, _vp]),
    "ak_newton_default_opts": (None, [C.POINTER(A.ak_newton_opts)]),
    "ak_newton_solve": (C.c_int, [_vp, C.POINTER(A.ak_problem), _vp, _vp, C.POINTER(A.ak_newton_opts),
                                  C.POINTER(A.ak_newton_stats), _dp, A.c_int64_p, _dp, C.c_int32,
                                  A.NEWTON_CALLBACK, _vp]),
    "ak_newton_solve_host": (C.c_int, [_vp, C.POINTER(A.ak_problem), _vp, _vp, C.POINTER(A.ak_newton_opts),
                                       C.POINTER(A.ak_newton_stats), _dp, A.c_int64_p, C.c_int32]),
    "ak_forcing_ew": (C.c_double, [C.c_double] * 6),
    "ak_implicit_solve": (C.c_int, [_vp, C.POINTER(A.ak_problem), _vp, C.c_int32, C.POINTER(A.ak_newton_opts),
                                    A.c_int32_p, A.c_int64_p, A.c_int32_p]),
}


def load():
    """dlopen the CUDA library (once) and declare every prototype.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (or newtonkrylov.jl_b200/build.py). "
            "This package has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.ak_abi_version() != A.ABI_VERSION:
        raise ImportError("libariadne_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def last_error():
    return load().ak_last_error().decode(errors="replace")


def check(rc):
    """Raise on a negative status; return the (non-negative) numerical flags otherwise."""
    if rc < 0:
        raise AriadneError(rc, last_error())
    return rc
