"""ctypes binding of librnnt_b200.so -- the C ABI declared in include/rnnt_b200.h.

No pybind / torch extension: plain ``extern "C"`` symbols, raw device pointers
(``tensor.data_ptr()``) and the caller's CUDA stream.  There is no CPU fallback: if the shared
library is missing this raises, it never routes anywhere else.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librnnt_b200.so")

_c_int, _c_void_p, _c_size_t = ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t
_P = _c_void_p  # every device pointer and the stream travel as void*

# name -> (restype, argtypes); mirrors include/rnnt_b200.h one to one
SIGNATURES = {
    "rnntb200_version": (_c_int, []),
    "rnntb200_status_string": (ctypes.c_char_p, [_c_int]),
    "rnntb200_lattice_sweep": (_c_int, [_P, _P, _P, _c_int, _c_int, _c_int, _P, _P, _P, _P, _P]),
    "rnntb200_loss_dense_fwd": (_c_int, [_P, _c_int, _P, _P, _P] + [_c_int] * 5 + [_P] * 6),
    "rnntb200_loss_dense_bwd": (_c_int, [_P, _c_int, _P, _P, _P] + [_c_int] * 5 + [_P] * 6),
    "rnntb200_joint_cg_project_workspace_bytes": (_c_size_t, [_c_int] * 3),
    "rnntb200_joint_cg_project": (_c_int, [_P] * 2 + [_c_int] + [_P] * 2 + [_c_int] * 5 + [_P] * 3 + [_c_size_t, _P]),
    "rnntb200_joint_cg_project_bwd_workspace_bytes": (_c_size_t, [_c_int] * 3),
    "rnntb200_joint_cg_project_bwd": (_c_int, [_P] * 2 + [_c_int] + [_P] * 3 + [_c_int] * 5 + [_P] * 5 + [_c_size_t, _c_int, _P]),
    "rnntb200_joint_cg_factors_bytes": (_c_size_t, [_c_int] * 4),
    "rnntb200_joint_cg_fwd": (_c_int, [_P] * 5 + [_c_int] * 5 + [_P] * 6 + [_c_size_t, _P]),
    "rnntb200_joint_cg_fwd_loss": (_c_int, [_P] * 5 + [_c_int] * 5 + [_P] * 6 + [_c_size_t, _P, _P, ctypes.c_float, _P]),
    "rnntb200_joint_cg_bwd_loss": (_c_int, [_P] * 5 + [_c_int] * 5 + [_P] * 4 + [ctypes.c_float, _P, _P, _c_int, _P, _c_size_t,
                                            _P, _c_size_t, _P]),
    "rnntb200_joint_cg_bwd_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "rnntb200_joint_cg_bwd": (_c_int, [_P] * 5 + [_c_int] * 5 + [_P] * 6 + [_c_int, _P, _c_size_t, _P, _c_size_t, _P]),
    "rnntb200_joint_at_workspace_bytes": (_c_size_t, [_c_int] * 3),
    "rnntb200_joint_at_fwd": (_c_int, [_P] * 4 + [_c_int] + [_P] * 3 + [_c_int] * 6 + [_P] * 6 + [_c_size_t, _P]),
    "rnntb200_dense_logprobs": (_c_int, [_P, _c_int, _P, _P, _P] + [_c_int] * 5 + [_P] * 3),
    "rnntb200_joint_cg_logprobs": (_c_int, [_P] * 5 + [_c_int] * 5 + [_P] * 3 + [_c_size_t, _P]),
    "rnntb200_joint_at_logprobs": (_c_int, [_P] * 4 + [_c_int] + [_P] * 3 + [_c_int] * 6 + [_P] * 3 + [_c_size_t, _P]),
    "rnntb200_comm_buffer_bytes": (_c_size_t, [_c_size_t, _c_int]),
    "rnntb200_comm_alloc": (_c_int, [_c_size_t, ctypes.POINTER(_c_void_p)]),
    "rnntb200_comm_free": (_c_int, [_P]),
    "rnntb200_comm_export": (_c_int, [_P, ctypes.c_char_p]),
    "rnntb200_comm_import": (_c_int, [ctypes.c_char_p, ctypes.POINTER(_c_void_p)]),
    "rnntb200_comm_release": (_c_int, [_P]),
    "rnntb200_comm_allreduce": (_c_int, [ctypes.POINTER(_c_void_p), _c_int, _c_int, ctypes.POINTER(_c_void_p),
                                         ctypes.POINTER(_c_int), _c_int, _c_size_t, ctypes.c_float, _P]),
    "rnntb200_joint_at_bwd": (_c_int, [_P] * 4 + [_c_int] + [_P] * 3 + [_c_int] * 6 + [_P] * 10 + [_c_size_t, _P]),
}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes library.  Fails loudly when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} is not built. Run `python -m rnntransducer_b200.build` (needs nvcc). "
            "rnntransducer_b200 has no CPU or eager fallback for the fused joint + RNN-T loss path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def status_string(status: int) -> str:
    return load().rnntb200_status_string(int(status)).decode()


def check(status: int, what: str = "") -> None:
    if status != 0:
        raise RuntimeError(f"{what}: {status_string(status)} (status {status})")
