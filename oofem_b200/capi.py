"""ctypes binding of liboofem_b200.so (the C ABI declared in include/oofem_b200.h).

This is the only module that touches the shared library.  It fails loudly when the
library is missing or cannot be loaded -- there is no CPU fallback anywhere in the
package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboofem_b200.so")

OK, ENODEVICE, ECUDA, EINVAL, ESTRUCT, EZERODIAG, ECAPACITY, ENCCL = 0, -1, -2, -3, -4, -5, -6, -7
LSPACE, LTRSPACE = 1, 2
MAT_ISOLE, MAT_MISES = 1, 2
MATPARAM_STRIDE = 8
MISES_STATE_DOUBLES = 29
PRECOND_VOID, PRECOND_DIAG = 0, 1

# every exported symbol of include/oofem_b200.h: name -> (restype, argtypes)
_vp, _i32, _i64, _dbl, _int = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_int
_pp = C.POINTER(C.c_void_p)
SYMBOLS = {
    "ob200_last_error": (C.c_char_p, []),
    "ob200_version": (C.c_char_p, []),
    "ob200_context_create": (_int, [_int, _pp]),
    "ob200_context_destroy": (None, [_vp]),
    "ob200_context_sync": (_int, [_vp]),
    "ob200_context_stream": (_vp, [_vp]),
    "ob200_context_launch_count": (_i64, [_vp]),
    "ob200_context_set_profiling": (_int, [_vp, _int]),
    "ob200_context_profile_reset": (_int, [_vp]),
    "ob200_context_profile_report": (_int, [_vp, C.c_char_p, _i64]),
    "ob200_malloc": (_int, [_vp, _i64, _pp]),
    "ob200_free": (_int, [_vp, _vp]),
    "ob200_memcpy_h2d": (_int, [_vp, _vp, _vp, _i64]),
    "ob200_memcpy_d2h": (_int, [_vp, _vp, _vp, _i64]),
    "ob200_memset": (_int, [_vp, _vp, _int, _i64]),
    "ob200_flush_l2": (_int, [_vp]),
    "ob200_csr_create": (_int, [_vp, _pp]),
    "ob200_csr_destroy": (None, [_vp]),
    "ob200_csr_build_structure": (_int, [_vp, _i32, _i64, _i32, _vp, _int]),
    "ob200_csr_rows": (_i32, [_vp]),
    "ob200_csr_nnz": (_i64, [_vp]),
    "ob200_csr_get_structure": (_int, [_vp, _vp, _vp, _int]),
    "ob200_csr_get_values": (_int, [_vp, _vp, _int]),
    "ob200_csr_set_values": (_int, [_vp, _vp, _int]),
    "ob200_csr_device_arrays": (_int, [_vp, _pp, _pp, _pp]),
    "ob200_csr_zero": (_int, [_vp]),
    "ob200_csr_scale": (_int, [_vp, _dbl]),
    "ob200_csr_assemble": (_int, [_vp, _i64, _i32, _vp, _vp, _int]),
    "ob200_csr_assemble_rect": (_int, [_vp, C.c_int32, C.c_int32, _vp, _vp, _vp, _int]),
    "ob200_csr_times": (_int, [_vp, _vp, _vp, _int]),
    "ob200_csr_times_t": (_int, [_vp, _vp, _vp, _int]),
    "ob200_csr_at": (_int, [_vp, _i32, _i32, C.POINTER(_dbl)]),
    "ob200_csr_version": (_i64, [_vp]),
    "ob200_elemset_create": (_int, [_vp, _int, _i64, _vp, _i64, _vp, _vp, _i32, _vp, _vp, _i32, _int, _pp]),
    "ob200_elemset_create_nodal": (_int, [_vp, _int, _i64, _vp, _i64, _vp, _vp, _i32, _vp, _vp, _i32, _int, _pp]),
    "ob200_elemset_destroy": (None, [_vp]),
    "ob200_elemset_size": (_i64, [_vp]),
    "ob200_elemset_stiffness": (_int, [_vp, _vp, _int]),
    "ob200_elemset_internal_forces": (_int, [_vp, _vp, _vp, _vp, _vp, _int]),
    "ob200_elemset_bind": (_int, [_vp, _vp]),
    "ob200_elemset_assemble_stiffness": (_int, [_vp, _vp]),
    "ob200_elemset_assemble_internal_forces": (_int, [_vp, _vp, _vp, _vp, _int]),
    "ob200_elemset_assemble_extrapolated_forces": (_int, [_vp, _vp, _vp, _int]),
    "ob200_elemset_commit": (_int, [_vp]),
    "ob200_elemset_get_state": (_int, [_vp, _vp, _int]),
    "ob200_elemset_set_state": (_int, [_vp, _vp, _int]),
    "ob200_cg_solve": (_int, [_vp, _vp, _vp, _int, _int, _dbl, C.POINTER(_int), C.POINTER(_dbl), _int]),
    "ob200_csr_spmv_layout": (_int, [_vp, _vp]),
    "ob200_comm_unique_id": (_int, [_vp]),
    "ob200_comm_create": (_int, [_vp, _int, _int, _vp, _pp]),
    "ob200_comm_destroy": (None, [_vp]),
    "ob200_comm_set_halo": (_int, [_vp, _i32, _int, _vp, _vp, _vp, _vp]),
    "ob200_comm_exchange_add": (_int, [_vp, _vp]),
    "ob200_comm_p2p_export": (_int, [_vp, C.c_int64, _vp]),
    "ob200_comm_p2p_open": (_int, [_vp, _vp]),
    "ob200_comm_p2p_enabled": (_int, [_vp]),
    "ob200_comm_p2p_disable": (_int, [_vp]),
    "ob200_cg_solve_dist": (_int, [_vp, _vp, _vp, _vp, _int, _int, _dbl, C.POINTER(_int), C.POINTER(_dbl), _int]),
}

_lib = None


class OofemB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[ob200 error {code}] {msg}")
        self.code = code


def lib():
    """Load the shared library (once).  Raises if it has not been built: the CUDA
    extension is the product; nothing in this package works without it."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OofemB200Error(ENODEVICE, f"{LIB_PATH} not built; run `python -m oofem_b200.build` "
                                            "(or __graft_entry__.build()). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> int:
    if rc < 0:
        raise OofemB200Error(rc, lib().ob200_last_error().decode())
    return rc


def ptr(a):
    """Device or host address of numpy arrays / torch tensors / raw ints (None -> NULL)."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())        # torch tensor


def on_device(a) -> int:
    """1 for torch CUDA tensors, 0 for numpy arrays / CPU tensors."""
    if a is None or isinstance(a, np.ndarray):
        return 0
    return 1 if getattr(a, "is_cuda", False) else 0


class Context:
    """ob200_context: one per process / GPU."""

    def __init__(self, device: int = 0):
        self.h = C.c_void_p()
        check(lib().ob200_context_create(device, C.byref(self.h)))
        self.device = device

    def sync(self):
        check(lib().ob200_context_sync(self.h))

    @property
    def stream(self) -> int:
        return lib().ob200_context_stream(self.h) or 0

    @property
    def launches(self) -> int:
        return lib().ob200_context_launch_count(self.h)

    def flush_l2(self):
        check(lib().ob200_flush_l2(self.h))

    def set_profiling(self, enable: bool):
        check(lib().ob200_context_set_profiling(self.h, 1 if enable else 0))

    def profile_reset(self):
        check(lib().ob200_context_profile_reset(self.h))

    def profile_report(self) -> dict:
        """{kernel name: (total ms, launches)} measured with CUDA events around every launch."""
        buf = C.create_string_buffer(1 << 16)
        check(lib().ob200_context_profile_report(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, n = line.split("\t")
            out[name] = (float(ms), int(n))
        return out

    def close(self):
        if self.h:
            lib().ob200_context_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
