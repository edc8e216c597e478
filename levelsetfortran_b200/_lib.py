"""ctypes loader of liblsf_b200.so (CUDA kernels + C ABI, include/lsf_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present the
compute entry points raise.  Build with `python -m levelsetfortran_b200.build` (or
__graft_entry__.build()).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LSF_LIB_PATH", os.path.join(_HERE, "liblsf_b200.so"))   # override: tuning variants only

c_double_p = C.POINTER(C.c_double)
c_i32_p = C.POINTER(C.c_int32)
c_int_p = C.POINTER(C.c_int)

LSF_OK, LSF_NAN = 0, 1
LSF_ERR_CUDA, LSF_ERR_ARG, LSF_ERR_BAND_ON_BOUNDARY, LSF_ERR_TIMEOUT, LSF_ERR_NODE_OFF_GRID = -1, -2, -3, -4, -5
IPC_HANDLE_BYTES = 64
ARITH_FAST, ARITH_EXACT, ARITH_AUTO = 0, 1, 2
SCHED_MARCH, SCHED_PLANE = 0, 1
MINMAX_LIST, MINMAX_MARCH = 0, 1
PREC_F64, PREC_F32 = 0, 1

# every symbol include/lsf_b200.h declares: name -> (restype, argtypes)
_I, _D, _V = C.c_int, C.c_double, C.c_void_p
SYMBOLS = {
    "lsf_init": (_I, [_I]),
    "lsf_finalize": (_I, []),
    "lsf_last_error": (C.c_char_p, []),
    "lsf_set_arith": (_I, [_I]),
    "lsf_last_arith": (_I, []),
    "lsf_set_sched": (_I, [_I]),
    "lsf_set_minmax_algo": (_I, [_I]),
    "lsf_last_minmax_active": (C.c_longlong, []),
    "lsf_set_precision": (_I, [_I]),
    "lsf_set_overlap": (_I, [_I]),
    "lsf_last_timing": (_I, [c_double_p, c_int_p]),
    "lsf_set_profile": (_I, [_I]),
    "lsf_last_sweep_timing": (_I, [c_double_p, c_int_p]),
    "lsf_sign_init": (_I, [c_double_p, _I, _I, _I, c_double_p, _D, c_double_p, _I, c_i32_p, _I] + [_I] * 6),
    "lsf_reinit": (_I, [c_double_p, c_double_p, c_double_p, _I, _I, _I, _I, _D, _D, c_int_p, c_double_p]),
    "lsf_narrowband": (_I, [_I, _I, _I, _D, c_double_p, c_i32_p, c_i32_p]),
    "lsf_minmax": (_I, [c_double_p, c_double_p, c_i32_p, c_i32_p, _I, _I, _I, _I, _D, _D, _D, c_int_p, c_double_p]),
    "lsf_advect_nodes": (_I, [c_double_p, c_i32_p, _I, _I, _I, c_double_p, _D, c_double_p, _I, c_double_p, c_double_p, _I,
                              C.POINTER(C.c_longlong)]),
    "lsf_grid_advect_nodes": (_I, [_V, c_double_p, _D, c_double_p, _I, c_double_p, c_double_p, _I, C.POINTER(C.c_longlong)]),
    "lsf_grid_create": (_I, [C.POINTER(_V), _I, _I, _I]),
    "lsf_grid_create_f32": (_I, [C.POINTER(_V), _I, _I, _I]),
    "lsf_grid_is_f32": (_I, [_V]),
    "lsf_grid_destroy": (_I, [_V]),
    "lsf_grid_fill": (_I, [_V, _D]),
    "lsf_grid_upload": (_I, [_V, _V]),
    "lsf_grid_download": (_I, [_V, _V]),
    "lsf_grid_download_phiN": (_I, [_V, _V]),
    "lsf_grid_device_ptr": (_V, [_V]),
    "lsf_grid_checksum": (_I, [_V, C.POINTER(C.c_uint64)]),
    "lsf_host_register": (_I, [_V, C.c_size_t]),
    "lsf_host_unregister": (_I, [_V]),
    "lsf_write_vti": (_I, [C.c_char_p, c_double_p, _I, _I, _I, c_double_p, _D]),
    "lsf_write_s3d": (_I, [C.c_char_p, _I, _I, _I, _I, c_i32_p, c_i32_p, c_i32_p, c_double_p, c_double_p]),
    "lsf_stl_count": (_I, [C.c_char_p, c_int_p]),
    "lsf_stl_read_triangles": (_I, [C.c_char_p, _I, C.POINTER(C.c_float)]),
    "lsf_stl_dedup": (_I, [C.POINTER(C.c_float), _I, C.POINTER(C.c_float), c_i32_p, c_int_p]),
    "lsf_grid_sign_init": (_I, [_V, c_double_p, _D, c_double_p, _I, c_i32_p, _I] + [_I] * 6),
    "lsf_grid_reinit": (_I, [_V, _I, _D, _D, _D, c_int_p, c_double_p]),
    "lsf_grid_narrowband": (_I, [_V, _D, c_i32_p, c_i32_p]),
    "lsf_grid_minmax": (_I, [_V, _I, _D, _D, _D, c_int_p, c_double_p]),
    "lsf_grid_reinit_rk3": (_I, [_V, _I, _D, _D, _D, c_int_p, c_double_p]),
    "lsf_slab_range": (_I, [_I, _I, _I, c_int_p, c_int_p]),
    "lsf_sgrid_create": (_I, [C.POINTER(_V), _I, _I, _I, _I, _I]),
    "lsf_sgrid_create_f32": (_I, [C.POINTER(_V), _I, _I, _I, _I, _I]),
    "lsf_sgrid_ipc_handle": (_I, [_V, _V]),
    "lsf_sgrid_attach": (_I, [_V, _V]),
    "lsf_sgrid_sync_ghosts": (_I, [_V]),
}

_lib = None


class LsfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"liblsf_b200 error {code}: {msg}")
        self.code = code


def lib():
    """Load the shared library and bind every exported symbol (no device needed for this)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: the CUDA extension is not built "
                              "(python -m levelsetfortran_b200.build); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            if "LSF_LIB_PATH" in os.environ and not hasattr(L, name):
                continue           # a tuning variant built from an older source tree (tools/build_variant.sh)
            f = getattr(L, name)   # AttributeError if the .so does not export it
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> int:
    """Raise on a negative return code; pass 0 / LSF_NAN through."""
    if rc < 0:
        raise LsfError(rc, lib().lsf_last_error().decode(errors="replace"))
    return rc


def last_sweep_timing():
    ms, n = C.c_double(0), C.c_int(0)
    lib().lsf_last_sweep_timing(C.byref(ms), C.byref(n))
    return ms.value, n.value


def last_timing():
    ms, n = C.c_double(0), C.c_int(0)
    lib().lsf_last_timing(C.byref(ms), C.byref(n))
    return ms.value, n.value
