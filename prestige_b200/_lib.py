"""ctypes binding of libprestige_b200.so (the C ABI in include/prestige_b200.h).

There is no CPU path: if the shared library is missing or no CUDA device is
present, the calls fail loudly (ImportError here, PST_ECUDA from pst_create).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PRESTIGE_B200_LIB selects another build of the same library (kernel A/B experiments); never a CPU path
LIB_PATH = os.environ.get("PRESTIGE_B200_LIB") or os.path.join(_HERE, "libprestige_b200.so")

PST_OK, PST_EINVAL, PST_ENOMEM, PST_ECUDA, PST_ENCCL, PST_EOVERFLOW, PST_ESTATE = range(7)
STATUS_NAMES = ["PST_OK", "PST_EINVAL", "PST_ENOMEM", "PST_ECUDA", "PST_ENCCL", "PST_EOVERFLOW", "PST_ESTATE"]
PST_F32, PST_F64, PST_U32, PST_I32, PST_REAL = 0, 1, 2, 3, 15
PST_KEY_LINEAR, PST_KEY_MORTON = 0, 1
PST_PHYS_NONE, PST_PHYS_WCSPH, PST_PHYS_DEM = 0, 1, 2
PST_ARRAY_PERSISTENT, PST_ARRAY_OUTPUT = 1, 2
PST_COMM_ID_BYTES = 128


class PstConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("dim", C.c_int32), ("real", C.c_int32),
                ("key", C.c_int32), ("max_contacts", C.c_int32), ("physics", C.c_uint32), ("reserved", C.c_uint32),
                ("capacity", C.c_uint64), ("ghost_capacity", C.c_uint64), ("lo", C.c_double * 3),
                ("hi", C.c_double * 3), ("cell_size", C.c_double)]


# every symbol include/prestige_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "pst_version", "pst_create", "pst_destroy", "pst_last_error", "pst_stream", "pst_sync", "pst_set_param",
    "pst_get_param", "pst_set_count", "pst_get_count", "pst_array_create", "pst_array", "pst_upload",
    "pst_download", "pst_upload_async", "pst_download_async", "pst_wait_transfers", "pst_host_alloc", "pst_host_free", "pst_build_neighbours", "pst_apply", "pst_dump_pairs",
    "pst_step", "pst_integrate", "pst_get_stat", "pst_kernel_name", "pst_set_option", "pst_comm_unique_id", "pst_comm_init",
    "pst_halo_exchange", "pst_bodies_create", "pst_bodies_setup", "pst_bodies_restore", "pst_bodies_state",
]

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise ImportError with the build hint if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C prestige_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, cp, st = C.c_void_p, C.c_char_p, C.c_int
    lib.pst_version.restype = cp
    lib.pst_create.argtypes = [C.POINTER(PstConfig), C.POINTER(vp)]; lib.pst_create.restype = st
    lib.pst_destroy.argtypes = [vp]; lib.pst_destroy.restype = None
    lib.pst_last_error.argtypes = [vp]; lib.pst_last_error.restype = cp
    lib.pst_stream.argtypes = [vp]; lib.pst_stream.restype = vp
    lib.pst_sync.argtypes = [vp]; lib.pst_sync.restype = st
    lib.pst_set_param.argtypes = [vp, cp, C.c_double]; lib.pst_set_param.restype = st
    lib.pst_get_param.argtypes = [vp, cp, C.POINTER(C.c_double)]; lib.pst_get_param.restype = st
    lib.pst_set_count.argtypes = [vp, C.c_uint64]; lib.pst_set_count.restype = st
    lib.pst_get_count.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]; lib.pst_get_count.restype = st
    lib.pst_array_create.argtypes = [vp, cp, C.c_int, C.c_uint32]; lib.pst_array_create.restype = st
    lib.pst_array.argtypes = [vp, cp, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.pst_array.restype = st
    lib.pst_upload.argtypes = [vp, cp, vp, C.c_size_t]; lib.pst_upload.restype = st
    lib.pst_download.argtypes = [vp, cp, vp, C.c_size_t]; lib.pst_download.restype = st
    lib.pst_upload_async.argtypes = [vp, cp, vp, C.c_size_t]; lib.pst_upload_async.restype = st
    lib.pst_download_async.argtypes = [vp, cp, vp, C.c_size_t]; lib.pst_download_async.restype = st
    lib.pst_wait_transfers.argtypes = [vp]; lib.pst_wait_transfers.restype = st
    lib.pst_host_alloc.argtypes = [C.c_size_t]; lib.pst_host_alloc.restype = vp
    lib.pst_host_free.argtypes = [vp]; lib.pst_host_free.restype = None
    lib.pst_build_neighbours.argtypes = [vp]; lib.pst_build_neighbours.restype = st
    lib.pst_apply.argtypes = [vp, C.POINTER(cp), C.c_int]; lib.pst_apply.restype = st
    lib.pst_dump_pairs.argtypes = [vp, C.c_int, vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]; lib.pst_dump_pairs.restype = st
    lib.pst_step.argtypes = [vp, C.c_double, C.c_int]; lib.pst_step.restype = st
    lib.pst_integrate.argtypes = [vp, C.c_double]; lib.pst_integrate.restype = st
    lib.pst_get_stat.argtypes = [vp, cp, C.POINTER(C.c_double)]; lib.pst_get_stat.restype = st
    lib.pst_set_option.argtypes = [vp, cp, C.c_int]; lib.pst_set_option.restype = st
    lib.pst_kernel_name.argtypes = [vp, cp, C.c_char_p, C.c_size_t]; lib.pst_kernel_name.restype = st
    lib.pst_comm_unique_id.argtypes = [vp]; lib.pst_comm_unique_id.restype = st
    lib.pst_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]; lib.pst_comm_init.restype = st
    lib.pst_halo_exchange.argtypes = [vp]; lib.pst_halo_exchange.restype = st
    lib.pst_bodies_create.argtypes = [vp, C.c_uint32]; lib.pst_bodies_create.restype = st
    lib.pst_bodies_setup.argtypes = [vp]; lib.pst_bodies_setup.restype = st
    lib.pst_bodies_restore.argtypes = [vp]; lib.pst_bodies_restore.restype = st
    lib.pst_bodies_state.argtypes = [vp, cp, vp, C.c_size_t, C.c_int]; lib.pst_bodies_state.restype = st
    _lib = lib
    return lib


class PstError(RuntimeError):
    def __init__(self, status: int, message: str):
        self.status = status
        name = STATUS_NAMES[status] if 0 <= status < len(STATUS_NAMES) else str(status)
        super().__init__(f"{name}: {message}")
