"""ctypes binding of libminarrow_b200.so (C ABI: include/minarrow_b200.h).

The shared library is built in-tree by `__graft_entry__.build()` (or `make -C minarrow_b200/csrc`).  There is
no fallback: if the library is missing, or no sm_100 device is usable, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MINARROW_B200_LIB") or os.path.join(_HERE, "libminarrow_b200.so")   # override: tuning builds

c_ctx = C.c_void_p
c_buf = C.c_void_p
c_bits = C.c_void_p
c_sz = C.c_size_t
c_int = C.c_int
c_vp = C.c_void_p
PP = C.POINTER(C.c_void_p)


class Scalar64(C.Union):
    _fields_ = [("i64", C.c_int64), ("u64", C.c_uint64), ("f64", C.c_double)]


class Agg(C.Structure):
    """mnr_agg: null-aware aggregate of one column."""
    _fields_ = [("sum", Scalar64), ("min", Scalar64), ("max", Scalar64), ("count", C.c_uint64)]


# name -> (restype, argtypes).  Must list every symbol declared in include/minarrow_b200.h
# (tests/test_abi.py parses the header and checks both directions).
SIGNATURES = {
    "mnr_abi_version": (c_int, []),
    "mnr_last_error": (C.c_char_p, []),
    "mnr_device_count": (c_int, []),
    "mnr_ctx_create": (c_int, [c_int, PP]),
    "mnr_ctx_create_on_stream": (c_int, [c_int, c_vp, PP]),
    "mnr_ctx_destroy": (None, [c_ctx]),
    "mnr_ctx_synchronize": (c_int, [c_ctx]),
    "mnr_ctx_device": (c_int, [c_ctx]),
    "mnr_ctx_stream": (c_vp, [c_ctx]),
    "mnr_ctx_launch_count": (C.c_uint64, [c_ctx]),
    "mnr_ctx_set_option": (c_int, [c_ctx, C.c_char_p, C.c_int64]),
    "mnr_buf_alloc": (c_int, [c_ctx, c_int, c_sz, PP]),
    "mnr_buf_upload": (c_int, [c_ctx, c_int, c_vp, c_sz, PP]),
    "mnr_buf_upload_async": (c_int, [c_ctx, c_int, c_vp, c_sz, PP]),
    "mnr_buf_wrap": (c_int, [c_ctx, c_int, c_vp, c_sz, PP]),
    "mnr_buf_slice": (c_int, [c_buf, c_sz, c_sz, PP]),
    "mnr_buf_download": (c_int, [c_ctx, c_buf, c_vp]),
    "mnr_buf_len": (c_sz, [c_buf]),
    "mnr_buf_dtype": (c_int, [c_buf]),
    "mnr_buf_device_ptr": (c_vp, [c_buf]),
    "mnr_buf_free": (None, [c_buf]),
    "mnr_bits_alloc": (c_int, [c_ctx, c_sz, PP]),
    "mnr_bits_new_set_all": (c_int, [c_ctx, c_sz, c_int, PP]),
    "mnr_bits_upload": (c_int, [c_ctx, c_vp, c_sz, PP]),
    "mnr_bits_upload_async": (c_int, [c_ctx, c_vp, c_sz, PP]),
    "mnr_bits_wrap": (c_int, [c_ctx, c_vp, c_sz, PP]),
    "mnr_bits_download": (c_int, [c_ctx, c_bits, c_vp]),
    "mnr_bits_len": (c_sz, [c_bits]),
    "mnr_bits_device_ptr": (c_vp, [c_bits]),
    "mnr_bits_free": (None, [c_bits]),
    "mnr_ew_binary": (c_int, [c_ctx, c_int, c_buf, c_buf, c_bits, c_bits, c_int, PP, PP]),
    "mnr_ew_binary_into": (c_int, [c_ctx, c_int, c_buf, c_buf, c_bits, c_bits, c_int, c_buf, c_bits]),
    "mnr_ew_scalar": (c_int, [c_ctx, c_int, c_buf, c_vp, c_int, c_bits, PP, PP]),
    "mnr_ew_scalar_into": (c_int, [c_ctx, c_int, c_buf, c_vp, c_int, c_bits, c_buf, c_bits]),
    "mnr_ew_fma": (c_int, [c_ctx, c_buf, c_buf, c_buf, c_bits, PP, PP]),
    "mnr_ew_fma_into": (c_int, [c_ctx, c_buf, c_buf, c_buf, c_bits, c_buf, c_bits]),
    "mnr_ew_binary_promote": (c_int, [c_ctx, c_int, c_buf, c_buf, c_bits, c_bits, c_int, PP, PP]),
    "mnr_ew_binary_batch": (c_int, [c_ctx, c_int, c_sz, PP, PP, PP, PP, c_int, PP, PP]),
    "mnr_ew_binary_batch_into": (c_int, [c_ctx, c_int, c_sz, PP, PP, PP, PP, c_int, PP, PP]),
    "mnr_ew_scalar_batch_into": (c_int, [c_ctx, c_int, c_sz, PP, PP, c_int, PP, PP, PP]),
    "mnr_bits_binop": (c_int, [c_ctx, c_int, c_bits, c_sz, c_bits, c_sz, c_sz, PP]),
    "mnr_bits_binop_into": (c_int, [c_ctx, c_int, c_bits, c_sz, c_bits, c_sz, c_sz, c_bits]),
    "mnr_bits_not": (c_int, [c_ctx, c_bits, c_sz, c_sz, PP]),
    "mnr_bits_not_into": (c_int, [c_ctx, c_bits, c_sz, c_sz, c_bits]),
    "mnr_bits_popcount": (c_int, [c_ctx, c_bits, c_sz, c_sz, C.POINTER(C.c_uint64)]),
    "mnr_bits_popcount_async": (c_int, [c_ctx, c_bits, c_sz, c_sz, c_vp]),
    "mnr_bits_all_true": (c_int, [c_ctx, c_bits, C.POINTER(c_int)]),
    "mnr_bits_all_false": (c_int, [c_ctx, c_bits, C.POINTER(c_int)]),
    "mnr_bits_merge": (c_int, [c_ctx, c_bits, c_bits, c_sz, c_int, PP]),
    "mnr_bits_eq": (c_int, [c_ctx, c_bits, c_sz, c_bits, c_sz, c_sz, c_int, PP]),
    "mnr_bits_all_eq": (c_int, [c_ctx, c_bits, c_sz, c_bits, c_sz, c_sz, C.POINTER(c_int)]),
    "mnr_bits_in": (c_int, [c_ctx, c_bits, c_sz, c_bits, c_sz, c_sz, c_int, PP]),
    "mnr_bits_slice": (c_int, [c_ctx, c_bits, c_sz, c_sz, PP]),
    "mnr_concat": (c_int, [c_ctx, c_sz, PP, PP, PP, PP]),
    "mnr_eq_mask": (c_int, [c_ctx, c_buf, c_vp, c_vp, PP]),
    "mnr_reduce_stats": (c_int, [c_ctx, c_buf, c_bits, C.POINTER(Agg)]),
    "mnr_reduce_sum": (c_int, [c_ctx, c_buf, c_bits, C.POINTER(Scalar64), C.POINTER(C.c_uint64)]),
    "mnr_reduce_stats_async": (c_int, [c_ctx, c_buf, c_bits, c_int, c_vp]),
    "mnr_reduce_stats_batch": (c_int, [c_ctx, c_sz, PP, PP, c_int, C.POINTER(Agg)]),
    "mnr_reduce_stats_batch_async": (c_int, [c_ctx, c_sz, PP, PP, c_int, c_vp]),
    "mnr_xchg_create": (c_int, [c_ctx, c_int, c_int, PP]),
    "mnr_xchg_local_handle": (c_int, [c_vp, c_vp]),
    "mnr_xchg_connect": (c_int, [c_vp, c_vp]),
    "mnr_xchg_connect_local": (c_int, [c_vp, PP]),
    "mnr_xchg_status": (c_int, [c_vp, c_int, C.POINTER(c_int)]),
    "mnr_xchg_destroy": (None, [c_vp]),
    "mnr_reduce_stats_batch_exchange": (c_int, [c_ctx, c_vp, c_sz, PP, PP, c_int, c_sz, C.POINTER(C.c_uint32),
                                                C.POINTER(c_int), c_vp]),
    "mnr_reduce_stats_batch_exchange_sync": (c_int, [c_ctx, c_vp, c_sz, PP, PP, c_int, c_sz, C.POINTER(C.c_uint32),
                                                     C.POINTER(c_int), C.POINTER(Agg)]),
    "mnr_shard_owner": (c_int, [c_sz, c_sz, c_int]),
    "mnr_shard_chunk_range": (c_int, [c_sz, c_int, c_int, C.POINTER(c_sz), C.POINTER(c_sz)]),
    "mnr_shard_row_range": (c_int, [c_sz, c_int, c_int, c_sz, C.POINTER(c_sz), C.POINTER(c_sz)]),
    "mnr_group_create": (c_int, [c_int, C.POINTER(c_int), PP]),
    "mnr_group_destroy": (None, [c_vp]),
    "mnr_group_world": (c_int, [c_vp]),
    "mnr_group_ctx": (c_vp, [c_vp, c_int]),
    "mnr_group_xchg": (c_vp, [c_vp, c_int]),
    "mnr_group_synchronize": (c_int, [c_vp]),
    "mnr_group_upload": (c_int, [c_vp, c_int, c_sz, PP, C.POINTER(c_sz), PP, PP, PP]),
    "mnr_group_ew_binary": (c_int, [c_vp, c_int, c_sz, PP, PP, PP, PP, c_int, PP, PP]),
    "mnr_group_ew_scalar": (c_int, [c_vp, c_int, c_sz, PP, PP, c_int, PP, PP, PP]),
    "mnr_group_reduce_stats": (c_int, [c_vp, c_sz, PP, PP, c_int, c_sz, C.POINTER(C.c_uint32), C.POINTER(c_int),
                                       C.POINTER(Agg)]),
    "mnr_reduce_stats_exchange": (c_int, [c_ctx, c_vp, c_buf, c_bits, c_int, c_vp]),
    "mnr_reduce_stats_exchange_sync": (c_int, [c_ctx, c_vp, c_buf, c_bits, c_int, C.POINTER(Agg)]),
    "mnr_agg_mean": (C.c_double, [c_int, C.POINTER(Agg)]),
    "mnr_agg_combine": (c_int, [c_int, C.POINTER(Agg), c_sz, C.POINTER(Agg)]),
    "mnr_apply_host": (c_int, [c_ctx, c_int, c_int, c_vp, c_sz, c_vp, c_sz, c_vp, c_vp, c_vp]),
    "mnr_apply_int_i32": (c_int, [c_ctx, c_vp, c_sz, c_vp, c_sz, c_int, c_vp, c_vp, c_vp]),
    "mnr_apply_int_u32": (c_int, [c_ctx, c_vp, c_sz, c_vp, c_sz, c_int, c_vp, c_vp, c_vp]),
    "mnr_apply_int_i64": (c_int, [c_ctx, c_vp, c_sz, c_vp, c_sz, c_int, c_vp, c_vp, c_vp]),
    "mnr_apply_int_u64": (c_int, [c_ctx, c_vp, c_sz, c_vp, c_sz, c_int, c_vp, c_vp, c_vp]),
    "mnr_apply_float_f32": (c_int, [c_ctx, c_vp, c_sz, c_vp, c_sz, c_int, c_vp, c_vp, c_vp]),
    "mnr_apply_float_f64": (c_int, [c_ctx, c_vp, c_sz, c_vp, c_sz, c_int, c_vp, c_vp, c_vp]),
    "mnr_apply_fma_host": (c_int, [c_ctx, c_int, c_vp, c_sz, c_vp, c_sz, c_vp, c_sz, c_vp, c_vp, c_vp]),
    "mnr_stats_host": (c_int, [c_ctx, c_int, c_vp, c_sz, c_vp, c_int, C.POINTER(Agg)]),
    "mnr_bitmask_binop_host": (c_int, [c_ctx, c_int, c_vp, c_sz, c_vp, c_sz, c_sz, c_vp]),
    "mnr_arrow_import": (c_int, [c_ctx, c_vp, c_vp, PP, PP, PP]),
    "mnr_arrow_export": (c_int, [c_ctx, c_buf, c_bits, c_vp, c_vp]),
    "mnr_arrow_export_bool": (c_int, [c_ctx, c_bits, c_bits, c_vp, c_vp]),
    "mnr_arrow_stream_import": (c_int, [c_ctx, c_vp, c_sz, c_sz, c_sz, c_vp, c_vp, C.POINTER(c_sz), C.POINTER(c_sz)]),
    "mnr_host_register": (c_int, [c_vp, c_sz]),
    "mnr_host_unregister": (c_int, [c_vp]),
    "mnr_host_alloc": (c_int, [c_sz, PP]),
    "mnr_host_free": (None, [c_vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library.  Raises — never falls back — when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the sm_100a kernels first (python -c 'import __graft_entry__ as g; "
            "g.build()' or make -C minarrow_b200/csrc).  minarrow_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here = header/library drift
        fn.restype = res
        fn.argtypes = args
    if lib.mnr_abi_version() != 1:
        raise ImportError("libminarrow_b200.so ABI version mismatch")
    _lib = lib
    return lib
