"""Arrow C Data Interface <-> device buffers (the reference's own C ABI, src/ffi/arrow_c_ffi.rs:87-98,432-470,640;
PyArrow is its canonical foreign peer, pyo3/tests/test_roundtrip.py).  Any producer that can `_export_to_c` feeds the
GPU path without a Minarrow-specific copy; `offset` (element offset for values, arbitrary bit offset for validity /
boolean data) is honoured on import, results come back as a host ArrowArray that the consumer releases."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

from .core import Context, DeviceBitmask, DeviceBuffer, check


class ArrowSchema(C.Structure):
    pass


class ArrowArray(C.Structure):
    pass


ArrowSchema._fields_ = [("format", C.c_char_p), ("name", C.c_char_p), ("metadata", C.c_char_p), ("flags", C.c_int64),
                        ("n_children", C.c_int64), ("children", C.c_void_p), ("dictionary", C.c_void_p),
                        ("release", C.CFUNCTYPE(None, C.POINTER(ArrowSchema))), ("private_data", C.c_void_p)]
ArrowArray._fields_ = [("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64), ("n_buffers", C.c_int64),
                       ("n_children", C.c_int64), ("buffers", C.c_void_p), ("children", C.c_void_p),
                       ("dictionary", C.c_void_p), ("release", C.CFUNCTYPE(None, C.POINTER(ArrowArray))),
                       ("private_data", C.c_void_p)]


def from_arrow(ctx: Context, arr) -> Tuple[Optional[DeviceBuffer], Optional[DeviceBitmask], Optional[DeviceBitmask]]:
    """Upload a pyarrow numeric / boolean Array (sliced or not).  Returns (values, data_bits, validity): `values` for
    numeric arrays, `data_bits` for boolean arrays, `validity` None when the array has no nulls."""
    a, s = ArrowArray(), ArrowSchema()
    arr._export_to_c(C.addressof(a), C.addressof(s))
    try:
        v, d, m = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(ctx.lib.mnr_arrow_import(ctx.h, C.byref(a), C.byref(s), C.byref(v), C.byref(d), C.byref(m)))
    finally:
        if a.release:
            a.release(C.byref(a))
        if s.release:
            s.release(C.byref(s))
    return (DeviceBuffer(ctx, v) if v.value else None, DeviceBitmask(ctx, d) if d.value else None,
            DeviceBitmask(ctx, m) if m.value else None)


def to_arrow(ctx: Context, values: DeviceBuffer, validity: Optional[DeviceBitmask] = None):
    """Download a device column as a pyarrow Array (ownership of the host buffers moves to pyarrow)."""
    import pyarrow as pa
    a, s = ArrowArray(), ArrowSchema()
    check(ctx.lib.mnr_arrow_export(ctx.h, values.h, None if validity is None else validity.h, C.byref(a), C.byref(s)))
    return pa.Array._import_from_c(C.addressof(a), C.addressof(s))


def to_arrow_bool(ctx: Context, data_bits: DeviceBitmask, validity: Optional[DeviceBitmask] = None):
    import pyarrow as pa
    a, s = ArrowArray(), ArrowSchema()
    check(ctx.lib.mnr_arrow_export_bool(ctx.h, data_bits.h, None if validity is None else validity.h, C.byref(a), C.byref(s)))
    return pa.Array._import_from_c(C.addressof(a), C.addressof(s))


def shard_range(n_chunks: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the chunks `chunk i -> rank floor(i*world/n_chunks)` gives to `rank` (sharded.py uses the same map)."""
    lo = -(-rank * n_chunks // world)
    hi = -(-(rank + 1) * n_chunks // world)
    return lo, hi


def from_arrow_stream(ctx: Context, chunked, rank: int = 0, world: int = 1):
    """Drain an Arrow C stream (anything with `__arrow_c_stream__`: pyarrow.ChunkedArray, a polars Series, the reference's
    own PyCapsule export, src/ffi/arrow_c_ffi.rs:153-168) and upload this rank's contiguous block of chunks.
    Returns ([DeviceBuffer], [DeviceBitmask | None], chunks_in_stream)."""
    n = chunked.num_chunks if hasattr(chunked, "num_chunks") else None
    if n is None:
        raise TypeError("from_arrow_stream needs the chunk count (num_chunks) to place the shard boundaries")
    lo, hi = shard_range(n, rank, world)
    cap = max(hi - lo, 0)
    capsule = chunked.__arrow_c_stream__()
    C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
    C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    stream = C.pythonapi.PyCapsule_GetPointer(capsule, b"arrow_array_stream")
    vals, masks = (C.c_void_p * max(cap, 1))(), (C.c_void_p * max(cap, 1))()
    got, seen = C.c_size_t(), C.c_size_t()
    check(ctx.lib.mnr_arrow_stream_import(ctx.h, C.c_void_p(stream), lo, hi, cap, vals, masks, C.byref(got), C.byref(seen)))
    del capsule   # the capsule's destructor releases the (now drained) stream
    bufs = [DeviceBuffer(ctx, C.c_void_p(vals[k])) for k in range(got.value)]
    vms = [DeviceBitmask(ctx, C.c_void_p(masks[k])) if masks[k] else None for k in range(got.value)]
    return bufs, vms, int(seen.value)
