"""Device-resident operations over `DeviceBuffer` / `DeviceBitmask` (thin, typed wrappers of the C ABI).

These are the calls a GPU-resident pipeline chains without leaving HBM; the host-slice drop-ins in
`minarrow_b200.kernels.*` are built from the same kernels plus upload/download.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib
from .core import (ArithmeticOperator, Context, DeviceBitmask, DeviceBuffer, LogicalOperator, MaskMode, check,
                   dtype_code)


def _h(x):
    return None if x is None else x.h


def _outs(ctx: Context, ob, om) -> Tuple[DeviceBuffer, Optional[DeviceBitmask]]:
    return DeviceBuffer(ctx, ob), (DeviceBitmask(ctx, om) if om.value else None)


def ew_binary(ctx: Context, op: int, lhs: DeviceBuffer, rhs: DeviceBuffer, lhs_mask: Optional[DeviceBitmask] = None,
              rhs_mask: Optional[DeviceBitmask] = None, mode: int = MaskMode.And):
    """Fused null-aware `lhs op rhs` (apply_int_* / apply_float_*, dispatch.rs:65-206 + mask merge)."""
    ob, om = C.c_void_p(), C.c_void_p()
    check(ctx.lib.mnr_ew_binary(ctx.h, int(op), lhs.h, rhs.h, _h(lhs_mask), _h(rhs_mask), int(mode), C.byref(ob),
                                C.byref(om)))
    return _outs(ctx, ob, om)


def ew_binary_into(ctx: Context, op: int, lhs, rhs, lhs_mask, rhs_mask, mode, out: DeviceBuffer,
                   out_mask: Optional[DeviceBitmask]) -> None:
    check(ctx.lib.mnr_ew_binary_into(ctx.h, int(op), lhs.h, rhs.h, _h(lhs_mask), _h(rhs_mask), int(mode), out.h,
                                     _h(out_mask)))


def ew_binary_promote(ctx: Context, op: int, lhs, rhs, lhs_mask=None, rhs_mask=None, mode: int = MaskMode.And):
    """Mixed (i32, f64)/(i32, f32) operands, cast on load (routing/arithmetic.rs:244-269,342-373)."""
    ob, om = C.c_void_p(), C.c_void_p()
    check(ctx.lib.mnr_ew_binary_promote(ctx.h, int(op), lhs.h, rhs.h, _h(lhs_mask), _h(rhs_mask), int(mode),
                                        C.byref(ob), C.byref(om)))
    return _outs(ctx, ob, om)


def _scalar_arg(arr: DeviceBuffer, scalar):
    s = np.array([scalar], dtype=arr.dtype)
    return s, s.ctypes.data_as(C.c_void_p)


def ew_scalar(ctx: Context, op: int, arr: DeviceBuffer, scalar, scalar_is_lhs: bool,
              mask: Optional[DeviceBitmask] = None):
    """`arr op scalar` / `scalar op arr` with the scalar in a register (no broadcast_length_1_array)."""
    keep, sp = _scalar_arg(arr, scalar)
    ob, om = C.c_void_p(), C.c_void_p()
    check(ctx.lib.mnr_ew_scalar(ctx.h, int(op), arr.h, sp, int(scalar_is_lhs), _h(mask), C.byref(ob), C.byref(om)))
    return _outs(ctx, ob, om)


def ew_scalar_into(ctx: Context, op: int, arr, scalar, scalar_is_lhs: bool, mask, out, out_mask) -> None:
    keep, sp = _scalar_arg(arr, scalar)
    check(ctx.lib.mnr_ew_scalar_into(ctx.h, int(op), arr.h, sp, int(scalar_is_lhs), _h(mask), out.h, _h(out_mask)))


def ew_fma(ctx: Context, a, b, acc, mask: Optional[DeviceBitmask] = None):
    ob, om = C.c_void_p(), C.c_void_p()
    check(ctx.lib.mnr_ew_fma(ctx.h, a.h, b.h, acc.h, _h(mask), C.byref(ob), C.byref(om)))
    return _outs(ctx, ob, om)


def ew_fma_into(ctx: Context, a, b, acc, mask, out, out_mask) -> None:
    check(ctx.lib.mnr_ew_fma_into(ctx.h, a.h, b.h, acc.h, _h(mask), out.h, _h(out_mask)))


def bits_binop(ctx: Context, op: int, lhs: DeviceBitmask, lhs_off: int, rhs: DeviceBitmask, rhs_off: int,
               length: int) -> DeviceBitmask:
    o = C.c_void_p()
    check(ctx.lib.mnr_bits_binop(ctx.h, int(op), lhs.h, lhs_off, rhs.h, rhs_off, length, C.byref(o)))
    return DeviceBitmask(ctx, o)


def bits_binop_into(ctx: Context, op: int, lhs, lhs_off, rhs, rhs_off, length, out: DeviceBitmask) -> None:
    check(ctx.lib.mnr_bits_binop_into(ctx.h, int(op), lhs.h, lhs_off, rhs.h, rhs_off, length, out.h))


def bits_not(ctx: Context, src: DeviceBitmask, off: int, length: int) -> DeviceBitmask:
    o = C.c_void_p()
    check(ctx.lib.mnr_bits_not(ctx.h, src.h, off, length, C.byref(o)))
    return DeviceBitmask(ctx, o)


def bits_not_into(ctx: Context, src, off, length, out: DeviceBitmask) -> None:
    check(ctx.lib.mnr_bits_not_into(ctx.h, src.h, off, length, out.h))


def bits_popcount(ctx: Context, m: DeviceBitmask, off: int, length: int) -> int:
    ones = C.c_uint64()
    check(ctx.lib.mnr_bits_popcount(ctx.h, m.h, off, length, C.byref(ones)))
    return int(ones.value)


def bits_popcount_async(ctx: Context, m: DeviceBitmask, off: int, length: int, out_device_ptr: int) -> None:
    """Set-bit count of the window -> one uint64 in device memory on the context stream (no sync)."""
    check(ctx.lib.mnr_bits_popcount_async(ctx.h, m.h, off, length, C.c_void_p(out_device_ptr)))


def bits_all_true(ctx: Context, m: DeviceBitmask) -> bool:
    r = C.c_int()
    check(ctx.lib.mnr_bits_all_true(ctx.h, m.h, C.byref(r)))
    return bool(r.value)


def bits_all_false(ctx: Context, m: DeviceBitmask) -> bool:
    r = C.c_int()
    check(ctx.lib.mnr_bits_all_false(ctx.h, m.h, C.byref(r)))
    return bool(r.value)


def bits_merge(ctx: Context, lhs: Optional[DeviceBitmask], rhs: Optional[DeviceBitmask], length: int,
               mode: int = MaskMode.And) -> Optional[DeviceBitmask]:
    o = C.c_void_p()
    check(ctx.lib.mnr_bits_merge(ctx.h, _h(lhs), _h(rhs), length, int(mode), C.byref(o)))
    return DeviceBitmask(ctx, o) if o.value else None


def bits_eq(ctx: Context, a, a_off, b, b_off, length, negate: bool = False) -> DeviceBitmask:
    o = C.c_void_p()
    check(ctx.lib.mnr_bits_eq(ctx.h, a.h, a_off, b.h, b_off, length, int(negate), C.byref(o)))
    return DeviceBitmask(ctx, o)


def bits_all_eq(ctx: Context, a, a_off, b, b_off, length) -> bool:
    r = C.c_int()
    check(ctx.lib.mnr_bits_all_eq(ctx.h, a.h, a_off, b.h, b_off, length, C.byref(r)))
    return bool(r.value)


def bits_in(ctx: Context, lhs, lhs_off, rhs, rhs_off, length, negate: bool = False) -> DeviceBitmask:
    o = C.c_void_p()
    check(ctx.lib.mnr_bits_in(ctx.h, lhs.h, lhs_off, rhs.h, rhs_off, length, int(negate), C.byref(o)))
    return DeviceBitmask(ctx, o)


def _agg_dict(dtype, agg: _lib.Agg, lib) -> dict:
    kind = np.dtype(dtype).kind
    f = {"i": "i64", "u": "u64", "f": "f64"}[kind]
    return {"sum": getattr(agg.sum, f), "min": getattr(agg.min, f), "max": getattr(agg.max, f),
            "count": int(agg.count), "mean": float(lib.mnr_agg_mean(dtype_code(dtype), C.byref(agg)))}


def reduce_stats(ctx: Context, buf: DeviceBuffer, validity: Optional[DeviceBitmask] = None) -> dict:
    """{sum, min, max, count, mean} of a device column in one pass (synchronises)."""
    agg = _lib.Agg()
    check(ctx.lib.mnr_reduce_stats(ctx.h, buf.h, _h(validity), C.byref(agg)))
    return _agg_dict(buf.dtype, agg, ctx.lib)


def reduce_sum(ctx: Context, buf: DeviceBuffer, validity: Optional[DeviceBitmask] = None):
    """(sum, valid_count): the sum the reference's benches time (benchmark_parallel_simd.rs:81-97), null-aware."""
    s, c = _lib.Scalar64(), C.c_uint64()
    check(ctx.lib.mnr_reduce_sum(ctx.h, buf.h, _h(validity), C.byref(s), C.byref(c)))
    f = {"i": "i64", "u": "u64", "f": "f64"}[buf.dtype.kind]
    return getattr(s, f), int(c.value)


def reduce_stats_async(ctx: Context, buf: DeviceBuffer, validity: Optional[DeviceBitmask], with_minmax: bool,
                       out_device_ptr: int) -> None:
    """Writes the 32-byte `mnr_agg` partial to device memory on the context stream (no sync)."""
    check(ctx.lib.mnr_reduce_stats_async(ctx.h, buf.h, _h(validity), int(with_minmax), C.c_void_p(out_device_ptr)))


def _handle_array(items):
    arr = (C.c_void_p * len(items))()
    for i, x in enumerate(items):
        arr[i] = None if x is None else x.h
    return arr


def reduce_stats_batch(ctx: Context, bufs, validities=None, with_minmax: bool = True) -> list:
    """{sum, min, max, count, mean} of many device columns/chunks with one launch per (dtype, alignment, masked)
    class (the per-chunk / per-column loops of broadcast/super_array.rs:180-249 and table.rs:31-62 in one call)."""
    n = len(bufs)
    if n == 0:
        return []
    aggs = (_lib.Agg * n)()
    vals = None if validities is None else _handle_array(list(validities))
    check(ctx.lib.mnr_reduce_stats_batch(ctx.h, n, _handle_array(list(bufs)), vals, int(with_minmax), aggs))
    return [_agg_dict(b.dtype, a, ctx.lib) for b, a in zip(bufs, aggs)]


def reduce_stats_batch_async(ctx: Context, bufs, validities, with_minmax: bool, out_device_ptr: int) -> None:
    """Writes len(bufs) x 32-byte `mnr_agg` to device memory on the context stream (no sync)."""
    n = len(bufs)
    vals = None if validities is None else _handle_array(list(validities))
    check(ctx.lib.mnr_reduce_stats_batch_async(ctx.h, n, _handle_array(list(bufs)), vals, int(with_minmax),
                                               C.c_void_p(out_device_ptr)))


def ew_binary_batch(ctx: Context, op: int, lhs, rhs, lhs_masks=None, rhs_masks=None, mode: int = MaskMode.And):
    """Chunk-wise `lhs[i] op rhs[i]` for a whole SuperArray / SuperTable in one call: fresh outputs per chunk, one launch
    per (dtype, alignment, masked) class.  Returns ([DeviceBuffer], [DeviceBitmask | None])."""
    n = len(lhs)
    ob, om = (C.c_void_p * n)(), (C.c_void_p * n)()
    lm = None if lhs_masks is None else _handle_array(list(lhs_masks))
    rm = None if rhs_masks is None else _handle_array(list(rhs_masks))
    check(ctx.lib.mnr_ew_binary_batch(ctx.h, int(op), n, _handle_array(list(lhs)), _handle_array(list(rhs)), lm, rm, int(mode),
                                      ob, om))
    return ([DeviceBuffer(ctx, C.c_void_p(ob[i])) for i in range(n)],
            [DeviceBitmask(ctx, C.c_void_p(om[i])) if om[i] else None for i in range(n)])


class EwBatchPlan:
    """Pre-marshalled argument arrays for a repeated batched call (the ctypes arrays are built once)."""

    def __init__(self, *lists):
        self.arrays = [None if x is None else _handle_array(list(x)) for x in lists]
        self.keep = lists


def ew_binary_batch_into(ctx: Context, op: int, lhs, rhs, lhs_masks, rhs_masks, mode, out, out_masks, plan: EwBatchPlan = None):
    p = plan or EwBatchPlan(lhs, rhs, lhs_masks, rhs_masks, out, out_masks)
    a = p.arrays
    check(ctx.lib.mnr_ew_binary_batch_into(ctx.h, int(op), len(p.keep[0]), a[0], a[1], a[2], a[3], int(mode), a[4], a[5]))
    return p


def ew_scalar_batch_into(ctx: Context, op: int, arrs, scalars, scalar_is_lhs: bool, masks, out, out_masks):
    """`arrs[i] op scalars[i]` (or the reverse) for every chunk in one call; `scalars[i]` is cast to arrs[i]'s dtype."""
    n = len(arrs)
    keep = [np.array([s], dtype=a.dtype) for a, s in zip(arrs, scalars)]
    sp = (C.c_void_p * n)(*[k.ctypes.data for k in keep])
    check(ctx.lib.mnr_ew_scalar_batch_into(ctx.h, int(op), n, _handle_array(list(arrs)), sp, int(scalar_is_lhs),
                                           None if masks is None else _handle_array(list(masks)),
                                           _handle_array(list(out)), None if out_masks is None else _handle_array(list(out_masks))))


def eq_mask(ctx: Context, data: DeviceBuffer, field_mask, target) -> DeviceBitmask:
    """simd_eq_mask_u{8,16,32,64} (bitmask/simd.rs:741-788): bit i = ((data[i] & field_mask) == target)."""
    fm, tg = np.array([field_mask], dtype=data.dtype), np.array([target], dtype=data.dtype)
    o = C.c_void_p()
    check(ctx.lib.mnr_eq_mask(ctx.h, data.h, fm.ctypes.data_as(C.c_void_p), tg.ctypes.data_as(C.c_void_p), C.byref(o)))
    return DeviceBitmask(ctx, o)


def bits_slice(ctx: Context, src: DeviceBitmask, offset: int, length: int) -> DeviceBitmask:
    """Bitmask::slice_clone (bitmask.rs:604-626) on the device: exact bit offset, result starts at bit 0."""
    o = C.c_void_p()
    check(ctx.lib.mnr_bits_slice(ctx.h, src.h, offset, length, C.byref(o)))
    return DeviceBitmask(ctx, o)


def concat(ctx: Context, bufs, validities=None):
    """consolidate(): all chunks -> one contiguous device column (+ validity iff any chunk has one)."""
    n = len(bufs)
    ob, om = C.c_void_p(), C.c_void_p()
    vals = None if validities is None else _handle_array(list(validities))
    check(ctx.lib.mnr_concat(ctx.h, n, _handle_array(list(bufs)), vals, C.byref(ob), C.byref(om)))
    return DeviceBuffer(ctx, ob), (DeviceBitmask(ctx, om) if om.value else None)


def rechunk(ctx: Context, bufs, validities, chunk_rows: int):
    """SuperArray::rechunk(Count(chunk_rows)) on the device (super_array.rs:674-787): chunks of exactly `chunk_rows` rows
    plus one remainder chunk.  One consolidate (two launches), then value chunks are zero-copy windows of the
    consolidated buffer and validity chunks are bit-offset slices."""
    if chunk_rows <= 0:
        from .core import KernelError
        raise KernelError("OutOfBounds", "Count chunk size must be greater than 0")
    total = sum(len(b) for b in bufs)
    if not bufs or total == 0:
        return list(bufs), list(validities) if validities is not None else [None] * len(bufs)
    whole, wmask = concat(ctx, bufs, validities)
    out_b, out_v = [], []
    for r0 in range(0, total, chunk_rows):
        ln = min(chunk_rows, total - r0)
        out_b.append(whole.slice(r0, ln))
        out_v.append(None if wmask is None else bits_slice(ctx, wmask, r0, ln))
    return out_b, out_v
