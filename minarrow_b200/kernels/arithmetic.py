"""Leaf arithmetic kernels with the reference's names and signatures
(src/kernels/arithmetic/dispatch.rs:74-79,147-152,221-226):

    apply_int_i64(lhs: &[i64], rhs: &[i64], op, mask: Option<&Bitmask>) -> Result<IntegerArray<i64>, KernelError>

Host slices in, a fresh `IntegerArray`/`FloatArray` out; the work happens on the GPU through
`mnr_apply_host` (chunked upload -> fused kernel -> download).  Errors follow the reference:
LengthMismatch, and DivideByZero where the reference's dense integer kernels panic.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from ..core import (ArithmeticOperator, Bitmask, Context, FloatArray, IntegerArray, KernelError, _vp, check,
                    default_context, dtype_code)


def _mask_arg(mask: Optional[Bitmask], n: int):
    if mask is None:
        return None
    need = (n + 7) // 8
    bits = np.ascontiguousarray(mask.bits, dtype=np.uint8)
    if bits.size < need:
        raise KernelError("InvalidArguments", f"mask has {bits.size} bytes, need {need}")
    return bits


def _apply(dtype, lhs, rhs, op, mask: Optional[Bitmask], ctx: Optional[Context]):
    ctx = ctx or default_context()
    lhs = np.ascontiguousarray(lhs, dtype=dtype)
    rhs = np.ascontiguousarray(rhs, dtype=dtype)
    n = lhs.size
    m = _mask_arg(mask, n)
    out = np.empty(n, dtype=dtype)
    om = np.zeros((n + 7) // 8, dtype=np.uint8) if m is not None else None
    check(ctx.lib.mnr_apply_host(ctx.h, dtype_code(dtype), int(op), _vp(lhs), lhs.size, _vp(rhs), rhs.size, _vp(m),
                                 _vp(out), _vp(om)))
    null_mask = Bitmask(om, n) if om is not None else None   # Some(mask) iff a mask was passed (dispatch.rs:90-104)
    return (FloatArray if np.dtype(dtype).kind == "f" else IntegerArray)(out, null_mask)


def apply_int_i32(lhs, rhs, op: ArithmeticOperator, mask: Optional[Bitmask] = None, ctx=None) -> IntegerArray:
    return _apply(np.int32, lhs, rhs, op, mask, ctx)


def apply_int_u32(lhs, rhs, op: ArithmeticOperator, mask: Optional[Bitmask] = None, ctx=None) -> IntegerArray:
    return _apply(np.uint32, lhs, rhs, op, mask, ctx)


def apply_int_i64(lhs, rhs, op: ArithmeticOperator, mask: Optional[Bitmask] = None, ctx=None) -> IntegerArray:
    return _apply(np.int64, lhs, rhs, op, mask, ctx)


def apply_int_u64(lhs, rhs, op: ArithmeticOperator, mask: Optional[Bitmask] = None, ctx=None) -> IntegerArray:
    return _apply(np.uint64, lhs, rhs, op, mask, ctx)


# `extended_numeric_types` (dispatch.rs:380-387)
def apply_int_i16(lhs, rhs, op, mask=None, ctx=None) -> IntegerArray:
    return _apply(np.int16, lhs, rhs, op, mask, ctx)


def apply_int_u16(lhs, rhs, op, mask=None, ctx=None) -> IntegerArray:
    return _apply(np.uint16, lhs, rhs, op, mask, ctx)


def apply_int_i8(lhs, rhs, op, mask=None, ctx=None) -> IntegerArray:
    return _apply(np.int8, lhs, rhs, op, mask, ctx)


def apply_int_u8(lhs, rhs, op, mask=None, ctx=None) -> IntegerArray:
    return _apply(np.uint8, lhs, rhs, op, mask, ctx)


def apply_float_f32(lhs, rhs, op: ArithmeticOperator, mask: Optional[Bitmask] = None, ctx=None) -> FloatArray:
    return _apply(np.float32, lhs, rhs, op, mask, ctx)


def apply_float_f64(lhs, rhs, op: ArithmeticOperator, mask: Optional[Bitmask] = None, ctx=None) -> FloatArray:
    return _apply(np.float64, lhs, rhs, op, mask, ctx)


APPLY = {np.dtype(np.int32): apply_int_i32, np.dtype(np.uint32): apply_int_u32, np.dtype(np.int64): apply_int_i64,
         np.dtype(np.uint64): apply_int_u64, np.dtype(np.int16): apply_int_i16, np.dtype(np.uint16): apply_int_u16,
         np.dtype(np.int8): apply_int_i8, np.dtype(np.uint8): apply_int_u8, np.dtype(np.float32): apply_float_f32,
         np.dtype(np.float64): apply_float_f64}


def _fma(dtype, lhs, rhs, acc, mask, ctx) -> FloatArray:
    ctx = ctx or default_context()
    lhs, rhs, acc = (np.ascontiguousarray(x, dtype=dtype) for x in (lhs, rhs, acc))
    n = lhs.size
    m = _mask_arg(mask, n)
    out = np.empty(n, dtype=dtype)
    om = np.zeros((n + 7) // 8, dtype=np.uint8) if m is not None else None
    check(ctx.lib.mnr_apply_fma_host(ctx.h, dtype_code(dtype), _vp(lhs), lhs.size, _vp(rhs), rhs.size, _vp(acc),
                                     acc.size, _vp(m), _vp(out), _vp(om)))
    return FloatArray(out, Bitmask(om, n) if om is not None else None)


def apply_fma_f32(lhs, rhs, acc, mask: Optional[Bitmask] = None, ctx=None) -> FloatArray:
    """apply_fma_f32 (dispatch.rs:404-410)."""
    return _fma(np.float32, lhs, rhs, acc, mask, ctx)


def apply_fma_f64(lhs, rhs, acc, mask: Optional[Bitmask] = None, ctx=None) -> FloatArray:
    """apply_fma_f64 (dispatch.rs:412-418)."""
    return _fma(np.float64, lhs, rhs, acc, mask, ctx)


# ---- datetime delegation (dispatch.rs:300-372, 420-427) ----------------------------------------------------------------
def _apply_datetime(dtype, lhs, rhs, op, ctx):
    """`apply_datetime_*(lhs: DatetimeAVT<T>, rhs: DatetimeAVT<T>, op)`: the integer kernels over the two data windows.
    Output validity = merge_bitmasks_to_new(lhs.null_mask, rhs.null_mask, llen) — bits [0, llen) of each array's mask, which
    the reference does not offset by the window start (dispatch.rs:321-322) — fused into the launch as two mask operands
    (per-row AND); `Some(mask)` iff either side has one.  LengthMismatch like the reference's `confirm_equal_len`; dense
    integer division by zero is DivideByZero where the reference panics."""
    from .. import device_ops as dev
    from ..core import DatetimeArray, DeviceBitmask, DeviceBuffer, MaskMode
    (larr, loff, llen), (rarr, roff, rlen) = lhs, rhs
    if llen != rlen:
        raise KernelError("LengthMismatch", f"apply_datetime: length mismatch (lhs: {llen}, rhs: {rlen})")
    ld = np.ascontiguousarray(larr.data, dtype=dtype)
    rd = np.ascontiguousarray(rarr.data, dtype=dtype)
    if loff + llen > ld.size or roff + rlen > rd.size:
        raise KernelError("OutOfBounds", f"apply_datetime: window ({loff}, {llen}) / ({roff}, {rlen}) outside the arrays ({ld.size}, {rd.size})")
    unit = larr.time_unit
    if llen == 0:
        both = larr.null_mask is not None or rarr.null_mask is not None
        return DatetimeArray(np.empty(0, dtype=dtype), Bitmask(np.zeros(0, dtype=np.uint8), 0) if both else None, unit)
    masks = []
    for arr in (larr, rarr):
        if arr.null_mask is None:
            masks.append(None)
            continue
        if arr.null_mask.len < llen:
            raise KernelError("InvalidArguments", f"Bitmask too short in merge ({arr.null_mask.len} < {llen})")
        masks.append(Bitmask(arr.null_mask.bits[:(llen + 7) // 8], llen))
    ctx = ctx or default_context()      # every argument error above is raised before a device is touched
    masks = [None if m is None else DeviceBitmask.upload(ctx, m) for m in masks]
    L = DeviceBuffer.upload(ctx, ld[loff:loff + llen])
    R = DeviceBuffer.upload(ctx, rd[roff:roff + rlen])
    ob, om = dev.ew_binary(ctx, int(op), L, R, masks[0], masks[1], MaskMode.And)
    return DatetimeArray(ob.download(), om.download() if om is not None else None, unit)


def apply_datetime_i32(lhs, rhs, op: ArithmeticOperator, ctx=None):
    """apply_datetime_i32 (dispatch.rs:420-421)."""
    return _apply_datetime(np.int32, lhs, rhs, op, ctx)


def apply_datetime_u32(lhs, rhs, op: ArithmeticOperator, ctx=None):
    """apply_datetime_u32 (dispatch.rs:422-423)."""
    return _apply_datetime(np.uint32, lhs, rhs, op, ctx)


def apply_datetime_i64(lhs, rhs, op: ArithmeticOperator, ctx=None):
    """apply_datetime_i64 (dispatch.rs:424-425)."""
    return _apply_datetime(np.int64, lhs, rhs, op, ctx)


def apply_datetime_u64(lhs, rhs, op: ArithmeticOperator, ctx=None):
    """apply_datetime_u64 (dispatch.rs:426-427)."""
    return _apply_datetime(np.uint64, lhs, rhs, op, ctx)
