"""`resolve_binary_arithmetic` (src/kernels/routing/arithmetic.rs:214-406): length-1 broadcast, dtype match,
int->float promotion, then the leaf kernel.  Differences from the reference are performance-only:
  * a length-1 operand is passed to the kernel by value instead of being materialised `len` times
    (routing/broadcast.rs:25-47) — same results, 2/3 of the HBM traffic;
  * (i32,f64)/(i32,f32) pairs are cast on load inside the kernel instead of through two copied Vec64s
    (routing/arithmetic.rs:244-269).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .. import device_ops as dev
from ..core import (ArithmeticOperator, Bitmask, Context, DeviceBitmask, DeviceBuffer, KernelError, default_context,
                    make_array)

_SAME = {np.dtype(t) for t in (np.int32, np.int64, np.uint32, np.uint64, np.float32, np.float64)}


def _values(x):
    return np.ascontiguousarray(getattr(x, "data", x))


def resolve_binary_arithmetic(op: ArithmeticOperator, lhs, rhs, null_mask: Optional[Bitmask] = None,
                              ctx: Optional[Context] = None):
    """Arrays (numpy or IntegerArray/FloatArray) in, a fresh typed array out.  `null_mask` is the single
    pre-merged mask of the leaf API, indexed from bit 0 (routing/arithmetic.rs:284-287); operands' own
    masks are NOT consulted here, exactly like the reference (that is the caller's job)."""
    ctx = ctx or default_context()
    l, r = _values(lhs), _values(rhs)
    ln, rn = l.size, r.size
    if ln != rn and ln != 1 and rn != 1:
        raise KernelError("LengthMismatch", f"cannot broadcast arrays of length {ln} and {rn}")
    lt, rt = l.dtype, r.dtype
    promote = None
    if lt != rt:
        pair = {lt, rt}
        if pair == {np.dtype(np.int32), np.dtype(np.float64)}:
            promote = np.dtype(np.float64)
        elif pair == {np.dtype(np.int32), np.dtype(np.float32)}:
            promote = np.dtype(np.float32)
        else:
            raise KernelError("UnsupportedType", "Unsupported array type combination for arithmetic operations")
    elif lt not in _SAME:
        raise KernelError("UnsupportedType", "Unsupported array type combination for arithmetic operations")
    n = max(ln, rn) if ln != rn else ln
    dmask = None
    if null_mask is not None:
        if null_mask.len < n:
            raise KernelError("InvalidArguments", f"mask has {null_mask.len} bits, need {n}")
        dmask = DeviceBitmask.upload(ctx, null_mask)
    if ln != rn:   # maybe_broadcast_scalar_array (routing/broadcast.rs:87-112)
        scalar_is_lhs = ln == 1
        arr, sc = (r, l) if scalar_is_lhs else (l, r)
        out_dt = promote or arr.dtype
        # the length-1 side's first element, cast like `x as f64` when promoting; a.data[0] (broadcast.rs:29-46)
        arr_c = arr.astype(out_dt) if promote is not None and arr.dtype != out_dt else arr
        scalar = sc.reshape(-1)[0].astype(out_dt)
        ob, om = dev.ew_scalar(ctx, op, DeviceBuffer.upload(ctx, arr_c), scalar, scalar_is_lhs, dmask)
    elif promote is not None:
        ob, om = dev.ew_binary_promote(ctx, op, DeviceBuffer.upload(ctx, l), DeviceBuffer.upload(ctx, r), dmask)
    else:
        ob, om = dev.ew_binary(ctx, op, DeviceBuffer.upload(ctx, l), DeviceBuffer.upload(ctx, r), dmask)
    return make_array(ob.download(), None if om is None else om.download())
