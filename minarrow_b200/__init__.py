"""minarrow_b200 — B200-native drop-in for Minarrow's columnar compute hot path.

Hand-written sm_100a CUDA kernels behind a C ABI (include/minarrow_b200.h); this package is the host-side
mirror of the reference's interface for that path.  Importing it loads libminarrow_b200.so and raises if
the library has not been built — there is no CPU fallback.
"""
from . import _lib

_lib.load()

from . import arrow, device_ops, kernels, sharded  # noqa: E402
from .core import (ArithmeticOperator, Bitmask, BooleanArray, Context, DeviceBitmask, DeviceBuffer, FloatArray,  # noqa: E402
                   DatetimeArray, IntegerArray, KernelError, LogicalOperator, MaskMode, ShapeError, default_context)
from .kernels.arithmetic import (apply_float_f32, apply_float_f64, apply_fma_f32, apply_fma_f64, apply_int_i32,  # noqa: E402
                                 apply_int_i64, apply_int_u32, apply_int_u64)
from .kernels.broadcast import SuperArray, route_super_array_broadcast  # noqa: E402
from .kernels.routing import resolve_binary_arithmetic  # noqa: E402

__all__ = ["ArithmeticOperator", "LogicalOperator", "MaskMode", "Bitmask", "IntegerArray", "FloatArray", "BooleanArray", "DatetimeArray",
           "Context", "DeviceBuffer", "DeviceBitmask", "KernelError", "ShapeError", "default_context", "device_ops",
           "kernels", "apply_int_i32", "apply_int_u32", "apply_int_i64", "apply_int_u64", "apply_float_f32",
           "apply_float_f64", "apply_fma_f32", "apply_fma_f64", "resolve_binary_arithmetic", "SuperArray",
           "route_super_array_broadcast"]
