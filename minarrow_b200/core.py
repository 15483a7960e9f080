"""Host-side mirror of the reference's data model for the hot path, over the C ABI.

Names follow the reference: `ArithmeticOperator`/`LogicalOperator` (src/enums/operators.rs:19-48,88-104),
`KernelError` (src/enums/error.rs:157-187), `Bitmask` (src/structs/bitmask.rs:66-71), `IntegerArray<T>` /
`FloatArray<T>` (src/structs/variants/{integer,float}.rs), `BooleanArray` (variants/boolean.rs:108-119).
The device-resident types (`DeviceBuffer`, `DeviceBitmask`) are the new subsystem the north-star asks for:
a buffer type alongside `Vec64` with Arrow-layout-preserving upload and download.
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib


class ArithmeticOperator(enum.IntEnum):
    Add = 0
    Subtract = 1
    Multiply = 2
    Divide = 3
    Remainder = 4
    Power = 5
    FloorDiv = 6


class LogicalOperator(enum.IntEnum):
    And = 0
    Or = 1
    Xor = 2


class MaskMode(enum.IntEnum):
    And = 0   # merge_bitmasks_to_new / Bitmask::intersect
    Or = 1    # Bitmask::union (route_super_array_broadcast)


DTYPES = {np.dtype(np.int32): 0, np.dtype(np.uint32): 1, np.dtype(np.int64): 2, np.dtype(np.uint64): 3,
          np.dtype(np.float32): 4, np.dtype(np.float64): 5, np.dtype(np.int8): 6, np.dtype(np.uint8): 7,
          np.dtype(np.int16): 8, np.dtype(np.uint16): 9}
NP_OF = {v: k for k, v in DTYPES.items()}

_KINDS = {-1: "TypeMismatch", -2: "LengthMismatch", -3: "BroadcastingError", -4: "OperatorMismatch",
          -5: "UnsupportedType", -6: "ColumnNotFound", -7: "InvalidArguments", -8: "Plan", -9: "OutOfBounds",
          -10: "DivideByZero", -100: "Cuda", -101: "NoDevice", -102: "OutOfMemory"}


class KernelError(Exception):
    """`KernelError` of the reference; `.kind` is the variant name.  The reference's dense-integer
    divide-by-zero *panic* surfaces as kind "DivideByZero"."""

    def __init__(self, kind: str, msg: str = ""):
        super().__init__(f"{kind}: {msg}")
        self.kind = kind


class ShapeError(KernelError):
    """`MinarrowError::ShapeError` raised by the container routes (broadcast/super_array.rs:203-211)."""

    def __init__(self, msg: str):
        super().__init__("ShapeError", msg)


def check(rc: int) -> None:
    if rc != 0:
        msg = _lib.load().mnr_last_error()
        raise KernelError(_KINDS.get(rc, f"Error{rc}"), msg.decode() if msg else "")


def dtype_code(dt) -> int:
    try:
        return DTYPES[np.dtype(dt)]
    except KeyError:
        raise KernelError("UnsupportedType", f"dtype {dt} is not a Minarrow numeric type") from None


def _vp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One device + stream (`mnr_ctx`).  `stream` borrows a raw cudaStream_t, e.g.
    `torch.cuda.current_stream().cuda_stream`, so torch events bracket the kernels."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.lib = _lib.load()
        h = C.c_void_p()
        if stream is None:
            check(self.lib.mnr_ctx_create(device, C.byref(h)))
        else:
            check(self.lib.mnr_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(h)))
        self.h = h
        self.device = device
        self._owned = True

    @classmethod
    def borrow(cls, handle) -> "Context":
        """Non-owning view of a context that belongs to a `mnr_group` (sharded.Group)."""
        self = cls.__new__(cls)
        self.lib = _lib.load()
        self.h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        self.device = int(self.lib.mnr_ctx_device(self.h))
        self._owned = False
        return self

    @property
    def stream(self) -> int:
        """Raw cudaStream_t of the context (wrap with torch.cuda.ExternalStream to put torch work on it)."""
        return int(self.lib.mnr_ctx_stream(self.h) or 0)

    def synchronize(self) -> None:
        check(self.lib.mnr_ctx_synchronize(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.mnr_ctx_launch_count(self.h))

    def set_option(self, key: str, value: int) -> None:
        check(self.lib.mnr_ctx_set_option(self.h, key.encode(), int(value)))

    def close(self) -> None:
        if getattr(self, "h", None):
            if getattr(self, "_owned", True):
                self.lib.mnr_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


class DeviceBuffer:
    """Device-resident values buffer (`mnr_buf`): the `Vec64<T>` / `Buffer<T>` analogue in HBM."""

    def __init__(self, ctx: Context, handle, keepalive=None):
        self.ctx, self.h, self._keep = ctx, handle, keepalive

    @classmethod
    def upload(cls, ctx: Context, data: np.ndarray) -> "DeviceBuffer":
        data = np.ascontiguousarray(data)
        h = C.c_void_p()
        check(ctx.lib.mnr_buf_upload(ctx.h, dtype_code(data.dtype), _vp(data), data.size, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def alloc(cls, ctx: Context, dtype, length: int) -> "DeviceBuffer":
        h = C.c_void_p()
        check(ctx.lib.mnr_buf_alloc(ctx.h, dtype_code(dtype), length, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def wrap(cls, ctx: Context, dtype, device_ptr: int, length: int, keepalive=None) -> "DeviceBuffer":
        h = C.c_void_p()
        check(ctx.lib.mnr_buf_wrap(ctx.h, dtype_code(dtype), C.c_void_p(device_ptr), length, C.byref(h)))
        return cls(ctx, h, keepalive)

    def slice(self, offset: int, length: int) -> "DeviceBuffer":
        """ArrayV window (src/structs/views/array_view.rs:79-94)."""
        h = C.c_void_p()
        check(self.ctx.lib.mnr_buf_slice(self.h, offset, length, C.byref(h)))
        return DeviceBuffer(self.ctx, h, keepalive=self)

    def __len__(self) -> int:
        return int(self.ctx.lib.mnr_buf_len(self.h))

    @property
    def dtype(self) -> np.dtype:
        return NP_OF[int(self.ctx.lib.mnr_buf_dtype(self.h))]

    @property
    def device_ptr(self) -> int:
        return int(self.ctx.lib.mnr_buf_device_ptr(self.h) or 0)

    def download(self) -> np.ndarray:
        out = np.empty(len(self), dtype=self.dtype)
        check(self.ctx.lib.mnr_buf_download(self.ctx.h, self.h, _vp(out)))
        return out

    def free(self) -> None:
        if self.h:
            self.ctx.lib.mnr_buf_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:  # noqa: BLE001
            pass


class DeviceBitmask:
    """Device-resident validity / boolean bitmask (`mnr_bits`): Arrow layout, LSB first, slack bits zero."""

    def __init__(self, ctx: Context, handle, keepalive=None):
        self.ctx, self.h, self._keep = ctx, handle, keepalive

    @classmethod
    def upload(cls, ctx: Context, mask: "Bitmask") -> "DeviceBitmask":
        h = C.c_void_p()
        bits = np.ascontiguousarray(mask.bits, dtype=np.uint8)
        if bits.size < (mask.len + 7) // 8:   # the reference would panic on the slice; never hand the C ABI a short buffer
            raise KernelError("OutOfBounds", f"Bitmask of {mask.len} bits is backed by {bits.size} bytes, need {(mask.len + 7) // 8}")
        check(ctx.lib.mnr_bits_upload(ctx.h, _vp(bits), mask.len, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def alloc(cls, ctx: Context, len_bits: int) -> "DeviceBitmask":
        h = C.c_void_p()
        check(ctx.lib.mnr_bits_alloc(ctx.h, len_bits, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def new_set_all(cls, ctx: Context, len_bits: int, value: bool) -> "DeviceBitmask":
        h = C.c_void_p()
        check(ctx.lib.mnr_bits_new_set_all(ctx.h, len_bits, int(value), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def wrap(cls, ctx: Context, device_ptr: int, len_bits: int, keepalive=None) -> "DeviceBitmask":
        h = C.c_void_p()
        check(ctx.lib.mnr_bits_wrap(ctx.h, C.c_void_p(device_ptr), len_bits, C.byref(h)))
        return cls(ctx, h, keepalive)

    def __len__(self) -> int:
        return int(self.ctx.lib.mnr_bits_len(self.h))

    @property
    def device_ptr(self) -> int:
        return int(self.ctx.lib.mnr_bits_device_ptr(self.h) or 0)

    def download(self) -> "Bitmask":
        n = len(self)
        out = np.zeros((n + 7) // 8, dtype=np.uint8)
        check(self.ctx.lib.mnr_bits_download(self.ctx.h, self.h, _vp(out)))
        return Bitmask(out, n)

    def free(self) -> None:
        if self.h:
            self.ctx.lib.mnr_bits_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:  # noqa: BLE001
            pass


@dataclass
class Bitmask:
    """Host `Bitmask {bits, len}` (src/structs/bitmask.rs:66-71): ceil(len/8) bytes, LSB first, 1 = valid."""
    bits: np.ndarray
    len: int

    @staticmethod
    def from_bools(b) -> "Bitmask":
        """Bitmask::from_bools (bitmask.rs:349-365)."""
        b = np.asarray(b, dtype=bool)
        return Bitmask(np.packbits(b, bitorder="little"), int(b.size))

    @staticmethod
    def new_set_all(length: int, value: bool) -> "Bitmask":
        """Bitmask::new_set_all (bitmask.rs:94-105)."""
        bits = np.full((length + 7) // 8, 0xFF if value else 0, dtype=np.uint8)
        if value and length % 8:
            bits[-1] &= (1 << (length % 8)) - 1
        return Bitmask(bits, length)

    def to_bools(self) -> np.ndarray:
        return np.unpackbits(self.bits, bitorder="little")[: self.len].astype(bool)

    def get(self, i: int) -> bool:
        return bool((self.bits[i >> 3] >> (i & 7)) & 1) if i < self.len else False

    def __len__(self) -> int:
        return self.len


@dataclass
class IntegerArray:
    """`IntegerArray<T> {data, null_mask}` (src/structs/variants/integer.rs:105-111)."""
    data: np.ndarray
    null_mask: Optional[Bitmask] = None

    def __len__(self) -> int:
        return int(self.data.size)

    def is_empty(self) -> bool:
        return self.data.size == 0


@dataclass
class FloatArray:
    """`FloatArray<T> {data, null_mask}` (src/structs/variants/float.rs:109-116)."""
    data: np.ndarray
    null_mask: Optional[Bitmask] = None

    def __len__(self) -> int:
        return int(self.data.size)

    def is_empty(self) -> bool:
        return self.data.size == 0


@dataclass
class DatetimeArray:
    """`DatetimeArray<T> {data, null_mask, time_unit}` (src/structs/variants/datetime/mod.rs:90-140): integer offsets from the
    epoch in `time_unit`; arithmetic on it is the integer kernels' (dispatch.rs:300-372)."""
    data: np.ndarray
    null_mask: Optional[Bitmask] = None
    time_unit: Optional[str] = None

    @classmethod
    def from_slice(cls, data, time_unit: Optional[str] = None) -> "DatetimeArray":
        return cls(np.ascontiguousarray(data), None, time_unit)

    def __len__(self) -> int:
        return int(self.data.size)

    def is_empty(self) -> bool:
        return self.data.size == 0


@dataclass
class BooleanArray:
    """`BooleanArray {data: Bitmask, null_mask, len}` (src/structs/variants/boolean.rs:108-119)."""
    data: Bitmask
    null_mask: Optional[Bitmask] = None

    def __len__(self) -> int:
        return self.data.len

    def __invert__(self) -> "BooleanArray":
        """`impl Not for BooleanArray` (boolean.rs:853-866): the data bits are inverted on the device (`not_mask`, slack bits
        stay zero), the validity is kept as is."""
        from .kernels.bitmask import not_mask
        return BooleanArray(not_mask((self.data, 0, self.data.len)), self.null_mask)


def make_array(data: np.ndarray, null_mask: Optional[Bitmask]):
    return (FloatArray if data.dtype.kind == "f" else IntegerArray)(data, null_mask)


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    """Process-wide context on the current rank's device (LOCAL_RANK, else 0)."""
    global _default_ctx
    if _default_ctx is None or _default_ctx.h is None:
        import os
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx
