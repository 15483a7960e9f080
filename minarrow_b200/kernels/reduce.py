"""Null-aware reductions over host columns: sum (the reduction the reference's benches define,
benches/benchmark_parallel_simd.rs:44-97, benches/hotloop_benchmark_simd.rs:56-174) plus count / min / max /
mean (reference-unpinned; definition in DESIGN.md "A.6")."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from .. import _lib
from ..core import Bitmask, Context, KernelError, _vp, check, default_context, dtype_code
from ..device_ops import _agg_dict


def stats(data, validity: Optional[Bitmask] = None, with_minmax: bool = True, ctx: Optional[Context] = None) -> dict:
    """{sum, min, max, count, mean} of a host column: chunked upload overlapped with the reduction kernel."""
    ctx = ctx or default_context()
    d = np.ascontiguousarray(getattr(data, "data", data))
    if validity is None:
        validity = getattr(data, "null_mask", None)
    v = None
    if validity is not None:
        if validity.len < d.size:
            raise KernelError("InvalidArguments", f"validity has {validity.len} bits, need {d.size}")
        v = np.ascontiguousarray(validity.bits, dtype=np.uint8)
    agg = _lib.Agg()
    check(ctx.lib.mnr_stats_host(ctx.h, dtype_code(d.dtype), _vp(d), d.size, _vp(v), int(with_minmax), C.byref(agg)))
    return _agg_dict(d.dtype, agg, ctx.lib)


def sum(data, validity: Optional[Bitmask] = None, ctx=None):  # noqa: A001 - reference vocabulary
    return stats(data, validity, False, ctx)["sum"]


def count(data, validity: Optional[Bitmask] = None, ctx=None) -> int:
    return stats(data, validity, False, ctx)["count"]


def mean(data, validity: Optional[Bitmask] = None, ctx=None) -> float:
    return stats(data, validity, False, ctx)["mean"]


def min(data, validity: Optional[Bitmask] = None, ctx=None):  # noqa: A001
    return stats(data, validity, True, ctx)["min"]


def max(data, validity: Optional[Bitmask] = None, ctx=None):  # noqa: A001
    return stats(data, validity, True, ctx)["max"]
