"""Bitmask kernels with the reference's names (src/kernels/bitmask/dispatch.rs:47-295, mod.rs:171-197).

A window is `BitmaskVT = (&Bitmask, offset, len)`.  Host `Bitmask`es are uploaded, combined on the GPU and
downloaded; chain `minarrow_b200.device_ops.bits_*` instead to stay resident.
"""
from __future__ import annotations

from typing import Optional, Tuple

from .. import device_ops as dev
from ..core import Bitmask, Context, DeviceBitmask, LogicalOperator, MaskMode, default_context

BitmaskVT = Tuple[Bitmask, int, int]


def _up(ctx, m: Bitmask) -> DeviceBitmask:
    return DeviceBitmask.upload(ctx, m)


def bitmask_binop(lhs: BitmaskVT, rhs: BitmaskVT, op: LogicalOperator, ctx: Optional[Context] = None) -> Bitmask:
    """bitmask_binop (dispatch.rs:47-56)."""
    ctx = ctx or default_context()
    (lm, lo, ln), (rm, ro, _) = lhs, rhs
    return dev.bits_binop(ctx, op, _up(ctx, lm), lo, _up(ctx, rm), ro, ln).download()


def and_masks(lhs: BitmaskVT, rhs: BitmaskVT, ctx=None) -> Bitmask:
    return bitmask_binop(lhs, rhs, LogicalOperator.And, ctx)


def or_masks(lhs: BitmaskVT, rhs: BitmaskVT, ctx=None) -> Bitmask:
    return bitmask_binop(lhs, rhs, LogicalOperator.Or, ctx)


def xor_masks(lhs: BitmaskVT, rhs: BitmaskVT, ctx=None) -> Bitmask:
    return bitmask_binop(lhs, rhs, LogicalOperator.Xor, ctx)


def not_mask(src: BitmaskVT, ctx=None) -> Bitmask:
    """not_mask (dispatch.rs:135-144)."""
    ctx = ctx or default_context()
    m, off, ln = src
    return dev.bits_not(ctx, _up(ctx, m), off, ln).download()


def popcount_mask(m: BitmaskVT, ctx=None) -> int:
    """popcount_mask (dispatch.rs:258-267)."""
    ctx = ctx or default_context()
    mask, off, ln = m
    return dev.bits_popcount(ctx, _up(ctx, mask), off, ln)


def count_ones(mask: Bitmask, ctx=None) -> int:
    """Bitmask::count_ones (src/structs/bitmask.rs:393-406)."""
    return popcount_mask((mask, 0, mask.len), ctx)


def null_count(mask: Bitmask, ctx=None) -> int:
    """Bitmask::null_count = count_zeros (bitmask.rs:409-417)."""
    return mask.len - count_ones(mask, ctx)


def all_true_mask(mask: Bitmask, ctx=None) -> bool:
    ctx = ctx or default_context()
    return dev.bits_all_true(ctx, _up(ctx, mask))


def all_false_mask(mask: Bitmask, ctx=None) -> bool:
    ctx = ctx or default_context()
    return dev.bits_all_false(ctx, _up(ctx, mask))


def eq_mask(a: BitmaskVT, b: BitmaskVT, ctx=None) -> Bitmask:
    ctx = ctx or default_context()
    (am, ao, ln), (bm, bo, _) = a, b
    return dev.bits_eq(ctx, _up(ctx, am), ao, _up(ctx, bm), bo, ln, False).download()


def ne_mask(a: BitmaskVT, b: BitmaskVT, ctx=None) -> Bitmask:
    ctx = ctx or default_context()
    (am, ao, ln), (bm, bo, _) = a, b
    return dev.bits_eq(ctx, _up(ctx, am), ao, _up(ctx, bm), bo, ln, True).download()


def all_eq(a: BitmaskVT, b: BitmaskVT, ctx=None) -> bool:
    ctx = ctx or default_context()
    (am, ao, ln), (bm, bo, _) = a, b
    return dev.bits_all_eq(ctx, _up(ctx, am), ao, _up(ctx, bm), bo, ln)


def all_ne(a: BitmaskVT, b: BitmaskVT, ctx=None) -> bool:
    """all_ne_mask_simd is `!all_eq_mask_simd` in the reference (simd.rs:490-494); kept as is."""
    return not all_eq(a, b, ctx)


def in_mask(lhs: BitmaskVT, rhs: BitmaskVT, ctx=None) -> Bitmask:
    ctx = ctx or default_context()
    (lm, lo, ln), (rm, ro, _) = lhs, rhs
    return dev.bits_in(ctx, _up(ctx, lm), lo, _up(ctx, rm), ro, ln, False).download()


def not_in_mask(lhs: BitmaskVT, rhs: BitmaskVT, ctx=None) -> Bitmask:
    ctx = ctx or default_context()
    (lm, lo, ln), (rm, ro, _) = lhs, rhs
    return dev.bits_in(ctx, _up(ctx, lm), lo, _up(ctx, rm), ro, ln, True).download()


def merge_bitmasks_to_new(lhs: Optional[Bitmask], rhs: Optional[Bitmask], length: int, ctx=None) -> Optional[Bitmask]:
    """merge_bitmasks_to_new — per-row AND (src/kernels/bitmask/mod.rs:171-197)."""
    ctx = ctx or default_context()
    out = dev.bits_merge(ctx, None if lhs is None else _up(ctx, lhs), None if rhs is None else _up(ctx, rhs), length,
                         MaskMode.And)
    return None if out is None else out.download()


def union(a: Bitmask, b: Bitmask, ctx=None) -> Bitmask:
    """Bitmask::union — bitwise OR (src/structs/bitmask.rs:661-669)."""
    assert a.len == b.len, "Bitmask::union length mismatch"
    ctx = ctx or default_context()
    return dev.bits_merge(ctx, _up(ctx, a), _up(ctx, b), a.len, MaskMode.Or).download()


def intersect(a: Bitmask, b: Bitmask, ctx=None) -> Bitmask:
    """Bitmask::intersect — bitwise AND (bitmask.rs:673-681)."""
    assert a.len == b.len, "Bitmask::intersect length mismatch"
    ctx = ctx or default_context()
    return dev.bits_merge(ctx, _up(ctx, a), _up(ctx, b), a.len, MaskMode.And).download()


def invert(a: Bitmask, ctx=None) -> Bitmask:
    """Bitmask::invert (bitmask.rs:685-692)."""
    return not_mask((a, 0, a.len), ctx)
