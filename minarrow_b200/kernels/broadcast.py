"""Container fan-out of the broadcast layer that feeds the leaf kernels: the SuperArray route
(src/kernels/broadcast/super_array.rs:180-249), the Table route (table.rs:31-62) and the SuperTable route
(super_table.rs:38-73).  Only the decomposition semantics live here — per chunk / per column, operand order,
chunk-length checks, and which validity merge each route uses; the arithmetic is one fused kernel per chunk.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from .. import device_ops as dev
from ..core import (ArithmeticOperator, Bitmask, Context, DeviceBitmask, DeviceBuffer, KernelError, MaskMode, ShapeError,
                    default_context, make_array)
from .routing import resolve_binary_arithmetic


@dataclass
class SuperArray:
    """`SuperArray {chunks: Vec<Array>, ..}` (src/structs/chunked/super_array.rs:96-103): equal-dtype chunks."""
    chunks: List = field(default_factory=list)

    def __len__(self) -> int:
        return sum(len(c) for c in self.chunks)

    def n_chunks(self) -> int:
        return len(self.chunks)

    def shape_1d(self):
        return [len(c) for c in self.chunks]


def route_super_array_broadcast(op: ArithmeticOperator, lhs: SuperArray, rhs: SuperArray,
                                null_mask_override: Optional[Bitmask] = None, ctx: Optional[Context] = None) -> SuperArray:
    """Per-chunk arithmetic.  Validity per chunk: override if given, else the OR-union of the two chunks'
    masks, else the one present mask (super_array.rs:214-230) — fused into the kernel (MaskMode.Or)."""
    ctx = ctx or default_context()
    out = SuperArray()
    # Chunk pairs that can share one batched launch (same dtype and length, at least one mask, no override):
    # upload them all, one mnr_ew_binary_batch call, download.  Everything else takes the per-chunk router.
    slots, bl, br, blm, brm = [], [], [], [], []
    for i, lc in enumerate(lhs.chunks):
        rc = rhs.chunks[i]
        if len(lc) != len(rc):
            raise ShapeError(f"Super Array broadcasting error for {op!r} - Chunk: LHS {len(lc)} RHS {len(rc)}, "
                             f"Shape: LHS {lhs.shape_1d()} RHS {rhs.shape_1d()}")
        if null_mask_override is not None:
            out.chunks.append(resolve_binary_arithmetic(op, lc, rc, null_mask_override, ctx))
            continue
        lm, rm = getattr(lc, "null_mask", None), getattr(rc, "null_mask", None)
        if lm is None and rm is None:
            out.chunks.append(resolve_binary_arithmetic(op, lc, rc, None, ctx))
            continue
        ld, rd = np.ascontiguousarray(lc.data), np.ascontiguousarray(rc.data)
        if ld.dtype != rd.dtype or ld.size != rd.size:
            merged = lm if rm is None else rm if lm is None else None
            if merged is None:
                from .bitmask import union
                merged = union(lm, rm, ctx)
            out.chunks.append(resolve_binary_arithmetic(op, lc, rc, merged, ctx))
            continue
        slots.append(len(out.chunks))
        out.chunks.append(None)
        bl.append(DeviceBuffer.upload(ctx, ld))
        br.append(DeviceBuffer.upload(ctx, rd))
        blm.append(None if lm is None else DeviceBitmask.upload(ctx, lm))
        brm.append(None if rm is None else DeviceBitmask.upload(ctx, rm))
    if slots:
        obs, oms = dev.ew_binary_batch(ctx, op, bl, br, blm, brm, MaskMode.Or)
        for slot, ob, om in zip(slots, obs, oms):
            out.chunks[slot] = make_array(ob.download(), om.download())
    return out


def broadcast_table_with_operator(op: ArithmeticOperator, lhs_cols: Sequence, rhs_cols: Sequence, ctx=None) -> list:
    """Table route (table.rs:31-62): column i of lhs against column i of rhs, no mask passed (None)."""
    if len(lhs_cols) != len(rhs_cols):
        raise KernelError("BroadcastingError", f"Table column count mismatch: LHS {len(lhs_cols)} RHS {len(rhs_cols)}")
    return [resolve_binary_arithmetic(op, l, r, None, ctx) for l, r in zip(lhs_cols, rhs_cols)]


def broadcast_super_table_with_operator(op: ArithmeticOperator, lhs_batches: Sequence[Sequence],
                                        rhs_batches: Sequence[Sequence], ctx=None) -> list:
    """SuperTable route (super_table.rs:38-73): batch by batch through the Table route."""
    if len(lhs_batches) != len(rhs_batches):
        raise KernelError("BroadcastingError",
                          f"SuperTable batch count mismatch: LHS {len(lhs_batches)} RHS {len(rhs_batches)}")
    return [broadcast_table_with_operator(op, l, r, ctx) for l, r in zip(lhs_batches, rhs_batches)]
