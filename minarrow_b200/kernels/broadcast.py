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


@dataclass
class Table:
    """`Table {cols, n_rows, name}` (src/structs/table.rs:103-115) reduced to what the route needs: named columns."""
    name: str = ""
    cols: List = field(default_factory=list)

    def n_cols(self) -> int:
        return len(self.cols)

    def n_rows(self) -> int:
        return len(self.cols[0]) if self.cols else 0


@dataclass
class SuperTable:
    """`SuperTable {batches: Vec<Arc<Table>>, ..}` (src/structs/chunked/super_table.rs:78-83)."""
    batches: List[Table] = field(default_factory=list)
    name: str = ""

    def n_batches(self) -> int:
        return len(self.batches)

    def n_rows(self) -> int:
        return sum(b.n_rows() for b in self.batches)

    def n_cols(self) -> int:
        return self.batches[0].n_cols() if self.batches else 0


def _cols(t):
    return t.cols if isinstance(t, Table) else list(t)


def broadcast_table_with_operator(op: ArithmeticOperator, lhs, rhs, ctx=None):
    """Table route (table.rs:31-62): column i of lhs against column i of rhs through the router with NO mask (so the
    result columns carry no validity and a dense integer zero divisor is the reference's panic).  Returns a `Table`
    named after the left one (or a list when plain column lists were passed)."""
    lc, rc = _cols(lhs), _cols(rhs)
    if len(lc) != len(rc):
        raise ShapeError(f"Table column count mismatch: {len(lc)} vs {len(rc)}")
    out = [resolve_binary_arithmetic(op, l, r, None, ctx) for l, r in zip(lc, rc)]
    return Table(lhs.name, out) if isinstance(lhs, Table) else out


def broadcast_table_to_array(op: ArithmeticOperator, table: Table, arr, ctx=None) -> Table:
    """`table op array`: every column against the same array (table.rs broadcast_table_to_array)."""
    return Table(table.name, [resolve_binary_arithmetic(op, c, arr, None, ctx) for c in table.cols])


def broadcast_array_to_table(op: ArithmeticOperator, arr, table: Table, ctx=None) -> Table:
    """`array op table`: operand order is significant (array.rs broadcast_array_to_table)."""
    return Table(table.name, [resolve_binary_arithmetic(op, arr, c, None, ctx) for c in table.cols])


def _scalar_array(scalar, like):
    """Scalar -> length-1 array of the column's dtype (scalar.rs:169-210, array.rs:139-184)."""
    return np.array([scalar], dtype=np.asarray(getattr(like, "data", like)).dtype)


def broadcast_table_to_scalar(op: ArithmeticOperator, table: Table, scalar, ctx=None) -> Table:
    """`table op scalar`: the same scalar against every column (table.rs:230-261)."""
    return Table(table.name, [resolve_binary_arithmetic(op, c, _scalar_array(scalar, c), None, ctx) for c in table.cols])


def broadcast_scalar_to_table(op: ArithmeticOperator, scalar, table: Table, ctx=None) -> Table:
    return Table(table.name, [resolve_binary_arithmetic(op, _scalar_array(scalar, c), c, None, ctx) for c in table.cols])


def broadcast_super_table_with_operator(op: ArithmeticOperator, lhs, rhs, ctx=None):
    """SuperTable route (super_table.rs:38-73): batch by batch through the Table route."""
    lb = lhs.batches if isinstance(lhs, SuperTable) else list(lhs)
    rb = rhs.batches if isinstance(rhs, SuperTable) else list(rhs)
    if len(lb) != len(rb):
        raise ShapeError(f"SuperTable chunk count mismatch: {len(lb)} vs {len(rb)}")
    out = [broadcast_table_with_operator(op, l, r, ctx) for l, r in zip(lb, rb)]
    return SuperTable(out, getattr(lhs, "name", "")) if isinstance(lhs, SuperTable) else out


def _is_scalar(x) -> bool:
    return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool)


def broadcast_value(op: ArithmeticOperator, lhs, rhs, ctx=None):
    """`broadcast_value(op, Value, Value)` (src/kernels/broadcast/mod.rs:152-...) for the Value variants on this path:
    Scalar (python / numpy number), Array (numpy, IntegerArray, FloatArray), SuperArray, Table, SuperTable.
    Array-level routes pass no mask (mod.rs:166-209); the SuperArray route merges chunk masks by OR-union."""
    if isinstance(lhs, SuperTable) and isinstance(rhs, SuperTable):
        return broadcast_super_table_with_operator(op, lhs, rhs, ctx)
    if isinstance(lhs, SuperTable) and _is_scalar(rhs):
        return SuperTable([broadcast_table_to_scalar(op, b, rhs, ctx) for b in lhs.batches], lhs.name)
    if _is_scalar(lhs) and isinstance(rhs, SuperTable):
        return SuperTable([broadcast_scalar_to_table(op, lhs, b, ctx) for b in rhs.batches], rhs.name)
    if isinstance(lhs, Table) and isinstance(rhs, Table):
        return broadcast_table_with_operator(op, lhs, rhs, ctx)
    if isinstance(lhs, Table):
        return broadcast_table_to_scalar(op, lhs, rhs, ctx) if _is_scalar(rhs) else broadcast_table_to_array(op, lhs, rhs, ctx)
    if isinstance(rhs, Table):
        return broadcast_scalar_to_table(op, lhs, rhs, ctx) if _is_scalar(lhs) else broadcast_array_to_table(op, lhs, rhs, ctx)
    if isinstance(lhs, SuperArray) and isinstance(rhs, SuperArray):
        return route_super_array_broadcast(op, lhs, rhs, None, ctx)
    if isinstance(lhs, SuperArray):
        rr = (lambda c: _scalar_array(rhs, c)) if _is_scalar(rhs) else (lambda c: rhs)
        return SuperArray([resolve_binary_arithmetic(op, c, rr(c), None, ctx) for c in lhs.chunks])
    if isinstance(rhs, SuperArray):
        ll = (lambda c: _scalar_array(lhs, c)) if _is_scalar(lhs) else (lambda c: lhs)
        return SuperArray([resolve_binary_arithmetic(op, ll(c), c, None, ctx) for c in rhs.chunks])
    if _is_scalar(lhs) and _is_scalar(rhs):
        raise KernelError("UnsupportedType", "Scalar op Scalar is host arithmetic, not a kernel route")
    if _is_scalar(lhs):
        return resolve_binary_arithmetic(op, _scalar_array(lhs, rhs), rhs, None, ctx)
    if _is_scalar(rhs):
        return resolve_binary_arithmetic(op, lhs, _scalar_array(rhs, lhs), None, ctx)
    return resolve_binary_arithmetic(op, lhs, rhs, None, ctx)


def value_add(lhs, rhs, ctx=None):
    """value_add .. value_power (src/kernels/broadcast/mod.rs:116-150)."""
    return broadcast_value(ArithmeticOperator.Add, lhs, rhs, ctx)


def value_subtract(lhs, rhs, ctx=None):
    return broadcast_value(ArithmeticOperator.Subtract, lhs, rhs, ctx)


def value_multiply(lhs, rhs, ctx=None):
    return broadcast_value(ArithmeticOperator.Multiply, lhs, rhs, ctx)


def value_divide(lhs, rhs, ctx=None):
    return broadcast_value(ArithmeticOperator.Divide, lhs, rhs, ctx)


def value_remainder(lhs, rhs, ctx=None):
    return broadcast_value(ArithmeticOperator.Remainder, lhs, rhs, ctx)


def value_power(lhs, rhs, ctx=None):
    return broadcast_value(ArithmeticOperator.Power, lhs, rhs, ctx)
