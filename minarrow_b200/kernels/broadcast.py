"""Container fan-out of the broadcast layer that feeds the leaf kernels, on HOST containers: the SuperArray route
(src/kernels/broadcast/super_array.rs:180-249), the Table route (table.rs:31-62), the SuperTable route
(super_table.rs:38-73), the Array <-> SuperArray re-chunk arms (mod.rs:1351-1375 + utils.rs:367-481) and the view
variants (ArrayV / SuperArrayV / TableV).  Each route uploads its operands once, runs the device-resident route of
`minarrow_b200.containers` — every leaf call of the operation in batched launches — and downloads the result.  Callers
that keep their data in HBM use `minarrow_b200.containers` directly and skip both copies.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .. import containers as dc
from ..core import (ArithmeticOperator, Bitmask, Context, DeviceBitmask, KernelError, ShapeError, default_context)


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


@dataclass
class ArrayV:
    """`ArrayV = (Array, offset, len)` window (src/structs/views/array_view.rs:79-94)."""
    array: object
    offset: int = 0
    len: int = -1

    def __post_init__(self):
        n = len(self.array)
        if self.len < 0:
            self.len = n - self.offset
        if self.offset < 0 or self.offset + self.len > n:
            raise KernelError("OutOfBounds", f"ArrayV window [{self.offset}, {self.offset + self.len}) of an array of {n}")

    def __len__(self) -> int:
        return self.len

    def slice(self, offset: int, length: int) -> "ArrayV":
        """`ArrayV::slice` — a window of the window."""
        if offset + length > self.len:
            raise KernelError("OutOfBounds", "ArrayV::slice out of bounds")
        return ArrayV(self.array, self.offset + offset, length)


@dataclass
class SuperArrayV:
    """`SuperArrayV {slices: Vec<ArrayV>, len}` (src/structs/views/chunked/super_array_view.rs)."""
    slices: List[ArrayV] = field(default_factory=list)

    def __len__(self) -> int:
        return sum(len(s) for s in self.slices)


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
class TableV:
    """`TableV {cols: Vec<ArrayV>, offset, len}` (src/structs/views/table_view.rs): the same row window of every column."""
    table: Table
    offset: int = 0
    len: int = -1

    def __post_init__(self):
        if self.len < 0:
            self.len = self.table.n_rows() - self.offset
        if self.offset < 0 or self.offset + self.len > self.table.n_rows():
            raise KernelError("OutOfBounds", "TableV window exceeds the table")

    def n_cols(self) -> int:
        return self.table.n_cols()


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


@dataclass
class SuperTableV:
    """`SuperTableV {slices: Vec<TableV>, len}` (src/structs/views/chunked/super_table_view.rs): row windows of tables."""
    slices: List[TableV] = field(default_factory=list)

    @property
    def len(self) -> int:
        return sum(s.len for s in self.slices)

    def n_slices(self) -> int:
        return len(self.slices)


def _is_scalar(x) -> bool:
    return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool)


def _window(arr, offset: int, length: int):
    """Host copy of rows [offset, offset + len) of a typed array: values slice + validity bits re-based to bit 0
    (`slice_clone`)."""
    from ..core import make_array
    data = np.ascontiguousarray(getattr(arr, "data", arr))[offset:offset + length]
    m = getattr(arr, "null_mask", None)
    if m is not None:
        m = Bitmask.from_bools(m.to_bools()[offset:offset + length])
    return make_array(data, m)


def to_device(ctx: Context, v):
    """Host Value -> device-resident Value (scalars stay on the host)."""
    if _is_scalar(v) or isinstance(v, (dc.DeviceArray, dc.DeviceSuperArray, dc.DeviceTable, dc.DeviceSuperTable)):
        return v
    if isinstance(v, SuperTable):
        return dc.DeviceSuperTable.from_host(ctx, v)
    if isinstance(v, Table):
        return dc.DeviceTable.from_host(ctx, v)
    if isinstance(v, TableV):
        return dc.DeviceTable(v.table.name, [dc.DeviceArray.from_host(ctx, _window(c, v.offset, v.len)) for c in v.table.cols])
    if isinstance(v, SuperTableV):
        return dc.DeviceSuperTable([to_device(ctx, s) for s in v.slices])
    if isinstance(v, tuple):
        return tuple(to_device(ctx, x) for x in v)
    if isinstance(v, list):
        return [to_device(ctx, x) for x in v]
    if isinstance(v, SuperArray):
        return dc.DeviceSuperArray.from_host(ctx, v)
    if isinstance(v, SuperArrayV):
        return dc.DeviceSuperArray([dc.DeviceArray.from_host(ctx, _window(s.array, s.offset, s.len)) for s in v.slices])
    if isinstance(v, ArrayV):
        return dc.DeviceArray.from_host(ctx, _window(v.array, v.offset, v.len))
    return dc.DeviceArray.from_host(ctx, v)


def _to_host(v):
    if isinstance(v, tuple):
        return tuple(_to_host(x) for x in v)
    if isinstance(v, list):
        return [_to_host(x) for x in v]
    return v.to_host() if hasattr(v, "to_host") else v      # scalars stay as they are


def _run(op, lhs, rhs, ctx, fn=dc.broadcast_value, **kw):
    ctx = ctx or default_context()
    return _to_host(fn(op, to_device(ctx, lhs), to_device(ctx, rhs), ctx=ctx, **kw))


def route_super_array_broadcast(op: ArithmeticOperator, lhs: SuperArray, rhs: SuperArray,
                                null_mask_override: Optional[Bitmask] = None, ctx: Optional[Context] = None) -> SuperArray:
    """Per-chunk arithmetic.  Validity per chunk: override if given, else the OR-union of the two chunks'
    masks, else the one present mask (super_array.rs:214-230) — fused into the kernel (MaskMode.Or); one batched call."""
    ctx = ctx or default_context()
    ov = None if null_mask_override is None else DeviceBitmask.upload(ctx, null_mask_override)
    return dc.route_super_array_broadcast(op, to_device(ctx, lhs), to_device(ctx, rhs), ov, ctx).to_host()


def create_aligned_chunks_from_array(array, super_array: SuperArray, ctx: Optional[Context] = None) -> SuperArray:
    """src/utils.rs:417-481: `array` split to the SuperArray's chunk lengths, each chunk carrying its window of the union
    of the array's mask and the concatenated chunk masks (computed on the device)."""
    ctx = ctx or default_context()
    return dc.create_aligned_chunks_from_array(to_device(ctx, array), to_device(ctx, super_array)).to_host()


def broadcast_table_with_operator(op: ArithmeticOperator, lhs, rhs, ctx=None):
    """Table route (table.rs:31-62): column i of lhs against column i of rhs through the router with NO mask (so the
    result columns carry no validity and a dense integer zero divisor is the reference's panic).  Returns a `Table`
    named after the left one (or a list when plain column lists were passed)."""
    as_list = not isinstance(lhs, Table)
    lt = lhs if isinstance(lhs, Table) else Table("", list(lhs))
    rt = rhs if isinstance(rhs, Table) else Table("", list(rhs))
    if lt.n_cols() != rt.n_cols():
        raise ShapeError(f"Table column count mismatch: {lt.n_cols()} vs {rt.n_cols()}")
    out = _run(op, lt, rt, ctx, dc.broadcast_table_with_operator)
    return out.cols if as_list else out


def broadcast_table_to_array(op: ArithmeticOperator, table: Table, arr, ctx=None) -> Table:
    """`table op array`: every column against the same array (table.rs:179-228)."""
    return _run(op, table, arr, ctx)


def broadcast_array_to_table(op: ArithmeticOperator, arr, table: Table, ctx=None) -> Table:
    """`array op table`: operand order is significant (array.rs:187-236)."""
    return _run(op, arr, table, ctx)


def broadcast_table_to_scalar(op: ArithmeticOperator, table: Table, scalar, ctx=None) -> Table:
    """`table op scalar`: the same scalar against every column (table.rs:230-261)."""
    return _run(op, table, scalar, ctx)


def broadcast_scalar_to_table(op: ArithmeticOperator, scalar, table: Table, ctx=None) -> Table:
    return _run(op, scalar, table, ctx)


def broadcast_super_table_with_operator(op: ArithmeticOperator, lhs, rhs, ctx=None):
    """SuperTable route (super_table.rs:38-73): batch by batch through the Table route; all batches x columns go out in one
    batched call (one launch per column dtype)."""
    as_list = not isinstance(lhs, SuperTable)
    ls = lhs if isinstance(lhs, SuperTable) else SuperTable([b if isinstance(b, Table) else Table("", list(b)) for b in lhs])
    rs = rhs if isinstance(rhs, SuperTable) else SuperTable([b if isinstance(b, Table) else Table("", list(b)) for b in rhs])
    if ls.n_batches() != rs.n_batches():
        raise ShapeError(f"SuperTable chunk count mismatch: {ls.n_batches()} vs {rs.n_batches()}")
    out = _run(op, ls, rs, ctx, dc.broadcast_super_table_with_operator)
    return [b.cols for b in out.batches] if as_list else out


def _device_mask(ctx, null_mask):
    from ..core import DeviceBitmask
    return None if null_mask is None else DeviceBitmask.upload(ctx, null_mask)


def broadcast_table_add(lhs, rhs, null_mask: Optional[Bitmask] = None, ctx=None) -> Table:
    """`broadcast_table_add(lhs: impl Into<TableV>, rhs, null_mask)` (table.rs:69-128): Table or TableV operands; the optional
    mask is handed to every column's kernel (`broadcast_array_add`)."""
    ctx = ctx or default_context()
    return _to_host(dc.broadcast_table_add(to_device(ctx, lhs), to_device(ctx, rhs), _device_mask(ctx, null_mask), ctx))


def broadcast_super_table_add(lhs, rhs, null_mask: Optional[Bitmask] = None, ctx=None) -> SuperTable:
    """`broadcast_super_table_add(lhs: impl Into<SuperTableV>, rhs, null_mask)` (table.rs:135-176): SuperTable or SuperTableV
    operands, chunk by chunk; all chunks x columns in one batched call."""
    ctx = ctx or default_context()
    return _to_host(dc.broadcast_super_table_add(to_device(ctx, lhs), to_device(ctx, rhs), _device_mask(ctx, null_mask), ctx))


def broadcast_table_to_superarray(op: ArithmeticOperator, table, sa: SuperArray, ctx=None) -> SuperArray:
    """table.rs:382-406: every chunk of the SuperArray against the whole (single-column) table."""
    return _run(op, table, sa, ctx, dc.broadcast_table_to_superarray, table_is_lhs=True)


def broadcast_superarray_to_table(op: ArithmeticOperator, sa: SuperArray, table, ctx=None) -> SuperArray:
    """super_array.rs:153-176: the mirror, chunks on the left."""
    return _run(op, table, sa, ctx, dc.broadcast_table_to_superarray, table_is_lhs=False)


def broadcast_arrayview_to_superarray(op: ArithmeticOperator, view: ArrayV, sa, ctx=None) -> SuperArray:
    """ArrayView (op) SuperArray / SuperArrayView (super_array.rs:255-309, 367-419): NO mask, chunk by chunk."""
    ctx = ctx or default_context()
    return dc.broadcast_arrayview_to_superarray(op, to_device(ctx, view), to_device(ctx, sa), True, ctx).to_host()


def broadcast_superarray_to_arrayview(op: ArithmeticOperator, sa, view: ArrayV, ctx=None) -> SuperArray:
    """SuperArray / SuperArrayView (op) ArrayView (super_array.rs:311-365, 421-470)."""
    ctx = ctx or default_context()
    return dc.broadcast_arrayview_to_superarray(op, to_device(ctx, view), to_device(ctx, sa), False, ctx).to_host()


def broadcast_tableview_to_tableview(op: ArithmeticOperator, lhs: TableV, rhs: TableV, ctx=None) -> Table:
    """table_view.rs:25-60: the windows of column i against each other, no mask; the result Table is unnamed."""
    if lhs.n_cols() != rhs.n_cols():
        raise ShapeError(f"TableView column count mismatch: {lhs.n_cols()} vs {rhs.n_cols()}")
    out = _run(op, lhs, rhs, ctx, dc.broadcast_table_with_operator)
    out.name = ""
    return out


def broadcast_supertableview_to_arrayview(op: ArithmeticOperator, stv: SuperTableV, view: ArrayV, ctx=None) -> SuperTable:
    """super_table_view.rs:66-105: table slice i against the aligned window of the ArrayView; the ArrayView must be as long
    as the SuperTableView (ShapeError "... does not match ...").  Slices come back materialised (`TableV::from_table(result,
    0, n)` in the reference), i.e. as the batches of a SuperTable."""
    return _run(op, stv, view, ctx, dc.broadcast_supertableview_to_arrayview, stv_is_lhs=True)


def broadcast_arrayview_to_supertableview(op: ArithmeticOperator, view: ArrayV, stv: SuperTableV, ctx=None) -> SuperTable:
    """array_view.rs `broadcast_arrayview_to_supertableview`: the mirror, ArrayView windows on the left."""
    return _run(op, stv, view, ctx, dc.broadcast_supertableview_to_arrayview, stv_is_lhs=False)


def broadcast_supertableview_to_scalar(op: ArithmeticOperator, stv: SuperTableV, scalar, ctx=None) -> SuperTable:
    """super_table_view.rs:28-62: every slice, every column against the typed scalar — one batched call."""
    return _run(op, stv, scalar, ctx)


def broadcast_superarrayview_to_tableview(op: ArithmeticOperator, sav, table, ctx=None) -> SuperTable:
    """super_array_view.rs:22-80 / super_table_view.rs:108-154 (`broadcast_superarrayview_to_table`): the Table / TableView
    is cut into slices aligned with the SuperArrayView's and slice i meets slice i, SuperArrayView on the left."""
    return _run(op, sav, table, ctx, dc.broadcast_superarrayview_to_tableview)


def broadcast_value(op: ArithmeticOperator, lhs, rhs, ctx=None):
    """`broadcast_value(op, Value, Value)` (src/kernels/broadcast/mod.rs:152-...) for the Value variants on this path:
    Scalar (python / numpy number), Array (numpy, IntegerArray, FloatArray), ArrayV, SuperArray, SuperArrayV, Table, TableV,
    SuperTable.  Array-level routes pass no mask (mod.rs:166-209); the SuperArray route merges chunk masks by OR-union;
    Array (op) SuperArray / SuperArrayV re-chunks the array to the chunk lengths with the union mask (mod.rs:1351-1375);
    ArrayV (op) SuperArray goes chunk by chunk with no mask (super_array.rs:255-365); Array (op) TableV takes the table
    view's window of the array (mod.rs:1386-1393)."""
    if _is_scalar(lhs) and _is_scalar(rhs):
        return dc.broadcast_value(op, lhs, rhs)      # scalar_arithmetic on the host: nothing to launch (mod.rs:161-163)
    ctx = ctx or default_context()
    L, R = lhs, rhs
    chunked = (SuperArray, SuperArrayV)
    plain = lambda x: not isinstance(x, (Table, TableV, SuperTable, SuperTableV, ArrayV, tuple, list)) and not _is_scalar(x) and not isinstance(x, chunked)
    # SuperTableView arms (mod.rs:521-528, 612-618, 644-660, 1394-...): aligned windows, slice by slice
    if isinstance(L, SuperTableV) or isinstance(R, SuperTableV):
        stv, other, stv_is_lhs = (L, R, True) if isinstance(L, SuperTableV) else (R, L, False)
        if _is_scalar(other):
            return _run(op, L, R, ctx)                                                  # broadcast_supertableview_to_scalar
        if isinstance(other, ArrayV):
            return broadcast_supertableview_to_arrayview(op, stv, other, ctx) if stv_is_lhs else broadcast_arrayview_to_supertableview(op, other, stv, ctx)
        if isinstance(other, Table):
            return _run(op, stv, other, ctx, dc.broadcast_supertableview_to_table, stv_is_lhs=stv_is_lhs)
        if plain(other):
            return _run(op, stv, other, ctx, dc.broadcast_supertableview_to_arrayview, stv_is_lhs=stv_is_lhs, check_len=False)
        raise KernelError("UnsupportedType", f"no route for {type(L).__name__} (op) {type(R).__name__}")
    if isinstance(L, SuperArrayV) and isinstance(R, (Table, TableV)):
        return broadcast_superarrayview_to_tableview(op, L, R, ctx)
    # an OWNING SuperArray against a Table / TableView (mod.rs:622-641): every chunk meets the whole table
    if isinstance(L, SuperArray) and isinstance(R, (Table, TableV)):
        return broadcast_superarray_to_table(op, L, R, ctx)
    if isinstance(L, (Table, TableV)) and isinstance(R, SuperArray):
        return broadcast_table_to_superarray(op, L, R, ctx)
    if isinstance(L, ArrayV) and isinstance(R, chunked):
        return broadcast_arrayview_to_superarray(op, L, R, ctx)
    if isinstance(L, chunked) and isinstance(R, ArrayV):
        return broadcast_superarray_to_arrayview(op, L, R, ctx)
    if isinstance(R, TableV) and plain(L):
        L = ArrayV(L, R.offset, R.len)          # Array (op) TableView: ArrayV::new(array, tv.offset, tv.len)
    if isinstance(L, TableV) and plain(R):
        R = ArrayV(R, L.offset, L.len)
    if isinstance(L, TableV) and isinstance(R, TableV):
        return broadcast_tableview_to_tableview(op, L, R, ctx)
    return _run(op, L, R, ctx)


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
