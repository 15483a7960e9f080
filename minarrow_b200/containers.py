"""Device-resident containers and the container routes over them — operands stay in HBM between calls.

The reference's containers are host objects whose routes call the leaf kernel once per chunk and per column
(`SuperArray` src/structs/chunked/super_array.rs:96-103, `Table` src/structs/table.rs:103-115, `SuperTable`
src/structs/chunked/super_table.rs:78-83; routes src/kernels/broadcast/super_array.rs:180-249, table.rs:31-62,
super_table.rs:38-73, mod.rs:152-...).  Here the same containers hold `DeviceBuffer` / `DeviceBitmask` handles, a route
gathers every (chunk, column) leaf call of the operation and issues them through the batched C-ABI entry points
(`mnr_ew_binary_batch`, `mnr_ew_scalar_batch_into`, `mnr_reduce_stats_batch_exchange`): one launch per (dtype, alignment,
masked) class instead of one per leaf, outputs are device-resident containers again, nothing visits the host.
`table * table` on a 64-batch x 4-column SuperTable is 4 launches (one per column dtype).

Views are free on the device: `DeviceArray.view(offset, len)` is the `ArrayV` window (src/structs/views/array_view.rs:79-94)
— a pointer offset for the values, a lazily materialised bit-offset slice for the validity — so `SuperArrayV` is a
`DeviceSuperArray` of views and `TableV` a `DeviceTable` of views (`DeviceTable.view`).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from . import device_ops as dev
from .core import (ArithmeticOperator, Bitmask, Context, DeviceBitmask, DeviceBuffer, KernelError, MaskMode, ShapeError,
                   check, default_context, dtype_code, make_array)

_ROUTED = {np.dtype(t) for t in (np.int32, np.int64, np.uint32, np.uint64, np.float32, np.float64)}


class DeviceArray:
    """Device-resident `IntegerArray<T>` / `FloatArray<T>`: values in a `DeviceBuffer` + optional validity bitmask.
    A length-1 array created from a host scalar keeps the value on the host too, so the scalar-broadcast route
    (`maybe_broadcast_scalar_array`, routing/broadcast.rs:87-112) never reads the device."""

    def __init__(self, buf: DeviceBuffer, mask: Optional[DeviceBitmask] = None, host_scalar=None, _lazy_mask=None):
        self.buf, self._mask, self.host_scalar, self._lazy = buf, mask, host_scalar, _lazy_mask

    @classmethod
    def from_host(cls, ctx: Context, arr) -> "DeviceArray":
        data = np.ascontiguousarray(getattr(arr, "data", arr))
        m = getattr(arr, "null_mask", None)
        hs = data.reshape(-1)[0] if data.size == 1 else None
        return cls(DeviceBuffer.upload(ctx, data), None if m is None else DeviceBitmask.upload(ctx, m), hs)

    @classmethod
    def scalar(cls, ctx: Context, value, dtype) -> "DeviceArray":
        """`Scalar` -> length-1 array of `dtype` (broadcast/scalar.rs:169-210, array.rs:139-184)."""
        v = np.array([value], dtype=dtype)
        return cls(DeviceBuffer.upload(ctx, v), None, v[0])

    @property
    def ctx(self) -> Context:
        return self.buf.ctx

    @property
    def dtype(self) -> np.dtype:
        return self.buf.dtype

    def __len__(self) -> int:
        return len(self.buf)

    @property
    def null_mask(self) -> Optional[DeviceBitmask]:
        if self._mask is None and self._lazy is not None:   # a view's validity: exact bit-offset slice, made on first use
            parent, off, ln = self._lazy
            self._mask = dev.bits_slice(self.ctx, parent, off, ln)
            self._lazy = None
        return self._mask

    def has_mask(self) -> bool:
        return self._mask is not None or self._lazy is not None

    def view(self, offset: int, length: int) -> "DeviceArray":
        """`ArrayV::new(array, offset, len)` / `Array::view` — zero-copy values window; validity sliced lazily."""
        if offset < 0 or length < 0 or offset + length > len(self):
            raise KernelError("OutOfBounds", f"view [{offset}, {offset + length}) of an array of {len(self)}")
        lazy = None
        if self._mask is not None:
            lazy = (self._mask, offset, length)
        elif self._lazy is not None:
            lazy = (self._lazy[0], self._lazy[1] + offset, length)
        hs = self.host_scalar if (length == 1 and offset == 0) else None
        return DeviceArray(self.buf.slice(offset, length), None, hs, lazy)

    def first_value(self):
        """data[0] of the underlying array (what `broadcast_length_1_array` reads, routing/broadcast.rs:29-46)."""
        if self.host_scalar is None:
            self.host_scalar = self.buf.slice(0, 1).download()[0]
        return self.host_scalar

    def to_host(self):
        m = self.null_mask
        return make_array(self.buf.download(), None if m is None else m.download())


@dataclass
class DeviceSuperArray:
    """`SuperArray` with device-resident chunks (also the `SuperArrayV` analogue: chunks may be views)."""
    chunks: List[DeviceArray] = field(default_factory=list)

    @classmethod
    def from_host(cls, ctx: Context, sa) -> "DeviceSuperArray":
        return cls([DeviceArray.from_host(ctx, c) for c in sa.chunks])

    @classmethod
    def from_slices(cls, slices: Sequence[DeviceArray]) -> "DeviceSuperArray":
        """`SuperArray::from_slices(&view.slices, field)` (used by the SuperArrayView arms, mod.rs:1362-1375)."""
        return cls(list(slices))

    def __len__(self) -> int:
        return sum(len(c) for c in self.chunks)

    def n_chunks(self) -> int:
        return len(self.chunks)

    def shape_1d(self):
        return [len(c) for c in self.chunks]

    def to_host(self):
        from .kernels.broadcast import SuperArray
        return SuperArray([c.to_host() for c in self.chunks])


@dataclass
class DeviceTable:
    """`Table {cols, name}` with device-resident columns (also the `TableV` analogue: columns may be views)."""
    name: str = ""
    cols: List[DeviceArray] = field(default_factory=list)

    @classmethod
    def from_host(cls, ctx: Context, t) -> "DeviceTable":
        return cls(t.name, [DeviceArray.from_host(ctx, c) for c in t.cols])

    def n_cols(self) -> int:
        return len(self.cols)

    def n_rows(self) -> int:
        return len(self.cols[0]) if self.cols else 0

    def view(self, offset: int, length: int) -> "DeviceTable":
        """`TableV::from_table(table, offset, len)` / `TableV::from_self`: the same row window of every column."""
        return DeviceTable(self.name, [c.view(offset, length) for c in self.cols])

    def to_host(self):
        from .kernels.broadcast import Table
        return Table(self.name, [c.to_host() for c in self.cols])


@dataclass
class DeviceSuperTable:
    """`SuperTable {batches, name}` with device-resident batches."""
    batches: List[DeviceTable] = field(default_factory=list)
    name: str = ""

    @classmethod
    def from_host(cls, ctx: Context, st) -> "DeviceSuperTable":
        return cls([DeviceTable.from_host(ctx, b) for b in st.batches], st.name)

    def n_batches(self) -> int:
        return len(self.batches)

    def n_rows(self) -> int:
        return sum(b.n_rows() for b in self.batches)

    def n_cols(self) -> int:
        return self.batches[0].n_cols() if self.batches else 0

    def to_host(self):
        from .kernels.broadcast import SuperTable
        return SuperTable([b.to_host() for b in self.batches], self.name)


# ---- the router over many leaf calls at once ------------------------------------------------------------------------------------
@dataclass
class _Leaf:
    lhs: DeviceArray
    rhs: DeviceArray
    lmask: Optional[DeviceBitmask] = None
    rmask: Optional[DeviceBitmask] = None


def route_leaves(op: ArithmeticOperator, leaves: Sequence[_Leaf], mode: int = MaskMode.And, ctx: Optional[Context] = None) -> List[DeviceArray]:
    """`resolve_binary_arithmetic` (routing/arithmetic.rs:214-406) for a whole list of leaf calls: per leaf the same checks
    and routes as the reference (length-1 broadcast, same dtype on the six routed types, (i32, f64) / (i32, f32)
    promotion, else UnsupportedType / LengthMismatch), then equal-length same-dtype leaves share `mnr_ew_binary_batch`
    launches, scalar-broadcast leaves share `mnr_ew_scalar_batch_into` launches, promotions go one by one."""
    if not leaves:
        return []
    ctx = ctx or leaves[0].lhs.ctx
    out: List[Optional[DeviceArray]] = [None] * len(leaves)
    same, scal_l, scal_r, promo = [], [], [], []
    for i, lf in enumerate(leaves):
        ln, rn = len(lf.lhs), len(lf.rhs)
        if ln != rn and ln != 1 and rn != 1:
            raise KernelError("LengthMismatch", f"cannot broadcast arrays of length {ln} and {rn}")
        lt, rt = lf.lhs.dtype, lf.rhs.dtype
        if lt != rt:
            if {lt, rt} not in ({np.dtype(np.int32), np.dtype(np.float64)}, {np.dtype(np.int32), np.dtype(np.float32)}):
                raise KernelError("UnsupportedType", "Unsupported array type combination for arithmetic operations")
            promo.append(i)
        elif lt not in _ROUTED:
            raise KernelError("UnsupportedType", "Unsupported array type combination for arithmetic operations")
        elif ln == rn:
            same.append(i)
        elif ln == 1:
            scal_l.append(i)
        else:
            scal_r.append(i)
    if same:
        obs, oms = dev.ew_binary_batch(ctx, op, [leaves[i].lhs.buf for i in same], [leaves[i].rhs.buf for i in same],
                                       [leaves[i].lmask for i in same], [leaves[i].rmask for i in same], mode)
        for i, ob, om in zip(same, obs, oms):
            out[i] = DeviceArray(ob, om)
    for idx, scalar_is_lhs in ((scal_l, True), (scal_r, False)):
        if not idx:
            continue
        arrs, scalars, masks, obs, oms = [], [], [], [], []
        for i in idx:
            lf = leaves[i]
            arr, sc = (lf.rhs, lf.lhs) if scalar_is_lhs else (lf.lhs, lf.rhs)
            m = lf.lmask if lf.lmask is not None else lf.rmask
            if lf.lmask is not None and lf.rmask is not None:
                m = dev.bits_merge(ctx, lf.lmask, lf.rmask, len(arr), mode)
            arrs.append(arr.buf); scalars.append(sc.first_value()); masks.append(m)
            obs.append(DeviceBuffer.alloc(ctx, arr.dtype, len(arr)))
            oms.append(None if m is None else DeviceBitmask.alloc(ctx, len(arr)))
        dev.ew_scalar_batch_into(ctx, op, arrs, scalars, scalar_is_lhs, masks, obs, oms)
        for i, ob, om in zip(idx, obs, oms):
            out[i] = DeviceArray(ob, om)
    for i in promo:
        lf = leaves[i]
        ln, rn = len(lf.lhs), len(lf.rhs)
        if ln == rn:
            ob, om = dev.ew_binary_promote(ctx, op, lf.lhs.buf, lf.rhs.buf, lf.lmask, lf.rmask, mode)
        else:   # scalar broadcast + promotion: the length-1 side becomes a typed scalar of the float dtype
            scalar_is_lhs = ln == 1
            arr, sc = (lf.rhs, lf.lhs) if scalar_is_lhs else (lf.lhs, lf.rhs)
            out_dt = lf.lhs.dtype if lf.lhs.dtype.kind == "f" else lf.rhs.dtype
            m = lf.lmask if lf.lmask is not None else lf.rmask
            if lf.lmask is not None and lf.rmask is not None:
                m = dev.bits_merge(ctx, lf.lmask, lf.rmask, len(arr), mode)
            if arr.dtype != out_dt:
                # i32 column against a float scalar: the router casts the column (`x as f64`, routing/arithmetic.rs:244-269)
                # and broadcasts the scalar.  Rare corner: the scalar is materialised once as a float column and the
                # promote kernel casts the i32 column on load.
                sc_col = DeviceBuffer.upload(ctx, np.full(len(arr), sc.first_value(), dtype=out_dt))
                if scalar_is_lhs:
                    ob, om = dev.ew_binary_promote(ctx, op, sc_col, arr.buf, None, m, MaskMode.And)
                else:
                    ob, om = dev.ew_binary_promote(ctx, op, arr.buf, sc_col, m, None, MaskMode.And)
            else:
                ob, om = dev.ew_scalar(ctx, op, arr.buf, np.asarray(sc.first_value()).astype(out_dt), scalar_is_lhs, m)
        out[i] = DeviceArray(ob, om)
    return out  # type: ignore[return-value]


def resolve_binary_arithmetic(op: ArithmeticOperator, lhs: DeviceArray, rhs: DeviceArray, null_mask: Optional[DeviceBitmask] = None,
                              ctx: Optional[Context] = None) -> DeviceArray:
    """The router on device-resident operands; the operands' own masks are NOT consulted (the caller's job, as in the
    reference)."""
    return route_leaves(op, [_Leaf(lhs, rhs, null_mask, None)], MaskMode.And, ctx)[0]


# ---- SuperArray routes --------------------------------------------------------------------------------------------------------------
def route_super_array_broadcast(op: ArithmeticOperator, lhs: DeviceSuperArray, rhs: DeviceSuperArray,
                                null_mask_override: Optional[DeviceBitmask] = None, ctx: Optional[Context] = None) -> DeviceSuperArray:
    """broadcast/super_array.rs:180-249 on the device: chunk i against chunk i; validity = the override, else the UNION
    of the two chunks' masks, else the one present (fused into the kernel, MaskMode.Or); ONE batched call."""
    leaves = []
    for i, lc in enumerate(lhs.chunks):
        if i >= len(rhs.chunks):
            raise ShapeError(f"Super Array broadcasting error for {op!r} - chunk count: LHS {lhs.n_chunks()} RHS {rhs.n_chunks()}")
        rc = rhs.chunks[i]
        if len(lc) != len(rc):
            raise ShapeError(f"Super Array broadcasting error for {op!r} - Chunk: LHS {len(lc)} RHS {len(rc)}, "
                             f"Shape: LHS {lhs.shape_1d()} RHS {rhs.shape_1d()}")
        if null_mask_override is not None:
            leaves.append(_Leaf(lc, rc, null_mask_override, None))
        else:
            leaves.append(_Leaf(lc, rc, lc.null_mask, rc.null_mask))
    return DeviceSuperArray(route_leaves(op, leaves, MaskMode.Or, ctx))


def union_array_superarray_masks(array: DeviceArray, sa: DeviceSuperArray) -> Optional[DeviceBitmask]:
    """src/utils.rs:367-413 on the device: the chunk masks concatenated at bit granularity (`mnr_concat`'s validity
    gather; a chunk without a mask counts as all valid once ANY chunk has one) OR-ed with the array's mask."""
    ctx = array.ctx
    sa_mask = None
    if any(c.has_mask() for c in sa.chunks):
        _, sa_mask = dev.concat(ctx, [c.buf for c in sa.chunks], [c.null_mask for c in sa.chunks])
    am = array.null_mask
    if am is not None and sa_mask is not None:
        if len(am) != len(sa_mask):
            raise ShapeError(f"Mask lengths must match for union: {len(am)} vs {len(sa_mask)}")
        return dev.bits_merge(ctx, am, sa_mask, len(am), MaskMode.Or)
    return am if am is not None else sa_mask


def create_aligned_chunks_from_array(array: DeviceArray, sa: DeviceSuperArray) -> DeviceSuperArray:
    """src/utils.rs:417-481 on the device: `array` re-chunked to `sa`'s chunk lengths — value chunks are zero-copy windows,
    every chunk carries its window of the full union mask (exact bit offsets, `mnr_bits_slice`).  This is the re-shard
    step that lets a plain column meet a chunked (sharded) one."""
    if len(array) != len(sa):
        raise ShapeError(f"Array and SuperArray must have same total length for broadcasting: {len(array)} vs {len(sa)}")
    full = union_array_superarray_masks(array, sa)
    carrier = DeviceArray(array.buf, full)
    out, start = [], 0
    for c in sa.chunks:
        out.append(carrier.view(start, len(c)))
        start += len(c)
    return DeviceSuperArray(out)


def broadcast_array_to_superarray(op: ArithmeticOperator, array: DeviceArray, sa: DeviceSuperArray, array_is_lhs: bool = True,
                                  ctx: Optional[Context] = None) -> DeviceSuperArray:
    """`Value::Array (op) Value::SuperArray` and its mirror (broadcast/mod.rs:1351-1361) without materialising the aligned
    SuperArray's union masks: chunk i of the result is valid where `array_mask[window i] | chunk_mask[i]` — exactly what
    create_aligned_chunks_from_array + route_super_array_broadcast produce, because (A|S)|S = A|S — so the array's window
    and the chunk's own mask go straight into the fused OR of the kernel.  A chunk WITHOUT a mask next to chunks with one
    is all-valid in the union (utils.rs:386-388) and so is its result."""
    if len(array) != len(sa):
        raise ShapeError(f"Array and SuperArray must have same total length for broadcasting: {len(array)} vs {len(sa)}")
    ctx = ctx or array.ctx
    sa_any = any(c.has_mask() for c in sa.chunks)
    leaves, start = [], 0
    for c in sa.chunks:
        w = array.view(start, len(c))
        start += len(c)
        if c.has_mask():
            am, cm = w.null_mask, c.null_mask
        elif sa_any:
            am, cm = DeviceBitmask.new_set_all(ctx, len(c), True), None   # all-valid in the union -> all-valid result mask
        else:
            am, cm = w.null_mask, None
        leaves.append(_Leaf(w, c, am, cm) if array_is_lhs else _Leaf(c, w, cm, am))
    return DeviceSuperArray(route_leaves(op, leaves, MaskMode.Or, ctx))


def broadcast_arrayview_to_superarray(op: ArithmeticOperator, view: DeviceArray, sa: DeviceSuperArray, view_is_lhs: bool = True,
                                      ctx: Optional[Context] = None) -> DeviceSuperArray:
    """ArrayView (op) SuperArray / SuperArrayView and the mirrors (broadcast/super_array.rs:255-470): the view is cut to the
    chunk lengths (`array_view.slice`) and every pair goes through the Array-level route, which passes NO mask."""
    if len(view) != len(sa):
        raise ShapeError(f"ArrayView length ({len(view)}) does not match SuperArray length ({len(sa)})")
    leaves, start = [], 0
    for c in sa.chunks:
        w = view.view(start, len(c))
        start += len(c)
        leaves.append(_Leaf(w, c) if view_is_lhs else _Leaf(c, w))
    return DeviceSuperArray(route_leaves(op, leaves, MaskMode.And, ctx))


# ---- Table / SuperTable routes ----------------------------------------------------------------------------------------------------
def _table_leaves(lhs: DeviceTable, rhs: DeviceTable) -> List[_Leaf]:
    if lhs.n_cols() != rhs.n_cols():
        raise ShapeError(f"Table column count mismatch: {lhs.n_cols()} vs {rhs.n_cols()}")
    return [_Leaf(l, r) for l, r in zip(lhs.cols, rhs.cols)]   # table.rs:54: no mask


def broadcast_table_with_operator(op: ArithmeticOperator, lhs: DeviceTable, rhs: DeviceTable, ctx: Optional[Context] = None) -> DeviceTable:
    """Table route (table.rs:31-62; TableView form table_view.rs:25-60): column i against column i, no mask; ONE batched call."""
    return DeviceTable(lhs.name, route_leaves(op, _table_leaves(lhs, rhs), MaskMode.And, ctx))


def broadcast_super_table_with_operator(op: ArithmeticOperator, lhs: DeviceSuperTable, rhs: DeviceSuperTable,
                                        ctx: Optional[Context] = None) -> DeviceSuperTable:
    """SuperTable route (super_table.rs:38-73): batch by batch through the Table route — all batches x columns in ONE
    batched call (one launch per column dtype)."""
    if lhs.n_batches() != rhs.n_batches():
        raise ShapeError(f"SuperTable chunk count mismatch: {lhs.n_batches()} vs {rhs.n_batches()}")
    leaves, shape = [], []
    for lb, rb in zip(lhs.batches, rhs.batches):
        tl = _table_leaves(lb, rb)
        leaves += tl
        shape.append((lb.name, len(tl)))
    res = route_leaves(op, leaves, MaskMode.And, ctx)
    out, k = [], 0
    for name, n in shape:
        out.append(DeviceTable(name, res[k:k + n]))
        k += n
    return DeviceSuperTable(out, lhs.name)


def broadcast_table_add(lhs: DeviceTable, rhs: DeviceTable, null_mask: Optional[DeviceBitmask] = None,
                        ctx: Optional[Context] = None) -> DeviceTable:
    """`broadcast_table_add(lhs, rhs, null_mask)` (table.rs:69-128): column i + column i through `broadcast_array_add`, i.e.
    `resolve_binary_arithmetic(Add, l, r, null_mask)` — the SAME optional mask for every column (one row count), the columns'
    own validity not consulted; ONE batched call.  Shape errors are the reference's BroadcastingError texts."""
    if lhs.n_cols() != rhs.n_cols():
        raise KernelError("BroadcastingError", f"Table column count mismatch: LHS {lhs.n_cols()} cols, RHS {rhs.n_cols()} cols")
    if lhs.n_rows() != rhs.n_rows():
        raise KernelError("BroadcastingError", f"Table row count mismatch: LHS {lhs.n_rows()} rows, RHS {rhs.n_rows()} rows")
    leaves = [_Leaf(l, r, null_mask, None) for l, r in zip(lhs.cols, rhs.cols)]
    return DeviceTable(lhs.name, route_leaves(ArithmeticOperator.Add, leaves, MaskMode.And, ctx))


def broadcast_super_table_add(lhs: DeviceSuperTable, rhs: DeviceSuperTable, null_mask: Optional[DeviceBitmask] = None,
                              ctx: Optional[Context] = None) -> DeviceSuperTable:
    """`broadcast_super_table_add` (table.rs:135-176): chunk i + chunk i through `broadcast_table_add` with the same optional
    mask; all chunks x columns in ONE batched call.  The result takes the first slice's name, or "SuperTable"."""
    if lhs.n_batches() != rhs.n_batches():
        raise KernelError("BroadcastingError", f"SuperTable chunk count mismatch: LHS {lhs.n_batches()} chunks, RHS {rhs.n_batches()} chunks")
    leaves, shape = [], []
    for i, (lb, rb) in enumerate(zip(lhs.batches, rhs.batches)):
        if lb.n_cols() != rb.n_cols():
            raise KernelError("BroadcastingError", f"Chunk {i} addition failed: Table column count mismatch: LHS {lb.n_cols()} cols, RHS {rb.n_cols()} cols")
        if lb.n_rows() != rb.n_rows():
            raise KernelError("BroadcastingError", f"Chunk {i} addition failed: Table row count mismatch: LHS {lb.n_rows()} rows, RHS {rb.n_rows()} rows")
        leaves += [_Leaf(l, r, null_mask, None) for l, r in zip(lb.cols, rb.cols)]
        shape.append((lb.name, lb.n_cols()))
    res = route_leaves(ArithmeticOperator.Add, leaves, MaskMode.And, ctx)
    out, k = [], 0
    for name, n in shape:
        out.append(DeviceTable(name, res[k:k + n]))
        k += n
    name = lhs.batches[0].name if lhs.batches and lhs.batches[0].name else "SuperTable"
    return DeviceSuperTable(out, name)


def broadcast_table_to_superarray(op: ArithmeticOperator, table: DeviceTable, sa: DeviceSuperArray, table_is_lhs: bool = True,
                                  ctx: Optional[Context] = None) -> DeviceSuperArray:
    """`table op superarray` (table.rs:382-406) / `superarray op table` (super_array.rs:153-176): every CHUNK against the whole
    table through the Table-Array route, which must leave a single column — so the table has one column as long as each
    chunk; one batched call over the chunks."""
    if table.n_cols() != 1:
        raise ShapeError(("Table-SuperArray" if table_is_lhs else "SuperArray-Table") + " broadcasting should result in single column")
    c = table.cols[0]
    leaves = [_Leaf(c, ch) if table_is_lhs else _Leaf(ch, c) for ch in sa.chunks]
    return DeviceSuperArray(route_leaves(op, leaves, MaskMode.And, ctx))


def broadcast_table_to_array(op: ArithmeticOperator, table: DeviceTable, arr: DeviceArray, table_is_lhs: bool = True,
                             ctx: Optional[Context] = None) -> DeviceTable:
    """`table op array` / `array op table` (table.rs:179-228, array.rs:187-236; view forms table_view.rs:108-146): every
    column against the same array (or length-1 scalar array), operand order kept."""
    leaves = [_Leaf(c, arr) if table_is_lhs else _Leaf(arr, c) for c in table.cols]
    return DeviceTable(table.name, route_leaves(op, leaves, MaskMode.And, ctx))


def broadcast_table_to_scalar(op: ArithmeticOperator, table: DeviceTable, scalar, table_is_lhs: bool = True,
                              ctx: Optional[Context] = None) -> DeviceTable:
    """`table op scalar` (table.rs:230-261): the same scalar, typed like each column, against every column."""
    ctx = ctx or (table.cols[0].ctx if table.cols else default_context())
    leaves = []
    for c in table.cols:
        s = DeviceArray.scalar(ctx, scalar, c.dtype)
        leaves.append(_Leaf(c, s) if table_is_lhs else _Leaf(s, c))
    return DeviceTable(table.name, route_leaves(op, leaves, MaskMode.And, ctx))


def broadcast_super_table_to_scalar(op: ArithmeticOperator, st: DeviceSuperTable, scalar, table_is_lhs: bool = True,
                                    ctx: Optional[Context] = None) -> DeviceSuperTable:
    """`supertable op scalar` (super_table.rs:77-91): every batch, every column, one batched call."""
    ctx = ctx or default_context()
    leaves, shape = [], []
    for b in st.batches:
        for c in b.cols:
            s = DeviceArray.scalar(ctx, scalar, c.dtype)
            leaves.append(_Leaf(c, s) if table_is_lhs else _Leaf(s, c))
        shape.append((b.name, b.n_cols()))
    res = route_leaves(op, leaves, MaskMode.And, ctx)
    out, k = [], 0
    for name, n in shape:
        out.append(DeviceTable(name, res[k:k + n]))
        k += n
    return DeviceSuperTable(out, st.name)


def broadcast_tableview_to_superarrayview(op: ArithmeticOperator, table_view: DeviceTable, sav: DeviceSuperArray,
                                          ctx: Optional[Context] = None) -> DeviceSuperTable:
    """table_view.rs:148-200: the TableView is promoted to aligned slices (`from_self(offset, chunk_len)`) and slice i
    meets SuperArrayView slice i column by column; all slices x columns in one batched call."""
    if table_view.n_rows() != len(sav):
        raise ShapeError(f"TableView length ({table_view.n_rows()}) does not match SuperArrayView length ({len(sav)})")
    leaves, start = [], 0
    for s in sav.chunks:
        leaves += [_Leaf(c.view(start, len(s)), s) for c in table_view.cols]
        start += len(s)
    res = route_leaves(op, leaves, MaskMode.And, ctx)
    nc = table_view.n_cols()
    return DeviceSuperTable([DeviceTable(table_view.name, res[k * nc:(k + 1) * nc]) for k in range(len(sav.chunks))], table_view.name)


def broadcast_superarrayview_to_tableview(op: ArithmeticOperator, sav: DeviceSuperArray, table_view: DeviceTable,
                                          ctx: Optional[Context] = None) -> DeviceSuperTable:
    """super_array_view.rs:22-80 and super_table_view.rs:108-154 (`broadcast_superarrayview_to_table`): the mirror of
    `broadcast_tableview_to_superarrayview` — SuperArrayView slice i on the LEFT of the aligned TableView slice i."""
    if len(sav) != table_view.n_rows():
        raise ShapeError(f"SuperArrayView length ({len(sav)}) does not match TableView length ({table_view.n_rows()})")
    leaves, start = [], 0
    for s in sav.chunks:
        leaves += [_Leaf(s, c.view(start, len(s))) for c in table_view.cols]
        start += len(s)
    res = route_leaves(op, leaves, MaskMode.And, ctx)
    nc = table_view.n_cols()
    return DeviceSuperTable([DeviceTable(table_view.name, res[k * nc:(k + 1) * nc]) for k in range(len(sav.chunks))], table_view.name)


def _split_like(res: List[DeviceArray], stv: DeviceSuperTable) -> DeviceSuperTable:
    out, k = [], 0
    for b in stv.batches:
        out.append(DeviceTable(b.name, res[k:k + b.n_cols()]))
        k += b.n_cols()
    return DeviceSuperTable(out, stv.name)


def broadcast_supertableview_to_arrayview(op: ArithmeticOperator, stv: DeviceSuperTable, av: DeviceArray, stv_is_lhs: bool = True,
                                          check_len: bool = True, ctx: Optional[Context] = None) -> DeviceSuperTable:
    """`supertableview op arrayview` and the mirror (super_table_view.rs:66-105, array_view.rs `broadcast_arrayview_to_
    supertableview`): table slice i meets the window [sum of earlier slice lengths, + its own length) of the ArrayView,
    column by column; all slices x columns in one batched call.  `check_len=False` is the Array form
    (`broadcast_supertableview_to_array`, :157-180; `broadcast_array_to_supertableview`, array.rs:451-476), which only needs
    the array to be long enough."""
    if check_len and len(av) != stv.n_rows():
        raise ShapeError(f"ArrayView length ({len(av)}) does not match SuperTableView length ({stv.n_rows()})")
    leaves, start = [], 0
    for b in stv.batches:
        w = av.view(start, b.n_rows())
        leaves += [_Leaf(c, w) if stv_is_lhs else _Leaf(w, c) for c in b.cols]
        start += b.n_rows()
    return _split_like(route_leaves(op, leaves, MaskMode.And, ctx), stv)


def broadcast_supertableview_to_table(op: ArithmeticOperator, stv: DeviceSuperTable, table: DeviceTable, stv_is_lhs: bool = True,
                                      ctx: Optional[Context] = None) -> DeviceSuperTable:
    """`supertableview op table` / `table op supertableview` (super_table_view.rs:183-250): the Table is promoted to slices
    aligned with the SuperTableView's and slice i meets slice i through the TableView route (equal column counts)."""
    if stv.n_rows() != table.n_rows():
        raise ShapeError(f"SuperTableView length ({stv.n_rows()}) does not match Table rows ({table.n_rows()})" if stv_is_lhs else
                         f"Table rows ({table.n_rows()}) does not match SuperTableView length ({stv.n_rows()})")
    leaves, start = [], 0
    for b in stv.batches:
        if b.n_cols() != table.n_cols():
            raise ShapeError(f"TableView column count mismatch: {b.n_cols() if stv_is_lhs else table.n_cols()} vs {table.n_cols() if stv_is_lhs else b.n_cols()}")
        for c, t in zip(b.cols, table.cols):
            w = t.view(start, b.n_rows())
            leaves.append(_Leaf(c, w) if stv_is_lhs else _Leaf(w, c))
        start += b.n_rows()
    return _split_like(route_leaves(op, leaves, MaskMode.And, ctx), stv)


# ---- Value-level dispatch -----------------------------------------------------------------------------------------------------------
def broadcast_array_to_supertable(op: ArithmeticOperator, arr: DeviceArray, st: DeviceSuperTable, array_is_lhs: bool = True,
                                 ctx: Optional[Context] = None) -> DeviceSuperTable:
    """`array op supertable` / `supertable op array` (array.rs:236-252, super_table.rs `broadcast_supertable_to_array`): the
    array against every column of every batch (`broadcast_array_to_table` per batch, so its length must be each batch's
    row count or 1); all batches x columns in one batched call."""
    leaves, shape = [], []
    for b in st.batches:
        for c in b.cols:
            leaves.append(_Leaf(arr, c) if array_is_lhs else _Leaf(c, arr))
        shape.append((b.name, b.n_cols()))
    res = route_leaves(op, leaves, MaskMode.And, ctx)
    out, k = [], 0
    for name, n in shape:
        out.append(DeviceTable(name, res[k:k + n]))
        k += n
    return DeviceSuperTable(out, st.name)


def scalar_arithmetic(lhs, rhs, op: ArithmeticOperator):
    """`scalar_arithmetic(Scalar, Scalar, op)` (routing/arithmetic.rs:34-209): two host scalars — no array, no launch.  Same
    type -> same type for Add / Subtract / Multiply / Divide (integer division truncates, integers wrap like the release
    build); int (op) float promotes the integer; anything else is NotImplemented."""
    a, b = np.asarray(lhs), np.asarray(rhs)
    if a.dtype.kind in "iu" and b.dtype.kind == "f":
        a = a.astype(b.dtype)
    elif a.dtype.kind == "f" and b.dtype.kind in "iu":
        b = b.astype(a.dtype)
    if a.dtype != b.dtype or a.dtype.kind not in "iuf" or int(op) > int(ArithmeticOperator.Divide):
        raise KernelError("NotImplemented", f"Scalar arithmetic operation {ArithmeticOperator(op).name} between {a.dtype} and {b.dtype}")
    if int(op) == int(ArithmeticOperator.Divide) and a.dtype.kind in "iu":
        if b == 0:
            raise KernelError("DivideByZero", "attempt to divide by zero")
        q = abs(int(a)) // abs(int(b))
        return a.dtype.type(np.array(-q if (int(a) < 0) != (int(b) < 0) else q).astype(a.dtype))
    with np.errstate(all="ignore"):
        r = {int(ArithmeticOperator.Add): np.add, int(ArithmeticOperator.Subtract): np.subtract,
             int(ArithmeticOperator.Multiply): np.multiply, int(ArithmeticOperator.Divide): np.divide}[int(op)](a, b)
    return a.dtype.type(r)


def _is_scalar(x) -> bool:
    return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool)


def broadcast_value(op: ArithmeticOperator, lhs, rhs, ctx: Optional[Context] = None):
    """`broadcast_value(op, Value, Value)` (broadcast/mod.rs:152-...) on device-resident Values: Scalar (python / numpy
    number), DeviceArray (also ArrayView), DeviceSuperArray (also SuperArrayView), DeviceTable (also TableView),
    DeviceSuperTable, python tuples of 2-6 Values (Tuple2..Tuple6) and lists (VecValue).  The result is device-resident as
    well, so `a * b + c` chains in HBM."""
    L, R = lhs, rhs
    if _is_scalar(L) and _is_scalar(R):
        # mod.rs:161-163 passes ArithmeticOperator::Add to scalar_arithmetic whatever `op` is; a drop-in returns what the
        # reference returns
        return scalar_arithmetic(L, R, ArithmeticOperator.Add)
    # Tuple2..Tuple6 (mod.rs:305-353) and VecValue (:356-372): element-wise; an array-like operand against a tuple meets every
    # element (array.rs:269-447, mod.rs:766-907)
    if isinstance(L, tuple) and isinstance(R, tuple):
        if len(L) != len(R) or not 2 <= len(L) <= 6:
            raise KernelError("UnsupportedType", f"no route for Tuple{len(L)} (op) Tuple{len(R)}")
        return tuple(broadcast_value(op, a, b, ctx) for a, b in zip(L, R))
    if isinstance(L, list) and isinstance(R, list):
        if len(L) != len(R):
            raise KernelError("LengthMismatch", f"ColumnLengthMismatch: col 0, expected {len(L)}, found {len(R)}")
        return [broadcast_value(op, a, b, ctx) for a, b in zip(L, R)]
    if isinstance(R, tuple) and (isinstance(L, (DeviceArray, DeviceSuperArray)) or _is_scalar(L)) and 2 <= len(R) <= 6:
        return tuple(broadcast_value(op, L, b, ctx) for b in R)       # array / scalar against every element (scalar.rs:300-420)
    if isinstance(L, tuple) and (isinstance(R, (DeviceArray, DeviceSuperArray)) or _is_scalar(R)) and 2 <= len(L) <= 6:
        return tuple(broadcast_value(op, a, R, ctx) for a in L)
    if isinstance(L, DeviceArray) and isinstance(R, DeviceSuperTable):
        return broadcast_array_to_supertable(op, L, R, True, ctx)
    if isinstance(L, DeviceSuperTable) and isinstance(R, DeviceArray):
        return broadcast_array_to_supertable(op, R, L, False, ctx)
    if isinstance(L, DeviceSuperTable) and isinstance(R, DeviceSuperTable):
        return broadcast_super_table_with_operator(op, L, R, ctx)
    if isinstance(L, DeviceSuperTable) and _is_scalar(R):
        return broadcast_super_table_to_scalar(op, L, R, True, ctx)
    if _is_scalar(L) and isinstance(R, DeviceSuperTable):
        return broadcast_super_table_to_scalar(op, R, L, False, ctx)
    if isinstance(L, DeviceTable) and isinstance(R, DeviceTable):
        return broadcast_table_with_operator(op, L, R, ctx)
    if isinstance(L, DeviceTable):
        if _is_scalar(R):
            return broadcast_table_to_scalar(op, L, R, True, ctx)
        if isinstance(R, DeviceArray):
            return broadcast_table_to_array(op, L, R, True, ctx)
        if isinstance(R, DeviceSuperArray):
            return broadcast_tableview_to_superarrayview(op, L, R, ctx)
    if isinstance(R, DeviceTable):
        if _is_scalar(L):
            return broadcast_table_to_scalar(op, R, L, False, ctx)
        if isinstance(L, DeviceArray):
            return broadcast_table_to_array(op, R, L, False, ctx)
    if isinstance(L, DeviceSuperArray) and isinstance(R, DeviceSuperArray):
        return route_super_array_broadcast(op, L, R, None, ctx)
    if isinstance(L, DeviceSuperArray):
        if isinstance(R, DeviceArray):
            return broadcast_array_to_superarray(op, R, L, False, ctx)
        if _is_scalar(R):   # broadcast_superarray_to_scalar (super_array.rs:87-118): every chunk, no mask
            return DeviceSuperArray(route_leaves(op, [_Leaf(c, DeviceArray.scalar(c.ctx, R, c.dtype)) for c in L.chunks], MaskMode.And, ctx))
    if isinstance(R, DeviceSuperArray):
        if isinstance(L, DeviceArray):
            return broadcast_array_to_superarray(op, L, R, True, ctx)
        if _is_scalar(L):
            return DeviceSuperArray(route_leaves(op, [_Leaf(DeviceArray.scalar(c.ctx, L, c.dtype), c) for c in R.chunks], MaskMode.And, ctx))
    if isinstance(L, DeviceArray) and isinstance(R, DeviceArray):
        return resolve_binary_arithmetic(op, L, R, None, ctx)
    if isinstance(L, DeviceArray) and _is_scalar(R):
        return resolve_binary_arithmetic(op, L, DeviceArray.scalar(L.ctx, R, L.dtype), None, ctx)
    if _is_scalar(L) and isinstance(R, DeviceArray):
        return resolve_binary_arithmetic(op, DeviceArray.scalar(R.ctx, L, R.dtype), R, None, ctx)
    raise KernelError("UnsupportedType", f"no device route for {type(L).__name__} (op) {type(R).__name__}")


# ---- null-aware aggregates over containers ---------------------------------------------------------------------------------------------
def _agg_dict(dtype, agg, lib) -> dict:
    f = {"i": "i64", "u": "u64", "f": "f64"}[np.dtype(dtype).kind]
    return {"sum": getattr(agg.sum, f), "min": getattr(agg.min, f), "max": getattr(agg.max, f), "count": int(agg.count),
            "mean": float(lib.mnr_agg_mean(dtype_code(dtype), C.byref(agg)))}


def column_stats(columns: Sequence[Sequence[DeviceArray]], with_minmax: bool = True, ctx: Optional[Context] = None) -> List[dict]:
    """{sum, min, max, count, mean} of every column, where column c is the chunk list `columns[c]` (a SuperArray's chunks, a
    Table's single column, a SuperTable's column across its batches): ALL chunks of ALL columns reduced by one batched
    call, folded per column in chunk order on the device (`mnr_reduce_stats_batch_exchange` with no exchange)."""
    ctx = ctx or next((ch.ctx for col in columns for ch in col), None) or default_context()
    bufs, vals, cols, dts = [], [], [], []
    for c, col in enumerate(columns):
        if not col:
            raise KernelError("InvalidArguments", f"column {c} has no chunks (its dtype is unknown)")
        dts.append(col[0].dtype)
        for ch in col:
            bufs.append(ch.buf); vals.append(ch.null_mask); cols.append(c)
    n = len(bufs)
    aggs = (_lib.Agg * len(dts))()
    hb = dev._handle_array(bufs)
    hv = dev._handle_array(vals)
    check(ctx.lib.mnr_reduce_stats_batch_exchange_sync(ctx.h, None, n, hb, hv, int(with_minmax), len(dts),
                                                       (C.c_uint32 * max(1, n))(*cols), (C.c_int * len(dts))(*[dtype_code(d) for d in dts]), aggs))
    return [_agg_dict(dt, a, ctx.lib) for dt, a in zip(dts, aggs)]


def super_array_stats(sa: DeviceSuperArray, with_minmax: bool = True) -> dict:
    return column_stats([sa.chunks], with_minmax)[0]


def table_stats(t: DeviceTable, with_minmax: bool = True) -> List[dict]:
    return column_stats([[c] for c in t.cols], with_minmax)


def super_table_stats(st: DeviceSuperTable, with_minmax: bool = True) -> List[dict]:
    """Per-column aggregates over all batches of a SuperTable (BASELINE configs[4]: per-column sum/min/max)."""
    return column_stats([[b.cols[c] for b in st.batches] for c in range(st.n_cols())], with_minmax)
