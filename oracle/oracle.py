"""numpy front-end for the CPU oracle (oracle/mnr_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (minarrow_b200/) never does.

Function names and argument order follow the reference's leaf API
(src/kernels/arithmetic/dispatch.rs:74-79,147-152,221-226; src/kernels/bitmask/dispatch.rs:96-295):
slices in, `(data, mask_bytes | None)` out, `KernelError` for the reference's `Err(KernelError::..)`
and for its panics (dense integer division by zero).
A bitmask is `(bytes: np.ndarray[uint8], len_bits: int)`, Arrow layout (LSB first, 1 = valid).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import NamedTuple, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmnr_oracle.so")

ADD, SUB, MUL, DIV, REM, POW, FLOORDIV = range(7)
OPS = {"add": ADD, "subtract": SUB, "multiply": MUL, "divide": DIV, "remainder": REM, "power": POW,
       "floordiv": FLOORDIV}
AND, OR, XOR = range(3)


class KernelError(Exception):
    def __init__(self, kind: str, msg: str = ""):
        super().__init__(f"{kind}: {msg}")
        self.kind = kind


class Bits(NamedTuple):
    """Arrow validity / boolean bitmask: ceil(len/8) bytes, LSB first, slack bits zero."""
    bits: np.ndarray
    len: int

    def to_bools(self) -> np.ndarray:
        return np.unpackbits(self.bits, bitorder="little")[: self.len].astype(bool)

    @staticmethod
    def from_bools(b) -> "Bits":
        b = np.asarray(b, dtype=bool)
        return Bits(np.packbits(b, bitorder="little"), int(b.size))


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mnr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_bits_count_ones.restype = C.c_uint64
        _lib.orc_bits_null_count.restype = C.c_uint64
        _lib.orc_popcount_mask.restype = C.c_uint64
        _lib.orc_simd_sum_i64.restype = C.c_int64
        _lib.orc_hotloop_sum_i64.restype = C.c_int64
        _lib.orc_rayon_simd_sum_i64.restype = C.c_int64
        _lib.orc_simd_sum_f64.restype = C.c_double
        _lib.orc_hotloop_sum_f64.restype = C.c_double
        _lib.orc_rayon_simd_sum_f64.restype = C.c_double
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _sz(n) -> C.c_size_t:
    return C.c_size_t(int(n))


_NAMES = {np.dtype(np.int8): "i8", np.dtype(np.uint8): "u8", np.dtype(np.int16): "i16",
          np.dtype(np.uint16): "u16", np.dtype(np.int32): "i32", np.dtype(np.uint32): "u32",
          np.dtype(np.int64): "i64", np.dtype(np.uint64): "u64", np.dtype(np.float32): "f32",
          np.dtype(np.float64): "f64"}


def _mask_bytes(mask, n: int) -> Optional[np.ndarray]:
    if mask is None:
        return None
    bits = mask.bits if isinstance(mask, Bits) else np.asarray(mask, dtype=np.uint8)
    need = (n + 7) // 8
    if bits.size < need:
        raise KernelError("InvalidArguments", f"mask has {bits.size} bytes, need {need}")
    return np.ascontiguousarray(bits)


def _check(rc: int):
    if rc == -2:
        raise KernelError("LengthMismatch", "apply numeric: length mismatch")
    if rc == -10:
        raise KernelError("DivideByZero", "dense integer kernel: division by zero (reference panics)")
    if rc != 0:
        raise KernelError("Unknown", str(rc))


def apply_int(lhs, rhs, op: int, mask=None):
    """apply_int_{i32,u32,i64,u64,...} — src/kernels/arithmetic/dispatch.rs:65-133."""
    lhs = np.ascontiguousarray(lhs)
    rhs = np.ascontiguousarray(rhs, dtype=lhs.dtype)
    name = _NAMES[lhs.dtype]
    n = lhs.size
    m = _mask_bytes(mask, n)
    out = np.empty(n, dtype=lhs.dtype)
    om = np.zeros((n + 7) // 8, dtype=np.uint8) if m is not None else None
    rc = getattr(lib(), f"orc_apply_int_{name}")(_p(lhs), _sz(n), _p(rhs), _sz(rhs.size), C.c_int(op),
                                                  _p(m), _p(out), _p(om))
    _check(rc)
    return out, (Bits(om, n) if om is not None else None)


def apply_datetime(l_data, l_mask, l_off: int, l_len: int, r_data, r_mask, r_off: int, r_len: int, op: int):
    """apply_datetime_{i32,u32,i64,u64}(lhs: DatetimeAVT, rhs: DatetimeAVT, op) — src/kernels/arithmetic/dispatch.rs:309-372:
    the integer kernels over the two data windows; the output validity is merge_bitmasks_to_new(lhs.null_mask, rhs.null_mask,
    llen), i.e. bits [0, llen) of each ARRAY's mask (the reference does not offset the masks by the window start, :321-322)."""
    if l_len != r_len:
        raise KernelError("LengthMismatch", f"apply_datetime: length mismatch (lhs: {l_len}, rhs: {r_len})")
    l = np.ascontiguousarray(l_data)[l_off:l_off + l_len]
    r = np.ascontiguousarray(r_data, dtype=l.dtype)[r_off:r_off + r_len]
    return apply_int(l, r, op, merge_bitmasks_to_new(l_mask, r_mask, l_len))


def apply_float(lhs, rhs, op: int, mask=None):
    """apply_float_{f32,f64} — src/kernels/arithmetic/dispatch.rs:138-206."""
    lhs = np.ascontiguousarray(lhs)
    rhs = np.ascontiguousarray(rhs, dtype=lhs.dtype)
    name = _NAMES[lhs.dtype]
    n = lhs.size
    m = _mask_bytes(mask, n)
    out = np.empty(n, dtype=lhs.dtype)
    om = np.zeros((n + 7) // 8, dtype=np.uint8) if m is not None else None
    rc = getattr(lib(), f"orc_apply_float_{name}")(_p(lhs), _sz(n), _p(rhs), _sz(rhs.size), C.c_int(op),
                                                    _p(m), _p(out), _p(om))
    _check(rc)
    return out, (Bits(om, n) if om is not None else None)


def apply(lhs, rhs, op: int, mask=None):
    lhs = np.asarray(lhs)
    return apply_float(lhs, rhs, op, mask) if lhs.dtype.kind == "f" else apply_int(lhs, rhs, op, mask)


def apply_fma(lhs, rhs, acc, mask=None, fused: bool = True):
    """apply_fma_{f32,f64} — src/kernels/arithmetic/dispatch.rs:211-290."""
    lhs = np.ascontiguousarray(lhs)
    rhs = np.ascontiguousarray(rhs, dtype=lhs.dtype)
    acc = np.ascontiguousarray(acc, dtype=lhs.dtype)
    name = _NAMES[lhs.dtype]
    n = lhs.size
    m = _mask_bytes(mask, n)
    out = np.empty(n, dtype=lhs.dtype)
    om = np.zeros((n + 7) // 8, dtype=np.uint8) if m is not None else None
    rc = getattr(lib(), f"orc_apply_fma_{name}")(_p(lhs), _sz(n), _p(rhs), _sz(rhs.size), _p(acc),
                                                  _sz(acc.size), _p(m), C.c_int(int(fused)), _p(out), _p(om))
    _check(rc)
    return out, (Bits(om, n) if om is not None else None)


# ---- scalar broadcast / promotion (routing layer) ---------------------------------------------

def broadcast_length_1(value, length: int, dtype) -> np.ndarray:
    """broadcast_length_1_array — src/kernels/routing/broadcast.rs:25-47 (materialises `len` copies)."""
    return np.full(length, value, dtype=dtype)


def resolve_binary_arithmetic(op: int, lhs, rhs, null_mask=None):
    """resolve_binary_arithmetic — src/kernels/routing/arithmetic.rs:214-406.

    Length-1 broadcast (routing/broadcast.rs:87-112), same-dtype dispatch, and the two mixed pairs
    (i32,f64)->f64 and (i32,f32)->f32 via `as` casts (:244-269,342-373).  The optional mask is the one
    pre-merged mask of the leaf API, indexed from bit 0."""
    lhs = np.asarray(lhs)
    rhs = np.asarray(rhs)
    if lhs.size != rhs.size:
        if lhs.size == 1:
            lhs = broadcast_length_1(lhs.reshape(-1)[0], rhs.size, lhs.dtype)
        elif rhs.size == 1:
            rhs = broadcast_length_1(rhs.reshape(-1)[0], lhs.size, rhs.dtype)
        else:
            raise KernelError("LengthMismatch", f"cannot broadcast arrays of length {lhs.size} and {rhs.size}")
    lt, rt = lhs.dtype, rhs.dtype
    if lt == rt and lt in (np.int32, np.int64, np.uint32, np.uint64, np.float32, np.float64):
        return apply(lhs, rhs, op, null_mask)
    pair = {lt, rt}
    if pair == {np.dtype(np.int32), np.dtype(np.float64)}:
        return apply_float(lhs.astype(np.float64), rhs.astype(np.float64), op, null_mask)
    if pair == {np.dtype(np.int32), np.dtype(np.float32)}:
        return apply_float(lhs.astype(np.float32), rhs.astype(np.float32), op, null_mask)
    raise KernelError("UnsupportedType", "Unsupported array type combination for arithmetic operations")


# ---- container routes: SuperArray fan-out and the Array <-> SuperArray re-chunk --------------------
# Chunks are (data: np.ndarray, mask: Optional[Bits]) pairs.

def route_super_array_broadcast(op: int, lhs_chunks, rhs_chunks, null_mask_override: Optional[Bits] = None):
    """route_super_array_broadcast — src/kernels/broadcast/super_array.rs:180-249: chunk i of lhs against chunk i of rhs
    through resolve_binary_arithmetic; the chunk's mask is the override if given, else the UNION of the two chunks'
    masks, else the one present mask, else None (:214-230).  Chunk lengths must match pairwise (:203-213)."""
    out = []
    for i, ((ld, lm), (rd, rm)) in enumerate(zip(lhs_chunks, rhs_chunks)):
        if len(ld) != len(rd):
            raise KernelError("ShapeError", f"Super Array broadcasting error - Chunk: LHS {len(ld)} RHS {len(rd)}")
        if null_mask_override is not None:
            mask = null_mask_override
        elif lm is not None and rm is not None:
            mask = union(lm, rm)
        else:
            mask = lm if lm is not None else rm
        out.append(resolve_binary_arithmetic(op, ld, rd, mask))
    return out


def union_array_superarray_masks(arr_mask: Optional[Bits], sa_chunks) -> Optional[Bits]:
    """union_array_superarray_masks — src/utils.rs:367-413: the SuperArray's chunk masks concatenated bit by bit (a chunk
    without a mask counts as all valid, but only if SOME chunk has one), OR-ed with the array's mask; lengths must match."""
    if any(m is not None for _, m in sa_chunks):
        bools = np.concatenate([np.unpackbits(m.bits, bitorder="little")[:m.len].astype(bool) if m is not None
                                else np.ones(len(d), dtype=bool) for d, m in sa_chunks])
        sa_mask = Bits.from_bools(bools)
    else:
        sa_mask = None
    if arr_mask is not None and sa_mask is not None:
        if arr_mask.len != sa_mask.len:
            raise KernelError("ShapeError", f"Mask lengths must match for union: {arr_mask.len} vs {sa_mask.len}")
        return union(arr_mask, sa_mask)
    return arr_mask if arr_mask is not None else sa_mask


def create_aligned_chunks_from_array(arr, arr_mask: Optional[Bits], sa_chunks):
    """create_aligned_chunks_from_array — src/utils.rs:417-481: split `arr` to the SuperArray's chunk lengths
    (slice_clone of the values) and give every produced chunk its window of the FULL union mask (bit by bit, :463-469)."""
    arr = np.asarray(arr)
    total = sum(len(d) for d, _ in sa_chunks)
    if arr.size != total:
        raise KernelError("ShapeError", f"Array and SuperArray must have same total length for broadcasting: {arr.size} vs {total}")
    full = union_array_superarray_masks(arr_mask, sa_chunks)
    fb = None if full is None else np.unpackbits(full.bits, bitorder="little")[:full.len].astype(bool)
    out, start = [], 0
    for d, _ in sa_chunks:
        n = len(d)
        out.append((arr[start:start + n].copy(), None if fb is None else Bits.from_bools(fb[start:start + n])))
        start += n
    return out


def broadcast_array_superarray(op: int, arr, arr_mask, sa_chunks, array_is_lhs: bool):
    """`Value::Array (op) Value::SuperArray` and the mirrored arm — src/kernels/broadcast/mod.rs:1351-1361."""
    aligned = create_aligned_chunks_from_array(arr, arr_mask, sa_chunks)
    return route_super_array_broadcast(op, aligned, sa_chunks) if array_is_lhs else route_super_array_broadcast(op, sa_chunks, aligned)


# ---- Table / view arms: columns are plain np.ndarrays (these arms hand NO mask to the kernels) ----------------------------
# A table is a list of columns; a table view is (columns, offset, len); a super table view is a list of table views.
def _tv(tv):
    cols, off, n = tv
    return [np.asarray(c)[off:off + n] for c in cols]


def broadcast_tableview_to_tableview(op: int, lhs_tv, rhs_tv):
    """table_view.rs:25-60 (Table form table.rs:31-62): column i against column i, no mask."""
    l, r = _tv(lhs_tv), _tv(rhs_tv)
    if len(l) != len(r):
        raise KernelError("ShapeError", f"TableView column count mismatch: {len(l)} vs {len(r)}")
    return [resolve_binary_arithmetic(op, a, b, None)[0] for a, b in zip(l, r)]


def broadcast_tableview_to_arrayview(op: int, tv, arr, table_is_lhs: bool = True):
    """table_view.rs:108-146 / array_view.rs: every column against the same array window, operand order kept."""
    arr = np.asarray(arr)
    return [resolve_binary_arithmetic(op, c, arr, None)[0] if table_is_lhs else resolve_binary_arithmetic(op, arr, c, None)[0] for c in _tv(tv)]


def broadcast_supertableview_to_arrayview(op: int, stv, arr, stv_is_lhs: bool = True, check_len: bool = True):
    """super_table_view.rs:66-105 (and the mirror in array_view.rs; Array forms :157-180, array.rs:451-476 with
    check_len=False): slice i meets arr[sum of earlier slice lengths ..][.. its own length]."""
    arr = np.asarray(arr)
    total = sum(n for _, _, n in stv)
    if check_len and arr.size != total:
        raise KernelError("ShapeError", f"ArrayView length ({arr.size}) does not match SuperTableView length ({total})")
    out, start = [], 0
    for tv in stv:
        out.append(broadcast_tableview_to_arrayview(op, tv, arr[start:start + tv[2]], stv_is_lhs))
        start += tv[2]
    return out


def broadcast_tableview_to_superarrayview(op: int, tv, slices, table_is_lhs: bool = True):
    """table_view.rs:148-200 / super_array_view.rs:22-80: the table view cut into windows aligned with the array slices."""
    cols, off, n = tv
    total = sum(np.asarray(s).size for s in slices)
    if n != total:
        raise KernelError("ShapeError", (f"TableView length ({n}) does not match SuperArrayView length ({total})" if table_is_lhs else
                                         f"SuperArrayView length ({total}) does not match TableView length ({n})"))
    out, start = [], 0
    for s in slices:
        s = np.asarray(s)
        out.append(broadcast_tableview_to_arrayview(op, (cols, off + start, s.size), s, table_is_lhs))
        start += s.size
    return out


def broadcast_table_to_superarray(op: int, table_cols, chunks, table_is_lhs: bool = True):
    """table.rs:382-406 / super_array.rs:153-176: every chunk against the whole table; the result must be one column."""
    if len(table_cols) != 1:
        raise KernelError("ShapeError", ("Table-SuperArray" if table_is_lhs else "SuperArray-Table") + " broadcasting should result in single column")
    return [broadcast_tableview_to_arrayview(op, (table_cols, 0, np.asarray(table_cols[0]).size), ch, table_is_lhs)[0] for ch in chunks]


# ---- bitmask kernels ----------------------------------------------------------------------------

def _win(m):
    """BitmaskVT = (&Bitmask, offset, len) — src/aliases.rs."""
    mask, off, ln = m
    return np.ascontiguousarray(mask.bits), int(mask.len), int(off), int(ln)


def new_set_all(length: int, value: bool) -> Bits:
    out = np.zeros((length + 7) // 8, dtype=np.uint8)
    lib().orc_bits_new_set_all(_p(out), _sz(length), C.c_int(int(value)))
    return Bits(out, length)


def bitmask_binop(lhs, rhs, op: int) -> Bits:
    """bitmask_binop — src/kernels/bitmask/dispatch.rs:47-56 -> simd.rs:95-139."""
    lb, _, lo, ln = _win(lhs)
    rb, _, ro, _ = _win(rhs)
    out = np.zeros((ln + 7) // 8, dtype=np.uint8)
    lib().orc_bitmask_binop(_p(lb), _sz(lo), _p(rb), _sz(ro), _sz(ln), C.c_int(op), _p(out))
    return Bits(out, ln)


def and_masks(lhs, rhs) -> Bits:
    return bitmask_binop(lhs, rhs, AND)


def or_masks(lhs, rhs) -> Bits:
    return bitmask_binop(lhs, rhs, OR)


def xor_masks(lhs, rhs) -> Bits:
    return bitmask_binop(lhs, rhs, XOR)


def not_mask(src) -> Bits:
    """not_mask — bitmask/dispatch.rs:135-144 -> simd.rs:169-203."""
    b, _, off, ln = _win(src)
    out = np.zeros((ln + 7) // 8, dtype=np.uint8)
    lib().orc_bitmask_not(_p(b), _sz(off), _sz(ln), _p(out))
    return Bits(out, ln)


def popcount_mask(m) -> int:
    """popcount_mask — bitmask/dispatch.rs:258-267 -> simd.rs:596-645."""
    b, mlen, off, ln = _win(m)
    return int(lib().orc_popcount_mask(_p(b), _sz(mlen), _sz(off), _sz(ln)))


def count_ones(mask: Bits) -> int:
    """Bitmask::count_ones — src/structs/bitmask.rs:393-406."""
    return int(lib().orc_bits_count_ones(_p(np.ascontiguousarray(mask.bits)), _sz(mask.len)))


def null_count(mask: Bits) -> int:
    return mask.len - count_ones(mask)


def all_true_mask(mask: Bits) -> bool:
    return bool(lib().orc_all_true_mask(_p(np.ascontiguousarray(mask.bits)), _sz(mask.len)))


def all_false_mask(mask: Bits) -> bool:
    return bool(lib().orc_all_false_mask(_p(np.ascontiguousarray(mask.bits)), _sz(mask.len)))


def _two(a, b):
    ab, al, ao, ln = _win(a)
    bb, bl, bo, _ = _win(b)
    return ab, al, ao, bb, bl, bo, ln


def eq_mask(a, b) -> Bits:
    ab, al, ao, bb, bl, bo, ln = _two(a, b)
    out = np.zeros((ln + 7) // 8, dtype=np.uint8)
    if lib().orc_eq_mask(_p(ab), _sz(al), _sz(ao), _p(bb), _sz(bl), _sz(bo), _sz(ln), _p(out)):
        raise KernelError("Panic", "eq_bits_mask: offsets must be 64-bit aligned")
    return Bits(out, ln)


def ne_mask(a, b) -> Bits:
    ab, al, ao, bb, bl, bo, ln = _two(a, b)
    out = np.zeros((ln + 7) // 8, dtype=np.uint8)
    if lib().orc_ne_mask(_p(ab), _sz(al), _sz(ao), _p(bb), _sz(bl), _sz(bo), _sz(ln), _p(out)):
        raise KernelError("Panic", "eq_bits_mask: offsets must be 64-bit aligned")
    return Bits(out, ln)


def all_eq(a, b) -> bool:
    ab, al, ao, bb, bl, bo, ln = _two(a, b)
    return bool(lib().orc_all_eq_mask(_p(ab), _sz(al), _sz(ao), _p(bb), _sz(bl), _sz(bo), _sz(ln)))


def all_ne(a, b) -> bool:
    ab, al, ao, bb, bl, bo, ln = _two(a, b)
    return bool(lib().orc_all_ne_mask(_p(ab), _sz(al), _sz(ao), _p(bb), _sz(bl), _sz(bo), _sz(ln)))


def in_mask(lhs, rhs) -> Bits:
    lb, _, lo, ln = _win(lhs)
    rb, rl, ro, _ = _win(rhs)
    out = np.zeros((ln + 7) // 8, dtype=np.uint8)
    lib().orc_in_mask(_p(lb), _sz(lo), _p(rb), _sz(rl), _sz(ro), _sz(ln), _p(out))
    return Bits(out, ln)


def not_in_mask(lhs, rhs) -> Bits:
    lb, _, lo, ln = _win(lhs)
    rb, rl, ro, _ = _win(rhs)
    out = np.zeros((ln + 7) // 8, dtype=np.uint8)
    lib().orc_not_in_mask(_p(lb), _sz(lo), _p(rb), _sz(rl), _sz(ro), _sz(ln), _p(out))
    return Bits(out, ln)


def merge_bitmasks_to_new(lhs: Optional[Bits], rhs: Optional[Bits], length: int) -> Optional[Bits]:
    """merge_bitmasks_to_new (per-row AND) — src/kernels/bitmask/mod.rs:171-197."""
    if lhs is None and rhs is None:
        return None
    out = np.zeros((length + 7) // 8, dtype=np.uint8)
    lib().orc_merge_bitmasks_to_new(_p(None if lhs is None else np.ascontiguousarray(lhs.bits)),
                                    _p(None if rhs is None else np.ascontiguousarray(rhs.bits)),
                                    _sz(length), _p(out))
    return Bits(out, length)


def union(a: Bits, b: Bits) -> Bits:
    """Bitmask::union (bitwise OR) — src/structs/bitmask.rs:661-669."""
    assert a.len == b.len, "Bitmask::union length mismatch"
    out = np.zeros((a.len + 7) // 8, dtype=np.uint8)
    lib().orc_bits_union(_p(np.ascontiguousarray(a.bits)), _p(np.ascontiguousarray(b.bits)), _sz(a.len), _p(out))
    return Bits(out, a.len)


def intersect(a: Bits, b: Bits) -> Bits:
    assert a.len == b.len
    out = np.zeros((a.len + 7) // 8, dtype=np.uint8)
    lib().orc_bits_intersect(_p(np.ascontiguousarray(a.bits)), _p(np.ascontiguousarray(b.bits)), _sz(a.len), _p(out))
    return Bits(out, a.len)


def invert(a: Bits) -> Bits:
    out = np.zeros((a.len + 7) // 8, dtype=np.uint8)
    lib().orc_bits_invert(_p(np.ascontiguousarray(a.bits)), _sz(a.len), _p(out))
    return Bits(out, a.len)


def union_opt(a: Optional[Bits], b: Optional[Bits]) -> Optional[Bits]:
    """Bitmask::union_opt — src/structs/bitmask.rs:651-657."""
    if a is None and b is None:
        return None
    if a is None:
        return b
    if b is None:
        return a
    return union(a, b)


# ---- sums (bench-defined) and null-aware aggregates -----------------------------------------------

def simd_sum_i64(data, lanes: int = 4) -> int:
    d = np.ascontiguousarray(data, dtype=np.int64)
    return int(lib().orc_simd_sum_i64(_p(d), _sz(d.size), C.c_int(lanes)))


def simd_sum_f64(data, lanes: int = 4) -> float:
    d = np.ascontiguousarray(data, dtype=np.float64)
    return float(lib().orc_simd_sum_f64(_p(d), _sz(d.size), C.c_int(lanes)))


def hotloop_sum_i64(data, lanes: int = 4) -> int:
    d = np.ascontiguousarray(data, dtype=np.int64)
    return int(lib().orc_hotloop_sum_i64(_p(d), _sz(d.size), C.c_int(lanes)))


def hotloop_sum_f64(data, lanes: int = 4) -> float:
    d = np.ascontiguousarray(data, dtype=np.float64)
    return float(lib().orc_hotloop_sum_f64(_p(d), _sz(d.size), C.c_int(lanes)))


def rayon_simd_sum_i64(data, lanes: int = 4, threads: int = 1) -> int:
    d = np.ascontiguousarray(data, dtype=np.int64)
    return int(lib().orc_rayon_simd_sum_i64(_p(d), _sz(d.size), C.c_int(lanes), C.c_int(threads)))


def rayon_simd_sum_f64(data, lanes: int = 4, threads: int = 1) -> float:
    d = np.ascontiguousarray(data, dtype=np.float64)
    partials = np.zeros((d.size + (1 << 20) - 1) >> 20, dtype=np.float64)
    return float(lib().orc_rayon_simd_sum_f64(_p(d), _sz(d.size), C.c_int(lanes), C.c_int(threads), _p(partials)))


class _AggI(C.Structure):
    _fields_ = [("sum", C.c_int64), ("min", C.c_int64), ("max", C.c_int64), ("count", C.c_uint64)]


class _AggU(C.Structure):
    _fields_ = [("sum", C.c_uint64), ("min", C.c_uint64), ("max", C.c_uint64), ("count", C.c_uint64)]


class _AggF(C.Structure):
    _fields_ = [("sum", C.c_double), ("min", C.c_double), ("max", C.c_double), ("count", C.c_uint64)]


def stats(data, validity: Optional[Bits] = None) -> dict:
    """Null-aware {sum, min, max, count, mean}; definition in DESIGN.md (reference-unpinned)."""
    d = np.ascontiguousarray(data)
    name = _NAMES[d.dtype]
    agg = {"i": _AggI, "u": _AggU, "f": _AggF}[d.dtype.kind]()
    v = None if validity is None else _mask_bytes(validity, d.size)
    getattr(lib(), f"orc_stats_{name}")(_p(d), _sz(d.size), _p(v), C.byref(agg))
    cnt = int(agg.count)
    out = {"sum": agg.sum, "min": agg.min, "max": agg.max, "count": cnt}
    out["mean"] = (float(agg.sum) / cnt) if cnt else float("nan")
    return out


def par_masked_sum_i64(data, validity: Optional[Bits], threads: int = 1):
    d = np.ascontiguousarray(data, dtype=np.int64)
    v = None if validity is None else _mask_bytes(validity, d.size)
    s, c = C.c_int64(0), C.c_uint64(0)
    lib().orc_par_masked_sum_i64(_p(d), _sz(d.size), _p(v), C.c_int(threads), C.byref(s), C.byref(c))
    return int(s.value), int(c.value)


def par_apply_float_f64(lhs, rhs, op: int, mask: Optional[Bits], threads: int, out=None, out_mask=None):
    lhs = np.ascontiguousarray(lhs, dtype=np.float64)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    n = lhs.size
    m = None if mask is None else _mask_bytes(mask, n)
    out = np.empty(n, dtype=np.float64) if out is None else out
    om = (np.zeros((n + 7) // 8, dtype=np.uint8) if out_mask is None else out_mask) if m is not None else None
    lib().orc_par_apply_float_f64(_p(lhs), _p(rhs), _sz(n), C.c_int(op), _p(m), C.c_int(threads), _p(out), _p(om))
    return out, (Bits(om, n) if om is not None else None)


def simd_eq_mask(data, field_mask, target) -> Bits:
    """simd_eq_mask_u{8,16,32,64}(data, field_mask, target) -> Bitmask (bitmask/simd.rs:741-788)."""
    d = np.ascontiguousarray(data)
    name = {1: "u8", 2: "u16", 4: "u32", 8: "u64"}[d.dtype.itemsize]
    ut = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[d.dtype.itemsize]
    ct = {1: C.c_uint8, 2: C.c_uint16, 4: C.c_uint32, 8: C.c_uint64}[d.dtype.itemsize]
    out = np.zeros((d.size + 7) // 8, dtype=np.uint8)
    fm = int(np.array([field_mask], dtype=d.dtype).view(ut)[0])
    tg = int(np.array([target], dtype=d.dtype).view(ut)[0])
    getattr(lib(), f"orc_simd_eq_mask_{name}")(_p(d), _sz(d.size), ct(fm), ct(tg), _p(out))
    return Bits(out, int(d.size))


def max_threads() -> int:
    return int(lib().orc_max_threads())
