"""SuperArray / SuperTable chunks as shards over the GPUs of one box (one process per GPU, torch.distributed).

The reference's chunked containers are already independent equal-dtype units that its own route walks chunk by chunk
(`SuperArray {chunks: Vec<Array>}`, src/structs/chunked/super_array.rs:96-103; per-chunk loop and length check,
src/kernels/broadcast/super_array.rs:180-249).  Here chunk `i` of `n` lives on rank `floor(i * G / n)` (contiguous
blocks, so global row order is preserved), element-wise and bitmask work runs shard-local with no communication, and a
reduction leaves one 32-byte `mnr_agg` partial per rank that is exchanged with ONE collective — an all-gather of 32 bytes
per rank — and folded in rank order on every rank (`mnr_agg_combine`: integer sums wrap and are order-free, float sums
use the documented rank-order add, min/max use the NaN-skipping combine).  The backend is whatever the process group was
created with: NCCL over NVLink on the B200 box, gloo in the CPU tests.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .core import KernelError, check, dtype_code


def chunk_owner(i: int, n_chunks: int, world: int) -> int:
    """Rank that owns chunk `i`: contiguous block assignment, floor(i * G / n) (`mnr_shard_owner`)."""
    r = _lib.load().mnr_shard_owner(i, n_chunks, world)
    if r < 0:
        check(r)
    return r


def shard_chunks(n_chunks: int, world: int) -> List[range]:
    """Chunk index range of every rank (may be empty when n_chunks < world) (`mnr_shard_chunk_range`)."""
    lib = _lib.load()
    out = []
    lo, hi = C.c_size_t(), C.c_size_t()
    for r in range(world):
        check(lib.mnr_shard_chunk_range(n_chunks, world, r, C.byref(lo), C.byref(hi)))
        out.append(range(lo.value, hi.value) if hi.value > lo.value else range(0, 0))
    return out


def shard_rows(n_rows: int, world: int, align: int = 64) -> List[Tuple[int, int]]:
    """Split one big Array into `world` contiguous (offset, len) windows cut on `align`-row boundaries
    (64 rows = one validity word, so every shard's bitmask starts on a word boundary) (`mnr_shard_row_range`)."""
    if world < 1 or align < 1:
        raise KernelError("InvalidArguments", "world and align must be >= 1")
    lib = _lib.load()
    off, ln = C.c_size_t(), C.c_size_t()
    out = []
    for r in range(world):
        check(lib.mnr_shard_row_range(n_rows, world, r, align, C.byref(off), C.byref(ln)))
        out.append((off.value, ln.value))
    return out


def shard_rows_weighted(n_rows: int, weights: Sequence[float], align: int = 64) -> List[Tuple[int, int]]:
    """Like `shard_rows`, but rank r gets a share of the rows proportional to `weights[r]` (cut on `align`-row boundaries,
    contiguous, in rank order).  For HOST-resident columns the right weights are the ranks' measured host->device copy
    rates: GPUs behind a shared PCIe uplink copy slower when all links are busy (23 vs 35 GB/s per GPU on the 8-GPU boxes
    measured here), and an even split makes everyone wait for the slowest link."""
    world = len(weights)
    if world < 1 or align < 1 or any(not (w > 0) for w in weights):
        raise KernelError("InvalidArguments", "weights must be positive, one per rank")
    units = (n_rows + align - 1) // align
    total = float(sum(weights))
    cuts, acc = [0], 0.0
    for r in range(world - 1):
        acc += weights[r]
        cuts.append(min(units, max(cuts[-1], int(round(units * acc / total)))))
    cuts.append(units)
    out = []
    for r in range(world):
        off = min(cuts[r] * align, n_rows)
        end = min(cuts[r + 1] * align, n_rows)
        out.append((off, end - off))
    return out


def rebalance_plan(rows: Sequence[int], align: int = 64) -> Tuple[List[List[Tuple[int, int, int]]], List[Tuple[int, int]]]:
    """Who sends which rows to whom when the shards of one column (rank order = row order, `rows[r]` rows on rank r) are
    re-cut into the even contiguous windows of `shard_rows` — the multi-GPU form of SuperArray::rechunk
    (src/structs/chunked/super_array.rs:674-787), which the reference needs whenever one operand's chunking has to match
    the other's (`create_aligned_chunks_from_array`, src/utils.rs:417-481).

    Returns (sends, targets): sends[r] = [(dst, lo, hi), ...] local row windows of rank r in ascending dst (= ascending
    row) order, covering [0, rows[r]) exactly once; targets[d] = (global_start, n_rows) of rank d afterwards."""
    world = len(rows)
    if world < 1 or any(r < 0 for r in rows):
        raise KernelError("InvalidArguments", "rows must list one non-negative count per rank")
    targets = shard_rows(sum(rows), world, align)
    sends: List[List[Tuple[int, int, int]]] = []
    start = 0
    for r in range(world):
        mine = []
        for d, (t0, tn) in enumerate(targets):
            lo, hi = max(start, t0), min(start + rows[r], t0 + tn)
            if hi > lo:
                mine.append((d, lo - start, hi - start))
        sends.append(mine)
        start += rows[r]
    return sends, targets


class _CudaView:
    """`__cuda_array_interface__` over library-owned device memory so torch can alias it (no copy)."""

    def __init__(self, ptr: int, nbytes: int, keep):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        self.keep = keep


def agg_to_words(agg: "_lib.Agg") -> np.ndarray:
    """The 32-byte `mnr_agg` as 4 x int64 (the wire format of the exchange)."""
    return np.frombuffer(bytes(agg), dtype=np.int64).copy()


def words_to_agg(words) -> "_lib.Agg":
    w = np.ascontiguousarray(words, dtype=np.int64)
    return _lib.Agg.from_buffer_copy(w.tobytes())


def combine_partials(dtype, partials: Sequence) -> dict:
    """Fold `mnr_agg` partials (4 x int64 rows) in index order through the C ABI (`mnr_agg_combine`)."""
    lib = _lib.load()
    p = np.ascontiguousarray(np.asarray(partials, dtype=np.int64).reshape(-1, 4))
    if p.shape[0] == 0:
        raise KernelError("InvalidArguments", "need at least one partial")
    arr = (_lib.Agg * p.shape[0]).from_buffer_copy(p.tobytes())
    out = _lib.Agg()
    code = dtype_code(dtype)
    check(lib.mnr_agg_combine(code, arr, p.shape[0], C.byref(out)))
    f = {"i": "i64", "u": "u64", "f": "f64"}[np.dtype(dtype).kind]
    return {"sum": getattr(out.sum, f), "min": getattr(out.min, f), "max": getattr(out.max, f),
            "count": int(out.count), "mean": float(lib.mnr_agg_mean(code, C.byref(out)))}


def exchange_partials(local, group=None):
    """All-gather one 4 x int64 partial per rank -> tensor [world, 4] on the device of `local`.
    `local` is a torch int64 tensor of 4 elements (on the GPU for NCCL, on the CPU for gloo)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local.view(1, 4).clone()
    world = dist.get_world_size(group)
    out = torch.empty(world * 4, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous().view(-1), group=group)
    return out.view(world, 4)


class ShardedColumn:
    """This rank's shard of a chunked numeric column: device-resident chunks (+ optional validity) on one GPU."""

    def __init__(self, ctx, dtype, chunks: Sequence, validities: Optional[Sequence] = None):
        self.ctx, self.dtype = ctx, np.dtype(dtype)
        self.chunks = list(chunks)                                   # DeviceBuffer per local chunk
        self.validities = list(validities) if validities is not None else [None] * len(self.chunks)
        if len(self.validities) != len(self.chunks):
            raise KernelError("InvalidArguments", "one validity (or None) per chunk")

    @classmethod
    def from_host_chunks(cls, ctx, chunks: Sequence, rank: int, world: int) -> "ShardedColumn":
        """Upload the chunks this rank owns.  `chunks` = the whole SuperArray's host chunks (IntegerArray/FloatArray)."""
        from .core import DeviceBitmask, DeviceBuffer
        mine = shard_chunks(len(chunks), world)[rank]
        bufs, vals = [], []
        dtype = np.asarray(chunks[0].data).dtype if chunks else np.dtype(np.int64)
        for i in mine:
            c = chunks[i]
            bufs.append(DeviceBuffer.upload(ctx, np.ascontiguousarray(c.data)))
            vals.append(None if c.null_mask is None else DeviceBitmask.upload(ctx, c.null_mask))
        return cls(ctx, dtype, bufs, vals)

    def stats(self, with_minmax: bool = True, group=None, exchange: "FusedExchange" = None) -> dict:
        """Global {sum, min, max, count, mean} of the column (see `sharded_stats`)."""
        return sharded_stats([self], with_minmax, group, exchange)[0]

    def rebalance(self, group=None, align: int = 64) -> "ShardedColumn":
        """Re-cut the column into even contiguous shards (one chunk per rank, cut on `align`-row boundaries), moving rows
        between GPUs.  Local side: one device consolidate (`mnr_concat`), validity windows cut at their exact bit
        offsets (`mnr_bits_slice`), received pieces stitched with the bit-granular gather of `mnr_concat`.  Exchange:
        ONE all-to-all of value bytes and one of validity bytes (NCCL over NVLink on the box; the only place on this path
        where link bandwidth, not latency, matters).  Row order, values and validity bits are preserved exactly."""
        import torch
        import torch.distributed as dist
        from . import device_ops as dev
        from .core import DeviceBitmask, DeviceBuffer
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        dev_ = torch.device("cuda", self.ctx.device)
        es = self.dtype.itemsize
        # 1. one contiguous local shard
        if self.chunks:
            vals = self.validities if any(v is not None for v in self.validities) else None
            whole, wmask = dev.concat(self.ctx, self.chunks, vals)
        else:
            whole, wmask = DeviceBuffer.alloc(self.ctx, self.dtype, 0), None
        n_local = len(whole)
        # 2. every rank learns every rank's row count and whether any shard carries validity
        meta = torch.tensor([n_local, 0 if wmask is None else 1], dtype=torch.int64, device=dev_)
        allm = torch.empty(world * 2, dtype=torch.int64, device=dev_)
        if world > 1:
            dist.all_gather_into_tensor(allm, meta, group=group)
        else:
            allm.copy_(meta)
        allm = allm.cpu().numpy().reshape(world, 2)
        rows = [int(x) for x in allm[:, 0]]
        any_mask = bool(allm[:, 1].any())
        sends, targets = rebalance_plan(rows, align)
        my_sends = sends[rank]
        recv_rows = [0] * world                      # rows this rank receives from each source, in source (= row) order
        for src in range(world):
            for d, lo, hi in sends[src]:
                if d == rank:
                    recv_rows[src] += hi - lo
        send_rows = [0] * world
        for d, lo, hi in my_sends:
            send_rows[d] = hi - lo
        # 3. values: the consolidated shard already is the send buffer (windows in ascending destination order)
        self.ctx.synchronize()
        send_v = torch.as_tensor(_CudaView(whole.device_ptr, n_local * es, whole), device=dev_) if n_local else \
            torch.empty(0, dtype=torch.uint8, device=dev_)
        recv_v = torch.empty(sum(recv_rows) * es, dtype=torch.uint8, device=dev_)
        if world > 1:
            dist.all_to_all_single(recv_v, send_v, [r * es for r in recv_rows], [r * es for r in send_rows], group=group)
        else:
            recv_v.copy_(send_v)
        # 4. validity: each window re-based to bit 0 (exact bit offset), byte-padded per destination
        recv_m = None
        if any_mask:
            nb = lambda r: (r + 7) // 8   # noqa: E731
            send_m = torch.zeros(sum(nb(r) for r in send_rows), dtype=torch.uint8, device=dev_)
            pos, pieces = 0, []
            for d, lo, hi in my_sends:
                n = hi - lo
                piece = dev.bits_slice(self.ctx, wmask, lo, n) if wmask is not None else DeviceBitmask.new_set_all(self.ctx, n, True)
                pieces.append((pos, nb(n), piece))
                pos += nb(n)
            self.ctx.synchronize()
            for pos, k, piece in pieces:
                send_m[pos:pos + k].copy_(torch.as_tensor(_CudaView(piece.device_ptr, k, piece), device=dev_))
            recv_m = torch.empty(sum(nb(r) for r in recv_rows), dtype=torch.uint8, device=dev_)
            if world > 1:
                dist.all_to_all_single(recv_m, send_m, [nb(r) for r in recv_rows], [nb(r) for r in send_rows], group=group)
            else:
                recv_m.copy_(send_m)
        torch.cuda.synchronize(dev_)
        # 5. stitch: values are already contiguous in row order; validity pieces are byte-padded per source -> bit gather
        bufs, vms, voff, moff = [], [], 0, 0
        for src in range(world):
            r = recv_rows[src]
            if r == 0:
                continue
            bufs.append(DeviceBuffer.wrap(self.ctx, self.dtype, recv_v.data_ptr() + voff, r, recv_v))
            if recv_m is not None:
                vms.append(DeviceBitmask.wrap(self.ctx, recv_m.data_ptr() + moff, r, recv_m))
                moff += (r + 7) // 8
            voff += r * es
        if not bufs:
            return ShardedColumn(self.ctx, self.dtype, [], [])
        out, om = dev.concat(self.ctx, bufs, vms if recv_m is not None else None)
        self.ctx.synchronize()
        assert len(out) == targets[rank][1]
        return ShardedColumn(self.ctx, self.dtype, [out], [om])


class FusedExchange:
    """Mailboxes for the fused reduction + cross-GPU exchange kernel (`mnr_reduce_stats_exchange`): ONE kernel per
    reduction does the shard's aggregate, the P2P all-gather of the 32-byte partials over NVLink and the rank-order
    combine.  torch.distributed only carries the 64-byte CUDA IPC handles once, at construction."""

    def __init__(self, ctx, group=None):
        import torch
        import torch.distributed as dist
        self.ctx = ctx
        if dist.is_initialized():
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0
        self.h = None
        err = None
        mine = (C.c_uint8 * 64)()
        try:
            h = C.c_void_p()
            check(ctx.lib.mnr_xchg_create(ctx.h, self.world, self.rank, C.byref(h)))
            self.h = h
            if self.world > 1:
                check(ctx.lib.mnr_xchg_local_handle(self.h, mine))
        except KernelError as e:
            err = e
        if self.world > 1:
            # Every rank walks the same collectives whether or not its own setup worked, then all agree on the outcome.
            dev_ = torch.device("cuda", ctx.device)
            local = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=dev_)
            allh = torch.empty(self.world * 64, dtype=torch.uint8, device=dev_)
            dist.all_gather_into_tensor(allh, local, group=group)
            if err is None:
                try:
                    raw = bytes(allh.cpu().numpy().tobytes())
                    buf = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
                    check(ctx.lib.mnr_xchg_connect(self.h, buf))
                except KernelError as e:
                    err = e
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=dev_)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # also the barrier: all mailboxes are mapped
            if int(ok) == 0:
                self.close()
                raise err or KernelError("Cuda", "a peer rank could not set up its exchange mailbox (CUDA IPC)")
        elif err is not None:
            raise err

    def reduce_stats_async(self, buf, validity, with_minmax: bool, out_device_ptr: int) -> None:
        """Global aggregate of the sharded column -> 32 bytes at `out_device_ptr` on this rank (no sync, no NCCL)."""
        check(self.ctx.lib.mnr_reduce_stats_exchange(self.ctx.h, self.h, buf.h, None if validity is None else validity.h,
                                                     int(with_minmax), C.c_void_p(out_device_ptr)))

    def reduce_stats(self, buf, validity=None, with_minmax: bool = True) -> dict:
        agg = _lib.Agg()
        check(self.ctx.lib.mnr_reduce_stats_exchange_sync(self.ctx.h, self.h, buf.h, None if validity is None else validity.h,
                                                          int(with_minmax), C.byref(agg)))
        f = {"i": "i64", "u": "u64", "f": "f64"}[buf.dtype.kind]
        return {"sum": getattr(agg.sum, f), "min": getattr(agg.min, f), "max": getattr(agg.max, f), "count": int(agg.count),
                "mean": float(self.ctx.lib.mnr_agg_mean(dtype_code(buf.dtype), C.byref(agg)))}

    def reduce_stats_batch_async(self, bufs, validities, with_minmax: bool, col_of_chunk, col_dtypes, out_device_ptr: int,
                                 plan: "BatchExchangePlan" = None):
        """Sharded SuperArray / SuperTable reduction, asynchronous: this rank's chunks (chunk i belongs to column
        col_of_chunk[i]) -> len(col_dtypes) global aggregates at `out_device_ptr` on every rank.  One batched launch
        per (dtype, alignment, masked) class; the fold and the cross-GPU exchange ride in the last block."""
        _batch_exchange(self.ctx, self.h, bufs, validities, with_minmax, col_of_chunk, col_dtypes, out_device_ptr, None, plan)

    def reduce_stats_batch(self, bufs, validities, with_minmax: bool, col_of_chunk, col_dtypes) -> list:
        aggs = (_lib.Agg * len(col_dtypes))()
        _batch_exchange(self.ctx, self.h, bufs, validities, with_minmax, col_of_chunk, col_dtypes, None, aggs)
        return [_agg_dict(dt, a, self.ctx.lib) for dt, a in zip(col_dtypes, aggs)]

    def status(self, clear: bool = True) -> bool:
        """True if an exchange on this handle timed out since the last clear (`mnr_xchg_status`)."""
        r = C.c_int()
        check(self.ctx.lib.mnr_xchg_status(self.h, int(clear), C.byref(r)))
        return bool(r.value)

    def close(self) -> None:
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.mnr_xchg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def _agg_dict(dtype, agg, lib) -> dict:
    f = {"i": "i64", "u": "u64", "f": "f64"}[np.dtype(dtype).kind]
    return {"sum": getattr(agg.sum, f), "min": getattr(agg.min, f), "max": getattr(agg.max, f), "count": int(agg.count),
            "mean": float(lib.mnr_agg_mean(dtype_code(dtype), C.byref(agg)))}


def _handles(items):
    arr = (C.c_void_p * max(1, len(items)))()
    for i, x in enumerate(items):
        arr[i] = None if x is None else x.h
    return arr


class BatchExchangePlan:
    """Pre-marshalled argument arrays of a repeated sharded reduction (the ctypes arrays are built once)."""

    def __init__(self, bufs, validities, col_of_chunk, col_dtypes):
        self.n, self.n_cols = len(bufs), len(col_dtypes)
        self.bufs = _handles(list(bufs))
        self.vals = None if validities is None else _handles(list(validities))
        self.cols = (C.c_uint32 * max(1, self.n))(*[int(c) for c in col_of_chunk])
        self.dts = (C.c_int * self.n_cols)(*[dtype_code(d) for d in col_dtypes])
        self.keep = (bufs, validities)


def _batch_exchange(ctx, xh, bufs, validities, with_minmax, col_of_chunk, col_dtypes, out_device_ptr, out_aggs, plan=None):
    p = plan or BatchExchangePlan(bufs, validities, col_of_chunk, col_dtypes)
    if out_aggs is None:
        check(ctx.lib.mnr_reduce_stats_batch_exchange(ctx.h, xh, p.n, p.bufs, p.vals, int(with_minmax), p.n_cols, p.cols, p.dts,
                                                      C.c_void_p(out_device_ptr)))
    else:
        check(ctx.lib.mnr_reduce_stats_batch_exchange_sync(ctx.h, xh, p.n, p.bufs, p.vals, int(with_minmax), p.n_cols, p.cols,
                                                           p.dts, out_aggs))


def sharded_stats(columns: Sequence["ShardedColumn"], with_minmax: bool = True, group=None,
                  exchange: "FusedExchange" = None) -> List[dict]:
    """Global {sum, min, max, count, mean} of every column of a sharded SuperArray / SuperTable (`columns[c]` = this
    rank's chunks of column c; a rank may own none).  ONE call per rank on the device:
      * `exchange` given (FusedExchange): batched reduction + per-column fold + NVLink mailbox exchange + rank-order
        combine inside the kernels (`mnr_reduce_stats_batch_exchange`) — no NCCL call, no host round trip;
      * otherwise: the same batched reduction + on-device per-column fold, then ONE all-gather of 32 bytes per column
        and rank over the process group (NCCL on the box, gloo in the CPU tests) and the rank-order combine on the host.
    Integer sums wrap (order-free); float sums fold chunks in chunk order, then ranks in rank order."""
    import torch
    import torch.distributed as dist
    if not columns:
        return []
    ctx = columns[0].ctx
    bufs, vals, cols = [], [], []
    for c, col in enumerate(columns):
        bufs += col.chunks
        vals += col.validities
        cols += [c] * len(col.chunks)
    dts = [col.dtype for col in columns]
    if exchange is not None:
        return exchange.reduce_stats_batch(bufs, vals, with_minmax, cols, dts)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        aggs = (_lib.Agg * len(dts))()
        _batch_exchange(ctx, None, bufs, vals, with_minmax, cols, dts, None, aggs)
        return [_agg_dict(dt, a, ctx.lib) for dt, a in zip(dts, aggs)]
    dev_ = torch.device("cuda", ctx.device)
    local = torch.zeros(len(dts) * 4, dtype=torch.int64, device=dev_)
    torch.cuda.current_stream(dev_).synchronize()       # torch zero-filled it on ITS stream; the context has its own
    _batch_exchange(ctx, None, bufs, vals, with_minmax, cols, dts, local.data_ptr(), None)
    ctx.synchronize()
    allp = torch.empty(world * len(dts) * 4, dtype=torch.int64, device=dev_)
    dist.all_gather_into_tensor(allp, local, group=group)
    allp = allp.cpu().numpy().reshape(world, len(dts), 4)
    return [combine_partials(dt, allp[:, c, :]) for c, dt in enumerate(dts)]


class Group:
    """All GPUs of the box driven from ONE process (`mnr_group`): one context + mailbox per device, peer access instead
    of CUDA IPC.  `devices` may repeat a device ("virtual ranks"), which runs the whole exchange path on one GPU."""

    def __init__(self, world: Optional[int] = None, devices: Optional[Sequence[int]] = None):
        from .core import Context
        self.lib = _lib.load()
        if devices is not None:
            world = len(devices)
        elif world is None:
            world = max(1, int(self.lib.mnr_device_count()))
        dv = None if devices is None else (C.c_int * world)(*[int(d) for d in devices])
        h = C.c_void_p()
        check(self.lib.mnr_group_create(world, dv, C.byref(h)))
        self.h, self.world = h, world
        self.ctxs = [Context.borrow(self.lib.mnr_group_ctx(h, r)) for r in range(world)]

    def ctx(self, rank: int):
        return self.ctxs[rank]

    def synchronize(self) -> None:
        check(self.lib.mnr_group_synchronize(self.h))

    def upload(self, chunks: Sequence) -> Tuple[list, list]:
        """Host SuperArray chunks (IntegerArray / FloatArray / numpy) -> device chunks on their owning ranks."""
        from .core import DeviceBitmask, DeviceBuffer
        n = len(chunks)
        data = [np.ascontiguousarray(getattr(c, "data", c)) for c in chunks]
        masks = [getattr(c, "null_mask", None) for c in chunks]
        dt = data[0].dtype
        hp = (C.c_void_p * n)(*[d.ctypes.data for d in data])
        lens = (C.c_size_t * n)(*[d.size for d in data])
        mkeep = [None if m is None else np.ascontiguousarray(m.bits, dtype=np.uint8) for m in masks]
        for m, d, k in zip(masks, data, mkeep):
            if m is not None and (m.len < d.size or k.size < (d.size + 7) // 8):
                raise KernelError("InvalidArguments", f"mask has {m.len} bits / {k.size} bytes, need {d.size} bits")
        mp = (C.c_void_p * n)(*[None if k is None else k.ctypes.data for k in mkeep])
        ob, om = (C.c_void_p * n)(), (C.c_void_p * n)()
        check(self.lib.mnr_group_upload(self.h, dtype_code(dt), n, hp, lens, mp, ob, om))
        bufs = [DeviceBuffer(self.ctxs[chunk_owner(i, n, self.world)], C.c_void_p(ob[i])) for i in range(n)]
        vals = [DeviceBitmask(self.ctxs[chunk_owner(i, n, self.world)], C.c_void_p(om[i])) if om[i] else None for i in range(n)]
        return bufs, vals

    def _wrap_outs(self, like, ob, om):
        from .core import DeviceBitmask, DeviceBuffer
        return ([DeviceBuffer(x.ctx, C.c_void_p(ob[i])) for i, x in enumerate(like)],
                [DeviceBitmask(x.ctx, C.c_void_p(om[i])) if om[i] else None for i, x in enumerate(like)])

    def ew_binary(self, op: int, lhs, rhs, lhs_masks=None, rhs_masks=None, mode: int = 0):
        """Shard-local `lhs[i] op rhs[i]` on the owning devices (no communication), one batched launch per device."""
        n = len(lhs)
        ob, om = (C.c_void_p * max(1, n))(), (C.c_void_p * max(1, n))()
        check(self.lib.mnr_group_ew_binary(self.h, int(op), n, _handles(lhs), _handles(rhs),
                                           None if lhs_masks is None else _handles(lhs_masks),
                                           None if rhs_masks is None else _handles(rhs_masks), int(mode), ob, om))
        return self._wrap_outs(lhs, ob, om)

    def ew_scalar(self, op: int, arrs, scalars, scalar_is_lhs: bool = False, masks=None):
        n = len(arrs)
        keep = [np.array([s], dtype=a.dtype) for a, s in zip(arrs, scalars)]
        sp = (C.c_void_p * max(1, n))(*[k.ctypes.data for k in keep])
        ob, om = (C.c_void_p * max(1, n))(), (C.c_void_p * max(1, n))()
        check(self.lib.mnr_group_ew_scalar(self.h, int(op), n, _handles(arrs), sp, int(scalar_is_lhs),
                                           None if masks is None else _handles(masks), ob, om))
        return self._wrap_outs(arrs, ob, om)

    def reduce_stats(self, bufs, validities, with_minmax: bool, col_of_chunk, col_dtypes) -> list:
        n, n_cols = len(bufs), len(col_dtypes)
        cols = (C.c_uint32 * max(1, n))(*[int(c) for c in col_of_chunk])
        dts = (C.c_int * n_cols)(*[dtype_code(d) for d in col_dtypes])
        aggs = (_lib.Agg * n_cols)()
        check(self.lib.mnr_group_reduce_stats(self.h, n, _handles(bufs), None if validities is None else _handles(validities),
                                              int(with_minmax), n_cols, cols, dts, aggs))
        return [_agg_dict(dt, a, self.lib) for dt, a in zip(col_dtypes, aggs)]

    def close(self) -> None:
        if getattr(self, "h", None):
            for c in self.ctxs:
                c.h = None
            self.lib.mnr_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
