#!/usr/bin/env python
"""Run under torchrun (one rank per GPU): a SuperArray of host chunks sharded over the ranks, reduced with
reduce_stats_kernel per chunk + ONE NCCL all-gather of the 32-byte partials, checked against the oracle on the
whole column; element-wise ops run shard-local and are checked per chunk.  Exit code 0 = parity on every rank."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import minarrow_b200 as mnr
    from minarrow_b200 import sharded as sh
    from oracle import oracle as orc
    ctx = mnr.Context(local)
    fx = sh.FusedExchange(ctx)                           # mailboxes over CUDA IPC, once
    rng = np.random.default_rng(123)                     # every rank builds the same SuperArray, uploads only its shard
    n_chunks, rows = 13, 250_007
    table_cols, table_exp = [], []
    for dt in (np.int64, np.int32, np.uint64, np.float64, np.float32):
        chunks = []
        for _ in range(n_chunks):
            if np.dtype(dt).kind == "f":
                d = (rng.standard_normal(rows) * 100).astype(dt)
            else:
                d = rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, rows, dtype=dt, endpoint=True)
            chunks.append(mnr.core.make_array(d, mnr.Bitmask.from_bools(rng.random(rows) < 0.9)))
        col = sh.ShardedColumn.from_host_chunks(ctx, chunks, rank, world)
        whole = np.concatenate([c.data for c in chunks])
        valid = np.concatenate([c.null_mask.to_bools() for c in chunks])
        exp = orc.stats(whole, orc.Bits.from_bools(valid))
        table_cols.append(col)
        table_exp.append((dt, whole, valid, exp))
        # both finishes: the fused one (batched kernels + per-column fold + NVLink mailbox exchange, no NCCL) and the
        # NCCL one (same kernels + on-device fold, one all-gather of 32 bytes per rank); they must agree bit for bit
        got = col.stats(True, exchange=fx)
        got_nccl = col.stats(True)
        assert got == got_nccl or (got["sum"] != got["sum"]), (dt, got, got_nccl)
        assert got["count"] == exp["count"], (dt, got, exp)
        assert got["min"] == exp["min"] and got["max"] == exp["max"], (dt, got, exp)
        if np.dtype(dt).kind == "f":
            assert abs(got["sum"] - exp["sum"]) <= 1e-12 * np.abs(whole[valid].astype(np.float64)).sum(), (dt, got, exp)
        else:
            assert got["sum"] == exp["sum"], (dt, got, exp)
        # every rank must hold the same bits (rank-order combine)
        bits = int(np.float64(got["sum"]).view(np.int64)) if np.dtype(dt).kind == "f" else \
            (int(got["sum"]) + 2 ** 63) % 2 ** 64 - 2 ** 63
        t = torch.tensor([bits], dtype=torch.int64, device="cuda")
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        assert all(int(x) == int(t) for x in g), "ranks disagree on the combined sum"
        # shard-local element-wise op: chunk * chunk with fused validity, no communication
        mine = sh.shard_chunks(n_chunks, world)[rank]
        for k, i in enumerate(mine):
            ob, om = mnr.device_ops.ew_binary(ctx, mnr.ArithmeticOperator.Multiply, col.chunks[k], col.chunks[k],
                                              col.validities[k], col.validities[k], mnr.MaskMode.And)
            m = orc.Bits(chunks[i].null_mask.bits, rows)
            ed, em = orc.apply(chunks[i].data, chunks[i].data, orc.MUL, m)
            assert ob.download().tobytes() == ed.tobytes() and np.array_equal(om.download().bits, em.bits), (dt, i)
    # the whole 5-column "SuperTable" in ONE exchange epoch (configs[4] shape): 5 aggregates per rank through the mailbox
    for rep in range(3):
        allg = sh.sharded_stats(table_cols, True, exchange=fx)
    for got, (dt, whole, valid, exp) in zip(allg, table_exp):
        assert (got["count"], got["min"], got["max"]) == (exp["count"], exp["min"], exp["max"]), (dt, got, exp)
        if np.dtype(dt).kind == "f":
            assert abs(got["sum"] - exp["sum"]) <= 1e-12 * np.abs(whole[valid].astype(np.float64)).sum(), (dt, got, exp)
        else:
            assert got["sum"] == exp["sum"], (dt, got, exp)
    # fewer chunks than ranks: ranks that own nothing join the exchange with identity aggregates
    few = [mnr.core.make_array(np.arange(1000, dtype=np.int64) + 7 * k, None) for k in range(max(1, world - 1))]
    fcol = sh.ShardedColumn.from_host_chunks(ctx, few, rank, world)
    g = fcol.stats(True, exchange=fx)
    fw = np.concatenate([c.data for c in few])
    assert (g["sum"], g["min"], g["max"], g["count"]) == (int(fw.sum()), int(fw.min()), int(fw.max()), fw.size), g
    del table_cols, fcol
    # rebalance (multi-GPU rechunk): deliberately uneven shards -> even 64-row-aligned windows; rows move between GPUs
    # with one all-to-all of value bytes + one of validity bytes, stitched by the device consolidate
    for dt in (np.int64, np.int8, np.float32):
        sizes = [int(x) for x in rng.integers(0, 200_000, world)]
        sizes[0] = 300_001                                   # rank 0 is overloaded, odd row count -> odd bit offsets
        n = sum(sizes)
        whole = rng.integers(-100, 100, n).astype(dt)
        valid = rng.random(n) < 0.85
        st = int(np.sum(sizes[:rank]))
        cuts = [0, sizes[rank] // 3, sizes[rank] // 3, sizes[rank]]           # three local chunks, one empty
        chunks_l, vals_l = [], []
        for a, b in zip(cuts[:-1], cuts[1:]):
            chunks_l.append(mnr.DeviceBuffer.upload(ctx, whole[st + a:st + b]))
            vals_l.append(mnr.DeviceBitmask.upload(ctx, mnr.Bitmask.from_bools(valid[st + a:st + b])))
        col = sh.ShardedColumn(ctx, dt, chunks_l, vals_l)
        before = col.stats(True)
        bal = col.rebalance()
        t0, tn = sh.shard_rows(n, world)[rank]
        assert len(bal.chunks) == (1 if tn else 0)
        if tn:
            assert len(bal.chunks[0]) == tn
            assert bal.chunks[0].download().tobytes() == whole[t0:t0 + tn].tobytes(), (dt, rank)
            assert np.array_equal(bal.validities[0].download().bits, np.packbits(valid[t0:t0 + tn], bitorder="little")), (dt, rank)
        after = bal.stats(True)
        assert (after["count"], after["min"], after["max"]) == (before["count"], before["min"], before["max"])
        assert after["sum"] == before["sum"] or np.dtype(dt).kind == "f"
    # fused reduction + exchange kernel (P2P mailboxes): every rank reduces its window of one big column and must end
    # with the oracle's aggregate of the WHOLE column, bit-identical on all ranks, over many back-to-back epochs.
    n = 3_000_017
    for dt in (np.int64, np.int32, np.float64, np.float32):
        if np.dtype(dt).kind == "f":
            whole = (rng.standard_normal(n) * 100).astype(dt)
        else:
            whole = rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, n, dtype=dt, endpoint=True)
        valid = rng.random(n) < 0.9
        off, ln = sh.shard_rows(n, world)[rank]
        B = mnr.DeviceBuffer.upload(ctx, whole[off:off + ln])
        V = mnr.DeviceBitmask.upload(ctx, mnr.Bitmask.from_bools(valid[off:off + ln]))
        exp = orc.stats(whole, orc.Bits.from_bools(valid))
        for rep in range(25):
            for mm in (True, False):
                got = fx.reduce_stats(B, V, mm)
                assert got["count"] == exp["count"], (dt, rep, got, exp)
                if mm:
                    assert got["min"] == exp["min"] and got["max"] == exp["max"], (dt, rep, got, exp)
                if np.dtype(dt).kind == "f":
                    assert abs(got["sum"] - exp["sum"]) <= 1e-12 * np.abs(whole[valid].astype(np.float64)).sum()
                else:
                    assert got["sum"] == exp["sum"], (dt, rep, got, exp)
        bits = int(np.float64(got["sum"]).view(np.int64)) if np.dtype(dt).kind == "f" else \
            (int(got["sum"]) + 2 ** 63) % 2 ** 64 - 2 ** 63
        t = torch.tensor([bits], dtype=torch.int64, device="cuda")
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        assert all(int(x) == int(t) for x in g), "fused exchange: ranks disagree"
        # asynchronous form, 200 epochs back to back without host synchronisation — first serialised, then with
        # consecutive reductions overlapped (programmatic dependent launch + late dependency wait, as bench.py runs it)
        outd = torch.zeros(4, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        for overlap in (0, 1):
            ctx.set_option("reduce_overlap", overlap)
            ctx.synchronize()
            for k in range(200):
                fx.reduce_stats_async(B, V, k % 3 == 0, outd.data_ptr())
            ctx.synchronize()
            assert int(outd[3]) == exp["count"], (dt, overlap)
            if np.dtype(dt).kind != "f":
                assert int(outd[0]) == exp["sum"], (dt, overlap)
            assert not fx.status()
        ctx.set_option("reduce_overlap", 0)
    fx.close()
    dist.barrier()
    if rank == 0:
        print(f"multigpu_check ok: world={world}, {n_chunks} chunks x {rows} rows, 5 dtypes")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
