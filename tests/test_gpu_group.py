"""The multi-GPU path on whatever box runs the tests: `sharded.Group` (one process, one context + mailbox per rank)
with VIRTUAL ranks — several contexts on device 0 — so the fused reduction + exchange kernels, the batched per-column
fold, the programmatic-dependent-launch overlap and the timeout handling all run on a 1-GPU box.  The real 2/4/8-GPU runs
are tests/multigpu_check.py (torchrun, CUDA IPC) and bench.py --gpus N."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def _table(rng, n_chunks, rows, dtypes):
    import minarrow_b200 as mnr
    cols = []
    for dt in dtypes:
        chunks = []
        for k in range(n_chunks):
            n = rows + 5 * k
            if np.dtype(dt).kind == "f":
                d = (rng.standard_normal(n) * 100).astype(dt)
            else:
                d = rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, n, dtype=dt, endpoint=True)
            m = mnr.Bitmask.from_bools(rng.random(n) < 0.9) if k % 4 != 3 else None      # some chunks carry no validity
            chunks.append(mnr.core.make_array(d, m))
        cols.append(chunks)
    return cols


def _expect(orc, chunks):
    whole = np.concatenate([c.data for c in chunks])
    valid = np.concatenate([c.null_mask.to_bools() if c.null_mask is not None else np.ones(len(c), bool) for c in chunks])
    return whole, valid, orc.stats(whole, orc.Bits.from_bools(valid))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_group_supertable_stats_and_elementwise(gpu_ctx, oracle, world):
    """configs[4] shape in small: 4 typed columns x 9 chunks over `world` ranks — per-column sum/min/max/count in ONE
    call (batched kernels + per-column fold + mailbox exchange), table * table and a typed scalar broadcast shard-local."""
    import minarrow_b200 as mnr
    from minarrow_b200 import sharded as sh
    rng = np.random.default_rng(100 + world)
    dts = [np.int32, np.int64, np.float32, np.float64]
    nd = _ndev()
    g = sh.Group(devices=[r % nd for r in range(world)])
    lt, rt = _table(rng, 9, 20_011, dts), _table(rng, 9, 20_011, dts)
    lb, lv, rb, rv = [], [], [], []
    for c in range(4):
        b, v = g.upload(lt[c]); lb.append(b); lv.append(v)
        b, v = g.upload(rt[c]); rb.append(b); rv.append(v)
    flat_b = [x for c in lb for x in c]
    flat_v = [x for c in lv for x in c]
    cols = [c for c in range(4) for _ in range(9)]
    for rep in range(3):
        got = g.reduce_stats(flat_b, flat_v, True, cols, dts)
    for c, dt in enumerate(dts):
        whole, valid, exp = _expect(oracle, lt[c])
        assert got[c]["count"] == exp["count"] and got[c]["min"] == exp["min"] and got[c]["max"] == exp["max"], (dt, got[c], exp)
        if np.dtype(dt).kind == "f":
            assert abs(got[c]["sum"] - exp["sum"]) <= 1e-12 * np.abs(whole[valid].astype(np.float64)).sum()
        else:
            assert got[c]["sum"] == exp["sum"]
    # sum + count only (the cheaper kernel): same counts and integer sums; float sums take another vector tier = another
    # (documented) summation order, so they agree to the tolerance only
    got2 = g.reduce_stats(flat_b, flat_v, False, cols, dts)
    for c, dt in enumerate(dts):
        assert got2[c]["count"] == got[c]["count"]
        if np.dtype(dt).kind == "f":
            assert abs(got2[c]["sum"] - got[c]["sum"]) <= 1e-12 * np.abs(np.concatenate([x.data for x in lt[c]]).astype(np.float64)).sum()
        else:
            assert got2[c]["sum"] == got[c]["sum"]
    # table * table: chunk pairs on their owning ranks, OR-union validity like route_super_array_broadcast
    for c, dt in enumerate(dts):
        ob, om = g.ew_binary(mnr.ArithmeticOperator.Multiply, lb[c], rb[c], lv[c], rv[c], mnr.MaskMode.Or)
        for k in (0, 3, 8):
            l, r = lt[c][k], rt[c][k]
            lm = None if l.null_mask is None else oracle.Bits(l.null_mask.bits, len(l))
            rm = None if r.null_mask is None else oracle.Bits(r.null_mask.bits, len(r))
            merged = oracle.union_opt(lm, rm)   # the SuperArray route: union, or the one present mask
            ed, em = oracle.apply(l.data, r.data, oracle.MUL, merged)
            assert ob[k].download().tobytes() == ed.tobytes(), (dt, k)
            if merged is None:
                assert om[k] is None
            else:
                assert np.array_equal(om[k].download().bits, em.bits), (dt, k)
    ob, om = g.ew_scalar(mnr.ArithmeticOperator.Add, lb[1], [3] * 9, False, lv[1])
    l = lt[1][4]
    ed, em = oracle.apply(l.data, np.full(len(l), 3, np.int64), oracle.ADD, oracle.Bits(l.null_mask.bits, len(l)))
    assert ob[4].download().tobytes() == ed.tobytes() and np.array_equal(om[4].download().bits, em.bits)
    g.close()


def test_group_fewer_chunks_than_ranks(gpu_ctx, oracle):
    """Ranks that own no chunk still take part in the exchange (identity aggregates)."""
    import minarrow_b200 as mnr
    from minarrow_b200 import sharded as sh
    rng = np.random.default_rng(7)
    nd = _ndev()
    g = sh.Group(devices=[r % nd for r in range(5)])
    for dt in (np.int16, np.uint64, np.float64):
        chunks = _table(rng, 2, 3001, [dt])[0]
        b, v = g.upload(chunks)
        got = g.reduce_stats(b, v, True, [0, 0], [dt])[0]
        whole, valid, exp = _expect(oracle, chunks)
        assert (got["count"], got["min"], got["max"]) == (exp["count"], exp["min"], exp["max"])
        assert got["sum"] == exp["sum"] or np.dtype(dt).kind == "f"
    g.close()


def _xchg(g, r):
    return C.c_void_p(g.lib.mnr_group_xchg(g.h, r))


@pytest.mark.parametrize("world,overlap", [(4, 0), (1, 0), (1, 1), (0, 1)])
def test_single_column_exchange_back_to_back_epochs(gpu_ctx, oracle, world, overlap):
    """mnr_reduce_stats_exchange, 300 launches per rank queued without any host synchronisation: every rank must end with
    the oracle's aggregate of the whole column; mailbox parity, ticket re-arming and the launch-order dependency all have
    to hold for that.  world 4 = virtual ranks on the box's GPUs (co-located ranks launch without the programmatic
    attribute); world 1 runs the programmatic-dependent-launch path on one GPU, without and with the reduce_overlap option
    (late dependency wait); world 0 = one rank per GPU of the box with overlap (the 8-GPU bench configuration)."""
    import minarrow_b200 as mnr
    from minarrow_b200 import sharded as sh
    from minarrow_b200.core import check
    nd = _ndev()
    world = world or nd
    g = sh.Group(devices=[r % nd for r in range(world)])
    rng = np.random.default_rng(42)
    n = 2_000_003
    for dt in (np.int64, np.float64, np.int8):
        whole = (rng.standard_normal(n) * 10).astype(dt) if np.dtype(dt).kind == "f" else \
            rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, n, dtype=dt, endpoint=True)
        valid = rng.random(n) < 0.9
        exp = oracle.stats(whole, oracle.Bits.from_bools(valid))
        B, V, outs = [], [], []
        for r, (off, ln) in enumerate(sh.shard_rows(n, world)):
            ctx = g.ctx(r)
            ctx.set_option("reduce_overlap", overlap)
            B.append(mnr.DeviceBuffer.upload(ctx, whole[off:off + ln]))
            V.append(mnr.DeviceBitmask.upload(ctx, mnr.Bitmask.from_bools(valid[off:off + ln])))
            outs.append(mnr.DeviceBuffer.alloc(ctx, np.int64, 4))
        g.synchronize()
        for rep in range(300):
            for r in range(world):
                ctx = g.ctx(r)
                check(g.lib.mnr_reduce_stats_exchange(ctx.h, _xchg(g, r), B[r].h, V[r].h, rep % 2, C.c_void_p(outs[r].device_ptr)))
        g.synchronize()
        res = [o.download() for o in outs]
        for r in range(world):
            assert np.array_equal(res[r], res[0]), "ranks disagree"
            assert int(res[r][3]) == exp["count"], (dt, r)
            if np.dtype(dt).kind == "f":
                s = float(res[r][:1].view(np.float64)[0])
                assert abs(s - exp["sum"]) <= 1e-12 * np.abs(whole[valid].astype(np.float64)).sum()
            else:
                assert int(res[r][0]) == exp["sum"], (dt, r)
            t = C.c_int()
            check(g.lib.mnr_xchg_status(_xchg(g, r), 1, C.byref(t)))
            assert t.value == 0
    g.close()


def test_exchange_timeout_poisons_the_result_and_recovers(gpu_ctx):
    """A peer that never launches: after the bounded wait the kernel marks the aggregate unusable (count = 2^64 - 1), the
    sync API returns an error and clears the error word, and the exchange keeps working once the peer catches up."""
    import minarrow_b200 as mnr
    from minarrow_b200 import _lib, sharded as sh
    from minarrow_b200.core import KernelError, check
    nd = _ndev()
    g = sh.Group(devices=[r % nd for r in range(2)])
    a = np.arange(1000, dtype=np.int64)
    B = [mnr.DeviceBuffer.upload(g.ctx(r), a[r * 500:(r + 1) * 500]) for r in range(2)]
    agg = _lib.Agg()
    with pytest.raises(KernelError) as e:       # rank 1 never joins epoch 1
        check(g.lib.mnr_reduce_stats_exchange_sync(g.ctx(0).h, _xchg(g, 0), B[0].h, None, 0, C.byref(agg)))
    assert e.value.kind == "Cuda" and "timed out" in str(e.value)
    t = C.c_int()
    check(g.lib.mnr_xchg_status(_xchg(g, 0), 1, C.byref(t)))
    assert t.value == 0, "the sync call must have cleared the error word"
    # rank 1 catches up with epoch 1 (rank 0's epoch-1 partial is still in its mailbox), then both run epoch 2
    check(g.lib.mnr_reduce_stats_exchange_sync(g.ctx(1).h, _xchg(g, 1), B[1].h, None, 0, C.byref(agg)))
    assert agg.sum.i64 == 499500 and agg.count == 1000
    out0 = mnr.DeviceBuffer.alloc(g.ctx(0), np.int64, 4)
    check(g.lib.mnr_reduce_stats_exchange(g.ctx(0).h, _xchg(g, 0), B[0].h, None, 0, C.c_void_p(out0.device_ptr)))
    check(g.lib.mnr_reduce_stats_exchange_sync(g.ctx(1).h, _xchg(g, 1), B[1].h, None, 0, C.byref(agg)))
    g.synchronize()
    assert agg.sum.i64 == 499500 and agg.count == 1000
    r0 = out0.download()
    assert int(r0[0]) == 499500 and int(r0[3]) == 1000
    g.close()
