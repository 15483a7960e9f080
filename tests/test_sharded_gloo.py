"""Host-side logic of the multi-GPU path on CPU: shard maps, and the exchange + rank-order combine of per-rank
`mnr_agg` partials over torch.distributed (gloo, world_size 2).  The partials themselves come from the oracle here
(tests may use it); on the B200 box they come from reduce_stats_kernel (tests/test_gpu_multigpu.py, bench.py)."""
import os
import socket

import numpy as np
import pytest

import minarrow_b200.sharded as sh
from oracle import oracle as orc


def test_shard_chunks_contiguous_and_complete():
    for n in (1, 2, 3, 7, 8, 64, 65):
        for g in (1, 2, 4, 8):
            parts = sh.shard_chunks(n, g)
            flat = [i for r in parts for i in r]
            assert flat == list(range(n)), (n, g)
            assert all(sh.chunk_owner(i, n, g) == r for r, p in enumerate(parts) for i in p)
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1 or n < g


def test_shard_rows_word_aligned():
    for n in (0, 1, 63, 64, 65, 1000, 1_000_000_007):
        for g in (1, 2, 4, 8):
            w = sh.shard_rows(n, g)
            assert sum(l for _, l in w) == n
            pos = 0
            for off, ln in w:
                assert off == pos and (off % 64 == 0 or ln == 0)
                pos += ln


def test_shard_rows_weighted_is_contiguous_aligned_and_proportional():
    for n in (0, 1, 63, 1000, 1_000_000_007):
        for w in ([1.0], [1, 1], [23.2, 23.2, 23.2, 23.2, 35.2, 35.2, 35.2, 35.2], [1e-3, 5, 2]):
            parts = sh.shard_rows_weighted(n, w)
            assert sum(l for _, l in parts) == n and len(parts) == len(w)
            pos = 0
            for off, ln in parts:
                assert off == pos and (off % 64 == 0 or ln == 0)
                pos += ln
            if n > 10_000:
                for (off, ln), wi in zip(parts, w):
                    assert abs(ln / n - wi / sum(w)) < 1e-3
    assert sh.shard_rows_weighted(1 << 20, [1, 1, 1, 1]) == sh.shard_rows(1 << 20, 4)
    with pytest.raises(sh.KernelError):
        sh.shard_rows_weighted(10, [1, 0])


def test_combine_partials_matches_whole_column():
    rng = np.random.default_rng(5)
    for dt in (np.int64, np.uint32, np.int32, np.float64, np.float32):
        n = 100_003
        d = (rng.standard_normal(n) * 100).astype(dt) if np.dtype(dt).kind == "f" else \
            rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, n, dtype=dt, endpoint=True)
        valid = rng.random(n) < 0.9
        parts = []
        for off, ln in sh.shard_rows(n, 4):
            a = orc.stats(d[off:off + ln], orc.Bits.from_bools(valid[off:off + ln]))
            parts.append(_words(dt, a))
        got = sh.combine_partials(dt, parts)
        exp = orc.stats(d, orc.Bits.from_bools(valid))
        assert got["count"] == exp["count"] and got["min"] == exp["min"] and got["max"] == exp["max"]
        if np.dtype(dt).kind == "f":
            assert abs(got["sum"] - exp["sum"]) <= 1e-12 * np.abs(d[valid].astype(np.float64)).sum()
        else:
            assert got["sum"] == exp["sum"]


def _words(dt, a):
    k = np.dtype(dt).kind
    t = {"i": np.int64, "u": np.uint64, "f": np.float64}[k]
    w = np.zeros(4, dtype=np.int64)
    w[0:3] = np.array([a["sum"], a["min"], a["max"]], dtype=t).view(np.int64)
    w[3] = a["count"]
    return w


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)          # same column on every rank; each takes its own window
        n = 50_001
        d = rng.integers(-2 ** 62, 2 ** 62, n, dtype=np.int64)
        valid = rng.random(n) < 0.9
        off, ln = sh.shard_rows(n, world)[rank]
        a = orc.stats(d[off:off + ln], orc.Bits.from_bools(valid[off:off + ln]))
        local = torch.from_numpy(_words(np.int64, a))
        allp = sh.exchange_partials(local).numpy()
        got = sh.combine_partials(np.int64, allp)
        exp = orc.stats(d, orc.Bits.from_bools(valid))
        ok = (got["sum"], got["min"], got["max"], got["count"]) == (exp["sum"], exp["min"], exp["max"], exp["count"])
        # float partials: rank-order add must give the same bits on every rank
        x = rng.standard_normal(n)
        fa = orc.stats(x[off:off + ln], None)
        fall = sh.exchange_partials(torch.from_numpy(_words(np.float64, fa))).numpy()
        fs = sh.combine_partials(np.float64, fall)["sum"]
        sums = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(sums, torch.tensor([fs], dtype=torch.float64))
        same = all(float(s) == fs for s in sums)
        q.put((rank, bool(ok), bool(same), allp.shape))
    finally:
        dist.destroy_process_group()


def test_exchange_and_combine_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok, same, shape in res:
        assert ok and same and shape == (2, 4), (rank, ok, same, shape)


def test_stream_shard_ranges_match_the_chunk_map():
    """arrow.shard_range (which chunks of an Arrow C stream a rank uploads) is the inverse of sharded.shard_chunks."""
    from minarrow_b200.arrow import shard_range
    from minarrow_b200.sharded import shard_chunks
    for n in range(0, 40):
        for w in range(1, 9):
            rs = shard_chunks(n, w)
            for r in range(w):
                lo, hi = shard_range(n, r, w)
                assert (len(rs[r]) == 0 and lo == hi) or (rs[r].start, rs[r].stop) == (lo, hi), (n, w, r)


def test_rebalance_plan_moves_every_row_once_and_in_order():
    """sharded.rebalance_plan (multi-GPU rechunk): simulate the exchange on host arrays — sends cover each shard exactly
    once, every rank ends with its even 64-row-aligned window of the global column, validity bits included."""
    import numpy as np
    from minarrow_b200.sharded import rebalance_plan, shard_rows
    rng = np.random.default_rng(9)
    for rows in ([10, 1000, 3, 0], [0, 0, 5], [129], [64, 64], [1, 1, 1, 1, 1, 1, 1, 1000], [7000, 1, 0, 0, 0, 3, 9, 11]):
        world = len(rows)
        whole = rng.integers(0, 1 << 60, sum(rows))
        valid = rng.random(sum(rows)) < 0.7
        starts = np.concatenate([[0], np.cumsum(rows)])
        shards = [whole[starts[r]:starts[r + 1]] for r in range(world)]
        vshards = [valid[starts[r]:starts[r + 1]] for r in range(world)]
        sends, targets = rebalance_plan(rows)
        assert targets == shard_rows(sum(rows), world)
        inbox = [[] for _ in range(world)]
        for src in range(world):
            covered = 0
            for d, lo, hi in sends[src]:
                assert lo == covered and hi > lo          # ascending, gap-free
                covered = hi
                # validity travels re-based to bit 0 and byte-padded, like the device path
                inbox[d].append((shards[src][lo:hi], np.packbits(vshards[src][lo:hi], bitorder="little")))
            assert covered == rows[src]
        for d, (t0, tn) in enumerate(targets):
            got = np.concatenate([p[0] for p in inbox[d]]) if inbox[d] else np.zeros(0, whole.dtype)
            gv = np.concatenate([np.unpackbits(p[1], bitorder="little")[:len(p[0])] for p in inbox[d]]).astype(bool) if inbox[d] else np.zeros(0, bool)
            assert np.array_equal(got, whole[t0:t0 + tn]) and np.array_equal(gv, valid[t0:t0 + tn])
