"""Alignment tiers (256-bit / 128-bit / element loads) and the batched fan-out, against the oracle.

The launchers pick the widest vector every operand pointer allows — the analogue of the reference's
"64-byte aligned -> SIMD body, else scalar body" (src/kernels/arithmetic/dispatch.rs:86,108-111), whose two bodies must
agree bit for bit.  ArrayV windows (src/structs/views/array_view.rs:79-94) at element offsets 0 / 16 B / 8 B / odd hit
every tier.  The batched calls must equal the one-by-one calls exactly (same bits, floats included)."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mnr(gpu_ctx):
    import minarrow_b200 as m
    return m


def _bits_equal(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()


@pytest.mark.parametrize("dt", [np.int8, np.uint16, np.int32, np.uint32, np.int64, np.float32, np.float64])
def test_every_alignment_tier_matches_oracle(mnr, gpu_ctx, dt):
    dev = mnr.device_ops
    rng = np.random.default_rng(31)
    es = np.dtype(dt).itemsize
    n = 50_000
    is_f = np.dtype(dt).kind == "f"
    if is_f:
        a = (rng.standard_normal(n) * 50).astype(dt)
        b = (rng.standard_normal(n) * 50).astype(dt)
        b[rng.integers(0, n, 50)] = 0
    else:
        info = np.iinfo(dt)
        a = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)
        b = rng.integers(max(info.min, -5), min(info.max, 5), n, dtype=dt, endpoint=True)
    A, B = mnr.DeviceBuffer.upload(gpu_ctx, a), mnr.DeviceBuffer.upload(gpu_ctx, b)
    oapply = orc.apply_float if is_f else orc.apply_int
    # element offsets whose byte offsets are 0 mod 32, 16 mod 32, and not a multiple of 16 (when the type allows)
    offs = sorted({0, 32 // es, 16 // es, max(1, 8 // es), 1, 3})
    for oa in offs:
        for ob_ in (0, 16 // es, 1):
            ln = n - 64 - 7
            va = rng.random(ln) < 0.9
            M = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(va))
            for op in (orc.ADD, orc.MUL, orc.DIV):
                exp, em = oapply(a[oa:oa + ln], b[ob_:ob_ + ln], op, orc.Bits.from_bools(va))
                got, gm = dev.ew_binary(gpu_ctx, op, A.slice(oa, ln), B.slice(ob_, ln), M, None, mnr.MaskMode.And)
                assert _bits_equal(got.download(), exp), (dt, oa, ob_, op)
                assert np.array_equal(gm.download().bits, em.bits), (dt, oa, ob_, op)
            st = dev.reduce_stats(gpu_ctx, A.slice(oa, ln), M)
            ex = orc.stats(a[oa:oa + ln], orc.Bits.from_bools(va))
            assert st["count"] == ex["count"] and st["min"] == ex["min"] and st["max"] == ex["max"], (dt, oa)
            if is_f:
                assert abs(st["sum"] - ex["sum"]) <= 1e-12 * np.abs(a[oa:oa + ln][va].astype(np.float64)).sum()
            else:
                assert st["sum"] == ex["sum"], (dt, oa)
    if is_f:
        c = (rng.standard_normal(n)).astype(dt)
        Cc = mnr.DeviceBuffer.upload(gpu_ctx, c)
        for oa in offs:
            ln = n - 71
            va = rng.random(ln) < 0.9
            M = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(va))
            exp, em = orc.apply_fma(a[oa:oa + ln], b[:ln], c[oa:oa + ln], orc.Bits.from_bools(va))
            got, gm = dev.ew_fma(gpu_ctx, A.slice(oa, ln), B.slice(0, ln), Cc.slice(oa, ln), M)
            assert _bits_equal(got.download(), exp) and np.array_equal(gm.download().bits, em.bits), (dt, oa)


def test_bitmask_vector_tiers(mnr, gpu_ctx):
    """Window starts at 0 / 128 / 256 / 8 bits select the 256-bit, 128-bit and byte paths of bits_op_kernel."""
    dev = mnr.device_ops
    rng = np.random.default_rng(32)
    n = 1_000_003
    x, y = rng.random(n) < 0.5, rng.random(n) < 0.5
    X = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(x))
    Y = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(y))
    ox, oy = orc.Bits.from_bools(x), orc.Bits.from_bools(y)
    for xo in (0, 128, 256, 8, 64):
        for yo in (0, 128, 8):
            ln = n - 300
            for op, f in ((mnr.LogicalOperator.And, orc.and_masks), (mnr.LogicalOperator.Or, orc.or_masks),
                          (mnr.LogicalOperator.Xor, orc.xor_masks)):
                got = dev.bits_binop(gpu_ctx, op, X, xo, Y, yo, ln).download()
                exp = f((ox, xo, ln), (oy, yo, ln))
                assert got.len == exp.len and np.array_equal(got.bits, exp.bits), (xo, yo, op)
        got = dev.bits_not(gpu_ctx, X, xo, n - 300).download()
        assert np.array_equal(got.bits, orc.not_mask((ox, xo, n - 300)).bits)
        if xo % 64 == 0:
            assert dev.bits_popcount(gpu_ctx, X, xo, n - 300) == orc.popcount_mask((ox, xo, n - 300))


def test_batched_reductions_equal_one_by_one(mnr, gpu_ctx):
    """mnr_reduce_stats_batch over a mixed list (dtypes, lengths incl. 0, masked and dense, unaligned views) must give
    exactly the aggregates of mnr_reduce_stats called per item — the SuperTable per-batch x per-column fan-out."""
    dev = mnr.device_ops
    rng = np.random.default_rng(33)
    bufs, vals, hosts = [], [], []
    for dt in (np.int32, np.int64, np.uint32, np.float32, np.float64, np.int16):
        for n in (0, 1, 1000, 65_537, 300_001):
            if np.dtype(dt).kind == "f":
                d = (rng.standard_normal(n) * 10).astype(dt)
            else:
                d = rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, n, dtype=dt, endpoint=True)
            B = mnr.DeviceBuffer.upload(gpu_ctx, d)
            for masked in (False, True):
                v = rng.random(n) < 0.9
                bufs.append(B)
                vals.append(mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(v)) if masked else None)
                hosts.append((d, v if masked else None))
            if n > 100:   # an unaligned view joins the batch too
                bufs.append(B.slice(3, n - 10))
                vals.append(None)
                hosts.append((d[3:n - 7], None))
    for minmax in (True, False):
        got = dev.reduce_stats_batch(gpu_ctx, bufs, vals, minmax)
        assert len(got) == len(bufs)
        for i, (b, v) in enumerate(zip(bufs, vals)):
            one = dev.reduce_stats(gpu_ctx, b, v) if minmax else None
            d, hv = hosts[i]
            ex = orc.stats(d, None if hv is None else orc.Bits.from_bools(hv))
            g = got[i]
            assert g["count"] == ex["count"], i
            if minmax:
                # identical bits to the single call (floats: same summation order by construction)
                assert np.float64(g["sum"]).tobytes() == np.float64(one["sum"]).tobytes() or g["sum"] == one["sum"], i
                same_mm = lambda p, q: (p == q) or (p != p and q != q)   # noqa: E731
                assert same_mm(g["min"], one["min"]) and same_mm(g["max"], one["max"]), i
            if d.dtype.kind != "f":
                assert g["sum"] == ex["sum"], i
                if minmax:
                    assert (g["min"], g["max"]) == (ex["min"], ex["max"]), i
            else:
                sel = d if hv is None else d[hv]
                assert abs(g["sum"] - ex["sum"]) <= 1e-12 * max(1.0, np.abs(sel.astype(np.float64)).sum()), i


def test_batched_elementwise_equals_one_by_one(mnr, gpu_ctx):
    """mnr_ew_binary_batch / mnr_ew_scalar_batch_into over a mixed chunk list (dtypes, lengths incl. 0 and ragged,
    one/two/no masks, unaligned views) == the oracle leaf per chunk, for a cheap op and a divide (integer zero
    divisors null the row); dense integer division by zero anywhere in the batch is the reference's panic."""
    dev = mnr.device_ops
    rng = np.random.default_rng(34)
    for op in (orc.ADD, orc.MUL, orc.DIV, orc.REM):
        L, R, LM, RM, host = [], [], [], [], []
        for dt in (np.int32, np.int64, np.uint64, np.float32, np.float64, np.uint8):
            is_f = np.dtype(dt).kind == "f"
            for n in (0, 1, 63, 4096, 100_003):
                if is_f:
                    a, b = (rng.standard_normal(n) * 9).astype(dt), (rng.standard_normal(n) * 9).astype(dt)
                else:
                    info = np.iinfo(dt)
                    a = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)
                    b = rng.integers(max(info.min, -4), min(info.max, 4), n, dtype=dt, endpoint=True)
                A, B = mnr.DeviceBuffer.upload(gpu_ctx, a), mnr.DeviceBuffer.upload(gpu_ctx, b)
                for kind in ("two", "lhs", "rhs"):
                    la, lb = rng.random(n) < 0.9, rng.random(n) < 0.9
                    L.append(A); R.append(B)
                    LM.append(mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(la)) if kind != "rhs" else None)
                    RM.append(mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(lb)) if kind != "lhs" else None)
                    merged = la & lb if kind == "two" else la if kind == "lhs" else lb
                    host.append((a, b, merged))
                if n > 100:   # unaligned views (element-load tier, launched one by one inside the batch call)
                    m = rng.random(n - 9) < 0.9
                    L.append(A.slice(3, n - 9)); R.append(B.slice(1, n - 9))
                    LM.append(mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(m))); RM.append(None)
                    host.append((a[3:n - 6], b[1:n - 8], m))
        obs, oms = dev.ew_binary_batch(gpu_ctx, op, L, R, LM, RM, mnr.MaskMode.And)
        assert len(obs) == len(L)
        for i, (a, b, merged) in enumerate(host):
            exp, em = orc.apply(a, b, op, orc.Bits.from_bools(merged))
            assert _bits_equal(obs[i].download(), exp), (op, i, a.dtype, a.size)
            assert np.array_equal(oms[i].download().bits, em.bits), (op, i)
        # scalar broadcast over the same chunks, both operand orders
        outs = [mnr.DeviceBuffer.alloc(gpu_ctx, x.dtype, len(x)) for x in L]
        outm = [mnr.DeviceBitmask.alloc(gpu_ctx, len(x)) for x in L]
        for s_lhs in (False, True):
            dev.ew_scalar_batch_into(gpu_ctx, op, L, [3] * len(L), s_lhs, LM if all(m is not None for m in LM) else
                                     [m if m is not None else r for m, r in zip(LM, RM)], outs, outm)
            for i, (a, b, merged) in enumerate(host):
                mask_used = LM[i] if LM[i] is not None else RM[i]
                mb = mask_used.download().to_bools()
                full = np.full(a.size, 3, dtype=a.dtype)
                l, r = (full, a) if s_lhs else (a, full)
                exp, em = orc.apply(l, r, op, orc.Bits.from_bools(mb))
                assert _bits_equal(outs[i].download(), exp), (op, s_lhs, i)
                assert np.array_equal(outm[i].download().bits, em.bits), (op, s_lhs, i)
    # dense integer chunks, one zero divisor in the third chunk -> DivideByZero for the batch
    a = [mnr.DeviceBuffer.upload(gpu_ctx, np.arange(1, 1001, dtype=np.int64)) for _ in range(4)]
    d = np.ones(1000, dtype=np.int64)
    bs = [mnr.DeviceBuffer.upload(gpu_ctx, d) for _ in range(4)]
    obs, oms = dev.ew_binary_batch(gpu_ctx, orc.DIV, a, bs)
    assert all(m is None for m in oms) and _bits_equal(obs[0].download(), np.arange(1, 1001, dtype=np.int64))
    d2 = d.copy(); d2[777] = 0
    bs[2] = mnr.DeviceBuffer.upload(gpu_ctx, d2)
    with pytest.raises(mnr.KernelError) as ei:
        dev.ew_binary_batch(gpu_ctx, orc.DIV, a, bs)
    assert ei.value.kind == "DivideByZero"
    # ragged pair -> LengthMismatch, like the SuperArray route's per-chunk check (broadcast/super_array.rs:203-213)
    with pytest.raises(mnr.KernelError) as ei:
        dev.ew_binary_batch(gpu_ctx, orc.ADD, [a[0]], [mnr.DeviceBuffer.upload(gpu_ctx, np.ones(7, dtype=np.int64))])
    assert ei.value.kind == "LengthMismatch"
