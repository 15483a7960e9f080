"""Parity of the CUDA path (through the C ABI) with the CPU oracle on seeded inputs.

Bar: bit-exact for integer values, float add/sub/mul/div/rem/floordiv values, every validity bitmask, counts,
integer sums / min / max; float Power within the reference's own test tolerance (arithmetic/mod.rs:328-340:
1e-6 f32 / 1e-12 f64, relative); float sums within 1e-12 relative (f64) of the oracle order.
"""
import math

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 2, 3, 7, 8, 9, 31, 63, 64, 65, 127, 128, 129, 255, 256, 257, 1023, 1024, 1025, 4097, 32768 + 5, 100_003]
INT_DT = [np.int32, np.uint32, np.int64, np.uint64, np.int8, np.uint8, np.int16, np.uint16]
FLT_DT = [np.float32, np.float64]
OPS = list(range(7))


@pytest.fixture(scope="module")
def mnr(gpu_ctx):
    import minarrow_b200 as m
    return m


def rand_int(rng, dt, n, small=False):
    info = np.iinfo(dt)
    if small:
        return rng.integers(max(info.min, -9), min(info.max, 9), n, dtype=dt, endpoint=True)
    a = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)
    if n:
        edges = np.array([info.min, info.max, 0, 1, 2] + ([-1] if info.min < 0 else []), dtype=dt)
        idx = rng.integers(0, n, max(1, n // 16))
        a[idx] = edges[rng.integers(0, len(edges), idx.size)]
    return a


def rand_float(rng, dt, n):
    a = rng.standard_normal(n).astype(dt) * dt(100.0)
    if n:
        sp = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, np.finfo(dt).tiny / 4, np.finfo(dt).max, 1.0, -1.0], dtype=dt)
        idx = rng.integers(0, n, max(1, n // 16))
        a[idx] = sp[rng.integers(0, len(sp), idx.size)]
    return a


def bits_equal(a, b):
    """Bit-for-bit equality, NaN payloads included."""
    return a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()


def same_float(got, exp, op, dt, a=None, b=None):
    if op == orc.POW:
        # Power = exp(b * ln a) through libm in the reference (std.rs:153-154): not bit-reproducible across libms, its own
        # test (arithmetic/mod.rs:328-340) uses a relative tolerance of 1e-12 (f64) / 1e-6 (f32) on tame values (2^3 ...).
        # The bar here is exactly that tolerance, RELATIVE to the result, with one derived exception: both sides evaluate
        # exp(x) at x = fl(b * fl(ln a)) in the working precision, and two correctly-behaving libms (ln <= 1 ulp each, one
        # rounding of the product each, exp <= 1 ulp each) can differ by 3 ulp in x, which exp() turns into a relative
        # difference of 3 |x| eps (+ 2 eps).  For f64 that is 5e-14 at |x| = 74 — far inside 1e-12, so no allowance.  For f32
        # (eps = 2^-23) it exceeds 1e-6 as soon as |x| > 2.5, so the f32 bar is max(1e-6, (4 |x| + 4) * 2^-24).
        g, e = got.astype(np.float64), exp.astype(np.float64)
        nan_ok = np.isnan(g) == np.isnan(e)
        fin = np.isfinite(e) & np.isfinite(g)
        inf_ok = np.where(~fin & ~np.isnan(e), g == e, True)
        if dt == np.float32:
            with np.errstate(all="ignore"):
                x = np.abs(b.astype(np.float64) * np.log(a.astype(np.float64))) if a is not None else np.zeros_like(e)
            x = np.nan_to_num(x, nan=0.0, posinf=0.0)      # 0 / Inf / negative bases give exact 0 / Inf / NaN on both sides
            tol = np.maximum(1e-6, (4.0 * x + 4.0) * 2.0 ** -24)[fin]
        else:
            tol = 1e-12
        close = np.abs(g[fin] - e[fin]) <= tol * np.abs(e[fin]) + np.finfo(dt).tiny
        return nan_ok.all() and inf_ok.all() and close.all()
    nan = np.isnan(exp)
    return np.array_equal(np.isnan(got), nan) and bits_equal(got[~nan], exp[~nan])


def check_mask(got, exp):
    assert (got is None) == (exp is None)
    if exp is not None:
        assert got.len == exp.len and np.array_equal(got.bits, exp.bits), "validity bytes differ (incl. slack bits)"


@pytest.mark.parametrize("dt", INT_DT)
def test_int_leaf_all_ops_all_sizes(mnr, gpu_ctx, dt):
    rng = np.random.default_rng(10)
    f = mnr.kernels.arithmetic.APPLY[np.dtype(dt)]
    for n in SIZES:
        a = rand_int(rng, dt, n)
        for op in OPS:
            b = rand_int(rng, dt, n, small=op in (orc.DIV, orc.REM, orc.FLOORDIV, orc.POW))
            mask = orc.Bits.from_bools(rng.random(n) < 0.8)
            exp, em = orc.apply_int(a, b, op, mask)
            got = f(a, b, op, mnr.Bitmask(mask.bits, n), gpu_ctx)
            assert bits_equal(got.data, exp), (dt, n, op)
            check_mask(got.null_mask, em)
            # dense: zero divisors must raise like the reference panics; otherwise bit-exact, no mask
            if op in (orc.DIV, orc.REM, orc.FLOORDIV):
                if n and (b == 0).any():
                    with pytest.raises(mnr.KernelError) as ei:
                        f(a, b, op, None, gpu_ctx)
                    assert ei.value.kind == "DivideByZero"
                b = np.where(b == 0, dt(1), b)
            exp, em = orc.apply_int(a, b, op, None)
            got = f(a, b, op, None, gpu_ctx)
            assert bits_equal(got.data, exp) and got.null_mask is None and em is None, (dt, n, op)


@pytest.mark.parametrize("dt", FLT_DT)
def test_float_leaf_all_ops_all_sizes(mnr, gpu_ctx, dt):
    rng = np.random.default_rng(11)
    f = mnr.kernels.arithmetic.APPLY[np.dtype(dt)]
    for n in SIZES:
        a, b = rand_float(rng, dt, n), rand_float(rng, dt, n)
        for op in OPS:
            if op == orc.POW:
                a2 = np.abs(a) + dt(0.5)
                b2 = np.clip(b, -8, 8).astype(dt)
                b2[np.isnan(b2)] = dt(1.5)
                a2[~np.isfinite(a2)] = dt(2.0)
                a2 = np.minimum(a2, dt(1e4))
            else:
                a2, b2 = a, b
            mask = orc.Bits.from_bools(rng.random(n) < 0.8)
            exp, em = orc.apply_float(a2, b2, op, mask)
            got = f(a2, b2, op, mnr.Bitmask(mask.bits, n), gpu_ctx)
            assert same_float(got.data, exp, op, dt, a2, b2), (dt, n, op)
            check_mask(got.null_mask, em)
            exp, _ = orc.apply_float(a2, b2, op, None)
            got = f(a2, b2, op, None, gpu_ctx)
            assert same_float(got.data, exp, op, dt, a2, b2) and got.null_mask is None, (dt, n, op)


@pytest.mark.parametrize("dt", FLT_DT)
def test_fma_matches_oracle(mnr, gpu_ctx, dt):
    rng = np.random.default_rng(12)
    f = mnr.apply_fma_f32 if dt == np.float32 else mnr.apply_fma_f64
    for n in SIZES:
        a, b, c = (rand_float(rng, dt, n) for _ in range(3))
        mask = orc.Bits.from_bools(rng.random(n) < 0.7)
        exp, em = orc.apply_fma(a, b, c, mask)
        got = f(a, b, c, mnr.Bitmask(mask.bits, n), gpu_ctx)
        assert same_float(got.data, exp, 0, dt)
        check_mask(got.null_mask, em)
        exp, _ = orc.apply_fma(a, b, c, None)
        got = f(a, b, c, None, gpu_ctx)
        assert same_float(got.data, exp, 0, dt) and got.null_mask is None
    with pytest.raises(mnr.KernelError) as ei:
        f(np.zeros(3, dt), np.zeros(3, dt), np.zeros(2, dt), None, gpu_ctx)
    assert ei.value.kind == "LengthMismatch"


def test_length_mismatch_and_empty(mnr, gpu_ctx):
    with pytest.raises(mnr.KernelError) as ei:
        mnr.apply_int_i64(np.arange(4), np.arange(3), mnr.ArithmeticOperator.Add, None, gpu_ctx)
    assert ei.value.kind == "LengthMismatch"
    out = mnr.apply_float_f64(np.array([]), np.array([]), mnr.ArithmeticOperator.Add, None, gpu_ctx)
    assert out.is_empty() and out.null_mask is None


@pytest.mark.parametrize("dt", [np.int32, np.int64, np.uint64, np.float32, np.float64])
def test_device_resident_two_masks_scalar_and_views(mnr, gpu_ctx, dt):
    """Fused two-mask merge (AND = merge_bitmasks_to_new, OR = Bitmask::union), scalar operands on either
    side, and ArrayV windows at odd offsets (unaligned pointers -> scalar-path kernel)."""
    dev = mnr.device_ops
    rng = np.random.default_rng(13)
    is_f = np.dtype(dt).kind == "f"
    gen = (lambda n, small=False: rand_float(rng, dt, n)) if is_f else (lambda n, small=False: rand_int(rng, dt, n, small))
    oapply = orc.apply_float if is_f else orc.apply_int
    for n in [1, 5, 64, 67, 1000, 4099, 70_001]:
        a, b = gen(n), gen(n, True)
        la, lb = rng.random(n) < 0.85, rng.random(n) < 0.85
        A, B = mnr.DeviceBuffer.upload(gpu_ctx, a), mnr.DeviceBuffer.upload(gpu_ctx, b)
        LA = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(la))
        LB = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(lb))
        for op in (orc.ADD, orc.MUL, orc.DIV, orc.REM):
            for mode, merged in ((mnr.MaskMode.And, la & lb), (mnr.MaskMode.Or, la | lb)):
                exp, em = oapply(a, b, op, orc.Bits.from_bools(merged))
                ob, om = dev.ew_binary(gpu_ctx, op, A, B, LA, LB, mode)
                assert same_float(ob.download(), exp, op, dt) if is_f else bits_equal(ob.download(), exp)
                check_mask(om.download(), em)
            # one mask on either side behaves the same in both modes
            exp, em = oapply(a, b, op, orc.Bits.from_bools(lb))
            ob, om = dev.ew_binary(gpu_ctx, op, A, B, None, LB, mnr.MaskMode.And)
            assert same_float(ob.download(), exp, op, dt) if is_f else bits_equal(ob.download(), exp)
            check_mask(om.download(), em)
            # scalar broadcast, both operand orders, == materialised broadcast_length_1_array + leaf kernel
            s = dt(3)
            full = np.full(n, s, dtype=dt)
            for s_lhs in (False, True):
                l, r = (full, a) if s_lhs else (a, full)
                exp, em = oapply(l, r, op, orc.Bits.from_bools(la))
                ob, om = dev.ew_scalar(gpu_ctx, op, A, s, s_lhs, LA)
                assert same_float(ob.download(), exp, op, dt) if is_f else bits_equal(ob.download(), exp)
                check_mask(om.download(), em)
                try:
                    exp, _ = oapply(l, r, op, None)
                except orc.KernelError as e:   # dense integer zero divisor: both sides must refuse
                    assert e.kind == "DivideByZero"
                    with pytest.raises(mnr.KernelError) as ei:
                        dev.ew_scalar(gpu_ctx, op, A, s, s_lhs, None)
                    assert ei.value.kind == "DivideByZero"
                    continue
                ob, om = dev.ew_scalar(gpu_ctx, op, A, s, s_lhs, None)
                assert om is None
                assert same_float(ob.download(), exp, op, dt) if is_f else bits_equal(ob.download(), exp)
        # views: data sliced at an odd offset, mask indexed from bit 0 (routing/arithmetic.rs:284-287)
        if n > 8:
            off, ln = 3, n - 5
            exp, em = oapply(a[off:off + ln], b[1:1 + ln], orc.ADD, orc.Bits.from_bools(la[:ln]))
            LAv = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(la[:ln]))
            ob, om = dev.ew_binary(gpu_ctx, orc.ADD, A.slice(off, ln), B.slice(1, ln), LAv, None, mnr.MaskMode.And)
            assert same_float(ob.download(), exp, 0, dt) if is_f else bits_equal(ob.download(), exp)
            check_mask(om.download(), em)
            st = dev.reduce_stats(gpu_ctx, A.slice(off, ln), LAv)
            es = orc.stats(a[off:off + ln], orc.Bits.from_bools(la[:ln]))
            assert st["count"] == es["count"]
            if not is_f:
                assert (st["sum"], st["min"], st["max"]) == (es["sum"], es["min"], es["max"])


def test_router_broadcast_promotion_and_errors(mnr, gpu_ctx):
    rng = np.random.default_rng(14)
    R = mnr.resolve_binary_arithmetic
    n = 1000
    i32 = rand_int(rng, np.int32, n, small=True)
    for fdt in (np.float64, np.float32):
        f = rand_float(rng, fdt, n)
        for op in (orc.ADD, orc.SUB, orc.MUL, orc.DIV):
            for l, r in ((i32, f), (f, i32)):
                exp, _ = orc.resolve_binary_arithmetic(op, l, r)
                got = R(op, l, r, None, gpu_ctx)
                assert got.data.dtype == np.dtype(fdt) and same_float(got.data, exp, op, fdt)
        # promoted length-1 broadcast
        exp, _ = orc.resolve_binary_arithmetic(orc.MUL, i32[:1], f)
        got = R(orc.MUL, i32[:1], f, None, gpu_ctx)
        assert same_float(got.data, exp, orc.MUL, fdt)
    mask = mnr.Bitmask.from_bools(rng.random(n) < 0.5)
    exp, em = orc.resolve_binary_arithmetic(orc.SUB, np.array([7], np.int32), i32, orc.Bits(mask.bits, n))
    got = R(orc.SUB, np.array([7], np.int32), i32, mask, gpu_ctx)
    assert bits_equal(got.data, exp) and np.array_equal(got.null_mask.bits, em.bits)
    with pytest.raises(mnr.KernelError) as ei:
        R(orc.ADD, np.arange(3, dtype=np.int32), np.arange(2, dtype=np.int32), None, gpu_ctx)
    assert ei.value.kind == "LengthMismatch"
    with pytest.raises(mnr.KernelError) as ei:
        R(orc.ADD, np.arange(3, dtype=np.int64), np.arange(3, dtype=np.float64), None, gpu_ctx)
    assert ei.value.kind == "UnsupportedType"


def test_super_array_route_or_union_of_chunk_masks(mnr, gpu_ctx):
    """route_super_array_broadcast ORs the two chunks' masks (super_array.rs:214-229) — reproduced fused."""
    rng = np.random.default_rng(15)
    lhs, rhs, exp = [], [], []
    for n in (100, 257, 64, 1):
        a, b = rand_int(rng, np.int64, n), rand_int(rng, np.int64, n, small=True)
        la, lb = rng.random(n) < 0.7, rng.random(n) < 0.7
        lhs.append(mnr.IntegerArray(a, mnr.Bitmask.from_bools(la)))
        rhs.append(mnr.IntegerArray(b, mnr.Bitmask.from_bools(lb) if n != 64 else None))
        merged = (la | lb) if n != 64 else la
        exp.append(orc.apply_int(a, b, orc.DIV, orc.Bits.from_bools(merged)))
    out = mnr.route_super_array_broadcast(orc.DIV, mnr.SuperArray(lhs), mnr.SuperArray(rhs), None, gpu_ctx)
    for got, (ed, em) in zip(out.chunks, exp):
        assert bits_equal(got.data, ed) and np.array_equal(got.null_mask.bits, em.bits)


def test_bitmask_kernels_windows_and_offsets(mnr, gpu_ctx):
    bm = mnr.kernels.bitmask
    rng = np.random.default_rng(16)
    for n in [1, 7, 8, 9, 63, 64, 65, 127, 128, 129, 1000, 4099, 131072 + 77, 1 << 21]:
        a, b = rng.random(n) < 0.5, rng.random(n) < 0.5
        A, B = mnr.Bitmask.from_bools(a), mnr.Bitmask.from_bools(b)
        oA, oB = orc.Bits(A.bits, n), orc.Bits(B.bits, n)
        for f, g in ((bm.and_masks, orc.and_masks), (bm.or_masks, orc.or_masks), (bm.xor_masks, orc.xor_masks)):
            got, exp = f((A, 0, n), (B, 0, n), gpu_ctx), g((oA, 0, n), (oB, 0, n))
            assert got.len == n and np.array_equal(got.bits, exp.bits)
        assert np.array_equal(bm.not_mask((A, 0, n), gpu_ctx).bits, orc.not_mask((oA, 0, n)).bits)
        assert bm.popcount_mask((A, 0, n), gpu_ctx) == int(a.sum())
        assert bm.null_count(A, gpu_ctx) == n - int(a.sum())
        assert np.array_equal(bm.eq_mask((A, 0, n), (B, 0, n), gpu_ctx).bits, orc.eq_mask((oA, 0, n), (oB, 0, n)).bits)
        assert np.array_equal(bm.ne_mask((A, 0, n), (B, 0, n), gpu_ctx).bits, orc.ne_mask((oA, 0, n), (oB, 0, n)).bits)
        assert bm.all_eq((A, 0, n), (A, 0, n), gpu_ctx) and bm.all_eq((A, 0, n), (B, 0, n), gpu_ctx) == bool((a == b).all())
        assert np.array_equal(bm.merge_bitmasks_to_new(A, B, n, gpu_ctx).bits, orc.merge_bitmasks_to_new(oA, oB, n).bits)
        assert np.array_equal(bm.merge_bitmasks_to_new(None, B, n, gpu_ctx).bits, B.bits)
        assert bm.merge_bitmasks_to_new(None, None, n, gpu_ctx) is None
        assert np.array_equal(bm.union(A, B, gpu_ctx).bits, orc.union(oA, oB).bits)
        assert np.array_equal(bm.in_mask((A, 0, n), (B, 0, n), gpu_ctx).bits, orc.in_mask((oA, 0, n), (oB, 0, n)).bits)
        assert np.array_equal(bm.not_in_mask((A, 0, n), (B, 0, n), gpu_ctx).bits,
                              orc.not_in_mask((oA, 0, n), (oB, 0, n)).bits)
        assert bm.all_true_mask(A, gpu_ctx) == bool(a.all()) and bm.all_false_mask(A, gpu_ctx) == bool((~a).all())
        if n >= 256:   # windows: word-aligned, byte-aligned and sub-byte (floored like the reference) offsets
            for lo, ro, ln in ((64, 128, n - 200), (8, 72, n - 100), (67, 3, n - 80), (129, 1, 17)):
                got, exp = bm.and_masks((A, lo, ln), (B, ro, ln), gpu_ctx), orc.and_masks((oA, lo, ln), (oB, ro, ln))
                assert np.array_equal(got.bits, exp.bits), (n, lo, ro, ln)
                assert np.array_equal(bm.not_mask((A, lo, ln), gpu_ctx).bits, orc.not_mask((oA, lo, ln)).bits)
                assert bm.popcount_mask((A, lo, ln), gpu_ctx) == orc.popcount_mask((oA, lo, ln))
            ones, zeros = mnr.Bitmask.new_set_all(n, True), mnr.Bitmask.new_set_all(n, False)
            for rhs in (ones, zeros):   # in_mask with one-valued rhs: slice_clone keeps the exact bit offset
                got = bm.in_mask((A, 5, n - 9), (rhs, 0, n - 9), gpu_ctx)
                exp = orc.in_mask((oA, 5, n - 9), (orc.Bits(rhs.bits, n), 0, n - 9))
                assert np.array_equal(got.bits, exp.bits)
    assert bm.all_true_mask(mnr.Bitmask.new_set_all(1000, True), gpu_ctx)
    assert bm.all_false_mask(mnr.Bitmask.new_set_all(1000, False), gpu_ctx)
    with pytest.raises(mnr.KernelError):
        bm.eq_mask((A, 3, 10), (B, 0, 10), gpu_ctx)   # reference panics on non-word-aligned eq offsets


@pytest.mark.parametrize("dt", INT_DT + FLT_DT)
def test_null_aware_stats_match_oracle(mnr, gpu_ctx, dt):
    rng = np.random.default_rng(17)
    red = mnr.kernels.reduce
    is_f = np.dtype(dt).kind == "f"
    for n in SIZES + [(1 << 22) + 4099]:
        d = rand_float(rng, dt, n) if is_f else rand_int(rng, dt, n)
        if is_f:
            d[np.isinf(d)] = dt(1.0)
            d[np.abs(d) > 1e30] = dt(2.0)
        for valid in (None, rng.random(n) < 0.9, np.zeros(n, bool)):
            V = None if valid is None else mnr.Bitmask.from_bools(valid)
            oV = None if valid is None else orc.Bits(V.bits, n)
            exp, got = orc.stats(d, oV), red.stats(d, V, True, gpu_ctx)
            assert got["count"] == exp["count"], (dt, n)
            if is_f:
                for k in ("min", "max"):
                    assert (math.isnan(got[k]) and math.isnan(exp[k])) or \
                        (got[k] == exp[k] and math.copysign(1, got[k]) == math.copysign(1, exp[k])), (dt, n, k)
                clean = np.where(np.isnan(d), dt(0), d)
                e2, g2 = orc.stats(clean, oV), red.stats(clean, V, False, gpu_ctx)
                sel = clean if valid is None else clean[valid]
                scale = math.fsum(np.abs(sel.astype(np.float64))) or 1.0
                assert abs(g2["sum"] - e2["sum"]) <= 1e-12 * scale, (dt, n)
                assert abs(g2["sum"] - math.fsum(sel.astype(np.float64))) <= 1e-12 * scale
                # The bound above is the forward-error bound of ANY reordered sum (cancellation makes |sum| arbitrarily
                # smaller than sum|x|).  Where the sum is well conditioned — same-sign data, sum|x| = |sum| — the stated
                # 1e-12 holds relative to the result itself, vs the oracle's order and vs the exact sum:
                pos = np.abs(clean)
                e3, g3 = orc.stats(pos, oV), red.stats(pos, V, False, gpu_ctx)
                exact = math.fsum((pos if valid is None else pos[valid]).astype(np.float64))
                assert abs(g3["sum"] - e3["sum"]) <= 1e-12 * abs(e3["sum"]) and abs(g3["sum"] - exact) <= 1e-12 * abs(exact), (dt, n)
                assert math.isnan(got["sum"]) == math.isnan(exp["sum"])
            else:
                assert (got["sum"], got["min"], got["max"]) == (exp["sum"], exp["min"], exp["max"]), (dt, n)
                if exp["count"]:
                    assert got["mean"] == exp["mean"]
            s2 = red.stats(d if not is_f else np.where(np.isnan(d), dt(0), d), V, False, gpu_ctx)
            assert s2["count"] == exp["count"]


def test_float_sum_is_run_to_run_deterministic(mnr, gpu_ctx):
    rng = np.random.default_rng(18)
    d = rng.standard_normal(3_000_017)
    D = mnr.DeviceBuffer.upload(gpu_ctx, d)
    vals = {mnr.device_ops.reduce_sum(gpu_ctx, D)[0] for _ in range(5)}
    assert len(vals) == 1
    assert abs(vals.pop() - math.fsum(d)) <= 1e-12 * math.fsum(np.abs(d))


def test_host_pipeline_multi_chunk_matches_single_pass(mnr, gpu_ctx):
    """The host drop-ins stream in chunks; results must not depend on the chunking."""
    rng = np.random.default_rng(19)
    n = 3 * 8192 + 1234
    gpu_ctx.set_option("host_chunk_rows", 8192)
    try:
        a, b = rand_int(rng, np.int64, n), rand_int(rng, np.int64, n, small=True)
        mask = orc.Bits.from_bools(rng.random(n) < 0.9)
        exp, em = orc.apply_int(a, b, orc.DIV, mask)
        got = mnr.apply_int_i64(a, b, orc.DIV, mnr.Bitmask(mask.bits, n), gpu_ctx)
        assert bits_equal(got.data, exp) and np.array_equal(got.null_mask.bits, em.bits)
        st = mnr.kernels.reduce.stats(a, mnr.Bitmask(mask.bits, n), True, gpu_ctx)
        es = orc.stats(a, mask)
        assert (st["sum"], st["min"], st["max"], st["count"]) == (es["sum"], es["min"], es["max"], es["count"])
        x, y = rng.random(n * 8 + 3) < 0.5, rng.random(n * 8 + 3) < 0.5
        X, Y = mnr.Bitmask.from_bools(x), mnr.Bitmask.from_bools(y)
        out = np.zeros(X.bits.size, np.uint8)
        import ctypes as C
        from minarrow_b200.core import check
        check(gpu_ctx.lib.mnr_bitmask_binop_host(gpu_ctx.h, 2, X.bits.ctypes.data_as(C.c_void_p), 0,
                                                 Y.bits.ctypes.data_as(C.c_void_p), 0, x.size,
                                                 out.ctypes.data_as(C.c_void_p)))
        assert np.array_equal(out, orc.xor_masks((orc.Bits(X.bits, x.size), 0, x.size), (orc.Bits(Y.bits, x.size), 0, x.size)).bits)
    finally:
        gpu_ctx.set_option("host_chunk_rows", 1 << 22)
