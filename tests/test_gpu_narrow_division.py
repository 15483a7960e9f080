"""8/16-bit integer Div / Rem / FloorDiv run through the f32 pipe (`narrow_quot`, minarrow_b200/csrc/ew_kernels.cuh:
trunc((|l| + 0.5) * rcp(|r|)) instead of the generic 32-bit divide), and float `%` through one division + one fma
(minarrow_b200/csrc/fastmod.h) instead of the libm loop.  Both must stay bit-identical to the reference semantics
(src/kernels/arithmetic/std.rs:54-77,96-136,150): checked here over the WHOLE operand domain for 8-bit and 16-bit
columns, and for floats over every exponent distance plus the special values."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt", [np.int8, np.uint8])
def test_8bit_division_exhaustive(gpu_ctx, dt):
    import minarrow_b200 as mnr
    dev = mnr.device_ops
    info = np.iinfo(dt)
    v = np.arange(info.min, info.max + 1).astype(dt)
    a, b = np.repeat(v, 256), np.tile(v, 256)          # all 65 536 (dividend, divisor) pairs, zero divisors included
    valid = np.ones(a.size, dtype=bool)
    valid[::7] = False
    A, B = mnr.DeviceBuffer.upload(gpu_ctx, a), mnr.DeviceBuffer.upload(gpu_ctx, b)
    V = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(valid))
    for op in (orc.DIV, orc.REM, orc.FLOORDIV):
        exp, em = orc.apply_int(a, b, op, orc.Bits.from_bools(valid))
        ob, om = dev.ew_binary(gpu_ctx, op, A, B, V, None, mnr.MaskMode.And)
        assert ob.download().tobytes() == exp.tobytes(), (dt, op)
        assert np.array_equal(om.download().bits, em.bits)
    # dense, non-zero divisors only (a zero divisor is DivideByZero there)
    nz = b != 0
    a2, b2 = a[nz], b[nz]
    A2, B2 = mnr.DeviceBuffer.upload(gpu_ctx, a2), mnr.DeviceBuffer.upload(gpu_ctx, b2)
    for op in (orc.DIV, orc.REM, orc.FLOORDIV):
        exp, _ = orc.apply_int(a2, b2, op, None)
        ob, om = dev.ew_binary(gpu_ctx, op, A2, B2, None, None, mnr.MaskMode.And)
        assert om is None and ob.download().tobytes() == exp.tobytes(), (dt, op, "dense")
        ob, _ = dev.ew_binary(gpu_ctx, op, A2.slice(3, a2.size - 5), B2.slice(3, a2.size - 5), None, None, mnr.MaskMode.And)
        assert ob.download().tobytes() == orc.apply_int(a2[3:-2], b2[3:-2], op, None)[0].tobytes(), (dt, op, "unaligned")
    with pytest.raises(mnr.KernelError) as ei:
        dev.ew_binary(gpu_ctx, orc.DIV, A, B, None, None, mnr.MaskMode.And)
    assert ei.value.kind == "DivideByZero"
    # every scalar divisor over every dividend (the reciprocal is hoisted out of the loop on this route)
    D = mnr.DeviceBuffer.upload(gpu_ctx, v)
    for s in v:
        if s == 0:
            continue
        for op in (orc.DIV, orc.REM, orc.FLOORDIV):
            exp, _ = orc.apply_int(v, np.full(v.size, s, dtype=dt), op, None)
            ob, _ = dev.ew_scalar(gpu_ctx, op, D, dt(s), False, None)
            assert ob.download().tobytes() == exp.tobytes(), (dt, int(s), op, "scalar")


@pytest.mark.parametrize("dt", [np.int8, np.uint8])
def test_8bit_power_exhaustive(gpu_ctx, dt):
    """8-bit Power goes through a shared-memory table of base^e mod 256 (ew_kernels.cuh packed_pow8_vec: odd bases have
    order | 64, even bases vanish from e = 8 on; negative exponents are 0 -> 1): all 65 536 (base, exponent) pairs, masked,
    dense, unaligned (per-element path) and with a scalar on either side, against the oracle's repeated multiply."""
    import minarrow_b200 as mnr
    dev = mnr.device_ops
    info = np.iinfo(dt)
    v = np.arange(info.min, info.max + 1).astype(dt)
    a, b = np.repeat(v, 256), np.tile(v, 256)
    valid = np.ones(a.size, dtype=bool)
    valid[::5] = False
    A, B = mnr.DeviceBuffer.upload(gpu_ctx, a), mnr.DeviceBuffer.upload(gpu_ctx, b)
    V = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(valid))
    exp, em = orc.apply_int(a, b, orc.POW, orc.Bits.from_bools(valid))
    ob, om = dev.ew_binary(gpu_ctx, orc.POW, A, B, V, None, mnr.MaskMode.And)
    assert ob.download().tobytes() == exp.tobytes() and np.array_equal(om.download().bits, em.bits), dt
    exp, _ = orc.apply_int(a, b, orc.POW, None)
    ob, om = dev.ew_binary(gpu_ctx, orc.POW, A, B, None, None, mnr.MaskMode.And)
    assert om is None and ob.download().tobytes() == exp.tobytes(), (dt, "dense")
    ob, _ = dev.ew_binary(gpu_ctx, orc.POW, A.slice(3, a.size - 5), B.slice(3, a.size - 5), None, None, mnr.MaskMode.And)
    assert ob.download().tobytes() == orc.apply_int(a[3:-2], b[3:-2], orc.POW, None)[0].tobytes(), (dt, "unaligned")
    D = mnr.DeviceBuffer.upload(gpu_ctx, np.tile(v, 64))
    for s in (info.min, -3 if info.min < 0 else 3, 0, 1, 2, 7, 9, 64, info.max):
        full = np.full(D.__len__(), s, dtype=dt)
        for s_lhs in (False, True):
            l, r = (full, np.tile(v, 64)) if s_lhs else (np.tile(v, 64), full)
            exp, _ = orc.apply_int(l, r, orc.POW, None)
            ob, _ = dev.ew_scalar(gpu_ctx, orc.POW, D, dt(s), s_lhs, None)
            assert ob.download().tobytes() == exp.tobytes(), (dt, int(s), s_lhs, "scalar")


def _check_pairs(gpu_ctx, a, b, ops):
    import minarrow_b200 as mnr
    dev = mnr.device_ops
    ones = np.ones(a.size, dtype=bool)
    A, B = mnr.DeviceBuffer.upload(gpu_ctx, a), mnr.DeviceBuffer.upload(gpu_ctx, b)
    V = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(ones))
    for op in ops:
        exp, em = orc.apply_int(a, b, op, orc.Bits.from_bools(ones))
        ob, om = dev.ew_binary(gpu_ctx, op, A, B, V, None, mnr.MaskMode.And)
        got = ob.download()
        assert got.tobytes() == exp.tobytes(), (a.dtype, op, [(int(a[i]), int(b[i]), int(got[i]), int(exp[i])) for i in np.flatnonzero(got != exp)[:5]])
        assert np.array_equal(om.download().bits, em.bits), (a.dtype, op)


@pytest.mark.parametrize("dt", [np.uint16, np.int16])
def test_16bit_division_boundaries_and_blocks(gpu_ctx, dt):
    """(i) For EVERY divisor magnitude d and every multiple k*d in range: dividends k*d - 1, k*d, k*d + 1 with all sign
    combinations — the only places a truncated approximate quotient could land on the wrong integer.  (ii) Whole
    divisor blocks (1 024 divisors x all 65 536 dividends) at the ends and the middle of the domain.  The sweep over all
    2^32 pairs is tests/sweep_div16.py (run per GPU round, result under profiles/)."""
    info = np.iinfo(dt)
    top = int(info.max) + (1 if info.min < 0 else 0)       # largest magnitude: 65 535 or 32 768
    ds, ls = [], []
    for d in range(1, top + 1):
        k = np.arange(0, top // d + 1, dtype=np.int64) * d
        m = np.concatenate([k - 1, k, k + 1])
        m = m[(m >= 0) & (m <= top)]
        ls.append(m); ds.append(np.full(m.size, d, dtype=np.int64))
    l, d = np.concatenate(ls), np.concatenate(ds)
    if info.min < 0:
        l = np.concatenate([l, -l, l, -l]); d = np.concatenate([d, d, -d, -d])
        keep = (l <= info.max) & (d <= info.max)            # +32 768 does not exist
        l, d = l[keep], d[keep]
    _check_pairs(gpu_ctx, l.astype(dt), d.astype(dt), (orc.DIV, orc.REM, orc.FLOORDIV))
    v = np.arange(info.min, info.max + 1).astype(dt)
    a = np.tile(v, 1024)
    for blk, op in ((0, orc.DIV), (31, orc.FLOORDIV), (32, orc.REM), (63, orc.DIV)):
        _check_pairs(gpu_ctx, a, np.repeat(v[blk * 1024:(blk + 1) * 1024], 65536), (op,))


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_float_remainder_is_exact_fmod(gpu_ctx, dt):
    import minarrow_b200 as mnr
    dev = mnr.device_ops
    rng = np.random.default_rng(11)
    n = 1 << 21
    fi = np.finfo(dt)
    # every exponent distance from 0 to beyond the mantissa width (fast path and library path), both signs
    a = rng.standard_normal(n).astype(dt) * np.exp2(rng.integers(-40, 41, n)).astype(dt)
    b = rng.standard_normal(n).astype(dt) * np.exp2(rng.integers(-40, 41, n)).astype(dt)
    # exact multiples and their neighbours (where the rounded quotient lands one too high)
    k = rng.integers(0, 1 << 20, n // 4).astype(dt)
    a[:n // 4] = k * b[:n // 4]
    a[n // 8:n // 4] = np.nextafter(a[n // 8:n // 4], dt(0))
    specials = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, fi.tiny, -fi.tiny, fi.smallest_subnormal, fi.max, -fi.max, 1.0, -1.0, 2.0 ** 23,
                         2.0 ** 52, 0.5, 3.0], dtype=dt)
    sa, sb = np.repeat(specials, specials.size), np.tile(specials, specials.size)
    a[-sa.size:], b[-sb.size:] = sa, sb
    sub = slice(n // 2, n // 2 + 4096)                     # subnormal operands
    a[sub] = (rng.integers(1, 1 << 20, 4096) * float(fi.smallest_subnormal)).astype(dt)
    b[sub.start:sub.start + 2048] = (rng.integers(1, 1 << 10, 2048) * float(fi.smallest_subnormal)).astype(dt)
    valid = rng.random(n) < 0.9
    exp, em = orc.apply_float(a, b, orc.REM, orc.Bits.from_bools(valid))
    ob, om = dev.ew_binary(gpu_ctx, orc.REM, mnr.DeviceBuffer.upload(gpu_ctx, a), mnr.DeviceBuffer.upload(gpu_ctx, b),
                           mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(valid)), None, mnr.MaskMode.And)
    got = ob.download()
    nan_e, nan_g = np.isnan(exp), np.isnan(got)
    assert np.array_equal(nan_e, nan_g)                    # NaN by position (payload is outside the contract)
    ui = np.uint64 if dt == np.float64 else np.uint32
    assert np.array_equal(exp.view(ui)[~nan_e], got.view(ui)[~nan_e])
    assert np.array_equal(om.download().bits, em.bits)
    # scalar divisor and scalar dividend
    for s, lhs in ((dt(0.37), False), (dt(1e6), True)):
        full = np.full(n, s, dtype=dt)
        exp, _ = orc.apply_float(full, a, orc.REM, None) if lhs else orc.apply_float(a, full, orc.REM, None)
        got = dev.ew_scalar(gpu_ctx, orc.REM, mnr.DeviceBuffer.upload(gpu_ctx, a), s, lhs, None)[0].download()
        nan_e = np.isnan(exp)
        assert np.array_equal(nan_e, np.isnan(got)) and np.array_equal(exp.view(ui)[~nan_e], got.view(ui)[~nan_e])


@pytest.mark.parametrize("dt", [np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64])
def test_integer_power_full_exponent_range(gpu_ctx, dt):
    """16 / 32 / 64-bit Power runs square-and-multiply with a trip count shared by the rows of one thread (ew_kernels.cuh
    packed_pow16_vec / int_pow_vec).  Exponents over the whole range of the type — bit lengths mixed inside a vector, a lone
    large exponent among zeros, negative ones (-> 0, std.rs:67) and, for 64-bit columns, ones beyond u32::MAX (-> 0) —
    masked, dense, unaligned and with a scalar on either side, against the oracle; a sample is also checked against
    Python's pow(base, e, 2^bits)."""
    import minarrow_b200 as mnr
    dev = mnr.device_ops
    rng = np.random.default_rng(77)
    info = np.iinfo(dt)
    bits = 8 * np.dtype(dt).itemsize
    n = 262_144 + 37
    a = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)
    a[::11] = rng.integers(-3 if info.min < 0 else 0, 3, a[::11].size).astype(dt)
    nb = rng.integers(0, bits + 1, n)                                    # bit length of each exponent: uniform, so vectors mix them
    b = (rng.integers(0, 1 << 62, n, dtype=np.int64).astype(np.uint64) * np.uint64(4) + np.uint64(3))
    b = np.where(nb == 0, np.uint64(0), b >> (np.uint64(64) - np.maximum(nb, 1).astype(np.uint64))).astype(np.uint64)
    b = b.astype(dt) if info.min == 0 else b.astype(np.dtype(f"u{bits // 8}")).view(dt)   # signed: top bit set = negative exponent
    b[1000:3000] = 0
    b[2017] = dt(info.max)                                                # a lone long exponent in a run of zeros
    b[5000:5100] = np.arange(100).astype(dt)
    valid = rng.random(n) < 0.8
    A, B = mnr.DeviceBuffer.upload(gpu_ctx, a), mnr.DeviceBuffer.upload(gpu_ctx, b)
    V = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(valid))
    exp, em = orc.apply_int(a, b, orc.POW, orc.Bits.from_bools(valid))
    ob, om = dev.ew_binary(gpu_ctx, orc.POW, A, B, V, None, mnr.MaskMode.And)
    got = ob.download()
    assert got.tobytes() == exp.tobytes(), (dt, [(int(a[i]), int(b[i]), int(got[i]), int(exp[i])) for i in np.flatnonzero(got != exp)[:5]])
    assert np.array_equal(om.download().bits, em.bits), dt
    exp, _ = orc.apply_int(a, b, orc.POW, None)
    ob, om = dev.ew_binary(gpu_ctx, orc.POW, A, B, None, None, mnr.MaskMode.And)
    got = ob.download()
    assert om is None and got.tobytes() == exp.tobytes(), (dt, "dense")
    for i in rng.integers(0, n, 300):                                     # independent of the oracle's loop
        e = int(b[i])
        e = 0 if (e < 0 or e > 0xFFFFFFFF) else e
        assert int(got[i]) % (1 << bits) == pow(int(a[i]), e, 1 << bits), (dt, int(a[i]), int(b[i]))
    ob, _ = dev.ew_binary(gpu_ctx, orc.POW, A.slice(3, n - 5), B.slice(3, n - 5), None, None, mnr.MaskMode.And)
    assert ob.download().tobytes() == orc.apply_int(a[3:-2], b[3:-2], orc.POW, None)[0].tobytes(), (dt, "unaligned")
    for s in (info.min, 0, 1, 2, 3, 13, 1000, info.max):
        full = np.full(n, s, dtype=dt)
        for s_lhs in (False, True):
            l, r = (full, b) if s_lhs else (a, full)
            D = B if s_lhs else A
            exp, em = orc.apply_int(l, r, orc.POW, orc.Bits.from_bools(valid))
            ob, om = dev.ew_scalar(gpu_ctx, orc.POW, D, dt(s), s_lhs, V)
            assert ob.download().tobytes() == exp.tobytes() and np.array_equal(om.download().bits, em.bits), (dt, int(s), s_lhs, "scalar")
