"""Every selectable launch geometry of the division / remainder / power kernels (`mnr_ctx_set_option`: ew_sdiv64_cfg,
ew_fdiv_cfg, ew_heavy_cfg — DESIGN.md §3.11) must produce the same bits as the oracle, not only the library's default
choice: odd lengths (guarded tails), unaligned windows (narrower vector tiers), masked and dense."""
import numpy as np
import pytest

from oracle import oracle as orc
from test_gpu_parity import same_float

pytestmark = pytest.mark.gpu


def _same(got, exp, op=None, a=None, b=None):
    """Bit-exact (NaN by position) except float Power, which has the reference's own tolerance (tests/test_gpu_parity.py)."""
    if exp.dtype.kind == "f":
        return same_float(got, exp, orc.POW if op == "pow" else orc.DIV, exp.dtype.type, a, b)
    return got.tobytes() == exp.tobytes()


def _column(rng, dt, n, nonzero=False):
    dt = np.dtype(dt)
    if dt.kind == "f":
        a = (rng.standard_normal(n) * np.exp2(rng.integers(-8, 9, n))).astype(dt)
        if nonzero:
            a[a == 0] = 1
        return a
    info = np.iinfo(dt)
    a = rng.integers(max(info.min, -50000), min(info.max, 50000), n, dtype=dt, endpoint=True)
    a[:: 97] = info.min
    a[1:: 97] = info.max
    if nonzero:
        a[a == 0] = 1
    return a


@pytest.mark.parametrize("knob,values,dtypes,ops", [
    ("ew_sdiv64_cfg", (1, 2), (np.int64, np.uint64), ("scalar",)),
    ("ew_fdiv_cfg", (1, 2, 3), (np.float32, np.float64), ("div", "floordiv", "scalar")),
    ("ew_heavy_cfg", (1, 2, 3), (np.int8, np.uint16, np.int32, np.int64, np.uint64, np.float32, np.float64), ("div", "rem", "pow")),
])
def test_every_geometry_matches_the_oracle(gpu_ctx, knob, values, dtypes, ops):
    import minarrow_b200 as mnr
    dev = mnr.device_ops
    rng = np.random.default_rng(21)
    n = 300_007
    code = {"div": orc.DIV, "floordiv": orc.FLOORDIV, "rem": orc.REM, "pow": orc.POW}
    try:
        for dt in dtypes:
            is_f = np.dtype(dt).kind == "f"
            apply = orc.apply_float if is_f else orc.apply_int
            a, b = _column(rng, dt, n), _column(rng, dt, n)
            b[:: 13] = 0                                           # zero divisors: nulls under a mask (ints), Inf / NaN (floats)
            e = rng.integers(0, 6, n).astype(dt)                   # exponents for Power
            p = (np.minimum(np.abs(a), 1e4) + 0.5).astype(dt) if is_f else a   # float Power = exp(b ln a): positive bases
            valid = rng.random(n) < 0.85
            V = mnr.Bitmask.from_bools(valid)
            A, B, E, P = (mnr.DeviceBuffer.upload(gpu_ctx, x) for x in (a, b, e, p))
            DV = mnr.DeviceBitmask.upload(gpu_ctx, V)
            bnz = b.copy(); bnz[bnz == 0] = 3
            BNZ = mnr.DeviceBuffer.upload(gpu_ctx, bnz)
            for v in values:
                gpu_ctx.set_option(knob, v)
                for op in ops:
                    if op == "scalar":
                        s = np.dtype(dt).type(2.5) if is_f else np.dtype(dt).type(86400 if np.dtype(dt).itemsize == 8 else 7)
                        full = np.full(n, s, dtype=dt)
                        for scode in (orc.DIV, orc.FLOORDIV):
                            exp, em = apply(a, full, scode, orc.Bits(V.bits, n))
                            ob, om = dev.ew_scalar(gpu_ctx, scode, A, s, False, DV)
                            assert _same(ob.download(), exp) and np.array_equal(om.download().bits, em.bits), (knob, v, dt, "scalar masked")
                            exp, _ = apply(a, full, scode, None)
                            ob, om = dev.ew_scalar(gpu_ctx, scode, A, s, False, None)
                            assert om is None and _same(ob.download(), exp), (knob, v, dt, "scalar dense")
                        continue
                    lhs, L = (p, P) if op == "pow" else (a, A)
                    rhs, R = (e, E) if op == "pow" else (b, B)
                    exp, em = apply(lhs, rhs, code[op], orc.Bits(V.bits, n))
                    ob, om = dev.ew_binary(gpu_ctx, code[op], L, R, DV, None, mnr.MaskMode.And)
                    assert _same(ob.download(), exp, op, lhs, rhs) and np.array_equal(om.download().bits, em.bits), (knob, v, dt, op, "masked")
                    # dense (non-zero divisors for the integer kernels: a zero there is DivideByZero) + an unaligned window
                    rhs2, R2 = (e, E) if op == "pow" else ((b, B) if is_f else (bnz, BNZ))
                    exp, _ = apply(lhs, rhs2, code[op], None)
                    ob, om = dev.ew_binary(gpu_ctx, code[op], L, R2, None, None, mnr.MaskMode.And)
                    assert om is None and _same(ob.download(), exp, op, lhs, rhs2), (knob, v, dt, op, "dense")
                    exp, _ = apply(lhs[3:n - 2], rhs2[3:n - 2], code[op], None)
                    ob, _ = dev.ew_binary(gpu_ctx, code[op], L.slice(3, n - 5), R2.slice(3, n - 5), None, None, mnr.MaskMode.And)
                    assert _same(ob.download(), exp, op, lhs[3:n - 2], rhs2[3:n - 2]), (knob, v, dt, op, "unaligned")
    finally:
        gpu_ctx.set_option(knob, 0)


@pytest.mark.parametrize("dt", [np.int8, np.uint8])
def test_every_geometry_of_masked_1byte_add_mul_matches_the_oracle(gpu_ctx, dt):
    """ew_cheap8_cfg: masked add / sub / mul on 1-byte columns, all three geometries, two masks (AND / OR),
    one mask on either side, and windows at odd row offsets (narrower vector tiers, shifted validity bits)."""
    import minarrow_b200 as mnr
    dev = mnr.device_ops
    rng = np.random.default_rng(33)
    n = 400_009
    info = np.iinfo(dt)
    a = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)
    b = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)
    va, vb = rng.random(n) < 0.8, rng.random(n) < 0.7
    A, B = mnr.DeviceBuffer.upload(gpu_ctx, a), mnr.DeviceBuffer.upload(gpu_ctx, b)
    up = lambda v: mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(v))     # a window's validity is re-based to bit 0
    try:
        for v in (1, 2, 3):
            gpu_ctx.set_option("ew_cheap8_cfg", v)
            for op in (orc.ADD, orc.SUB, orc.MUL):
                for lo, cnt in ((0, n), (64, n - 64), (5, n - 11), (129, 70_001)):
                    sa, sb = a[lo:lo + cnt], b[lo:lo + cnt]
                    ma, mb = orc.Bits.from_bools(va[lo:lo + cnt]), orc.Bits.from_bools(vb[lo:lo + cnt])
                    VA, VB = up(va[lo:lo + cnt]), up(vb[lo:lo + cnt])
                    for (l, r, mode) in ((ma, mb, "and"), (ma, mb, "or"), (ma, None, "and"), (None, mb, "and")):
                        if mode == "or":
                            m = orc.Bits.from_bools(va[lo:lo + cnt] | vb[lo:lo + cnt])
                        else:
                            m = orc.Bits.from_bools(va[lo:lo + cnt] & vb[lo:lo + cnt]) if (l is not None and r is not None) else (l if l is not None else r)
                        exp, em = orc.apply_int(sa, sb, op, m)
                        ob, om = dev.ew_binary(gpu_ctx, op, A.slice(lo, cnt), B.slice(lo, cnt),
                                               VA if l is not None else None, VB if r is not None else None,
                                               mnr.MaskMode.Or if mode == "or" else mnr.MaskMode.And)
                        assert ob.download().tobytes() == exp.tobytes(), (v, op, lo, cnt, mode)
                        assert np.array_equal(om.download().bits, em.bits), (v, op, lo, cnt, mode)
    finally:
        gpu_ctx.set_option("ew_cheap8_cfg", 0)
