"""Datetime delegation through the product path (`apply_datetime_{i32,u32,i64,u64}`, dispatch.rs:300-372, 420-427): the
reference's own vectors (arithmetic/mod.rs:418-506) and seeded parity against the oracle — windows, one / two / no masks,
masks longer than the window, zero divisors under a mask and in the dense kernels."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_reference_vectors(gpu_ctx):
    import minarrow_b200 as mnr
    from minarrow_b200.kernels import arithmetic as ar
    A = mnr.ArithmeticOperator
    DA = mnr.DatetimeArray
    l, r = DA.from_slice(np.array([1000, 2000, 3000], np.int64)), DA.from_slice(np.array([10, 20, 30], np.int64))
    out = ar.apply_datetime_i64((l, 0, 3), (r, 0, 3), A.Add, gpu_ctx)
    assert out.data.tolist() == [1010, 2020, 3030] and out.null_mask is None
    a, b = DA.from_slice(np.array([10, 20, 30, 40], np.int64), "Milliseconds"), DA.from_slice(np.array([1, 2, 3, 4], np.int64))
    exp = {A.Add: [11, 22, 33, 44], A.Subtract: [9, 18, 27, 36], A.Multiply: [10, 40, 90, 160], A.Divide: [10, 10, 10, 10],
           A.Remainder: [0, 0, 0, 0], A.Power: [10, 400, 27000, 2560000]}
    for op, e in exp.items():
        out = ar.apply_datetime_i64((a, 0, 4), (b, 0, 4), op, gpu_ctx)
        assert out.data.tolist() == e and out.null_mask is None and out.time_unit == "Milliseconds", op
    am = DA(a.data, mnr.Bitmask.from_bools([True, False, True, True]))
    out = ar.apply_datetime_i64((am, 0, 4), (b, 0, 4), A.Add, gpu_ctx)
    assert out.data.tolist() == [11, 0, 33, 44] and out.null_mask.to_bools().tolist() == [True, False, True, True]
    e = DA.from_slice(np.zeros(0, np.int64))
    assert ar.apply_datetime_i64((e, 0, 0), (e, 0, 0), A.Add, gpu_ctx).is_empty()
    with pytest.raises(mnr.KernelError) as ei:
        ar.apply_datetime_i64((DA.from_slice(np.array([1000, 2000], np.int64)), 0, 2), (DA.from_slice(np.array([10], np.int64)), 0, 1), A.Add, gpu_ctx)
    assert ei.value.kind == "LengthMismatch"


@pytest.mark.parametrize("dt", [np.int32, np.uint32, np.int64, np.uint64])
def test_parity_with_the_oracle(gpu_ctx, dt):
    import minarrow_b200 as mnr
    from minarrow_b200.kernels import arithmetic as ar
    fn = {np.int32: ar.apply_datetime_i32, np.uint32: ar.apply_datetime_u32, np.int64: ar.apply_datetime_i64, np.uint64: ar.apply_datetime_u64}[dt]
    rng = np.random.default_rng(41)
    n = 70_003
    info = np.iinfo(dt)
    a = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)
    b = rng.integers(max(info.min, -1000), 1000, n).astype(dt)
    b[::17] = 0
    va, vb = rng.random(n) < 0.85, rng.random(n) < 0.9
    for lo, ro, ln in ((0, 0, n), (5, 11, 60_001), (64, 1, 1), (1000, 2000, 4097)):
        for ml, mr in ((va, vb), (va, None), (None, vb), (None, None)):
            L = mnr.DatetimeArray(a, None if ml is None else mnr.Bitmask.from_bools(ml))
            R = mnr.DatetimeArray(b, None if mr is None else mnr.Bitmask.from_bools(mr))
            om_l = None if ml is None else orc.Bits.from_bools(ml)
            om_r = None if mr is None else orc.Bits.from_bools(mr)
            for op in (orc.ADD, orc.SUB, orc.MUL, orc.DIV, orc.REM, orc.FLOORDIV):
                dense_div = ml is None and mr is None and op in (orc.DIV, orc.REM, orc.FLOORDIV)
                if dense_div and (b[ro:ro + ln] == 0).any():
                    with pytest.raises(mnr.KernelError) as ei:
                        fn((L, lo, ln), (R, ro, ln), op, gpu_ctx)
                    assert ei.value.kind == "DivideByZero"
                    continue
                exp, em = orc.apply_datetime(a, om_l, lo, ln, b, om_r, ro, ln, op)
                got = fn((L, lo, ln), (R, ro, ln), op, gpu_ctx)
                assert got.data.tobytes() == exp.tobytes(), (dt, lo, ro, ln, op)
                assert (got.null_mask is None) == (em is None)
                if em is not None:
                    assert got.null_mask.len == ln and np.array_equal(got.null_mask.bits, em.bits), (dt, lo, ro, ln, op)
