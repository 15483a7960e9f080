"""Randomised shapes / offsets / dtypes / operators against the oracle (hypothesis).  The reference has no property tests
(SURVEY §4); these search the space between the fixed sizes of test_gpu_parity.py: odd lengths, view offsets that pick
every alignment tier, sparse / dense / empty validity, every operator, both mask-merge modes, window offsets of the
bitmask kernels."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
DTS = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64]
COMMON = dict(deadline=None, max_examples=60, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow],
              derandomize=True)


def _data(rng, dt, n, small=False):
    if np.dtype(dt).kind == "f":
        a = (rng.standard_normal(n) * 50).astype(dt)
        if n:
            a[rng.integers(0, n, max(1, n // 9))] = rng.choice(np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0], dtype=dt), max(1, n // 9))
        return a
    info = np.iinfo(dt)
    lo, hi = (max(info.min, -6), min(info.max, 6)) if small else (info.min, info.max)
    return rng.integers(lo, hi, n, dtype=dt, endpoint=True)


@settings(**COMMON)
@given(n=st.integers(0, 6000), dti=st.integers(0, len(DTS) - 1), op=st.integers(0, 6), oa=st.integers(0, 9), ob=st.integers(0, 9),
       masks=st.sampled_from(["none", "lhs", "both_and", "both_or"]), p=st.sampled_from([0.0, 0.5, 0.9, 1.0]), seed=st.integers(0, 2 ** 16))
def test_elementwise_random_windows(gpu_ctx, n, dti, op, oa, ob, masks, p, seed):
    import minarrow_b200 as mnr
    dev = mnr.device_ops
    dt = DTS[dti]
    is_f = np.dtype(dt).kind == "f"
    rng = np.random.default_rng(seed)
    a, b = _data(rng, dt, n + 10), _data(rng, dt, n + 10, small=True)
    A, B = mnr.DeviceBuffer.upload(gpu_ctx, a), mnr.DeviceBuffer.upload(gpu_ctx, b)
    la, lb = rng.random(n) < p, rng.random(n) < p
    LA = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(la))
    LB = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(lb))
    ha, hb = a[oa:oa + n], b[ob:ob + n]
    if masks == "none":
        lm = rm = None; merged = None
    elif masks == "lhs":
        lm, rm, merged = LA, None, la
    else:
        lm, rm = LA, LB
        merged = (la & lb) if masks == "both_and" else (la | lb)
    mode = mnr.MaskMode.Or if masks == "both_or" else mnr.MaskMode.And
    try:
        exp, em = orc.apply(ha, hb, op, None if merged is None else orc.Bits.from_bools(merged))
    except orc.KernelError as e:
        assert e.kind == "DivideByZero" and merged is None
        with pytest.raises(mnr.KernelError) as ei:
            dev.ew_binary(gpu_ctx, op, A.slice(oa, n), B.slice(ob, n), lm, rm, mode)
        assert ei.value.kind == "DivideByZero"
        return
    got, gm = dev.ew_binary(gpu_ctx, op, A.slice(oa, n), B.slice(ob, n), lm, rm, mode)
    g = got.download()
    if is_f and op == orc.POW:
        # the reference's own tolerance (1e-12 / 1e-6 relative); f32 additionally gets the conditioning of exp() at
        # x = b ln a — see tests/test_gpu_parity.py::same_float for the derivation
        from test_gpu_parity import same_float
        with np.errstate(all="ignore"):
            assert same_float(g, exp, orc.POW, dt, np.abs(ha), hb), (dt, n, oa, ob, masks)
    elif is_f:
        # NaN is compared by position, everything else bit for bit: which NaN payload an operation returns is not part of
        # the reference's contract (Rust leaves NaN bit patterns unspecified; x86 propagates an input payload, the GPU
        # returns the canonical quiet NaN) — the same rule as tests/test_gpu_parity.py::same_float.
        nan = np.isnan(exp)
        assert np.array_equal(np.isnan(g), nan), (dt, op, n, oa, ob, masks)
        assert g[~nan].tobytes() == exp[~nan].tobytes(), (dt, op, n, oa, ob, masks)
    else:
        assert g.tobytes() == exp.tobytes(), (dt, op, n, oa, ob, masks)
    assert (gm is None) == (em is None)
    if em is not None:
        assert np.array_equal(gm.download().bits, em.bits)


@settings(**COMMON)
@given(n=st.integers(0, 20000), dti=st.integers(0, len(DTS) - 1), off=st.integers(0, 17), p=st.sampled_from([None, 0.0, 0.3, 0.95, 1.0]),
       seed=st.integers(0, 2 ** 16))
def test_reductions_random_windows(gpu_ctx, n, dti, off, p, seed):
    import minarrow_b200 as mnr
    dev = mnr.device_ops
    dt = DTS[dti]
    rng = np.random.default_rng(seed)
    a = _data(rng, dt, n + 20)
    if np.dtype(dt).kind == "f":
        a[~np.isfinite(a)] = dt(3.0)
    A = mnr.DeviceBuffer.upload(gpu_ctx, a)
    v = None if p is None else rng.random(n) < p
    V = None if v is None else mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(v))
    h = a[off:off + n]
    ex = orc.stats(h, None if v is None else orc.Bits.from_bools(v))
    for got in (dev.reduce_stats(gpu_ctx, A.slice(off, n), V), dev.reduce_stats_batch(gpu_ctx, [A.slice(off, n)], [V], True)[0]):
        assert got["count"] == ex["count"]
        same = lambda x, y: x == y or (x != x and y != y)   # noqa: E731
        assert same(got["min"], ex["min"]) and same(got["max"], ex["max"]), (dt, n, off, p)
        if np.dtype(dt).kind == "f":
            sel = h if v is None else h[v]
            assert abs(got["sum"] - ex["sum"]) <= 1e-12 * max(1.0, float(np.abs(sel.astype(np.float64)).sum()))
        else:
            assert got["sum"] == ex["sum"]


@settings(**COMMON)
@given(nbits=st.integers(1, 40000), ao=st.integers(0, 300), bo=st.integers(0, 300), op=st.integers(0, 2), seed=st.integers(0, 2 ** 16))
def test_bitmask_random_windows(gpu_ctx, nbits, ao, bo, op, seed):
    import minarrow_b200 as mnr
    dev = mnr.device_ops
    rng = np.random.default_rng(seed)
    x, y = rng.random(nbits + 400) < 0.5, rng.random(nbits + 400) < 0.5
    X = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(x))
    Y = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(y))
    ox, oy = orc.Bits.from_bools(x), orc.Bits.from_bools(y)
    got = dev.bits_binop(gpu_ctx, op, X, ao, Y, bo, nbits).download()
    exp = orc.bitmask_binop((ox, ao, nbits), (oy, bo, nbits), op)          # byte-floored offsets, like the reference
    assert got.len == exp.len and np.array_equal(got.bits, exp.bits)
    assert np.array_equal(dev.bits_not(gpu_ctx, X, ao, nbits).download().bits, orc.not_mask((ox, ao, nbits)).bits)
    assert np.array_equal(dev.bits_slice(gpu_ctx, X, ao, nbits).download().bits, np.packbits(x[ao:ao + nbits], bitorder="little"))
    w = (ao // 64) * 64
    assert dev.bits_popcount(gpu_ctx, X, w, nbits) == orc.popcount_mask((ox, w, nbits))
    cnt = mnr.DeviceBuffer.alloc(gpu_ctx, np.uint64, 1)                   # asynchronous form: count stays on the device
    dev.bits_popcount_async(gpu_ctx, X, w, nbits, cnt.device_ptr)
    assert int(cnt.download()[0]) == orc.popcount_mask((ox, w, nbits))
