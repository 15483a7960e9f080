"""Integer Div / Rem / FloorDiv by a broadcast scalar: the kernels evaluate these with a host-computed multiplicative
inverse (minarrow_b200/csrc/divmagic.h) instead of a divide.  Results must equal the reference route bit for bit —
`broadcast_length_1_array` (routing/broadcast.rs:25-47) materialises the scalar, then int_dense_body / int_masked_body
divide row by row (src/kernels/arithmetic/std.rs:54-77,96-136) — for every divisor, including the ones with special
inverses (±1, powers of two, MIN, MAX) and the wrapping MIN / -1."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
INT_DT = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64]


def _edge_column(rng, dt, n):
    info = np.iinfo(dt)
    a = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)
    edges = np.array([info.min, info.max, 0, 1, info.max // 2, info.max // 2 + 1] + ([-1, info.min + 1, -2] if info.min < 0 else [2, 3]),
                     dtype=dt)
    idx = rng.integers(0, n, max(1, n // 8))
    a[idx] = edges[rng.integers(0, len(edges), idx.size)]
    # a run of small values around zero: quotient sign / rounding direction
    k = min(n, 257)
    a[:k] = (np.arange(k) - (k // 2 if info.min < 0 else 0)).astype(dt)
    return a


def _scalars(dt):
    info = np.iinfo(dt)
    s = [1, 2, 3, 5, 7, 10, 16, 100, info.max, info.max - 1, info.max // 2 + 1, info.max // 3]
    if info.min < 0:
        s += [-1, -2, -3, -7, -16, info.min, info.min + 1, -(info.max // 2 + 1)]
    return [dt(x) for x in s]


@pytest.mark.parametrize("dt", INT_DT)
def test_scalar_divisor_matches_reference_route(gpu_ctx, dt):
    import minarrow_b200 as mnr
    dev = mnr.device_ops
    rng = np.random.default_rng(5)
    for n in (1, 67, 4099, 70_001):
        a = _edge_column(rng, dt, n)
        valid = rng.random(n) < 0.8
        A = mnr.DeviceBuffer.upload(gpu_ctx, a)
        V = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(valid))
        for s in _scalars(dt):
            full = np.full(n, s, dtype=dt)
            for op in (orc.DIV, orc.REM, orc.FLOORDIV):
                exp, em = orc.apply_int(a, full, op, orc.Bits.from_bools(valid))
                ob, om = dev.ew_scalar(gpu_ctx, op, A, s, False, V)
                assert ob.download().tobytes() == exp.tobytes(), (dt, n, s, op, "masked")
                assert np.array_equal(om.download().bits, em.bits)
                exp, _ = orc.apply_int(a, full, op, None)
                ob, om = dev.ew_scalar(gpu_ctx, op, A, s, False, None)
                assert om is None and ob.download().tobytes() == exp.tobytes(), (dt, n, s, op, "dense")
        # unaligned window (element-load tier) through the same inverse
        if n > 16:
            s = dt(7)
            exp, _ = orc.apply_int(a[3:n - 2], np.full(n - 5, s, dtype=dt), orc.DIV, None)
            ob, _ = dev.ew_scalar(gpu_ctx, orc.DIV, A.slice(3, n - 5), s, False, None)
            assert ob.download().tobytes() == exp.tobytes()


@pytest.mark.parametrize("dt", [np.int32, np.int64, np.uint16])
def test_zero_scalar_divisor_keeps_reference_behaviour(gpu_ctx, dt):
    """Masked: every row null, value 0 (std.rs:96-136).  Dense: the reference panics -> DivideByZero."""
    import minarrow_b200 as mnr
    dev = mnr.device_ops
    rng = np.random.default_rng(6)
    n = 1000
    a = _edge_column(rng, dt, n)
    valid = rng.random(n) < 0.8
    A = mnr.DeviceBuffer.upload(gpu_ctx, a)
    V = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(valid))
    for op in (orc.DIV, orc.REM, orc.FLOORDIV):
        ob, om = dev.ew_scalar(gpu_ctx, op, A, dt(0), False, V)
        assert not ob.download().any() and not om.download().bits.any()
        with pytest.raises(mnr.KernelError) as ei:
            dev.ew_scalar(gpu_ctx, op, A, dt(0), False, None)
        assert ei.value.kind == "DivideByZero"


def test_scalar_divisor_batch_route(gpu_ctx):
    """SuperTable / scalar route (mnr_ew_scalar_batch_into): per-chunk scalars, mixed dtypes, one launch per class."""
    import minarrow_b200 as mnr
    dev = mnr.device_ops
    rng = np.random.default_rng(7)
    cols, scal, exp, masks = [], [], [], []
    for dt, s in ((np.int32, 3), (np.int64, -7), (np.int32, 1000), (np.uint64, 86_400), (np.int64, 0), (np.int16, -3)):
        n = int(rng.integers(100, 9000))
        a = _edge_column(rng, dt, n)
        v = rng.random(n) < 0.9
        cols.append(a); scal.append(dt(s)); masks.append(v)
        exp.append(orc.apply_int(a, np.full(n, s, dtype=dt), orc.FLOORDIV, orc.Bits.from_bools(v)))
    A = [mnr.DeviceBuffer.upload(gpu_ctx, a) for a in cols]
    V = [mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(v)) for v in masks]
    O = [mnr.DeviceBuffer.alloc(gpu_ctx, a.dtype, len(a)) for a in cols]
    OM = [mnr.DeviceBitmask.alloc(gpu_ctx, len(a)) for a in cols]
    dev.ew_scalar_batch_into(gpu_ctx, orc.FLOORDIV, A, scal, False, V, O, OM)
    for o, om, (ed, em) in zip(O, OM, exp):
        assert o.download().tobytes() == ed.tobytes()
        assert np.array_equal(om.download().bits, em.bits)
