"""Row a11: container routes on DEVICE-RESIDENT containers (minarrow_b200/containers.py) and their host front-ends
(minarrow_b200/kernels/broadcast.py), against the oracle's restatement of the reference routes:

  * Array (op) SuperArray / SuperArray (op) Array with re-chunking and the union mask — src/kernels/broadcast/mod.rs:1351-1361,
    src/utils.rs:367-481.  The reference has no test with null masks for these arms, so parity is pinned on the oracle.
  * view variants — ArrayV / SuperArrayV / TableV (super_array.rs:255-470, table_view.rs:25-200, mod.rs:1362-1393).
  * chaining in HBM: table * table + table on a device-resident SuperTable, launches counted.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _chunks(rng, dt, lens, mask_pattern):
    """Host chunks (typed arrays); mask_pattern[k] True -> chunk k carries a validity mask."""
    import minarrow_b200 as mnr
    out = []
    for n, has in zip(lens, mask_pattern):
        if np.dtype(dt).kind == "f":
            d = (rng.standard_normal(n) * 10).astype(dt)
        else:
            d = rng.integers(-50, 50, n).astype(dt)
        out.append(mnr.core.make_array(d, mnr.Bitmask.from_bools(rng.random(n) < 0.8) if has else None))
    return out


def _orc_chunks(orc, chunks):
    return [(c.data, None if c.null_mask is None else orc.Bits(c.null_mask.bits, len(c))) for c in chunks]


def _same(got, exp_data, exp_mask, what):
    if exp_data.dtype.kind == "f":
        nan = np.isnan(exp_data)
        assert np.array_equal(nan, np.isnan(got.data)), what
        assert np.array_equal(exp_data[~nan].view(np.int64 if exp_data.itemsize == 8 else np.int32),
                              got.data[~nan].view(np.int64 if exp_data.itemsize == 8 else np.int32)), what
    else:
        assert got.data.tobytes() == exp_data.tobytes(), what
    if exp_mask is None:
        assert got.null_mask is None, what
    else:
        assert got.null_mask is not None and got.null_mask.len == exp_mask.len and np.array_equal(got.null_mask.bits, exp_mask.bits), what


LENS = [37, 64, 1, 1003, 8, 250]          # ragged: chunk starts fall on every bit offset


@pytest.mark.parametrize("dt", [np.int32, np.int64, np.float64, np.uint32])
@pytest.mark.parametrize("array_mask", [False, True])
@pytest.mark.parametrize("pattern", ["none", "all", "some"])
def test_array_superarray_rechunk_route_matches_oracle(gpu_ctx, oracle, dt, array_mask, pattern):
    import minarrow_b200 as mnr
    from minarrow_b200.kernels import broadcast as B
    import zlib
    rng = np.random.default_rng(zlib.crc32(repr((np.dtype(dt).name, array_mask, pattern)).encode()))
    pat = {"none": [False] * 6, "all": [True] * 6, "some": [True, False, True, False, False, True]}[pattern]
    sa = B.SuperArray(_chunks(rng, dt, LENS, pat))
    arr = _chunks(rng, dt, [sum(LENS)], [array_mask])[0]
    am = None if arr.null_mask is None else oracle.Bits(arr.null_mask.bits, len(arr))
    A = mnr.ArithmeticOperator
    ops = [(A.Add, oracle.ADD), (A.Multiply, oracle.MUL)]
    if array_mask or pattern == "all":      # masked everywhere: integer zero divisors become nulls instead of the dense-route error
        ops.append((A.Divide, oracle.DIV))
    for op, oop in ops:
        for array_is_lhs in (True, False):
            exp = oracle.broadcast_array_superarray(oop, arr.data, am, _orc_chunks(oracle, sa.chunks), array_is_lhs)
            got = B.broadcast_value(op, arr, sa, gpu_ctx) if array_is_lhs else B.broadcast_value(op, sa, arr, gpu_ctx)
            assert isinstance(got, B.SuperArray) and got.shape_1d() == LENS
            for k, (g, (ed, em)) in enumerate(zip(got.chunks, exp)):
                _same(g, ed, em, (np.dtype(dt).name, array_mask, pattern, int(op), array_is_lhs, k))
    # the standalone re-chunk: same chunk lengths, values are windows, masks are windows of the FULL union mask
    aligned = B.create_aligned_chunks_from_array(arr, sa, gpu_ctx)
    exp = oracle.create_aligned_chunks_from_array(arr.data, am, _orc_chunks(oracle, sa.chunks))
    for g, (ed, em) in zip(aligned.chunks, exp):
        _same(g, ed, em, "create_aligned_chunks_from_array")
    # SuperArrayView arm (mod.rs:1362-1375): slices of a bigger array materialise to the same chunks
    big = _chunks(rng, dt, [sum(LENS) + 11], [True])[0]
    sav = B.SuperArrayV([B.ArrayV(big, 5, 100), B.ArrayV(big, 108, sum(LENS) - 100)])
    sub = [B._window(big, 5, 100), B._window(big, 108, sum(LENS) - 100)]
    exp = oracle.broadcast_array_superarray(oracle.ADD, arr.data, am, _orc_chunks(oracle, sub), True)
    got = B.broadcast_value(A.Add, arr, sav, gpu_ctx)
    for g, (ed, em) in zip(got.chunks, exp):
        _same(g, ed, em, "Array + SuperArrayView")


def test_length_and_mask_mismatch_errors(gpu_ctx):
    import minarrow_b200 as mnr
    from minarrow_b200.kernels import broadcast as B
    rng = np.random.default_rng(3)
    sa = B.SuperArray(_chunks(rng, np.int32, [10, 20], [True, True]))
    with pytest.raises(mnr.ShapeError, match="same total length"):
        B.broadcast_value(mnr.ArithmeticOperator.Add, np.arange(31, dtype=np.int32), sa, gpu_ctx)
    zeros = B.SuperArray([mnr.IntegerArray(np.ones(10, np.int32)), mnr.IntegerArray(np.zeros(20, np.int32))])
    with pytest.raises(mnr.KernelError) as e:      # dense integer zero divisor through the no-mask ArrayView route = the reference's panic
        B.broadcast_value(mnr.ArithmeticOperator.Divide, B.ArrayV(np.ones(30, dtype=np.int32)), zeros, gpu_ctx)
    assert e.value.kind == "DivideByZero"


def test_view_routes(gpu_ctx, oracle):
    """ArrayV (op) SuperArray (no mask, super_array.rs:255-365), TableV (op) TableV (table_view.rs:25-60),
    Array (op) TableV (mod.rs:1386-1393), TableV (op) SuperArrayV (table_view.rs:148-200)."""
    import minarrow_b200 as mnr
    from minarrow_b200 import containers as dc
    from minarrow_b200.kernels import broadcast as B
    A = mnr.ArithmeticOperator
    rng = np.random.default_rng(11)
    big = rng.integers(-1000, 1000, 5000).astype(np.int64)
    sa = B.SuperArray(_chunks(rng, np.int64, LENS, [True] * 6))
    view = B.ArrayV(mnr.IntegerArray(big, mnr.Bitmask.from_bools(rng.random(5000) < 0.5)), 123, sum(LENS))
    got = B.broadcast_value(A.Subtract, view, sa, gpu_ctx)
    start = 123
    for g, c in zip(got.chunks, sa.chunks):      # Array-level route: no mask at all, the operands' own masks are ignored
        assert g.null_mask is None and np.array_equal(g.data, big[start:start + len(c)] - c.data)
        start += len(c)
    got = B.broadcast_value(A.Subtract, sa, view, gpu_ctx)
    assert np.array_equal(got.chunks[3].data, sa.chunks[3].data - big[123 + 102:123 + 102 + 1003])
    # tables
    t1 = B.Table("t1", [mnr.IntegerArray(rng.integers(-9, 9, 400).astype(np.int32)), mnr.FloatArray(rng.standard_normal(400))])
    t2 = B.Table("t2", [mnr.IntegerArray(rng.integers(1, 9, 400).astype(np.int32)), mnr.FloatArray(rng.standard_normal(400))])
    r = B.broadcast_value(A.Multiply, B.TableV(t1, 17, 300), B.TableV(t2, 50, 300), gpu_ctx)
    assert r.name == "" and np.array_equal(r.cols[0].data, t1.cols[0].data[17:317] * t2.cols[0].data[50:350])
    assert np.array_equal(r.cols[1].data, t1.cols[1].data[17:317] * t2.cols[1].data[50:350])
    with pytest.raises(mnr.ShapeError, match="column count mismatch"):
        B.broadcast_value(A.Add, B.TableV(t1, 0, 10), B.TableV(B.Table("x", t2.cols[:1]), 0, 10), gpu_ctx)
    arr = rng.integers(-5, 5, 400).astype(np.int32)
    ti = B.Table("ti", [t1.cols[0], t2.cols[0]])
    r = B.broadcast_value(A.Add, arr, B.TableV(ti, 100, 50), gpu_ctx)          # the table view's window of the array
    assert np.array_equal(r.cols[0].data, arr[100:150] + t1.cols[0].data[100:150])
    assert np.array_equal(r.cols[1].data, arr[100:150] + t2.cols[0].data[100:150])
    # TableView (op) SuperArrayView -> one result table per slice (device-resident form)
    dt_ = dc.DeviceTable.from_host(gpu_ctx, ti).view(20, 300)
    col = dc.DeviceArray.from_host(gpu_ctx, rng.integers(-5, 5, 1000).astype(np.int32))
    sav = dc.DeviceSuperArray.from_slices([col.view(0, 100), col.view(500, 200)])
    l0 = gpu_ctx.launch_count
    st = dc.broadcast_value(A.Add, dt_, sav, gpu_ctx)
    assert gpu_ctx.launch_count - l0 <= 3          # 2 slices x 2 columns in at most one launch per alignment class
    h = st.to_host()
    cv = col.to_host().data
    assert h.n_batches() == 2 and np.array_equal(h.batches[1].cols[1].data, t2.cols[0].data[120:320] + cv[500:700])


def test_device_resident_supertable_chains_in_hbm(gpu_ctx, oracle):
    """configs[4] in small, never leaving the device: (A * B) + A on a 6-batch x 4-column SuperTable = 2 route calls,
    at most 4 launches each (one per column dtype); per-column sum/min/max/count of the result in ONE more call."""
    import minarrow_b200 as mnr
    from minarrow_b200 import containers as dc
    from minarrow_b200.kernels import broadcast as B
    A = mnr.ArithmeticOperator
    rng = np.random.default_rng(21)
    dts = [np.int32, np.int64, np.float32, np.float64]

    def table(i):
        cols = []
        for dt in dts:
            n = 4096      # equal aligned lengths: every (dtype) class is one launch
            d = (rng.standard_normal(n)).astype(dt) if np.dtype(dt).kind == "f" else rng.integers(-100, 100, n).astype(dt)
            cols.append(mnr.core.make_array(d, None))
        return B.Table(f"b{i}", cols)
    ha = B.SuperTable([table(i) for i in range(6)], "A")
    hb = B.SuperTable([table(i) for i in range(6)], "B")
    da, db = dc.DeviceSuperTable.from_host(gpu_ctx, ha), dc.DeviceSuperTable.from_host(gpu_ctx, hb)
    gpu_ctx.synchronize()
    l0 = gpu_ctx.launch_count
    prod = dc.broadcast_value(A.Multiply, da, db, gpu_ctx)
    l1 = gpu_ctx.launch_count
    res = dc.broadcast_value(A.Add, prod, da, gpu_ctx)
    l2 = gpu_ctx.launch_count
    assert l1 - l0 <= 4 and l2 - l1 <= 4, (l1 - l0, l2 - l1)
    stats = dc.super_table_stats(res)
    l3 = gpu_ctx.launch_count
    assert l3 - l2 <= 4
    out = res.to_host()
    assert out.name == "A" and out.n_batches() == 6 and out.batches[2].name == "b2"
    for c, dt in enumerate(dts):
        whole = []
        for k in range(6):
            a, b = ha.batches[k].cols[c].data, hb.batches[k].cols[c].data
            e = (a * b + a).astype(dt)     # numpy: wrapping integer arithmetic, single IEEE operations
            assert out.batches[k].cols[c].data.tobytes() == e.tobytes(), (dt, k)
            assert out.batches[k].cols[c].null_mask is None
            whole.append(e)
        exp = oracle.stats(np.concatenate(whole), None)
        assert stats[c]["count"] == exp["count"] and stats[c]["min"] == exp["min"] and stats[c]["max"] == exp["max"]
        if np.dtype(dt).kind == "f":
            assert abs(stats[c]["sum"] - exp["sum"]) <= 1e-12 * np.abs(np.concatenate(whole).astype(np.float64)).sum()
        else:
            assert stats[c]["sum"] == exp["sum"]
    # scalar over the whole SuperTable: one typed scalar per column (table.rs:230-261), int columns and float columns
    r = dc.broadcast_value(A.Multiply, da, 3, gpu_ctx).to_host()
    assert np.array_equal(r.batches[5].cols[1].data, ha.batches[5].cols[1].data * 3)
    assert np.array_equal(r.batches[0].cols[3].data, ha.batches[0].cols[3].data * 3.0)


def test_super_array_stats_and_masked_route_on_device(gpu_ctx, oracle):
    import minarrow_b200 as mnr
    from minarrow_b200 import containers as dc
    from minarrow_b200.kernels import broadcast as B
    rng = np.random.default_rng(5)
    for dt in (np.int64, np.float32):
        l = B.SuperArray(_chunks(rng, dt, LENS, [True, False, True, True, False, True]))
        r = B.SuperArray(_chunks(rng, dt, LENS, [True, True, False, True, False, False]))
        dl, dr = dc.DeviceSuperArray.from_host(gpu_ctx, l), dc.DeviceSuperArray.from_host(gpu_ctx, r)
        got = dc.route_super_array_broadcast(mnr.ArithmeticOperator.Add, dl, dr)
        exp = oracle.route_super_array_broadcast(oracle.ADD, _orc_chunks(oracle, l.chunks), _orc_chunks(oracle, r.chunks))
        for g, (ed, em) in zip(got.to_host().chunks, exp):
            _same(g, ed, em, "route_super_array_broadcast")
        s = dc.super_array_stats(got)
        whole = np.concatenate([ed for ed, _ in exp])
        valid = np.concatenate([np.ones(len(ed), bool) if em is None else em.to_bools() for ed, em in exp])
        e = oracle.stats(whole, oracle.Bits.from_bools(valid))
        assert s["count"] == e["count"] and s["min"] == e["min"] and s["max"] == e["max"]
        assert s["sum"] == e["sum"] or abs(s["sum"] - e["sum"]) <= 1e-12 * np.abs(whole[valid].astype(np.float64)).sum()


def test_boolean_array_not(gpu_ctx):
    """BooleanArray `!` (boolean.rs:853-866): data inverted, validity unchanged, slack bits zero."""
    import minarrow_b200 as mnr
    rng = np.random.default_rng(9)
    for n in (1, 7, 64, 1001):
        d, v = rng.random(n) < 0.5, rng.random(n) < 0.8
        b = mnr.BooleanArray(mnr.Bitmask.from_bools(d), mnr.Bitmask.from_bools(v))
        r = ~b
        assert np.array_equal(r.data.to_bools(), ~d) and r.null_mask is b.null_mask
        assert np.array_equal(r.data.bits, np.packbits(~d, bitorder="little"))


def test_array_supertable_tuple_and_vecvalue_arms(gpu_ctx, oracle):
    """The compositional arms of `broadcast_value` (broadcast/mod.rs): Array (op) SuperTable and the mirror (:557-562, one
    batched call), Tuple2..6 and VecValue element-wise (:305-372), an array against every element of a tuple (:766-907), and
    Scalar (op) Scalar, which the reference always evaluates as Add (:161-163)."""
    import minarrow_b200 as mnr
    from minarrow_b200 import containers as dc
    from minarrow_b200.kernels import broadcast as B
    A = mnr.ArithmeticOperator
    rng = np.random.default_rng(5)
    n = 1000
    dt = np.int64

    def col(mask):
        return mnr.core.make_array(rng.integers(-1000, 1000, n).astype(dt), mnr.Bitmask.from_bools(rng.random(n) < 0.8) if mask else None)
    st = B.SuperTable([B.Table(f"b{i}", [col(False), col(True)]) for i in range(3)], "T")
    arr = col(True)
    dst, darr = dc.DeviceSuperTable.from_host(gpu_ctx, st), dc.DeviceArray.from_host(gpu_ctx, arr)
    for array_is_lhs in (True, False):
        l0 = gpu_ctx.launch_count
        got = (dc.broadcast_value(A.Subtract, darr, dst, gpu_ctx) if array_is_lhs else dc.broadcast_value(A.Subtract, dst, darr, gpu_ctx))
        assert gpu_ctx.launch_count - l0 == 1        # six leaves of one (dtype, class) in one launch
        got = got.to_host()
        assert got.name == "T" and got.n_batches() == 3
        for b in range(3):
            for c in range(2):
                # broadcast_array_to_table -> resolve_binary_arithmetic(op, array, col, None): the kernels get no mask, the
                # operands' own validity is not consulted and the result carries none (routing/arithmetic.rs:214-222)
                h = st.batches[b].cols[c]
                ed, em = oracle.apply_int(arr.data, h.data, oracle.SUB, None) if array_is_lhs else oracle.apply_int(h.data, arr.data, oracle.SUB, None)
                _same(got.batches[b].cols[c], ed, em, (array_is_lhs, b, c))
    # tuples / lists
    x, y = dc.DeviceArray.from_host(gpu_ctx, col(False)), dc.DeviceArray.from_host(gpu_ctx, col(False))
    t = dc.broadcast_value(A.Add, (x, y, 2), (y, x, 3), gpu_ctx)
    assert isinstance(t, tuple) and len(t) == 3 and t[2] == 5
    assert np.array_equal(t[0].to_host().data, x.to_host().data + y.to_host().data)
    v = dc.broadcast_value(A.Multiply, [x, y], [y, y], gpu_ctx)
    assert isinstance(v, list) and np.array_equal(v[1].to_host().data, y.to_host().data * y.to_host().data)
    with pytest.raises(mnr.KernelError):
        dc.broadcast_value(A.Multiply, [x, y], [y], gpu_ctx)
    at = dc.broadcast_value(A.Add, x, (y, x), gpu_ctx)
    assert isinstance(at, tuple) and np.array_equal(at[1].to_host().data, 2 * x.to_host().data)
