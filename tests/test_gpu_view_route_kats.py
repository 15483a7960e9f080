"""The reference's own tests of the view / cross-hierarchy broadcast arms, transcribed and run through the product path
(host front-end minarrow_b200/kernels/broadcast.py -> device routes -> batched CUDA launches):

  src/kernels/broadcast/array_view.rs:170-411        ArrayView (op) Table / TableView / SuperTableView
  src/kernels/broadcast/table_view.rs:200-480        TableView (op) TableView / Scalar / ArrayView / SuperArrayView
  src/kernels/broadcast/super_array_view.rs:89-160   SuperArrayView (op) TableView
  src/kernels/broadcast/super_table_view.rs:256-570  SuperTableView (op) Scalar / ArrayView / Array, SuperArrayView (op) Table
  src/kernels/broadcast/scalar.rs:937-1370           Scalar (op) Table / TableView / Array / SuperArray(View) / SuperTable / TupleN
  src/kernels/broadcast/array.rs:626-1050            Array (op) SuperTableView / Scalar / SuperTable / Tuple2..6
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def B(gpu_ctx):
    from minarrow_b200.kernels import broadcast as b
    return b


def i32(*v):
    import minarrow_b200 as mnr
    return mnr.core.make_array(np.array(v, dtype=np.int32), None)


def table(B, *cols, name="test"):
    return B.Table(name, [i32(*c) for c in cols])


def col(t, c=0):
    x = t.cols[c]
    assert x.data.dtype == np.int32 and x.null_mask is None
    return x.data.tolist()


def stv(B, *tables):
    return B.SuperTableV([B.TableV(t, 0, t.n_rows()) for t in tables])


def test_array_view_rs(B, gpu_ctx):
    import minarrow_b200 as mnr
    A = mnr.ArithmeticOperator
    r = B.broadcast_value(A.Add, B.ArrayV(i32(1, 2, 3)), table(B, (10, 20, 30), (100, 200, 300)), gpu_ctx)
    assert r.n_rows() == 3 and r.n_cols() == 2 and col(r, 0) == [11, 22, 33] and col(r, 1) == [101, 202, 303]
    tv = B.TableV(table(B, (10, 10, 10)), 0, 3)
    r = B.broadcast_value(A.Multiply, B.ArrayV(i32(2, 3, 4)), tv, gpu_ctx)
    assert r.n_rows() == 3 and col(r) == [20, 30, 40]
    tv = B.TableV(table(B, (10, 20, 30), (100, 200, 300)), 0, 3)
    r = B.broadcast_value(A.Subtract, B.ArrayV(i32(5, 5, 5)), tv, gpu_ctx)
    assert col(r, 0) == [-5, -15, -25] and col(r, 1) == [-95, -195, -295]
    s = stv(B, table(B, (10, 20, 30)), table(B, (40, 50, 60)))
    r = B.broadcast_arrayview_to_supertableview(A.Add, B.ArrayV(i32(1, 2, 3, 4, 5, 6)), s, gpu_ctx)
    assert r.n_rows() == 6 and r.n_batches() == 2 and col(r.batches[0]) == [11, 22, 33] and col(r.batches[1]) == [44, 55, 66]
    with pytest.raises(mnr.ShapeError) as ei:
        B.broadcast_value(A.Add, B.ArrayV(i32(1, 2, 3, 4, 5)), s, gpu_ctx)
    assert "does not match" in str(ei.value)


def test_table_view_rs(B, gpu_ctx):
    import minarrow_b200 as mnr
    A = mnr.ArithmeticOperator
    tv1 = B.TableV(table(B, (1, 2, 3), (10, 20, 30)), 0, 3)
    tv2 = B.TableV(table(B, (5, 5, 5), (100, 100, 100)), 0, 3)
    r = B.broadcast_tableview_to_tableview(A.Add, tv1, tv2, gpu_ctx)
    assert r.n_rows() == 3 and r.n_cols() == 2 and col(r, 0) == [6, 7, 8] and col(r, 1) == [110, 120, 130]
    with pytest.raises(mnr.ShapeError) as ei:
        B.broadcast_tableview_to_tableview(A.Add, B.TableV(table(B, (1, 2, 3)), 0, 3), B.TableV(table(B, (5, 5, 5), (10, 10, 10)), 0, 3), gpu_ctx)
    assert "column count mismatch" in str(ei.value)
    r = B.broadcast_value(A.Multiply, B.TableV(table(B, (2, 3, 4), (5, 6, 7)), 0, 3), np.int32(10), gpu_ctx)
    assert col(r, 0) == [20, 30, 40] and col(r, 1) == [50, 60, 70]
    r = B.broadcast_value(A.Subtract, B.TableV(table(B, (100, 200, 300)), 0, 3), B.ArrayV(i32(10, 20, 30)), gpu_ctx)
    assert r.n_rows() == 3 and col(r) == [90, 180, 270]
    arr = i32(10, 20, 30, 40, 50, 60)
    sav = B.SuperArrayV([B.ArrayV(arr).slice(0, 3), B.ArrayV(arr).slice(3, 3)])
    r = B.broadcast_value(A.Multiply, B.TableV(table(B, (1, 2, 3, 4, 5, 6)), 0, 6), sav, gpu_ctx)
    assert r.n_rows() == 6 and r.n_batches() == 2 and col(r.batches[0]) == [10, 40, 90] and col(r.batches[1]) == [160, 250, 360]
    sa2 = B.SuperArrayV([B.ArrayV(i32(10, 20, 30)), B.ArrayV(i32(40, 50, 60))])
    with pytest.raises(mnr.ShapeError) as ei:
        B.broadcast_value(A.Add, B.TableV(table(B, (1, 2, 3, 4, 5)), 0, 5), sa2, gpu_ctx)
    assert "does not match" in str(ei.value)


def test_super_array_view_rs_and_super_table_view_rs(B, gpu_ctx):
    import minarrow_b200 as mnr
    A = mnr.ArithmeticOperator
    r = B.broadcast_value(A.Add, B.SuperArrayV([B.ArrayV(i32(1, 2, 3))]), B.TableV(table(B, (10, 20, 30)), 0, 3), gpu_ctx)
    assert r.n_rows() == 3 and r.n_batches() == 1 and col(r.batches[0]) == [11, 22, 33]
    with pytest.raises(mnr.ShapeError) as ei:
        B.broadcast_value(A.Add, B.SuperArrayV([B.ArrayV(i32(1, 2, 3))]), B.TableV(table(B, (10, 20, 30, 40, 50)), 0, 5), gpu_ctx)
    assert "does not match" in str(ei.value)
    # super_table_view.rs
    s = stv(B, table(B, (1, 2, 3)), table(B, (4, 5, 6)))
    r = B.broadcast_supertableview_to_scalar(A.Add, s, np.int32(10), gpu_ctx)
    assert r.n_rows() == 6 and r.n_batches() == 2 and col(r.batches[0]) == [11, 12, 13] and col(r.batches[1]) == [14, 15, 16]
    s = stv(B, table(B, (2, 3, 4)), table(B, (5, 6, 7)))
    r = B.broadcast_supertableview_to_arrayview(A.Multiply, s, B.ArrayV(i32(10, 10, 10, 10, 10, 10)), gpu_ctx)
    assert r.n_rows() == 6 and r.n_batches() == 2 and col(r.batches[0]) == [20, 30, 40] and col(r.batches[1]) == [50, 60, 70]
    with pytest.raises(mnr.ShapeError) as ei:
        B.broadcast_supertableview_to_arrayview(A.Add, stv(B, table(B, (1, 2, 3)), table(B, (4, 5, 6))), B.ArrayV(i32(10, 10, 10, 10, 10)), gpu_ctx)
    assert "does not match" in str(ei.value)
    sav = B.SuperArrayV([B.ArrayV(i32(100, 200, 300)), B.ArrayV(i32(400, 500, 600))])
    r = B.broadcast_value(A.Subtract, sav, table(B, (10, 20, 30, 40, 50, 60)), gpu_ctx)             # broadcast_superarrayview_to_table
    assert r.n_rows() == 6 and r.n_batches() == 2 and col(r.batches[0]) == [90, 180, 270] and col(r.batches[1]) == [360, 450, 540]
    with pytest.raises(mnr.ShapeError) as ei:
        B.broadcast_value(A.Add, B.SuperArrayV([B.ArrayV(i32(1, 2, 3)), B.ArrayV(i32(4, 5, 6))]), table(B, (10, 20, 30, 40, 50)), gpu_ctx)
    assert "does not match" in str(ei.value)
    s = stv(B, table(B, (100, 200, 300)), table(B, (400, 500, 600)))
    r = B.broadcast_value(A.Divide, s, i32(10, 20, 30, 40, 50, 60), gpu_ctx)                        # broadcast_supertableview_to_array
    assert r.n_rows() == 6 and r.n_batches() == 2 and col(r.batches[0]) == [10, 10, 10] and col(r.batches[1]) == [10, 10, 10]
    r = B.broadcast_value(A.Add, i32(1, 2, 3, 4, 5, 6), stv(B, table(B, (10, 20, 30)), table(B, (40, 50, 60))), gpu_ctx)   # array.rs:628-685
    assert col(r.batches[0]) == [11, 22, 33] and col(r.batches[1]) == [44, 55, 66]
    # Table (op) SuperTableView and the mirror (super_table_view.rs:183-250; no reference test): aligned slices, operand order kept
    s = stv(B, table(B, (1, 2, 3)), table(B, (4, 5, 6)))
    t = table(B, (10, 20, 30, 40, 50, 60))
    r = B.broadcast_value(A.Subtract, t, s, gpu_ctx)
    assert col(r.batches[0]) == [9, 18, 27] and col(r.batches[1]) == [36, 45, 54]
    r = B.broadcast_value(A.Subtract, s, t, gpu_ctx)
    assert col(r.batches[0]) == [-9, -18, -27] and col(r.batches[1]) == [-36, -45, -54]


def test_scalar_rs_and_array_rs(B, gpu_ctx):
    import minarrow_b200 as mnr
    A = mnr.ArithmeticOperator
    s = np.int32
    r = B.broadcast_scalar_to_table(A.Add, s(5), table(B, (1, 2, 3), (10, 20, 30)), gpu_ctx)
    assert r.n_rows() == 3 and r.n_cols() == 2 and col(r, 0) == [6, 7, 8] and col(r, 1) == [15, 25, 35]
    assert col(B.broadcast_scalar_to_table(A.Multiply, s(10), table(B, (2, 3, 4)), gpu_ctx)) == [20, 30, 40]
    assert col(B.broadcast_value(A.Subtract, s(50), B.TableV(table(B, (100, 200, 300)), 0, 3), gpu_ctx)) == [-50, -150, -250]
    r = B.broadcast_value(A.Divide, s(1000), B.TableV(table(B, (10, 20, 30), (100, 200, 300)), 0, 3), gpu_ctx)
    assert col(r, 0) == [100, 50, 33] and col(r, 1) == [10, 5, 3]
    assert B.broadcast_value(A.Add, s(5), i32(10, 20, 30), gpu_ctx).data.tolist() == [15, 25, 35]
    assert B.broadcast_value(A.Multiply, s(10), i32(2, 3, 4), gpu_ctx).data.tolist() == [20, 30, 40]
    r = B.broadcast_value(A.Add, s(10), B.SuperArray([i32(1, 2, 3), i32(4, 5, 6)]), gpu_ctx)
    assert r.n_chunks() == 2 and r.chunks[0].data.tolist() == [11, 12, 13] and r.chunks[1].data.tolist() == [14, 15, 16]
    arr = i32(10, 20, 30, 40, 50, 60)
    r = B.broadcast_value(A.Multiply, s(5), B.SuperArrayV([B.ArrayV(arr).slice(0, 3), B.ArrayV(arr).slice(3, 3)]), gpu_ctx)
    assert r.n_chunks() == 2 and r.chunks[0].data.tolist() == [50, 100, 150] and r.chunks[1].data.tolist() == [200, 250, 300]
    st = B.SuperTable([table(B, (1, 2, 3)), table(B, (4, 5, 6))], "st")
    r = B.broadcast_value(A.Subtract, s(100), st, gpu_ctx)
    assert r.n_batches() == 2 and col(r.batches[0]) == [99, 98, 97] and col(r.batches[1]) == [96, 95, 94]
    t2 = B.broadcast_value(A.Add, s(5), (i32(1, 2, 3), i32(10, 20, 30)), gpu_ctx)
    assert [x.data.tolist() for x in t2] == [[6, 7, 8], [15, 25, 35]]
    t3 = B.broadcast_value(A.Multiply, s(2), (i32(2, 4, 6), i32(3, 6, 9), i32(4, 8, 12)), gpu_ctx)
    assert [x.data.tolist() for x in t3] == [[4, 8, 12], [6, 12, 18], [8, 16, 24]]
    assert B.broadcast_value(A.Divide, s(50), i32(100, 200, 300), gpu_ctx).data.tolist() == [0, 0, 0]        # scalar_to_fieldarray
    assert B.broadcast_value(A.Multiply, i32(10, 20, 30), s(5), gpu_ctx).data.tolist() == [50, 100, 150]    # fieldarray_to_scalar
    # array.rs
    assert B.broadcast_value(A.Multiply, i32(10, 20, 30), s(2), gpu_ctx).data.tolist() == [20, 40, 60]
    st = B.SuperTable([table(B, (10, 20, 30)), table(B, (100, 200, 300))], "st")
    r = B.broadcast_value(A.Add, i32(1, 2, 3), st, gpu_ctx)
    assert r.n_batches() == 2 and col(r.batches[0]) == [11, 22, 33] and col(r.batches[1]) == [101, 202, 303]
    t = B.broadcast_value(A.Add, i32(5, 10, 15), (i32(1, 2, 3), i32(10, 20, 30)), gpu_ctx)
    assert [x.data.tolist() for x in t] == [[6, 12, 18], [15, 30, 45]]
    t = B.broadcast_value(A.Multiply, i32(2, 3, 4), (i32(10, 10, 10), i32(5, 5, 5), i32(1, 1, 1)), gpu_ctx)
    assert [x.data.tolist() for x in t] == [[20, 30, 40], [10, 15, 20], [2, 3, 4]]
    t = B.broadcast_value(A.Add, i32(1, 1, 1), (i32(10, 20, 30), i32(100, 200, 300), i32(5, 10, 15), i32(2, 4, 6)), gpu_ctx)
    assert [x.data.tolist() for x in t] == [[11, 21, 31], [101, 201, 301], [6, 11, 16], [3, 5, 7]]
    t = B.broadcast_value(A.Multiply, i32(10, 10, 10), (i32(1, 2, 3), i32(2, 3, 4), i32(3, 4, 5), i32(4, 5, 6), i32(5, 6, 7)), gpu_ctx)
    assert t[0].data.tolist() == [10, 20, 30] and t[4].data.tolist() == [50, 60, 70] and len(t) == 5
    t = B.broadcast_value(A.Subtract, i32(5, 5, 5), (i32(10, 10, 10), i32(20, 20, 20), i32(15, 15, 15), i32(8, 8, 8), i32(12, 12, 12), i32(6, 6, 6)), gpu_ctx)
    assert t[0].data.tolist() == [-5, -5, -5] and t[5].data.tolist() == [-1, -1, -1] and len(t) == 6


def test_table_rs_and_super_array_rs_remaining(B, gpu_ctx):
    """table.rs:568-760 (table - arrayview, SuperTableV + SuperTableV, table + SuperArray, table * SuperArrayView) and
    super_array.rs:674-719 (SuperArray + table)."""
    import minarrow_b200 as mnr
    A = mnr.ArithmeticOperator
    t = table(B, (10, 20, 30), (100, 200, 300), name="table1")
    r = B.broadcast_value(A.Subtract, t, B.ArrayV(i32(2, 3, 4)), gpu_ctx)
    assert col(r, 0) == [8, 17, 26] and col(r, 1) == [98, 197, 296]
    t1, t2 = table(B, (1, 2, 3), (10, 20, 30), name="table1"), table(B, (4, 5, 6), (40, 50, 60), name="table2")
    t3, t4 = table(B, (7, 8, 9), (70, 80, 90), name="table3"), table(B, (1, 1, 1), (2, 2, 2), name="table4")
    r = B.broadcast_super_table_add(stv(B, t1, t2), stv(B, t3, t4), None, gpu_ctx)
    assert r.n_batches() == 2 and r.name == "table1"
    assert col(r.batches[0], 0) == [8, 10, 12] and col(r.batches[0], 1) == [80, 100, 120]
    assert col(r.batches[1], 0) == [5, 6, 7] and col(r.batches[1], 1) == [42, 52, 62]
    with pytest.raises(mnr.KernelError) as ei:
        B.broadcast_super_table_add(stv(B, t1, t2), stv(B, t3), None, gpu_ctx)
    assert ei.value.kind == "BroadcastingError" and "chunk count mismatch: LHS 2 chunks, RHS 1 chunks" in str(ei.value)
    # the optional mask goes to every column's kernel (table.rs:98-101): masked rows are 0 and null in every result column
    m = mnr.Bitmask.from_bools([True, False, True])
    r = B.broadcast_table_add(t1, t3, m, gpu_ctx)
    assert r.name == "table1" and [c.data.tolist() for c in r.cols] == [[8, 0, 12], [80, 0, 120]]
    assert all(c.null_mask.to_bools().tolist() == [True, False, True] for c in r.cols)
    with pytest.raises(mnr.KernelError) as ei:
        B.broadcast_table_add(t1, table(B, (1, 2, 3)), None, gpu_ctx)
    assert "Table column count mismatch: LHS 2 cols, RHS 1 cols" in str(ei.value)
    r = B.broadcast_value(A.Add, table(B, (2, 3, 4)), B.SuperArray([i32(10, 20, 30), i32(40, 50, 60)]), gpu_ctx)
    assert r.n_chunks() == 2 and r.chunks[0].data.tolist() == [12, 23, 34] and r.chunks[1].data.tolist() == [42, 53, 64]
    arr = i32(10, 20, 30, 40, 50, 60)
    sav = B.SuperArrayV([B.ArrayV(arr).slice(0, 3), B.ArrayV(arr).slice(3, 3)])
    r = B.broadcast_value(A.Multiply, table(B, (1, 2, 3, 4, 5, 6)), sav, gpu_ctx)
    assert r.n_batches() == 2 and r.n_rows() == 6 and col(r.batches[0]) == [10, 40, 90] and col(r.batches[1]) == [160, 250, 360]
    r = B.broadcast_value(A.Add, B.SuperArray([i32(1, 2, 3), i32(4, 5, 6)]), table(B, (10, 20, 30)), gpu_ctx)
    assert r.n_chunks() == 2 and r.chunks[0].data.tolist() == [11, 22, 33] and r.chunks[1].data.tolist() == [14, 25, 36]
    with pytest.raises(mnr.ShapeError) as ei:
        B.broadcast_value(A.Add, B.SuperArray([i32(1, 2, 3)]), table(B, (10, 20, 30), (1, 2, 3)), gpu_ctx)
    assert "should result in single column" in str(ei.value)


def test_supertableview_arms_on_ragged_offset_windows(B, gpu_ctx):
    """Beyond the reference's vectors (all of which use offset 0 and equal slices): SuperTableView slices that are real
    windows (non-zero offsets, ragged lengths, two dtypes), an ArrayView that is itself a window, masks present on the operands
    (these arms pass no mask to the kernels, so the result carries none and every row is computed).  Expected values straight
    from numpy on the same windows."""
    import minarrow_b200 as mnr
    A = mnr.ArithmeticOperator
    rng = np.random.default_rng(17)

    def tab(n, name):
        return B.Table(name, [mnr.core.make_array(rng.integers(-10 ** 6, 10 ** 6, n).astype(np.int64), mnr.Bitmask.from_bools(rng.random(n) < 0.7)),
                              mnr.core.make_array(rng.standard_normal(n), None)])
    tabs = [tab(500, "a"), tab(77, "b"), tab(1000, "c")]
    wins = [(13, 300), (0, 77), (421, 513)]
    s = B.SuperTableV([B.TableV(t, o, n) for t, (o, n) in zip(tabs, wins)])
    total = sum(n for _, n in wins)
    base_i = rng.integers(-1000, 1000, total + 50).astype(np.int64)
    base_f = rng.standard_normal(total + 50)
    for c, base in ((0, base_i), (1, base_f)):
        sub = B.SuperTableV([B.TableV(B.Table(t.name, [t.cols[c]]), o, n) for t, (o, n) in zip(tabs, wins)])
        view = B.ArrayV(mnr.core.make_array(base, mnr.Bitmask.from_bools(rng.random(base.size) < 0.5)), 29, total)
        for op, f in ((A.Subtract, np.subtract), (A.Multiply, np.multiply)):
            for stv_is_lhs in (True, False):
                r = B.broadcast_value(op, sub, view, gpu_ctx) if stv_is_lhs else B.broadcast_value(op, view, sub, gpu_ctx)
                assert r.n_batches() == 3 and r.n_rows() == total
                from oracle import oracle as orc           # the checker: the oracle's restatement of the same arm
                exp = orc.broadcast_supertableview_to_arrayview(int(op), [([t.cols[c].data], o, n) for t, (o, n) in zip(tabs, wins)],
                                                                base[29:29 + total], stv_is_lhs)
                for k in range(3):
                    assert r.batches[k].cols[0].data.tobytes() == exp[k][0].tobytes(), ("oracle", c, op, stv_is_lhs, k)
                start = 29
                for k, (t, (o, n)) in enumerate(zip(tabs, wins)):
                    x, y = t.cols[c].data[o:o + n], base[start:start + n]
                    with np.errstate(over="ignore"):
                        e = f(x, y) if stv_is_lhs else f(y, x)
                    got = r.batches[k].cols[0]
                    assert got.null_mask is None and got.data.tobytes() == e.tobytes(), (c, op, stv_is_lhs, k)
                    start += n
    # both columns at once against a typed scalar, and Table (op) SuperTableView with the table cut to the same windows
    r = B.broadcast_value(A.Add, s, 3, gpu_ctx)
    for k, (t, (o, n)) in enumerate(zip(tabs, wins)):
        assert np.array_equal(r.batches[k].cols[0].data, t.cols[0].data[o:o + n] + 3)
        assert r.batches[k].cols[1].data.tobytes() == (t.cols[1].data[o:o + n] + 3.0).tobytes()
    whole = B.Table("w", [mnr.core.make_array(base_i[:total].copy(), None), mnr.core.make_array(base_f[:total].copy(), None)])
    r = B.broadcast_value(A.Subtract, whole, s, gpu_ctx)
    start = 0
    for k, (t, (o, n)) in enumerate(zip(tabs, wins)):
        assert np.array_equal(r.batches[k].cols[0].data, base_i[start:start + n] - t.cols[0].data[o:o + n])
        assert r.batches[k].cols[1].data.tobytes() == (base_f[start:start + n] - t.cols[1].data[o:o + n]).tobytes()
        start += n
