"""The reference's container-route tests, transcribed (expected values are the reference's own literals), run through
the CUDA path: Array, Scalar, Table, SuperTable and `Value` operator routes.

Sources: src/kernels/broadcast/array.rs:485-553,560-626; table.rs:432-566; super_table.rs:684-900; scalar.rs:1332-1350;
src/kernels/arithmetic/types.rs:922-1054."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B(gpu_ctx):
    import minarrow_b200 as mnr
    from minarrow_b200.kernels import broadcast as b
    b._mnr, b._ctx = mnr, gpu_ctx
    return b


def i32(*v):
    return np.array(v, dtype=np.int32)


def table(B, name, c1, c2):     # create_test_table (table.rs tests): two Int32 columns
    return B.Table(name, [B._mnr.IntegerArray(i32(*c1)), B._mnr.IntegerArray(i32(*c2))])


def col(t, i):
    c = t.cols[i]
    assert c.data.dtype == np.int32 and c.null_mask is None     # Table routes pass no mask (table.rs:54)
    return c.data.tolist()


def test_array_routes(B):
    A = B._mnr.ArithmeticOperator
    ctx = B._ctx
    assert B.broadcast_value(A.Add, i32(1, 2, 3), i32(4, 5, 6), ctx).data.tolist() == [5, 7, 9]            # array.rs:485
    assert B.broadcast_value(A.Add, i32(1, 2, 3), i32(10), ctx).data.tolist() == [11, 12, 13]              # array.rs:500
    assert B.broadcast_value(A.Subtract, i32(10, 20, 30), i32(1, 2, 3), ctx).data.tolist() == [9, 18, 27]  # array.rs:515
    assert B.broadcast_value(A.Multiply, i32(2, 3, 4), i32(5, 6, 7), ctx).data.tolist() == [10, 18, 28]    # array.rs:530
    assert B.broadcast_value(A.Divide, i32(100, 200, 300), i32(10, 20, 30), ctx).data.tolist() == [10, 10, 10]  # array.rs:545
    # Value * Value with a length-1 left operand (types.rs:1035-1054)
    assert B.value_multiply(i32(10), i32(1, 2, 3, 4, 5), ctx).data.tolist() == [10, 20, 30, 40, 50]
    # operand order: Scalar / Array = 50 / [100,200,300] = [0,0,0] (scalar.rs:1332-1350)
    assert B.value_divide(50, i32(100, 200, 300), ctx).data.tolist() == [0, 0, 0]
    assert B.value_divide(i32(100, 200, 300), 50, ctx).data.tolist() == [2, 4, 6]
    # all operators on Values (types.rs:943-990): [10,20,30] op [2,4,6]
    a, b = i32(10, 20, 30), i32(2, 4, 6)
    assert B.value_add(a, b, ctx).data.tolist() == [12, 24, 36]
    assert B.value_subtract(a, b, ctx).data.tolist() == [8, 16, 24]
    assert B.value_multiply(a, b, ctx).data.tolist() == [20, 80, 180]
    assert B.value_divide(a, b, ctx).data.tolist() == [5, 5, 5]
    assert B.value_remainder(a, b, ctx).data.tolist() == [0, 0, 0]


def test_table_routes(B):
    A = B._mnr.ArithmeticOperator
    ctx = B._ctx
    r = B.broadcast_value(A.Add, table(B, "table1", [1, 2, 3], [10, 20, 30]), table(B, "table2", [4, 5, 6], [40, 50, 60]), ctx)
    assert (r.n_cols(), r.n_rows(), r.name) == (2, 3, "table1")                                     # table.rs:432-465
    assert col(r, 0) == [5, 7, 9] and col(r, 1) == [50, 70, 90]
    r = B.broadcast_table_with_operator(A.Multiply, table(B, "t1", [2, 3, 4], [5, 6, 7]), table(B, "t2", [10, 10, 10], [2, 2, 2]), ctx)
    assert col(r, 0) == [20, 30, 40] and col(r, 1) == [10, 12, 14]                                  # table.rs:498-519
    with pytest.raises(B._mnr.KernelError, match="column count mismatch"):                           # table.rs:467-479
        B.broadcast_table_with_operator(A.Add, B.Table("t", [B._mnr.IntegerArray(i32(1, 2, 3))]),
                                        table(B, "table2", [4, 5, 6], [40, 50, 60]), ctx)
    with pytest.raises(B._mnr.KernelError) as ei:                                                    # table.rs:481-496 (row count)
        B.broadcast_table_with_operator(A.Add, table(B, "t", [1, 2], [10, 20]), table(B, "table2", [4, 5, 6], [40, 50, 60]), ctx)
    assert ei.value.kind == "LengthMismatch"
    r = B.broadcast_table_to_array(A.Add, table(B, "table1", [10, 20, 30], [100, 200, 300]), i32(1, 2, 3), ctx)
    assert col(r, 0) == [11, 22, 33] and col(r, 1) == [101, 202, 303]                               # table.rs:521-541
    r = B.broadcast_table_to_scalar(A.Multiply, table(B, "table1", [10, 20, 30], [100, 200, 300]), 5, ctx)
    assert col(r, 0) == [50, 100, 150] and col(r, 1) == [500, 1000, 1500]                           # table.rs:544-565
    # array op table keeps the operand order (array.rs:560-626): [100,100,100] - cols
    r = B.broadcast_array_to_table(A.Subtract, i32(100, 100, 100), table(B, "t", [1, 2, 3], [10, 20, 30]), ctx)
    assert col(r, 0) == [99, 98, 97] and col(r, 1) == [90, 80, 70]


def test_super_table_routes(B):
    A = B._mnr.ArithmeticOperator
    ctx = B._ctx
    lhs = B.SuperTable([table(B, "batch1", [1, 2, 3], [10, 20, 30]), table(B, "batch2", [4, 5, 6], [40, 50, 60])])
    rhs = B.SuperTable([table(B, "batch1", [1, 1, 1], [5, 5, 5]), table(B, "batch2", [2, 2, 2], [10, 10, 10])])
    r = B.broadcast_value(A.Add, lhs, rhs, ctx)                                                     # super_table.rs:684-728
    assert (r.n_batches(), r.n_rows(), r.n_cols()) == (2, 6, 2)
    assert col(r.batches[0], 0) == [2, 3, 4] and col(r.batches[1], 0) == [6, 7, 8] and col(r.batches[1], 1) == [50, 60, 70]
    with pytest.raises(B._mnr.ShapeError, match="chunk count mismatch"):                              # super_table.rs:819-841
        B.broadcast_super_table_with_operator(A.Add, B.SuperTable([table(B, "b", [1, 2, 3], [10, 20, 30])]), rhs, ctx)
    l3 = B.SuperTable([table(B, "b1", [1, 2, 3], [10, 20, 30]), table(B, "b2", [4, 5, 6], [40, 50, 60]),
                       table(B, "b3", [7, 8, 9], [70, 80, 90])])
    r3 = B.SuperTable([table(B, "b1", [1, 1, 1], [1, 1, 1]), table(B, "b2", [2, 2, 2], [2, 2, 2]), table(B, "b3", [3, 3, 3], [3, 3, 3])])
    r = B.broadcast_super_table_with_operator(A.Add, l3, r3, ctx)                                    # super_table.rs:844-900
    assert (r.n_batches(), r.n_rows()) == (3, 9)
    assert [col(b, 0) for b in r.batches] == [[2, 3, 4], [6, 7, 8], [10, 11, 12]]
    # SuperTable op Scalar: every batch, every column (super_table.rs:77-91)
    r = B.broadcast_value(A.Multiply, lhs, 2, ctx)
    assert col(r.batches[1], 1) == [80, 100, 120]
