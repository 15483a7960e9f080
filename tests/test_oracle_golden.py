"""Pins the CPU oracle against the reference's own known-answer tests (transcribed, file:line in each
case) and against independent second opinions (numpy, pyarrow.compute, math.fsum)."""
import math

import numpy as np
import pytest

import kat_runner
from backends import OracleBackend
from oracle import oracle as orc

CASES = kat_runner.load_cases()


@pytest.mark.parametrize("case", CASES, ids=[kat_runner.case_id(c) for c in CASES])
def test_reference_kat(case):
    assert kat_runner.run_case(case, OracleBackend()) in ("ok", "skipped")


def test_every_kind_is_exercised():
    kinds = {c["kind"] for c in CASES}
    assert {"apply_int", "apply_float", "apply_fma", "merge_and", "bits_binop", "bits_not", "bits_popcount",
            "route", "super_route", "sum_arange"} <= kinds


# ---- independent cross-checks --------------------------------------------------------------------------

INT_DTYPES = [np.int32, np.uint32, np.int64, np.uint64]


def _rand_int(rng, dt, n, small=False):
    info = np.iinfo(dt)
    if small:
        return rng.integers(max(info.min, -50), min(info.max, 50), n, dtype=dt, endpoint=True)
    a = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)
    # sprinkle edge values
    edges = np.array([info.min, info.max, 0, 1, info.max - 1] + ([-1] if info.min < 0 else []), dtype=dt)
    idx = rng.integers(0, n, 64)
    a[idx] = edges[rng.integers(0, len(edges), 64)]
    return a


@pytest.mark.parametrize("dt", INT_DTYPES)
def test_int_wrapping_ops_match_numpy(dt):
    rng = np.random.default_rng(1)
    a, b = _rand_int(rng, dt, 5000), _rand_int(rng, dt, 5000)
    with np.errstate(over="ignore"):
        for op, f in ((orc.ADD, np.add), (orc.SUB, np.subtract), (orc.MUL, np.multiply)):
            got, m = orc.apply_int(a, b, op)
            assert m is None and np.array_equal(got, f(a, b))


@pytest.mark.parametrize("dt", INT_DTYPES)
def test_int_div_rem_floordiv_semantics(dt):
    rng = np.random.default_rng(2)
    n = 4000
    a, b = _rand_int(rng, dt, n), _rand_int(rng, dt, n, small=True)
    mask = rng.random(n) < 0.9
    bits = orc.Bits.from_bools(mask)
    A, B = a.astype(object), b.astype(object)
    info = np.iinfo(dt)

    def wrap(v):
        v &= (1 << info.bits) - 1
        return v - (1 << info.bits) if (info.min < 0 and v >= (1 << (info.bits - 1))) else v

    def trunc_div(x, y):
        q = abs(x) // abs(y)
        return q if (x < 0) == (y < 0) else -q

    for op in (orc.DIV, orc.REM, orc.FLOORDIV):
        got, m = orc.apply_int(a, b, op, bits)
        valid = m.to_bools()
        for i in range(n):
            if not mask[i] or B[i] == 0:
                assert got[i] == 0 and not valid[i]
                continue
            q = trunc_div(A[i], B[i])
            r = A[i] - q * B[i]
            exp = {orc.DIV: q, orc.REM: r, orc.FLOORDIV: A[i] // B[i]}[op]
            assert int(got[i]) == wrap(exp) and valid[i], (op, A[i], B[i], got[i], exp)


def test_int_min_over_minus_one_wraps():
    # ASSUMPTION documented in DESIGN.md (core::simd guard): MIN / -1 = MIN, MIN % -1 = 0.
    for dt in (np.int32, np.int64):
        mn = np.iinfo(dt).min
        a, b = np.array([mn], dtype=dt), np.array([-1], dtype=dt)
        assert orc.apply_int(a, b, orc.DIV)[0][0] == mn
        assert orc.apply_int(a, b, orc.REM)[0][0] == 0
        assert orc.apply_int(a, b, orc.FLOORDIV)[0][0] == mn


@pytest.mark.parametrize("dt", INT_DTYPES)
def test_int_power_is_wrapping_repeated_multiply(dt):
    rng = np.random.default_rng(3)
    a = _rand_int(rng, dt, 300)
    e = rng.integers(0, 70, 300).astype(dt)
    got, _ = orc.apply_int(a, e, orc.POW)
    bits = np.iinfo(dt).bits
    for x, k, g in zip(a.astype(object), e.astype(object), got.astype(object)):
        assert (g - pow(x, k)) % (1 << bits) == 0
    if np.iinfo(dt).min < 0:  # negative exponent -> to_u32() is None -> 0 -> 1   (std.rs:67)
        got, _ = orc.apply_int(np.array([7, -3], dtype=dt), np.array([-1, -5], dtype=dt), orc.POW)
        assert list(got) == [1, 1]
    if bits == 64:  # exponent above u32::MAX -> 0 -> 1
        got, _ = orc.apply_int(np.array([7], dtype=dt), np.array([1 << 33], dtype=dt), orc.POW)
        assert list(got) == [1]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_float_ops_match_numpy_ieee(dt):
    rng = np.random.default_rng(4)
    n = 20000
    a, b = rng.standard_normal(n).astype(dt), rng.standard_normal(n).astype(dt)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, np.finfo(dt).tiny / 4, np.finfo(dt).max], dtype=dt)
    a[rng.integers(0, n, 200)] = special[rng.integers(0, len(special), 200)]
    b[rng.integers(0, n, 200)] = special[rng.integers(0, len(special), 200)]
    with np.errstate(all="ignore"):
        for op, f in ((orc.ADD, np.add), (orc.SUB, np.subtract), (orc.MUL, np.multiply), (orc.DIV, np.divide),
                      (orc.REM, np.fmod)):
            got, _ = orc.apply_float(a, b, op)
            exp = f(a, b)
            assert np.array_equal(got.view(np.uint32 if dt == np.float32 else np.uint64)[~np.isnan(exp)],
                                  exp.view(np.uint32 if dt == np.float32 else np.uint64)[~np.isnan(exp)]), op
            assert np.array_equal(np.isnan(got), np.isnan(exp))
        got, _ = orc.apply_float(a, b, orc.FLOORDIV)
        exp = np.floor(a / b)
        assert np.array_equal(got[~np.isnan(exp)], exp[~np.isnan(exp)])


def test_float_masked_zeroes_nulls_and_keeps_nan_valid():
    a = np.array([1.0, np.nan, 3.0, 4.0, 5.0])
    b = np.array([0.0, 1.0, 1.0, 0.0, 2.0])
    m = orc.Bits.from_bools([True, True, False, True, True])
    got, om = orc.apply_float(a, b, orc.DIV, m)
    assert np.isinf(got[0]) and np.isnan(got[1]) and got[2] == 0.0 and not np.signbit(got[2])
    assert list(om.to_bools()) == [True, True, False, True, True]


def test_fma_is_single_rounding():
    a, b, c = np.array([1.0 + 2.0 ** -30]), np.array([1.0 - 2.0 ** -30]), np.array([-1.0])
    fused, _ = orc.apply_fma(a, b, c)
    unfused, _ = orc.apply_fma(a, b, c, fused=False)
    assert fused[0] == -(2.0 ** -60) and unfused[0] == 0.0


def test_bitmask_ops_match_numpy_on_odd_lengths():
    rng = np.random.default_rng(5)
    for n in (1, 7, 8, 9, 63, 64, 65, 127, 128, 129, 1000, 4099):
        a, b = rng.random(n) < 0.5, rng.random(n) < 0.5
        A, B = orc.Bits.from_bools(a), orc.Bits.from_bools(b)
        for f, g in ((orc.and_masks, np.logical_and), (orc.or_masks, np.logical_or), (orc.xor_masks, np.logical_xor)):
            out = f((A, 0, n), (B, 0, n))
            assert np.array_equal(out.to_bools(), g(a, b))
            assert out.bits.size == (n + 7) // 8
            if n % 8:
                assert out.bits[-1] >> (n % 8) == 0  # trailing bits cleared
        nt = orc.not_mask((A, 0, n))
        assert np.array_equal(nt.to_bools(), ~a) and (n % 8 == 0 or nt.bits[-1] >> (n % 8) == 0)
        assert orc.popcount_mask((A, 0, n)) == int(a.sum()) == orc.count_ones(A)
        assert orc.null_count(A) == n - int(a.sum())
        assert np.array_equal(orc.merge_bitmasks_to_new(A, B, n).to_bools(), a & b)
        assert np.array_equal(orc.merge_bitmasks_to_new(A, None, n).to_bools(), a)
        assert orc.merge_bitmasks_to_new(None, None, n) is None
        assert np.array_equal(orc.union(A, B).to_bools(), a | b)


def test_bitmask_window_offsets_floor_to_bytes_like_reference():
    # bitmask_window_bytes (bitmask/mod.rs:124-128): offset/8 floors sub-byte offsets.
    rng = np.random.default_rng(6)
    a, b = rng.random(256) < 0.5, rng.random(256) < 0.5
    A, B = orc.Bits.from_bools(a), orc.Bits.from_bools(b)
    out = orc.and_masks((A, 64, 100), (B, 128, 100))
    assert np.array_equal(out.to_bools(), a[64:164] & b[128:228])
    out = orc.and_masks((A, 67, 40), (B, 3, 40))  # floored to bits 64 and 0
    assert np.array_equal(out.to_bools(), a[64:104] & b[0:40])
    assert orc.popcount_mask((A, 64, 100)) == int(a[64:164].sum())


def test_sum_orders_agree_and_match_exact():
    rng = np.random.default_rng(7)
    d = rng.integers(-2 ** 63, 2 ** 63 - 1, 3_000_001, dtype=np.int64)
    exact = int(d.astype(object).sum())
    exp = (exact + 2 ** 63) % 2 ** 64 - 2 ** 63
    assert orc.simd_sum_i64(d) == orc.hotloop_sum_i64(d) == orc.rayon_simd_sum_i64(d, threads=4) == exp
    f = rng.standard_normal(3_000_001)
    ref = math.fsum(f)
    scale = math.fsum(np.abs(f))
    for got in (orc.simd_sum_f64(f), orc.hotloop_sum_f64(f), orc.rayon_simd_sum_f64(f, threads=4)):
        assert abs(got - ref) <= 1e-12 * scale


def test_null_aware_stats_match_pyarrow():
    pa = pytest.importorskip("pyarrow")
    pc = pytest.importorskip("pyarrow.compute")
    rng = np.random.default_rng(8)
    n = 100_003
    valid = rng.random(n) < 0.9
    V = orc.Bits.from_bools(valid)
    for dt in (np.int32, np.int64, np.uint32, np.uint64, np.float32, np.float64):
        if np.dtype(dt).kind == "f":
            d = rng.standard_normal(n).astype(dt)
            d[rng.integers(0, n, 50)] = np.nan
            d[rng.integers(0, n, 50)] = -0.0
        else:
            d = rng.integers(0 if np.dtype(dt).kind == "u" else -10 ** 6, 10 ** 6, n).astype(dt)
        arr = pa.array(d, mask=~valid)
        st = orc.stats(d, V)
        assert st["count"] == pc.count(arr).as_py() == int(valid.sum())
        mm = pc.min_max(arr).as_py()
        if np.dtype(dt).kind == "f":
            nn = pa.array(d, mask=~valid | np.isnan(d))
            mm = pc.min_max(nn).as_py()
            assert st["min"] == mm["min"] and st["max"] == mm["max"]
            assert math.isnan(pc.sum(arr).as_py()) and math.isnan(st["sum"])
            clean = np.where(np.isnan(d), 0, d)
            st2 = orc.stats(clean, V)
            ref = math.fsum(clean[valid].astype(np.float64))
            assert abs(st2["sum"] - ref) <= 1e-12 * math.fsum(np.abs(clean[valid].astype(np.float64)))
        else:
            assert st["min"] == mm["min"] and st["max"] == mm["max"]
            assert st["sum"] == pc.sum(arr).as_py()
            assert math.isclose(st["mean"], pc.mean(arr).as_py(), rel_tol=1e-12)
    # all-null and empty
    st = orc.stats(np.arange(10, dtype=np.int64), orc.Bits.from_bools(np.zeros(10, bool)))
    assert st["count"] == 0 and math.isnan(st["mean"])
    assert orc.stats(np.array([], dtype=np.float64))["count"] == 0
    # i64 sum wraps like pyarrow / Rust wrapping_add
    big = np.array([2 ** 63 - 1, 1], dtype=np.int64)
    assert orc.stats(big)["sum"] == -2 ** 63
    s, c = orc.par_masked_sum_i64(big, None, threads=2)
    assert (s, c) == (-2 ** 63, 2)


def test_par_masked_sum_matches_stats():
    rng = np.random.default_rng(9)
    n = (1 << 21) + 12345
    d = rng.integers(-2 ** 63, 2 ** 63 - 1, n, dtype=np.int64)
    valid = rng.random(n) < 0.9
    V = orc.Bits.from_bools(valid)
    st = orc.stats(d, V)
    assert orc.par_masked_sum_i64(d, V, threads=4) == (st["sum"], st["count"])
    with np.errstate(over="ignore"):
        assert st["sum"] == int(np.where(valid, d, 0).sum(dtype=np.int64))


def test_array_superarray_rechunk_route_of_the_oracle():
    """oracle.create_aligned_chunks_from_array / broadcast_array_superarray restate src/utils.rs:367-481 and
    src/kernels/broadcast/mod.rs:1351-1361.  The reference's tests for these arms carry no null masks, so the mask rules are
    pinned here on hand-worked literals (the same ones tests/cpp/test_container_routes.cpp runs on the GPU): chunk i is valid
    where array_mask[window i] | chunk_mask[i]; a chunk WITHOUT a mask next to one with a mask is all valid in the union;
    with no chunk masks at all the array's own window is used; no masks anywhere -> dense."""
    B = orc.Bits.from_bools
    arr = np.array([1, 2, 3, 4, 5, 6, 7], dtype=np.int32)
    am = B([True, False, True, True, False, False, True])
    sa = [(np.array([10, 20, 30], np.int32), B([False, False, True])), (np.array([40, 50, 60, 70], np.int32), None)]
    out = orc.broadcast_array_superarray(orc.ADD, arr, am, sa, True)
    assert out[0][0].tolist() == [11, 0, 33] and out[0][1].to_bools().tolist() == [True, False, True]
    assert out[1][0].tolist() == [44, 55, 66, 77] and out[1][1].to_bools().all()
    out = orc.broadcast_array_superarray(orc.SUB, arr, am, sa, False)          # SuperArray - Array
    assert out[0][0].tolist() == [9, 0, 27] and out[1][0].tolist() == [36, 45, 54, 63]
    plain = [(np.array([10, 20, 30], np.int32), None), (np.array([40, 50, 60, 70], np.int32), None)]
    out = orc.broadcast_array_superarray(orc.MUL, arr, am, plain, True)         # no chunk masks: the array's windows
    assert out[0][0].tolist() == [10, 0, 90] and out[1][0].tolist() == [160, 0, 0, 490]
    assert out[1][1].to_bools().tolist() == [True, False, False, True]
    out = orc.broadcast_array_superarray(orc.ADD, arr, None, plain, True)       # no masks anywhere
    assert out[1][0].tolist() == [44, 55, 66, 77] and out[1][1] is None
    al = orc.create_aligned_chunks_from_array(arr, am, sa)
    assert [len(d) for d, _ in al] == [3, 4] and al[1][0].tolist() == [4, 5, 6, 7]
    assert al[0][1].to_bools().tolist() == [True, False, True] and al[1][1].to_bools().all()
    with pytest.raises(orc.KernelError):
        orc.create_aligned_chunks_from_array(arr[:3], None, sa)
    with pytest.raises(orc.KernelError):                                        # mask lengths must match for the union
        orc.union_array_superarray_masks(B([True] * 6), sa)
    # the SuperArray route itself: union of the chunk masks, or the one present, or none (super_array.rs:214-230)
    r = orc.route_super_array_broadcast(orc.ADD, [(np.array([1, 2], np.int32), B([True, False]))], [(np.array([5, 5], np.int32), B([False, False]))])
    assert r[0][0].tolist() == [6, 0] and r[0][1].to_bools().tolist() == [True, False]
    with pytest.raises(orc.KernelError):
        orc.route_super_array_broadcast(orc.ADD, [(np.array([1, 2], np.int32), None)], [(np.array([5], np.int32), None)])


def test_datetime_delegation_reference_vectors():
    """apply_datetime_i64 — the reference's own tests (src/kernels/arithmetic/mod.rs:418-506: datetime_add, datetime_all_ops,
    datetime_masked_and_empty, datetime_len_mismatch_panics): integer kernels + merge of the two arrays' masks."""
    i64 = np.int64
    out, m = orc.apply_datetime(np.array([1000, 2000, 3000], i64), None, 0, 3, np.array([10, 20, 30], i64), None, 0, 3, orc.ADD)
    assert out.tolist() == [1010, 2020, 3030] and m is None
    a, b = np.array([10, 20, 30, 40], i64), np.array([1, 2, 3, 4], i64)
    exp = {orc.ADD: [11, 22, 33, 44], orc.SUB: [9, 18, 27, 36], orc.MUL: [10, 40, 90, 160], orc.DIV: [10, 10, 10, 10],
           orc.REM: [0, 0, 0, 0], orc.POW: [10, 20 ** 2, 30 ** 3, 40 ** 4]}
    for op, e in exp.items():
        out, m = orc.apply_datetime(a, None, 0, 4, b, None, 0, 4, op)
        assert out.tolist() == e and m is None, op
    mask = orc.Bits.from_bools([True, False, True, True])
    out, m = orc.apply_datetime(a, mask, 0, 4, b, None, 0, 4, orc.ADD)
    assert out.tolist() == [11, 0, 33, 44] and m.to_bools().tolist() == [True, False, True, True]
    out, m = orc.apply_datetime(np.zeros(0, i64), None, 0, 0, np.zeros(0, i64), None, 0, 0, orc.ADD)
    assert out.size == 0
    with pytest.raises(orc.KernelError) as ei:
        orc.apply_datetime(np.array([1000, 2000], i64), None, 0, 2, np.array([10], i64), None, 0, 1, orc.ADD)
    assert ei.value.kind == "LengthMismatch"
    # windows: the data is offset, the masks are merged from bit 0 (dispatch.rs:321-322)
    out, m = orc.apply_datetime(a, mask, 1, 2, b, None, 2, 2, orc.ADD)
    assert out.tolist() == [23, 0] and m.to_bools().tolist() == [True, False]


def test_view_arm_restatements_on_reference_vectors():
    """The oracle's table / view arms against the reference's own vectors (array_view.rs:170-411, table_view.rs:200-480,
    super_array_view.rs:89-160, super_table_view.rs:256-570, table.rs:568-760, super_array.rs:674-719)."""
    i = lambda *v: np.array(v, dtype=np.int32)
    T = lambda *cols: ([i(*c) for c in cols], 0, len(cols[0]))
    L = lambda res: [c.tolist() for c in res]
    assert L(orc.broadcast_tableview_to_arrayview(orc.ADD, T((10, 20, 30), (100, 200, 300)), i(1, 2, 3), False)) == [[11, 22, 33], [101, 202, 303]]
    assert L(orc.broadcast_tableview_to_arrayview(orc.MUL, T((10, 10, 10)), i(2, 3, 4), False)) == [[20, 30, 40]]
    assert L(orc.broadcast_tableview_to_arrayview(orc.SUB, T((10, 20, 30), (100, 200, 300)), i(5, 5, 5), False)) == [[-5, -15, -25], [-95, -195, -295]]
    stv = [T((10, 20, 30)), T((40, 50, 60))]
    assert [L(t) for t in orc.broadcast_supertableview_to_arrayview(orc.ADD, stv, i(1, 2, 3, 4, 5, 6), False)] == [[[11, 22, 33]], [[44, 55, 66]]]
    with pytest.raises(orc.KernelError, match="does not match"):
        orc.broadcast_supertableview_to_arrayview(orc.ADD, stv, i(1, 2, 3, 4, 5), False)
    assert L(orc.broadcast_tableview_to_tableview(orc.ADD, T((1, 2, 3), (10, 20, 30)), T((5, 5, 5), (100, 100, 100)))) == [[6, 7, 8], [110, 120, 130]]
    with pytest.raises(orc.KernelError, match="column count mismatch"):
        orc.broadcast_tableview_to_tableview(orc.ADD, T((1, 2, 3)), T((5, 5, 5), (10, 10, 10)))
    assert L(orc.broadcast_tableview_to_arrayview(orc.SUB, T((100, 200, 300)), i(10, 20, 30))) == [[90, 180, 270]]
    r = orc.broadcast_tableview_to_superarrayview(orc.MUL, T((1, 2, 3, 4, 5, 6)), [i(10, 20, 30), i(40, 50, 60)])
    assert [L(t) for t in r] == [[[10, 40, 90]], [[160, 250, 360]]]
    with pytest.raises(orc.KernelError, match="does not match"):
        orc.broadcast_tableview_to_superarrayview(orc.ADD, T((1, 2, 3, 4, 5)), [i(10, 20, 30), i(40, 50, 60)])
    assert [L(t) for t in orc.broadcast_tableview_to_superarrayview(orc.ADD, T((10, 20, 30)), [i(1, 2, 3)], False)] == [[[11, 22, 33]]]
    with pytest.raises(orc.KernelError, match="does not match"):
        orc.broadcast_tableview_to_superarrayview(orc.ADD, T((10, 20, 30, 40, 50)), [i(1, 2, 3)], False)
    r = orc.broadcast_supertableview_to_arrayview(orc.MUL, [T((2, 3, 4)), T((5, 6, 7))], i(10, 10, 10, 10, 10, 10))
    assert [L(t) for t in r] == [[[20, 30, 40]], [[50, 60, 70]]]
    r = orc.broadcast_tableview_to_superarrayview(orc.SUB, T((10, 20, 30, 40, 50, 60)), [i(100, 200, 300), i(400, 500, 600)], False)
    assert [L(t) for t in r] == [[[90, 180, 270]], [[360, 450, 540]]]
    r = orc.broadcast_supertableview_to_arrayview(orc.DIV, [T((100, 200, 300)), T((400, 500, 600))], i(10, 20, 30, 40, 50, 60), True, False)
    assert [L(t) for t in r] == [[[10, 10, 10]], [[10, 10, 10]]]
    assert L(orc.broadcast_tableview_to_arrayview(orc.SUB, T((10, 20, 30), (100, 200, 300)), i(2, 3, 4))) == [[8, 17, 26], [98, 197, 296]]
    assert L(orc.broadcast_table_to_superarray(orc.ADD, [i(2, 3, 4)], [i(10, 20, 30), i(40, 50, 60)])) == [[12, 23, 34], [42, 53, 64]]
    assert L(orc.broadcast_table_to_superarray(orc.ADD, [i(10, 20, 30)], [i(1, 2, 3), i(4, 5, 6)], False)) == [[11, 22, 33], [14, 25, 36]]
    with pytest.raises(orc.KernelError, match="single column"):
        orc.broadcast_table_to_superarray(orc.ADD, [i(1, 2, 3), i(1, 2, 3)], [i(1, 2, 3)])
