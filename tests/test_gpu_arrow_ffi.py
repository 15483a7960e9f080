"""Arrow C Data Interface round trips with PyArrow as the foreign peer (the reference's pyo3/tests/test_roundtrip.py
plays the same role for its own FFI, src/ffi/arrow_c_ffi.rs:432-470,640): import with element / bit offsets, compute on
the device, export, and compare with pyarrow.compute on the host."""
import numpy as np
import pytest

pa = pytest.importorskip("pyarrow")
import pyarrow.compute as pc  # noqa: E402

pytestmark = pytest.mark.gpu

TYPES = [pa.int8(), pa.uint8(), pa.int16(), pa.uint16(), pa.int32(), pa.uint32(), pa.int64(), pa.uint64(), pa.float32(),
         pa.float64()]


@pytest.fixture(scope="module")
def mnr(gpu_ctx):
    import minarrow_b200 as m
    import minarrow_b200.arrow  # noqa: F401
    return m


def _make(rng, typ, n, nulls=True):
    npdt = typ.to_pandas_dtype()
    if pa.types.is_floating(typ):
        d = (rng.standard_normal(n) * 100).astype(npdt)
    else:
        info = np.iinfo(npdt)
        d = rng.integers(info.min, info.max, n, dtype=npdt, endpoint=True)
    mask = (rng.random(n) < 0.2) if nulls else None      # pyarrow: True = null
    return pa.array(d, type=typ, mask=mask), d, mask


@pytest.mark.parametrize("typ", TYPES, ids=str)
def test_import_offsets_and_export_round_trip(mnr, gpu_ctx, typ):
    rng = np.random.default_rng(41)
    arr, d, mask = _make(rng, typ, 10_007)
    for off, ln in ((0, 10_007), (1, 10_000), (3, 77), (8, 64), (13, 9_990), (64, 0), (5, 1)):
        sl = arr.slice(off, ln)
        v, bits, val = mnr.arrow.from_arrow(gpu_ctx, sl)
        assert bits is None and len(v) == ln
        assert v.download().tobytes() == d[off:off + ln].tobytes()
        if sl.null_count:
            assert np.array_equal(val.download().to_bools(), ~mask[off:off + ln])
            assert val.download().bits.size == (ln + 7) // 8
        else:
            assert val is None
        back = mnr.arrow.to_arrow(gpu_ctx, v, val)
        assert back.type == typ and back.null_count == sl.null_count and back.equals(sl)
    dense, d2, _ = _make(rng, typ, 1000, nulls=False)
    v, _, val = mnr.arrow.from_arrow(gpu_ctx, dense)
    assert val is None and mnr.arrow.to_arrow(gpu_ctx, v, None).equals(dense)


def test_boolean_arrays_bit_offsets(mnr, gpu_ctx):
    rng = np.random.default_rng(42)
    n = 5_003
    data, mask = rng.random(n) < 0.5, rng.random(n) < 0.2
    arr = pa.array(data, type=pa.bool_(), mask=mask)
    for off, ln in ((0, n), (1, n - 1), (7, 100), (9, 4_000), (64, 640), (3, 0)):
        sl = arr.slice(off, ln)
        v, bits, val = mnr.arrow.from_arrow(gpu_ctx, sl)
        assert v is None and len(bits) == ln
        assert np.array_equal(bits.download().to_bools(), data[off:off + ln])
        if sl.null_count:
            assert np.array_equal(val.download().to_bools(), ~mask[off:off + ln])
        assert mnr.arrow.to_arrow_bool(gpu_ctx, bits, val).equals(sl)
        # BooleanArray `!` inverts data, keeps validity (src/structs/variants/boolean.rs:853-866) == pc.invert
        if ln:
            inv = mnr.device_ops.bits_not(gpu_ctx, bits, 0, ln)
            assert mnr.arrow.to_arrow_bool(gpu_ctx, inv, val).equals(pc.invert(sl))


@pytest.mark.parametrize("typ", [pa.int32(), pa.int64(), pa.uint64(), pa.float32(), pa.float64()], ids=str)
def test_device_arithmetic_matches_pyarrow_compute(mnr, gpu_ctx, typ):
    """Two sliced nullable Arrow columns -> fused null-aware add / multiply on the device -> Arrow == pyarrow.compute
    (wrapping integer arithmetic, null if either side is null)."""
    dev = mnr.device_ops
    rng = np.random.default_rng(43)
    a, _, _ = _make(rng, typ, 100_003)
    b, _, _ = _make(rng, typ, 100_003)
    sa, sb = a.slice(5, 99_990), b.slice(11, 99_990)
    va, _, ma = mnr.arrow.from_arrow(gpu_ctx, sa)
    vb, _, mb = mnr.arrow.from_arrow(gpu_ctx, sb)
    for op, f in ((mnr.ArithmeticOperator.Add, pc.add), (mnr.ArithmeticOperator.Multiply, pc.multiply),
                  (mnr.ArithmeticOperator.Subtract, pc.subtract)):
        ob, om = dev.ew_binary(gpu_ctx, op, va, vb, ma, mb, mnr.MaskMode.And)
        got = mnr.arrow.to_arrow(gpu_ctx, ob, om)
        exp = f(sa, sb)
        assert got.null_count == exp.null_count
        assert got.equals(exp), (typ, op)
    # null-aware aggregates vs pyarrow.compute (SURVEY A.6 cross-check): sum wraps, min/max/count skip nulls
    st = dev.reduce_stats(gpu_ctx, va, ma)
    assert st["count"] == pc.count(sa).as_py()
    mm = pc.min_max(sa).as_py()
    assert st["min"] == mm["min"] and st["max"] == mm["max"]
    if pa.types.is_integer(typ):
        assert st["sum"] == pc.sum(sa).as_py()
    else:
        assert abs(st["sum"] - pc.sum(sa).as_py()) <= 1e-9 * abs(pc.sum(pc.abs(sa)).as_py())


@pytest.mark.parametrize("world", [1, 2, 3])
def test_stream_import_shards_chunks_in_order(mnr, gpu_ctx, world):
    """Arrow C stream (pyarrow.ChunkedArray.__arrow_c_stream__, the PyCapsule route of src/ffi/arrow_c_ffi.rs:153-168) ->
    SuperArray chunks: every rank drains the stream and uploads its contiguous block; the blocks tile the column in
    order, sliced chunks keep their element / bit offsets, per-rank sums add up to pyarrow's."""
    rng = np.random.default_rng(43)
    whole, d, mask = _make(rng, pa.int64(), 50_000)
    cuts = [0, 7, 4_103, 4_103, 20_000, 33_333, 50_000]          # one empty chunk, odd offsets
    chunks = [whole.slice(a, b - a) for a, b in zip(cuts[:-1], cuts[1:])]
    ca = pa.chunked_array(chunks)
    total, count, seen_rows = 0, 0, 0
    for rank in range(world):
        bufs, vms, seen = mnr.arrow.from_arrow_stream(gpu_ctx, ca, rank, world)
        assert seen == len(chunks)
        lo, hi = mnr.arrow.shard_range(len(chunks), rank, world)
        assert len(bufs) == hi - lo
        for k, (b, v) in enumerate(zip(bufs, vms)):
            src = chunks[lo + k]
            assert len(b) == len(src)
            back = mnr.arrow.to_arrow(gpu_ctx, b, v)
            assert back.equals(src) or (len(src) == 0 and len(back) == 0)
            if len(b):
                st = mnr.device_ops.reduce_stats(gpu_ctx, b, v)
                total += st["sum"]; count += st["count"]
            seen_rows += len(b)
    assert seen_rows == 50_000
    wrap = lambda x: (x + 2 ** 63) % 2 ** 64 - 2 ** 63   # noqa: E731
    assert wrap(total) == wrap(int(np.sum(d[~mask].astype(object)))) and count == int((~mask).sum())


def test_stream_import_rejects_non_numeric(mnr, gpu_ctx):
    ca = pa.chunked_array([pa.array(["a", "b"]), pa.array(["c"])])
    with pytest.raises(mnr.KernelError) as ei:
        mnr.arrow.from_arrow_stream(gpu_ctx, ca)
    assert ei.value.kind == "UnsupportedType"
