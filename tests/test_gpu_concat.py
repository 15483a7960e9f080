"""Device consolidate / rechunk / bit-offset slices against the host definition.

Reference semantics: Array::concat = append_array (src/macros.rs:311-341; pinned by
src/structs/variants/integer.rs concat_tests `test_integer_array_concat[_with_nulls]`): values appended, validity
appended bit by bit, a chunk without a mask is all-valid, result mask present iff any chunk had one;
SuperArray::rechunk(Count(n)) (super_array.rs:674-787): chunks of exactly n rows + a remainder chunk;
Bitmask::slice_clone (bitmask.rs:604-626)."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mnr(gpu_ctx):
    import minarrow_b200 as m
    return m


def test_reference_concat_known_answers(mnr, gpu_ctx):
    dev = mnr.device_ops
    a = mnr.DeviceBuffer.upload(gpu_ctx, np.array([1, 2, 3], dtype=np.int32))
    b = mnr.DeviceBuffer.upload(gpu_ctx, np.array([4, 5, 6], dtype=np.int32))
    out, m = dev.concat(gpu_ctx, [a, b])
    assert m is None and out.download().tolist() == [1, 2, 3, 4, 5, 6]          # test_integer_array_concat
    x = mnr.DeviceBuffer.upload(gpu_ctx, np.array([10, 0], dtype=np.int32))
    y = mnr.DeviceBuffer.upload(gpu_ctx, np.array([0, 40], dtype=np.int32))
    mx = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools([True, False]))
    my = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools([False, True]))
    out, m = dev.concat(gpu_ctx, [x, y], [mx, my])                                # test_integer_array_concat_with_nulls
    assert out.download().tolist() == [10, 0, 0, 40] and m.download().to_bools().tolist() == [True, False, False, True]


@pytest.mark.parametrize("dt", [np.int8, np.uint16, np.int32, np.int64, np.float32, np.float64])
def test_concat_and_rechunk_ragged_chunks(mnr, gpu_ctx, dt):
    dev = mnr.device_ops
    rng = np.random.default_rng(61)
    lens = [0, 1, 7, 8, 9, 0, 63, 64, 65, 1000, 3, 3, 3, 1, 0, 4099, 70_001, 5]
    for mask_mode in ("all", "some", "none"):
        bufs, vals, hd, hv = [], [], [], []
        for i, n in enumerate(lens):
            d = (rng.standard_normal(n) * 100).astype(dt) if np.dtype(dt).kind == "f" else \
                rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, n, dtype=dt, endpoint=True)
            has = mask_mode == "all" or (mask_mode == "some" and i % 3 != 1)
            v = rng.random(n) < 0.8
            bufs.append(mnr.DeviceBuffer.upload(gpu_ctx, d))
            vals.append(mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(v)) if has else None)
            hd.append(d)
            hv.append(v if has else np.ones(n, bool))
        whole, wv = np.concatenate(hd), np.concatenate(hv)
        out, m = dev.concat(gpu_ctx, bufs, vals)
        assert out.download().tobytes() == whole.tobytes()
        if mask_mode == "none":
            assert m is None
        else:
            got = m.download()
            assert got.len == whole.size and np.array_equal(got.bits, np.packbits(wv, bitorder="little"))
        for chunk_rows in (1000, 8192, 7, whole.size, whole.size + 5):
            cb, cv = dev.rechunk(gpu_ctx, bufs, vals, chunk_rows)
            assert [len(b) for b in cb] == [min(chunk_rows, whole.size - r) for r in range(0, whole.size, chunk_rows)]
            r0 = 0
            for b, v in zip(cb, cv):
                ln = len(b)
                assert b.download().tobytes() == whole[r0:r0 + ln].tobytes()
                if mask_mode == "none":
                    assert v is None
                else:
                    assert np.array_equal(v.download().bits, np.packbits(wv[r0:r0 + ln], bitorder="little"))
                r0 += ln
            if chunk_rows == 7:
                break   # thousands of tiny downloads: once is enough
        # aggregates are invariant under rechunking (the reason rechunk may run before sharding)
        st0 = dev.reduce_stats(gpu_ctx, out, m)
        cb, cv = dev.rechunk(gpu_ctx, bufs, vals, 8192)
        parts = dev.reduce_stats_batch(gpu_ctx, cb, cv if mask_mode != "none" else None, True)
        assert sum(p["count"] for p in parts) == st0["count"]
        if np.dtype(dt).kind != "f":
            tot = sum(p["sum"] for p in parts)
            assert (tot - st0["sum"]) % 2 ** 64 == 0


def test_bits_slice_is_slice_clone(mnr, gpu_ctx):
    dev = mnr.device_ops
    rng = np.random.default_rng(62)
    n = 100_003
    b = rng.random(n) < 0.5
    B = mnr.DeviceBitmask.upload(gpu_ctx, mnr.Bitmask.from_bools(b))
    for off, ln in ((0, n), (1, n - 1), (7, 64), (8, 8), (63, 1), (64, 0), (12345, 54321), (n, 0), (n - 1, 1)):
        got = dev.bits_slice(gpu_ctx, B, off, ln).download()
        assert got.len == ln and np.array_equal(got.bits, np.packbits(b[off:off + ln], bitorder="little")), (off, ln)
    with pytest.raises(mnr.KernelError) as ei:
        dev.bits_slice(gpu_ctx, B, n - 3, 10)
    assert ei.value.kind == "OutOfBounds"
    with pytest.raises(mnr.KernelError) as ei:
        dev.concat(gpu_ctx, [mnr.DeviceBuffer.upload(gpu_ctx, np.zeros(3, np.int32)),
                             mnr.DeviceBuffer.upload(gpu_ctx, np.zeros(3, np.int64))])
    assert ei.value.kind == "TypeMismatch"
