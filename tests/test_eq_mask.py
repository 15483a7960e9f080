"""simd_eq_mask_u8/_u16/_u32/_u64 (src/kernels/bitmask/simd.rs:741-788): bit i = ((data[i] & field_mask) == target).
The reference has no test for these; the oracle restates the loop and is checked here against the definition in numpy,
and the CUDA kernel against the oracle (all alignment tiers, ragged lengths)."""
import numpy as np
import pytest

from oracle import oracle as orc

SIZES = [0, 1, 7, 8, 9, 63, 64, 65, 1000, 4099, 100_003]
UT = [np.uint8, np.uint16, np.uint32, np.uint64]


@pytest.mark.parametrize("dt", UT)
def test_oracle_eq_mask_is_the_definition(dt):
    rng = np.random.default_rng(51)
    for n in SIZES:
        d = rng.integers(0, 16, n, dtype=dt)
        for fm, tg in ((0xF, 3), (0x3, 1), (0, 0), (np.iinfo(dt).max, 7), (0x8, 0x8)):
            got = orc.simd_eq_mask(d, fm, tg)
            exp = ((d & dt(fm)) == dt(tg))
            assert got.len == n and np.array_equal(got.to_bools(), exp)
            assert got.bits.size == (n + 7) // 8
            if n % 8:
                assert got.bits[-1] >> (n % 8) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("dt", UT + [np.int32, np.int64])
def test_gpu_eq_mask_matches_oracle(gpu_ctx, dt):
    import minarrow_b200 as mnr
    dev = mnr.device_ops
    rng = np.random.default_rng(52)
    for n in SIZES + [(1 << 20) + 77]:
        d = rng.integers(0, 16, n, dtype=dt)
        D = mnr.DeviceBuffer.upload(gpu_ctx, d)
        for fm, tg in ((0xF, 3), (0x3, 1), (0, 0), (0x8, 0x8)):
            got = dev.eq_mask(gpu_ctx, D, fm, tg).download()
            exp = orc.simd_eq_mask(d, fm, tg)
            assert got.len == exp.len and np.array_equal(got.bits, exp.bits), (dt, n, fm, tg)
        if n > 64:   # element offsets 1, 2, 4: 128-bit and element-load tiers
            for off in (1, 2, 16 // d.itemsize):
                got = dev.eq_mask(gpu_ctx, D.slice(off, n - off - 3), 0xF, 5).download()
                assert np.array_equal(got.bits, orc.simd_eq_mask(d[off:n - 3], 0xF, 5).bits), (dt, n, off)
