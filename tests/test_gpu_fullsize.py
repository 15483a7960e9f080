"""BASELINE.json's full sizes, through size-independent properties and (where the oracle finishes in seconds with all
host cores) direct parity.  Inputs are generated on the device with torch so the suite stays within a minute or two.

C2: 1e9-row i64: bench self-check sum(0..1e9) = 499 999 999 500 000 000 (benches/benchmark_parallel_simd.rs:109-112);
    null-aware: sum(valid) + sum(!valid) == sum(all) (wrapping), count(valid) + count(!valid) == N.
C3: 2 x 256 Mi-row f64 with two validity masks: add / mul / div bit-exact vs the oracle leaf over the AND-merged mask.
C4: 4 Gi-bit masks: De Morgan, involution, popcount identities, trailing-bit hygiene at a ragged length."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = [pytest.mark.gpu, pytest.mark.heavy]


@pytest.fixture(scope="module")
def env(gpu_ctx):
    import torch
    import minarrow_b200 as mnr
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs 40 GB of free HBM")
    return torch, mnr, gpu_ctx


def _wrap_buf(mnr, ctx, t, npdt):
    # torch fills its tensors on torch's stream; the context launches on its own: order them here
    import torch
    torch.cuda.synchronize()
    return mnr.DeviceBuffer.wrap(ctx, npdt, t.data_ptr(), t.numel(), t)


def _wrap_bits(mnr, ctx, t, nbits):
    import torch
    torch.cuda.synchronize()
    return mnr.DeviceBitmask.wrap(ctx, t.data_ptr(), nbits, t)


def test_c2_one_billion_row_i64_sum(env):
    torch, mnr, ctx = env
    dev = mnr.device_ops
    n = 1_000_000_000
    data = torch.arange(n, dtype=torch.int64, device="cuda")
    B = _wrap_buf(mnr, ctx, data, np.int64)
    s, c = dev.reduce_sum(ctx, B)
    assert (s, c) == (499_999_999_500_000_000, n)
    st = dev.reduce_stats(ctx, B)
    assert (st["min"], st["max"], st["count"]) == (0, n - 1, n) and st["mean"] == (n - 1) / 2
    g = torch.Generator(device="cuda").manual_seed(3)
    vbytes = torch.randint(0, 256, ((n + 7) // 8,), dtype=torch.uint8, device="cuda", generator=g) | \
        torch.randint(0, 256, ((n + 7) // 8,), dtype=torch.uint8, device="cuda", generator=g)
    V = mnr.DeviceBitmask.upload(ctx, mnr.Bitmask(vbytes.cpu().numpy(), n))     # upload clears the slack bits
    NV = dev.bits_not(ctx, V, 0, n)
    s1, c1 = dev.reduce_sum(ctx, B, V)
    s0, c0 = dev.reduce_sum(ctx, B, NV)
    assert c1 + c0 == n and c1 == dev.bits_popcount(ctx, V, 0, n)
    assert (s1 + s0) % 2 ** 64 == 499_999_999_500_000_000
    # full-range values: the sum wraps exactly like the oracle's on a 2^26-row sample, and the two halves add up
    data.random_(-2 ** 63, 2 ** 63 - 1, generator=g)
    torch.cuda.synchronize()
    sa, _ = dev.reduce_sum(ctx, B)
    s1, _ = dev.reduce_sum(ctx, B, V)
    s0, _ = dev.reduce_sum(ctx, B, NV)
    assert (s1 + s0 - sa) % 2 ** 64 == 0
    m = 1 << 26
    hs = data[:m].cpu().numpy()
    hv = orc.Bits(vbytes[: m // 8].cpu().numpy(), m)
    assert dev.reduce_sum(ctx, B.slice(0, m), _wrap_bits(mnr, ctx, vbytes, m)) == orc.par_masked_sum_i64(hs, hv, 16)


def test_c3_f64_256Mi_rows_two_masks_bit_exact(env):
    torch, mnr, ctx = env
    dev = mnr.device_ops
    n = 1 << 28
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    idx = torch.randint(0, n, (n // 10_000,), device="cuda", generator=g)
    y[idx] = 0.0                      # zero divisors -> +-Inf / NaN stay valid (src/kernels/arithmetic/mod.rs:342-354)
    x[idx[::3]] = float("nan")
    x[idx[1::3]] = float("inf")
    y[idx[2::7]] = -0.0
    torch.cuda.synchronize()
    mx = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device="cuda", generator=g) | \
        torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device="cuda", generator=g)
    my = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device="cuda", generator=g) | \
        torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device="cuda", generator=g)
    X, Y = _wrap_buf(mnr, ctx, x, np.float64), _wrap_buf(mnr, ctx, y, np.float64)
    MX, MY = _wrap_bits(mnr, ctx, mx, n), _wrap_bits(mnr, ctx, my, n)
    hx, hy = x.cpu().numpy(), y.cpu().numpy()
    merged = orc.Bits((mx & my).cpu().numpy(), n)
    out, om = np.empty(n, dtype=np.float64), np.zeros(n // 8, dtype=np.uint8)
    for op in (orc.ADD, orc.MUL, orc.DIV):
        ob, obm = dev.ew_binary(ctx, op, X, Y, MX, MY, mnr.MaskMode.And)
        orc.par_apply_float_f64(hx, hy, op, merged, 16, out, om)
        got = ob.download()
        assert got.tobytes() == out.tobytes(), f"op {op}: values differ from the oracle"
        assert np.array_equal(obm.download().bits, om), f"op {op}: validity differs"
        del ob, obm, got
    # scalar broadcast both sides == the materialised broadcast the reference performs (routing/broadcast.rs:25-47)
    ob, obm = dev.ew_scalar(ctx, orc.MUL, X, 2.5, False, MX)
    orc.par_apply_float_f64(hx, np.full(n, 2.5), orc.MUL, orc.Bits(mx.cpu().numpy(), n), 16, out, om)
    assert ob.download().tobytes() == out.tobytes() and np.array_equal(obm.download().bits, om)


def test_c4_four_gibit_masks(env):
    torch, mnr, ctx = env
    dev = mnr.device_ops
    L = mnr.LogicalOperator
    n = 1 << 32
    g = torch.Generator(device="cuda").manual_seed(8)
    a = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device="cuda", generator=g)
    b = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device="cuda", generator=g)
    A, B = _wrap_bits(mnr, ctx, a, n), _wrap_bits(mnr, ctx, b, n)
    pa, pb = dev.bits_popcount(ctx, A, 0, n), dev.bits_popcount(ctx, B, 0, n)
    assert abs(pa - n // 2) < 1 << 20
    AND, OR, XOR = (dev.bits_binop(ctx, op, A, 0, B, 0, n) for op in (L.And, L.Or, L.Xor))
    p_and, p_or, p_xor = (dev.bits_popcount(ctx, m, 0, n) for m in (AND, OR, XOR))
    assert p_and + p_or == pa + pb and p_xor == p_or - p_and
    NA = dev.bits_not(ctx, A, 0, n)
    assert dev.bits_popcount(ctx, NA, 0, n) == n - pa                 # null_count = len - ones
    assert dev.bits_all_eq(ctx, dev.bits_not(ctx, NA, 0, n), 0, A, 0, n)          # involution
    NB = dev.bits_not(ctx, B, 0, n)
    assert dev.bits_all_eq(ctx, dev.bits_not(ctx, AND, 0, n), 0, dev.bits_binop(ctx, L.Or, NA, 0, NB, 0, n), 0, n)   # De Morgan
    # spot parity of the first 2^27 bits against the oracle's word loop
    m = 1 << 27
    ha, hb = orc.Bits(a[: m // 8].cpu().numpy(), m), orc.Bits(b[: m // 8].cpu().numpy(), m)
    got = dev.bits_binop(ctx, L.And, A, 0, B, 0, m).download()
    assert np.array_equal(got.bits, orc.and_masks((ha, 0, m), (hb, 0, m)).bits)
    # ragged length: slack bits of the last byte must be zero, popcount must ignore them
    r = n - 37
    Ar = dev.bits_not(ctx, A, 0, r)
    last = Ar.download().bits[-1]
    assert last >> (r % 8) == 0
    assert dev.bits_popcount(ctx, Ar, 0, r) + dev.bits_popcount(ctx, A, 0, r) == r
