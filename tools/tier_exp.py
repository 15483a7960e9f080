"""Tuning experiment: element-wise add per dtype x mask shape at the 256-bit and 128-bit tiers (ctx option ew_max_tier).
Usage: python tools/tier_exp.py [dtype-filter] [label-filter]"""
import sys, numpy as np, torch
DT_F = sys.argv[1] if len(sys.argv) > 1 else ""
LB_F = sys.argv[2] if len(sys.argv) > 2 else ""
sys.path.insert(0, '.')
import minarrow_b200 as mnr
from bench import event_time_ms
dev = torch.device("cuda:0"); ctx = mnr.Context(0, torch.cuda.current_stream().cuda_stream); ops = mnr.device_ops
A = mnr.ArithmeticOperator
g = torch.Generator(device=dev); g.manual_seed(1)
for name, sz, tdt in (("int8", 1, torch.int8), ("int16", 2, torch.int16), ("int32", 4, torch.int32), ("float64", 8, torch.float64)):
    if DT_F and DT_F != name:
        continue
    n = (1 << 30) // sz
    mk = lambda: (torch.randn(n, dtype=tdt, device=dev, generator=g) if tdt.is_floating_point else torch.randint(-100, 100, (n,), dtype=tdt, device=dev, generator=g))
    x, y = mk(), mk(); o = torch.empty_like(x)
    m1 = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device=dev, generator=g); m2 = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device=dev, generator=g); om = torch.empty_like(m1)
    W = lambda t: mnr.DeviceBuffer.wrap(ctx, np.dtype(name), t.data_ptr(), n, t)
    B = lambda t: mnr.DeviceBitmask.wrap(ctx, t.data_ptr(), n, t)
    X, Y, O, M1, M2, OM = W(x), W(y), W(o), B(m1), B(m2), B(om)
    for tier in (2, 1):
        ctx.set_option("ew_max_tier", tier)
        for label, fn, nb in (("two masks", lambda: ops.ew_binary_into(ctx, A.Add, X, Y, M1, M2, mnr.MaskMode.And, O, OM), n * (3 * sz + 0.375)),
                              ("one mask", lambda: ops.ew_binary_into(ctx, A.Add, X, Y, M1, None, mnr.MaskMode.And, O, OM), n * (3 * sz + 0.25)),
                              ("dense", lambda: ops.ew_binary_into(ctx, A.Add, X, Y, None, None, mnr.MaskMode.And, O, None), n * 3 * sz),
                              ("scalar masked", lambda: ops.ew_scalar_into(ctx, A.Add, X, 3, False, M1, O, OM), n * (2 * sz + 0.25))):
            if LB_F and LB_F != label:
                continue
            med, _ = event_time_ms(torch, fn, 15)
            print(f"{name:8s} tier={tier} add {label:14s} {med:8.4f} ms {nb / med / 1e6:8.1f} GB/s", flush=True)
    ctx.set_option("ew_max_tier", 2)
    del X, Y, O, M1, M2, OM, x, y, o, m1, m2, om
    torch.cuda.empty_cache()
