"""Tuning experiment: launch geometry of float Div / FloorDiv (ctx option ew_fdiv_cfg, elementwise.cu CfgFdiv*).
1 = 256 thr x 2 x 128-bit, <= 64 regs, resident grid (CfgHeavy); 2 = 128 thr x 2 x 256-bit, <= 85 regs, covering grid
(CfgFdiv2); 3 = 256 thr x 2 x 256-bit, <= 85 regs, resident (CfgFdiv3); 0 = the library's per-shape choice.  The r01zz
run numbered them 0 / 2 / 3 and also had CfgCheap (1) and CfgHeavy with a covering grid (4), both dropped.
Usage: python tools/fdiv_exp.py"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import minarrow_b200 as mnr
from bench import event_time_ms
dev = torch.device("cuda:0"); ctx = mnr.Context(0, torch.cuda.current_stream().cuda_stream); ops = mnr.device_ops
A = mnr.ArithmeticOperator
g = torch.Generator(device=dev); g.manual_seed(1)
for name, tdt, sz in (("float64", torch.float64, 8), ("float32", torch.float32, 4)):
    n = (1 << 30) // sz
    x = torch.randn(n, dtype=tdt, device=dev, generator=g); y = torch.randn(n, dtype=tdt, device=dev, generator=g); o = torch.empty_like(x)
    m1 = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device=dev, generator=g); m2 = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device=dev, generator=g)
    om = torch.empty_like(m1)
    W = lambda t: mnr.DeviceBuffer.wrap(ctx, np.dtype(name), t.data_ptr(), n, t)
    B = lambda t: mnr.DeviceBitmask.wrap(ctx, t.data_ptr(), n, t)
    X, Y, O, M1, M2, OM = W(x), W(y), W(o), B(m1), B(m2), B(om)
    ref = None
    for cfg in (1, 2, 3, 0):
        ctx.set_option("ew_fdiv_cfg", cfg)
        for label, fn, nb in (("scalar div masked", lambda: ops.ew_scalar_into(ctx, A.Divide, X, 2.5, False, M1, O, OM), n * (2 * sz + 0.25)),
                              ("scalar div dense", lambda: ops.ew_scalar_into(ctx, A.Divide, X, 2.5, False, None, O, None), n * 2 * sz),
                              ("div two masks", lambda: ops.ew_binary_into(ctx, A.Divide, X, Y, M1, M2, mnr.MaskMode.And, O, OM), n * (3 * sz + 0.375)),
                              ("div dense", lambda: ops.ew_binary_into(ctx, A.Divide, X, Y, None, None, mnr.MaskMode.And, O, None), n * 3 * sz)):
            med, _ = event_time_ms(torch, fn, 15)
            print(f"{name:7s} cfg={cfg} {label:18s} {med:8.4f} ms {nb / med / 1e6:8.1f} GB/s", flush=True)
        ops.ew_binary_into(ctx, A.FloorDiv, X, Y, M1, M2, mnr.MaskMode.And, O, OM)
        torch.cuda.synchronize()
        chk = (int(o.view(torch.int64 if sz == 8 else torch.int32).to(torch.int64).sum().item()), int(om.to(torch.int64).sum().item()))
        ref = chk if ref is None else ref
        print(f"{name:7s} cfg={cfg} checksum {'same' if chk == ref else 'DIFFERENT'}", flush=True)
    ctx.set_option("ew_fdiv_cfg", 0)
    del X, Y, O, M1, M2, OM, x, y, o, m1, m2, om
    torch.cuda.empty_cache()
