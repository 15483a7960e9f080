"""Tuning experiment: launch geometry of 64-bit `column / scalar` (ctx option ew_sdiv64_cfg, elementwise.cu CfgSdiv64*).
1 = 128 thr x 4 x 256-bit, <= 128 regs (CfgCheap); 2 = 128 thr x 2 x 256-bit, <= 85 regs (CfgSdiv64); 0 = the library's
choice (masked -> 2, dense -> 1).  The r01zz run also had 256-thread / 64-register variants (dropped, see elementwise.cu).
Usage: python tools/sdiv64_exp.py"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import minarrow_b200 as mnr
from bench import event_time_ms
dev = torch.device("cuda:0"); ctx = mnr.Context(0, torch.cuda.current_stream().cuda_stream); ops = mnr.device_ops
A = mnr.ArithmeticOperator
g = torch.Generator(device=dev); g.manual_seed(1)
n = (1 << 30) // 8
for name in ("int64", "uint64"):
    x = torch.randint(-2 ** 62, 2 ** 62, (n,), dtype=torch.int64, device=dev, generator=g)
    if name == "uint64":
        x = x.abs_()
    o = torch.empty_like(x)
    m1 = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device=dev, generator=g); om = torch.empty_like(m1)
    X = mnr.DeviceBuffer.wrap(ctx, np.dtype(name), x.data_ptr(), n, x); O = mnr.DeviceBuffer.wrap(ctx, np.dtype(name), o.data_ptr(), n, o)
    M1 = mnr.DeviceBitmask.wrap(ctx, m1.data_ptr(), n, m1); OM = mnr.DeviceBitmask.wrap(ctx, om.data_ptr(), n, om)
    ref = None
    for cfg in (1, 2, 0):
        ctx.set_option("ew_sdiv64_cfg", cfg)
        for label, fn, nb in (("div masked", lambda: ops.ew_scalar_into(ctx, A.Divide, X, 86400, False, M1, O, OM), n * 16.25),
                              ("floordiv masked", lambda: ops.ew_scalar_into(ctx, A.FloorDiv, X, 86400, False, M1, O, OM), n * 16.25),
                              ("div dense", lambda: ops.ew_scalar_into(ctx, A.Divide, X, 86400, False, None, O, None), n * 16.0)):
            med, _ = event_time_ms(torch, fn, 15)
            print(f"{name:7s} cfg={cfg} {label:16s} {med:8.4f} ms {nb / med / 1e6:8.1f} GB/s", flush=True)
        ops.ew_scalar_into(ctx, A.FloorDiv, X, -86400 if name == "int64" else 86400, False, M1, O, OM)
        torch.cuda.synchronize()
        chk = int(o.sum().item()) ^ int(om.to(torch.int64).sum().item())
        ref = chk if ref is None else ref
        print(f"{name:7s} cfg={cfg} checksum {'same' if chk == ref else 'DIFFERENT'}", flush=True)
    ctx.set_option("ew_sdiv64_cfg", 0)
