"""Tuning experiment: launch geometry of masked add / mul on 1-byte columns (ctx option ew_cheap8_cfg: 1 = CfgCheap 128 thr x 4 x
256-bit <= 128 regs covering; 2 = 128 thr x 2 x 256-bit <= 85 regs covering; 3 = 256 thr x 2 x 256-bit <= 85 regs resident).
Usage: python tools/cheap8_exp.py"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import minarrow_b200 as mnr
from bench import event_time_ms
dev = torch.device("cuda:0"); ctx = mnr.Context(0, torch.cuda.current_stream().cuda_stream); ops = mnr.device_ops
A = mnr.ArithmeticOperator
g = torch.Generator(device=dev); g.manual_seed(1)
for name in (sys.argv[1:] or ["int8", "uint8"]):
    nd = np.dtype(name); n = (1 << 30) // nd.itemsize
    carrier = {1: torch.int8, 2: torch.int16}[nd.itemsize]
    x = torch.randint(-100, 100, (n,), dtype=carrier, device=dev, generator=g); y = torch.randint(-100, 100, (n,), dtype=carrier, device=dev, generator=g)
    o = torch.empty_like(x)
    m1 = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device=dev, generator=g); m2 = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device=dev, generator=g)
    om = torch.empty_like(m1)
    W = lambda t: mnr.DeviceBuffer.wrap(ctx, nd, t.data_ptr(), n, t)
    B = lambda t: mnr.DeviceBitmask.wrap(ctx, t.data_ptr(), n, t)
    X, Y, O, M1, M2, OM = W(x), W(y), W(o), B(m1), B(m2), B(om)
    cases = [("add two masks", lambda: ops.ew_binary_into(ctx, A.Add, X, Y, M1, M2, mnr.MaskMode.And, O, OM), n * (3 * nd.itemsize + 0.375)),
             ("mul two masks", lambda: ops.ew_binary_into(ctx, A.Multiply, X, Y, M1, M2, mnr.MaskMode.And, O, OM), n * (3 * nd.itemsize + 0.375)),
             ("add one mask", lambda: ops.ew_binary_into(ctx, A.Add, X, Y, M1, None, mnr.MaskMode.And, O, OM), n * (3 * nd.itemsize + 0.25)),
             ("scalar add masked", lambda: ops.ew_scalar_into(ctx, A.Add, X, 3, False, M1, O, OM), n * (2 * nd.itemsize + 0.25))]
    ref = {}
    for cfg in (1, 2, 3):
        ctx.set_option("ew_cheap8_cfg", cfg)
        for label, fn, nb in cases:
            med, _ = event_time_ms(torch, fn, 11)
            torch.cuda.synchronize()
            chk = int(o.to(torch.int64).sum().item()) ^ int(om.to(torch.int64).sum().item())
            same = "same" if ref.setdefault(label, chk) == chk else "DIFFERENT"
            print(f"{name:6s} cfg={cfg} {label:18s} {med:8.4f} ms {nb / med / 1e6:8.1f} GB/s  checksum {same}", flush=True)
    ctx.set_option("ew_cheap8_cfg", 0)
