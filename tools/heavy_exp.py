"""Tuning experiment: launch geometry of the heavy element-wise classes — integer Div / Rem / FloorDiv of two columns,
float Rem, Power — (ctx option ew_heavy_cfg: 1 = CfgHeavy 256 thr x 2 x 128-bit <= 64 regs resident; 2 = 128 thr x 2 x
256-bit <= 85 regs covering; 3 = 256 thr x 2 x 256-bit <= 85 regs resident; 0 = the library's choice.  The r01zz run
still numbered CfgHeavy 0).  Usage: python tools/heavy_exp.py [dtype,dtype,...]"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import minarrow_b200 as mnr
from bench import event_time_ms
dev = torch.device("cuda:0"); ctx = mnr.Context(0, torch.cuda.current_stream().cuda_stream); ops = mnr.device_ops
A = mnr.ArithmeticOperator
g = torch.Generator(device=dev); g.manual_seed(1)
carrier = {1: torch.int8, 2: torch.int16, 4: torch.int32, 8: torch.int64}
NAMES = sys.argv[1].split(",") if len(sys.argv) > 1 else ("int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64", "float32", "float64")
for name in NAMES:
    nd = np.dtype(name); sz = nd.itemsize
    n = (1 << 30) // sz
    if nd.kind == "f":
        tdt = torch.float64 if sz == 8 else torch.float32
        x = torch.randn(n, dtype=tdt, device=dev, generator=g); y = torch.randn(n, dtype=tdt, device=dev, generator=g)
        e = torch.randint(0, 8, (n,), dtype=torch.int32, device=dev, generator=g).to(tdt)
    else:
        lo, hi = (-100, 100) if sz == 1 else (-30000, 30000)
        x = torch.randint(lo, hi, (n,), dtype=carrier[sz], device=dev, generator=g); y = torch.randint(lo, hi, (n,), dtype=carrier[sz], device=dev, generator=g) | 1
        if nd.kind == "u":
            x = x.abs_(); y = y.abs_()
        e = torch.randint(0, 8, (n,), dtype=carrier[sz], device=dev, generator=g)
    o = torch.empty_like(x)
    m1 = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device=dev, generator=g); m2 = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device=dev, generator=g)
    om = torch.empty_like(m1)
    W = lambda t: mnr.DeviceBuffer.wrap(ctx, nd, t.data_ptr(), n, t)
    B = lambda t: mnr.DeviceBitmask.wrap(ctx, t.data_ptr(), n, t)
    X, Y, E, O, M1, M2, OM = W(x), W(y), W(e), W(o), B(m1), B(m2), B(om)
    cases = [("rem two masks", lambda: ops.ew_binary_into(ctx, A.Remainder, X, Y, M1, M2, mnr.MaskMode.And, O, OM), n * (3 * sz + 0.375)),
             ("pow two masks", lambda: ops.ew_binary_into(ctx, A.Power, X, E, M1, M2, mnr.MaskMode.And, O, OM), n * (3 * sz + 0.375))]
    if nd.kind != "f":
        cases = [("div two masks", lambda: ops.ew_binary_into(ctx, A.Divide, X, Y, M1, M2, mnr.MaskMode.And, O, OM), n * (3 * sz + 0.375)),
                 ("floordiv two masks", lambda: ops.ew_binary_into(ctx, A.FloorDiv, X, Y, M1, M2, mnr.MaskMode.And, O, OM), n * (3 * sz + 0.375)),
                 ("div dense", lambda: ops.ew_binary_into(ctx, A.Divide, X, Y, None, None, mnr.MaskMode.And, O, None), n * 3 * sz)] + cases
    ref = {}
    for cfg in ((1, 2, 3, 4, 5, 0) if name == "float64" else (1, 2, 3, 0)):
        ctx.set_option("ew_heavy_cfg", cfg)
        for label, fn, nb in cases:
            med, _ = event_time_ms(torch, fn, 11)
            torch.cuda.synchronize()
            chk = int(o.view(carrier[sz]).to(torch.int64).sum().item())
            same = "same" if ref.setdefault(label, chk) == chk else "DIFFERENT"
            print(f"{name:7s} cfg={cfg} {label:18s} {med:8.4f} ms {nb / med / 1e6:8.1f} GB/s  checksum {same}", flush=True)
    ctx.set_option("ew_heavy_cfg", 0)
    del X, Y, E, O, M1, M2, OM, x, y, e, o, m1, m2, om
    torch.cuda.empty_cache()
