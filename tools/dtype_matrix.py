"""Throughput of every (kernel, dtype) on the path at a fixed column size (default 1 GiB per operand), one line each.

Measurement only (tuning tool): CUDA events around each launch on the context stream, median of 15 after 3 warm-ups,
inputs far larger than the 126 MB L2.  Algorithmic bytes follow SURVEY.md §8d (validity at 1 bit/row).

    python tools/dtype_matrix.py [--gib 1.0] [--out gpurun_out/matrix.md]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="", help="comma-separated substrings of kernel labels to run")
    ap.add_argument("--iters", type=int, default=15, help="timed launches per entry (median); 1 for ncu captures")
    ap.add_argument("--dtypes", default="", help="comma-separated numpy dtype names (default: all ten)")
    args = ap.parse_args()

    import numpy as np
    import torch

    import minarrow_b200 as mnr
    from bench import event_time_ms, peaks

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    ctx = mnr.Context(0, torch.cuda.current_stream().cuda_stream)
    ops = mnr.device_ops
    A = mnr.ArithmeticOperator
    peak, peak_src = peaks()
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    rows_out = []

    tdt = {"int8": torch.int8, "uint8": torch.uint8, "int16": torch.int16, "uint16": torch.uint16, "int32": torch.int32,
           "uint32": torch.uint32, "int64": torch.int64, "uint64": torch.uint64, "float32": torch.float32,
           "float64": torch.float64}

    def column(name, n, nonzero=False):
        """Random device column of numpy dtype `name` (torch storage of the same width, reinterpreted)."""
        nd = np.dtype(name)
        if nd.kind == "f":
            t = torch.randn(n, dtype=tdt[name], device=dev, generator=g)
        else:
            carrier = {1: torch.int8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[nd.itemsize]
            lo, hi = (-100, 100) if nd.itemsize == 1 else (-30000, 30000)
            t = torch.randint(lo, hi, (n,), dtype=carrier, device=dev, generator=g)
            if nd.kind == "u":
                t = t.abs_()
            if nonzero:
                t = t | 1
        return t, mnr.DeviceBuffer.wrap(ctx, nd, t.data_ptr(), n, t)

    def bitmask(n):
        t = torch.randint(0, 256, ((n + 7) // 8,), dtype=torch.uint8, device=dev, generator=g) | \
            torch.randint(0, 256, ((n + 7) // 8,), dtype=torch.uint8, device=dev, generator=g)
        return t, mnr.DeviceBitmask.wrap(ctx, t.data_ptr(), n, t)

    def entry(kernel, dtype, nbytes, fn):
        if args.only and not any(k in kernel for k in args.only.split(",")):
            return
        med, best = event_time_ms(torch, fn, args.iters, warmup=min(3, args.iters))
        gbs = nbytes / med / 1e6
        rows_out.append({"kernel": kernel, "dtype": dtype, "GB/s": round(gbs, 1), "frac": round(gbs / peak, 3),
                         "ms": round(med, 4), "bytes": int(nbytes)})
        print(f"{kernel:34s} {dtype:8s} {med:8.4f} ms {gbs:8.1f} GB/s  {gbs / peak:5.3f} of peak", flush=True)

    part = torch.zeros(4, dtype=torch.int64, device=dev)
    for name in ("int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64", "float32", "float64"):
        if args.dtypes and name not in args.dtypes.split(","):
            continue
        sz = np.dtype(name).itemsize
        n = int(args.gib * (1 << 30)) // sz
        n -= n % 64
        tx, X = column(name, n)
        ty, Y = column(name, n, nonzero=True)
        to = torch.empty_like(tx)
        O = mnr.DeviceBuffer.wrap(ctx, np.dtype(name), to.data_ptr(), n, to)
        tmx, MX = bitmask(n)
        tmy, MY = bitmask(n)
        tom = torch.empty_like(tmx)
        OM = mnr.DeviceBitmask.wrap(ctx, tom.data_ptr(), n, tom)
        entry("reduce sum+count masked", name, n * (sz + 0.125), lambda: ops.reduce_stats_async(ctx, X, MX, False, part.data_ptr()))
        entry("reduce sum+min+max+count masked", name, n * (sz + 0.125), lambda: ops.reduce_stats_async(ctx, X, MX, True, part.data_ptr()))
        entry("reduce sum dense", name, n * sz, lambda: ops.reduce_stats_async(ctx, X, None, False, part.data_ptr()))
        entry("reduce sum+min+max dense", name, n * sz, lambda: ops.reduce_stats_async(ctx, X, None, True, part.data_ptr()))
        for opn, op in (("add", A.Add), ("mul", A.Multiply), ("div", A.Divide), ("rem", A.Remainder), ("floordiv", A.FloorDiv)):
            entry(f"ew {opn} two masks", name, n * (3 * sz + 0.375),
                  lambda op=op: ops.ew_binary_into(ctx, op, X, Y, MX, MY, mnr.MaskMode.And, O, OM))
        # Power: exponents 0..7 (the reference's own tests use 2 and 3, arithmetic/mod.rs:507-537); integer pow is a
        # square-and-multiply loop whose trip count is log2(exponent)
        if np.dtype(name).kind == "f":
            te = torch.randint(0, 8, (n,), dtype=torch.int32, device=dev, generator=g).to(tdt[name])
        else:
            te = torch.randint(0, 8, (n,), dtype={1: torch.int8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[sz], device=dev, generator=g)
        E = mnr.DeviceBuffer.wrap(ctx, np.dtype(name), te.data_ptr(), n, te)
        entry("ew pow two masks (exponents 0..7)", name, n * (3 * sz + 0.375),
              lambda: ops.ew_binary_into(ctx, A.Power, X, E, MX, MY, mnr.MaskMode.And, O, OM))
        if np.dtype(name).kind == "f":
            # the same launch on |x|: float Power is exp(b ln a), NaN for every negative base — a column of positive bases
            # is the case that computes something
            tp = tx.abs() + 0.5
            P = mnr.DeviceBuffer.wrap(ctx, np.dtype(name), tp.data_ptr(), n, tp)
            entry("ew pow two masks, positive bases", name, n * (3 * sz + 0.375),
                  lambda: ops.ew_binary_into(ctx, A.Power, P, E, MX, MY, mnr.MaskMode.And, O, OM))
            del P, tp
        del E, te
        entry("ew add dense", name, n * 3 * sz, lambda: ops.ew_binary_into(ctx, A.Add, X, Y, None, None, mnr.MaskMode.And, O, None))
        entry("ew scalar add masked", name, n * (2 * sz + 0.25), lambda: ops.ew_scalar_into(ctx, A.Add, X, 3, False, MX, O, OM))
        entry("ew scalar div masked", name, n * (2 * sz + 0.25), lambda: ops.ew_scalar_into(ctx, A.Divide, X, 3, False, MX, O, OM))
        if np.dtype(name).kind == "f":
            tz, Z = column(name, n)
            entry("fma masked", name, n * (4 * sz + 0.25), lambda: ops.ew_fma_into(ctx, X, Y, Z, MX, O, OM))
            del tz, Z
        if np.dtype(name).kind == "u":
            def eqm():
                ops.eq_mask(ctx, X, 3, 1).free()
            entry("eq_mask (fresh output)", name, n * (sz + 0.125), eqm)
        if name in ("float32", "float64"):
            ti, I = column("int32", n)

            def prom():
                ob, om = ops.ew_binary_promote(ctx, A.Add, I, X, MX, MY, mnr.MaskMode.And)
                ob.free()
                om.free()
            entry("ew add i32 (cast on load) + T", name, n * (4 + 2 * sz + 0.375), prom)
            del ti, I
        # consolidate (device concat): 64 chunks with validity; aligned chunk lengths, then ragged ones (rows % 8 != 0, so
        # every chunk after the first lands on an odd element and an odd bit offset)
        if name in ("int8", "int32", "int64", "float64"):
            for label, rows in (("aligned", (n // 64) - (n // 64) % 64), ("ragged", (n // 64) - (n // 64) % 64 - 3)):
                cb = [X.slice(k * (n // 64), rows) for k in range(64)]
                cm = [ops.bits_slice(ctx, MX, k * (n // 64) - (k * (n // 64)) % 64, rows) for k in range(64)]

                def cat():
                    ob, om = ops.concat(ctx, cb, cm)
                    ob.free()
                    om.free()
                entry(f"concat 64 chunks + validity ({label})", name, 64 * rows * 2 * (sz + 0.125), cat)
                del cb, cm
        del X, Y, O, MX, MY, OM, tx, ty, to, tmx, tmy, tom
        torch.cuda.empty_cache()

    # bitmask family at 4 Gi bits
    nb = 1 << 32
    ta, Ab = bitmask(nb)
    tb, Bb = bitmask(nb)
    tr = torch.empty_like(ta)
    Rb = mnr.DeviceBitmask.wrap(ctx, tr.data_ptr(), nb, tr)
    L = mnr.LogicalOperator
    for opn, op in (("and", L.And), ("or", L.Or), ("xor", L.Xor)):
        entry(f"bits {opn}", "bit", nb * 3 / 8, lambda op=op: ops.bits_binop_into(ctx, op, Ab, 0, Bb, 0, nb, Rb))
    tr2 = torch.empty((nb - 64) // 8, dtype=torch.uint8, device=dev)
    Rb2 = mnr.DeviceBitmask.wrap(ctx, tr2.data_ptr(), nb - 64, tr2)
    entry("bits and (windows at bytes 1/2)", "bit", (nb - 64) * 3 / 8, lambda: ops.bits_binop_into(ctx, L.And, Ab, 8, Bb, 16, nb - 64, Rb2))
    entry("bits not", "bit", nb * 2 / 8, lambda: ops.bits_not_into(ctx, Ab, 0, nb, Rb))
    entry("bits popcount (sync API)", "bit", nb / 8, lambda: ops.bits_popcount(ctx, Ab, 0, nb))
    entry("bits all_eq (sync API)", "bit", nb * 2 / 8, lambda: ops.bits_all_eq(ctx, Ab, 0, Ab, 0, nb))

    def sl():
        ops.bits_slice(ctx, Ab, 13, nb - 64).free()
    entry("bits slice (bit offset 13, fresh)", "bit", (nb - 64) * 2 / 8, sl)

    out = {"peak": peak, "peak_source": peak_src, "gib_per_operand": args.gib, "rows": rows_out}
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            f.write(f"# (kernel, dtype) throughput matrix — {args.gib} GiB per operand, median of 15, peak = {peak} GB/s ({peak_src})\n\n")
            f.write("| kernel | dtype | ms | GB/s (algorithmic) | of peak |\n|---|---|---|---|---|\n")
            for r in rows_out:
                f.write(f"| {r['kernel']} | {r['dtype']} | {r['ms']} | {r['GB/s']} | {r['frac']} |\n")
        with open(os.path.splitext(args.out)[0] + ".json", "w") as f:
            json.dump(out, f)


if __name__ == "__main__":
    main()
