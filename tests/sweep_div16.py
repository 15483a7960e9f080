"""Whole-domain check of the 16-bit integer division path (narrow_quot, minarrow_b200/csrc/ew_kernels.cuh): all 2^32
(dividend, divisor) pairs of int16 and of uint16 against the CPU oracle, Div on every block, Rem / FloorDiv on
alternating blocks.  ~1.5 min on the GPU box; run by tools/gpu_round.sh, output kept under profiles/.
Usage: python tests/sweep_div16.py [blocks-per-dtype (default 64 = everything)]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import minarrow_b200 as mnr  # noqa: E402
from oracle import oracle as orc  # noqa: E402  (checker only)


def main():
    nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    ctx = mnr.Context(0)
    dev = mnr.device_ops
    t0 = time.time()
    for dt in (np.uint16, np.int16):
        info = np.iinfo(dt)
        v = np.arange(info.min, info.max + 1).astype(dt)
        a = np.tile(v, 1024)
        ones = np.ones(a.size, dtype=bool)
        A = mnr.DeviceBuffer.upload(ctx, a)
        V = mnr.DeviceBitmask.upload(ctx, mnr.Bitmask.from_bools(ones))
        obits = orc.Bits.from_bools(ones)
        pairs = bad = 0
        for blk in range(0, 64, 64 // nblk):
            b = np.repeat(v[blk * 1024:(blk + 1) * 1024], 65536)
            B = mnr.DeviceBuffer.upload(ctx, b)
            for op in (orc.DIV, (orc.REM, orc.FLOORDIV)[blk & 1]):
                exp, em = orc.apply_int(a, b, op, obits)
                ob, om = dev.ew_binary(ctx, op, A, B, V, None, mnr.MaskMode.And)
                got = ob.download()
                if got.tobytes() != exp.tobytes() or not np.array_equal(om.download().bits, em.bits):
                    bad += 1
                    i = np.flatnonzero(got != exp)[:5]
                    print(f"MISMATCH {np.dtype(dt).name} block {blk} op {op}: {[(int(a[j]), int(b[j]), int(got[j]), int(exp[j])) for j in i]}")
                pairs += a.size
        print(f"{np.dtype(dt).name}: {pairs} (dividend, divisor, op) evaluations over {nblk} of 64 divisor blocks, {bad} mismatching launches")
    print(f"done in {time.time() - t0:.1f} s, {ctx.launch_count} kernel launches")


if __name__ == "__main__":
    main()
