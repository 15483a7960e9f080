#!/usr/bin/env python
"""bench.py — the hot path's headline measurement (contract: task statement ④).

Workload (BASELINE.json configs[1], benches/benchmark_parallel_simd.rs shape): a 1 000 000 000-row
IntegerArray<i64> with a 10 %-null validity bitmask -> null-aware sum + valid count (-> avg = sum / count).
One "step" = one pass of that reduction over the column: ONE launch of reduce_stats_kernel per GPU, which at
N > 1 is followed by an NCCL all-gather of the 32-byte per-GPU partial aggregates, combined in rank order.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # product arm (CUDA, sm_100a)
    python bench.py --impl reference [...]                         # the reference's CPU algorithm on host cores

Multi-GPU: weak scaling by default — the SuperArray has N shards of `--rows` rows, one per GPU (per-GPU work
fixed; `--scaling strong` splits one `--rows`-row column over the N GPUs instead).

Numbers on the JSON line:
  value     whole-job GB/s of ALGORITHMIC bytes (8 B value + 1/8 B validity per row), device-resident inputs,
            CUDA events on the launching stream, max over ranks.
  e2e       same metric through the host-slice C ABI call (mnr_stats_host): pinned HOST column + validity in,
            32-byte aggregate out, host<->device copies inside the timed region.
  roofline  the dominant kernel (reduce_stats_kernel) timed per launch with CUDA events inside the timed region.
  cpu_baseline  the oracle's OpenMP restatement of the reference's SIMD+rayon sum, on this box's host cores.
  secondary the other BASELINE configs timed alone (f64 masked add / scalar broadcast, bitmask ops, dense sums).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "GB/s, 1B-row IntegerArray<i64> null-aware sum/avg (10% nulls)"
UNIT = "GB/s"
BYTES_PER_ROW = 8.125            # sizeof(i64) + 1 validity bit
PUBLISHED_DENSE_GBS = 8.0e9 / 0.113874 / 1e9   # 113.874 ms for the dense 1e9-row i64 sum (Ultra 7 155H)
P_VALID = 0.9


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000_000)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N > 1 finish: fused = one kernel (reduce + P2P mailbox all-gather + combine); nccl = kernel + all-gather")
    ap.add_argument("--cpu-rows", type=int, default=1 << 28, help="rows of the bounded CPU sample")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-supertable", action="store_true", help="skip the 73 GiB configs[4] part of the secondary set")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel_key: str):
    """dram bytes per launch from the committed ncu --set full capture (profiles/traffic.json), else None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(kernel_key)
    return None


# --------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU algorithm (oracle port; the Rust crate cannot be built in this image)
# --------------------------------------------------------------------------------------------------------------
def host_sample(rows: int, seed: int):
    """Seeded i64 column in [-2^31, 2^31) + 0.9-valid bitmask, generated on the host (numpy)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    data = rng.integers(-2 ** 31, 2 ** 31, rows, dtype=np.int64)
    valid = rng.random(rows) < P_VALID
    bits = np.packbits(valid, bitorder="little")
    return data, bits, valid


def host_threads() -> int:
    """All host cores this process may run on.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1,
    and the oracle's OpenMP loops take the thread count as an explicit num_threads() argument."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def time_cpu(orc, data, bits, rows, threads, min_reps, max_reps, budget_s):
    """Best and mean seconds per pass of orc_par_masked_sum_i64 (par_chunks(1<<20) -> chunk sums -> combine)."""
    v = orc.Bits(bits, rows)
    orc.par_masked_sum_i64(data, v, threads)          # warm-up (page-in, thread pool)
    ts = []
    t_all = time.perf_counter()
    while len(ts) < max_reps and (len(ts) < min_reps or time.perf_counter() - t_all < budget_s):
        t0 = time.perf_counter()
        r = orc.par_masked_sum_i64(data, v, threads)
        ts.append(time.perf_counter() - t0)
    return ts, r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle import oracle as orc
    orc.build()
    threads = host_threads()
    rows = min(args.cpu_rows, args.rows)
    data, bits, valid = host_sample(rows, 1)
    exp_sum = int(data[valid].sum())
    v = orc.Bits(bits, rows)
    for _ in range(max(1, args.warmup)):
        orc.par_masked_sum_i64(data, v, threads)
    steps = max(1, args.steps)
    # bounded: stop early when the timed run would exceed ~2 minutes
    ts = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        s, c = orc.par_masked_sum_i64(data, v, threads)
        ts.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > 120:
            break
    assert s == exp_sum and c == int(valid.sum()), "reference arm: wrong sum"
    sec = sum(ts) / len(ts)
    gbs = rows * BYTES_PER_ROW / sec / 1e9
    sample = (f"{rows} rows of the {args.rows}-row workload per step (i64 in [-2^31,2^31), 0.9-valid bitmask), "
              f"{len(ts)} timed passes")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(ts), "warmup": max(1, args.warmup), "ms_per_step": round(sec * 1e3, 4), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "configs[1]: 1B-row i64 null-aware sum/avg, 10% nulls (bounded sample per step)",
                   "rows_per_step": rows, "bytes_per_row": BYTES_PER_ROW,
                   "algorithm": "par_chunks(1<<20) -> per-chunk masked sum -> combine "
                                "(benches/benchmark_parallel_simd.rs:81-89 restated in C + OpenMP; the Rust crate "
                                "cannot be built here: no cargo/rustc)"},
        "rows_per_s": round(rows / sec, 1),
        "cpu_baseline": {"value": round(gbs, 3), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": round(gbs, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------------------
# product arm
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Polls SM clock + clock-event reasons of one GPU (NVML) while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, torch_dev: int):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self.h = None
        try:
            import pynvml
            import torch
            self.nv = pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(torch_dev).uuid)
                if not uuid.startswith("GPU-"):
                    uuid = "GPU-" + uuid
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:  # noqa: BLE001
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_dev)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.h is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if self.h is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": getattr(self, "err", "no samples")}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_max": round(max(self.power), 1) if self.power else None}


def gen_column(torch, n, seed, dev):
    """Seeded synthetic shard on the device: i64 values in [-2^31, 2^31), Bernoulli(0.9) validity packed LSB-first
    (Arrow layout, slack bits zero).  Returns (data, bits, expected wrapping sum, expected valid count)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    data = torch.empty(n, dtype=torch.int64, device=dev)
    nbytes = (n + 7) // 8
    bits = torch.zeros(nbytes + 64, dtype=torch.uint8, device=dev)[:nbytes]
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.int32, device=dev)
    chunk = 1 << 26
    exp_sum = torch.zeros((), dtype=torch.int64, device=dev)
    exp_cnt = torch.zeros((), dtype=torch.int64, device=dev)
    for r0 in range(0, n, chunk):
        rows = min(chunk, n - r0)
        pad = (rows + 7) // 8 * 8
        d = torch.randint(-2 ** 31, 2 ** 31, (rows,), dtype=torch.int64, device=dev, generator=g)
        data[r0:r0 + rows] = d
        v = torch.rand(pad, device=dev, generator=g) < P_VALID
        v[rows:] = False
        bits[r0 // 8:r0 // 8 + pad // 8] = (v.view(-1, 8).to(torch.int32) * w).sum(dim=1).to(torch.uint8)
        exp_sum += (d * v[:rows]).sum()
        exp_cnt += v.sum()
        del d, v
    return data, bits, int(exp_sum.item()), int(exp_cnt.item())


def event_time_ms(torch, fn, iters, warmup=3):
    """Median / best CUDA-event ms of fn() on the current stream (each call timed alone)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def secondary(torch, mnr, ctx, dev, peak, data_buf, n_rows, supertable=True):
    """The other BASELINE configs, each kernel timed alone (median of 20, inputs >> L2)."""
    import numpy as np
    devops = mnr.device_ops
    out = {}

    def entry(name, nbytes, fn, iters=20):
        med, best = event_time_ms(torch, fn, iters)
        out[name] = {"GB/s": round(nbytes / med / 1e6, 1), "frac_of_measured_peak": round(nbytes / med / 1e6 / peak, 4),
                     "ms_median": round(med, 4), "ms_best": round(best, 4), "algorithmic_bytes": int(nbytes)}

    # dense i64 sum of the same column (benches/benchmark_parallel_simd.rs shape; published: 113.874 ms on Ultra 7 155H)
    part = torch.zeros(4, dtype=torch.int64, device=dev)
    entry("i64_dense_sum", n_rows * 8, lambda: devops.reduce_stats_async(ctx, data_buf, None, False, part.data_ptr()))
    out["i64_dense_sum"]["vs_published_70.3GBps_cpu"] = round(out["i64_dense_sum"]["GB/s"] / PUBLISHED_DENSE_GBS, 1)

    # configs[2]: f64 element-wise over two 256 Mi-row columns with validity bitmasks + scalar broadcast
    n = 1 << 28
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    x = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    y = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    mx = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device=dev, generator=g) | \
        torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device=dev, generator=g)
    my = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device=dev, generator=g) | \
        torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device=dev, generator=g)
    o = torch.empty(n, dtype=torch.float64, device=dev)
    om = torch.empty(n // 8, dtype=torch.uint8, device=dev)
    X = mnr.DeviceBuffer.wrap(ctx, np.float64, x.data_ptr(), n, x)
    Y = mnr.DeviceBuffer.wrap(ctx, np.float64, y.data_ptr(), n, y)
    O = mnr.DeviceBuffer.wrap(ctx, np.float64, o.data_ptr(), n, o)
    MX = mnr.DeviceBitmask.wrap(ctx, mx.data_ptr(), n, mx)
    MY = mnr.DeviceBitmask.wrap(ctx, my.data_ptr(), n, my)
    OM = mnr.DeviceBitmask.wrap(ctx, om.data_ptr(), n, om)
    A = mnr.ArithmeticOperator
    for name, op in (("add", A.Add), ("mul", A.Multiply), ("div", A.Divide)):
        entry(f"f64_masked_{name}_two_masks", n * 24.375,
              lambda op=op: devops.ew_binary_into(ctx, op, X, Y, MX, MY, mnr.MaskMode.And, O, OM))
    entry("f64_masked_add_one_mask", n * 24.25,
          lambda: devops.ew_binary_into(ctx, A.Add, X, Y, MX, None, mnr.MaskMode.And, O, OM))
    entry("f64_dense_add", n * 24.0, lambda: devops.ew_binary_into(ctx, A.Add, X, Y, None, None, mnr.MaskMode.And, O, None))
    entry("f64_masked_scalar_mul", n * 16.25, lambda: devops.ew_scalar_into(ctx, A.Multiply, X, 2.5, False, MX, O, OM))
    # spot parity of the last full-size result against torch (bit-exact: single IEEE multiply, nulls -> +0.0)
    vb = ((mx[: 1 << 17].to(torch.int32).view(-1, 1) >> torch.arange(8, device=dev, dtype=torch.int32)) & 1).bool().view(-1)
    exp = torch.where(vb, x[: 1 << 20] * 2.5, torch.zeros((), dtype=torch.float64, device=dev))
    assert torch.equal(exp.view(torch.int64), o[: 1 << 20].view(torch.int64)), "f64 scalar-broadcast mismatch vs torch"
    assert torch.equal(om, mx), "scalar-broadcast output validity must equal the input validity"
    del X, Y, O, MX, MY, OM, x, y, o, om

    # configs[3]: bitmask and / not over 4 Gi bits, popcount (null_count)
    nb = 1 << 32
    a = torch.randint(0, 256, (nb // 8,), dtype=torch.uint8, device=dev, generator=g)
    b = torch.randint(0, 256, (nb // 8,), dtype=torch.uint8, device=dev, generator=g)
    r = torch.empty(nb // 8, dtype=torch.uint8, device=dev)
    Ab = mnr.DeviceBitmask.wrap(ctx, a.data_ptr(), nb, a)
    Bb = mnr.DeviceBitmask.wrap(ctx, b.data_ptr(), nb, b)
    Rb = mnr.DeviceBitmask.wrap(ctx, r.data_ptr(), nb, r)
    entry("bitmask_and_4Gbit", nb * 3 / 8, lambda: devops.bits_binop_into(ctx, mnr.LogicalOperator.And, Ab, 0, Bb, 0, nb, Rb))
    assert torch.equal(r[: 1 << 24], a[: 1 << 24] & b[: 1 << 24])
    entry("bitmask_not_4Gbit", nb * 2 / 8, lambda: devops.bits_not_into(ctx, Ab, 0, nb, Rb))
    entry("bitmask_popcount_4Gbit", nb / 8, lambda: devops.bits_popcount(ctx, Ab, 0, nb), iters=10)
    cnt_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    entry("bitmask_popcount_4Gbit_async", nb / 8, lambda: devops.bits_popcount_async(ctx, Ab, 0, nb, cnt_dev.data_ptr()))
    assert int(cnt_dev.item()) == devops.bits_popcount(ctx, Ab, 0, nb)
    # BitmaskVT windows that do not start on a vector boundary (byte offsets 1 and 2): aligned loads + funnel shift
    r2 = torch.empty((nb - 64) // 8, dtype=torch.uint8, device=dev)
    Rb2 = mnr.DeviceBitmask.wrap(ctx, r2.data_ptr(), nb - 64, r2)
    entry("bitmask_and_4Gbit_unaligned_windows", (nb - 64) * 3 / 8,
          lambda: devops.bits_binop_into(ctx, mnr.LogicalOperator.And, Ab, 8, Bb, 16, nb - 64, Rb2))
    assert torch.equal(r2[: 1 << 24], a[1: (1 << 24) + 1] & b[2: (1 << 24) + 2])
    del Ab, Bb, Rb, Rb2, a, b, r, r2
    torch.cuda.empty_cache()

    # integer column / broadcast scalar (Array o Scalar route): multiply-high by the host-computed inverse, no divide
    n = 1 << 28
    xi = torch.randint(-2 ** 40, 2 ** 40, (n,), dtype=torch.int64, device=dev, generator=g)
    mi = torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device=dev, generator=g) | \
        torch.randint(0, 256, (n // 8,), dtype=torch.uint8, device=dev, generator=g)
    oi = torch.empty_like(xi)
    omi = torch.empty_like(mi)
    XI = mnr.DeviceBuffer.wrap(ctx, np.int64, xi.data_ptr(), n, xi)
    OI = mnr.DeviceBuffer.wrap(ctx, np.int64, oi.data_ptr(), n, oi)
    MI = mnr.DeviceBitmask.wrap(ctx, mi.data_ptr(), n, mi)
    OMI = mnr.DeviceBitmask.wrap(ctx, omi.data_ptr(), n, omi)
    entry("i64_masked_div_by_scalar_86400", n * 16.25, lambda: devops.ew_scalar_into(ctx, A.Divide, XI, 86400, False, MI, OI, OMI))
    vb = ((mi[: 1 << 17].to(torch.int32).view(-1, 1) >> torch.arange(8, device=dev, dtype=torch.int32)) & 1).bool().view(-1)
    exp = torch.where(vb, torch.div(xi[: 1 << 20], 86400, rounding_mode="trunc"), torch.zeros((), dtype=torch.int64, device=dev))
    assert torch.equal(exp, oi[: 1 << 20]), "i64 / scalar mismatch vs torch"
    del XI, OI, MI, OMI, xi, mi, oi, omi
    torch.cuda.empty_cache()

    # configs[0]: IntegerArray<i64> sum of 1 000 elements, averaged over 1 000 runs (hotloop_benchmark_simd shape).
    # 8 KB is launch-latency-bound on any GPU: report us/call honestly, plus the batched form (1 000 arrays, 1 launch).
    small = np.arange(1000, dtype=np.int64)
    S = mnr.DeviceBuffer.upload(ctx, small)
    assert devops.reduce_sum(ctx, S)[0] == 499500
    t0 = time.perf_counter()
    for _ in range(1000):
        devops.reduce_sum(ctx, S)
    dev_us = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    for _ in range(1000):
        mnr.kernels.reduce.stats(small, None, False, ctx)
    host_us = (time.perf_counter() - t0) * 1e3
    many = [S] * 1000
    harr = devops._handle_array(many)
    aggs = (mnr._lib.Agg * 1000)()
    ctx.lib.mnr_reduce_stats_batch(ctx.h, 1000, harr, None, 0, aggs)
    t0 = time.perf_counter()
    for _ in range(20):
        mnr.core.check(ctx.lib.mnr_reduce_stats_batch(ctx.h, 1000, harr, None, 0, aggs))     # one C ABI call, one launch
    batch_us = (time.perf_counter() - t0) / 20 * 1e6
    assert all(a.sum.i64 == 499500 and a.count == 1000 for a in aggs)
    out["c1_i64_1000_sum"] = {"device_resident_us_per_call": round(dev_us, 2), "host_slice_us_per_call": round(host_us, 2),
                              "batched_1000_arrays_us_per_array": round(batch_us / 1000, 3),
                              "published_cpu_ns": {"Vec64<i64>": 55, "IntegerArray direct": 88, "Array enum": 170},
                              "note": "roofline N/A (8 KB): one kernel launch + one stream sync per call; result 499500 checked"}

    if not supertable:
        return out
    # configs[4]: SuperTable of 64 batches x 16 Mi rows x {i32, i64, f32, f64}: per-column sum/min/max(+count) over all
    # chunks in ONE batched call (4 launches), then table * table and per-column scalar broadcast, chunk by chunk.
    nb, rows_b = 64, 1 << 24
    cols = [(np.int32, torch.int32), (np.int64, torch.int64), (np.float32, torch.float32), (np.float64, torch.float64)]
    tabs = []
    for which in range(2):
        t = []
        for npdt, tdt in cols:
            if tdt.is_floating_point:
                d = torch.randn(nb * rows_b, dtype=tdt, device=dev, generator=g)
            else:
                d = torch.randint(-1000, 1000, (nb * rows_b,), dtype=tdt, device=dev, generator=g)
            v = torch.randint(0, 256, (nb * rows_b // 8,), dtype=torch.uint8, device=dev, generator=g) | \
                torch.randint(0, 256, (nb * rows_b // 8,), dtype=torch.uint8, device=dev, generator=g)
            t.append((npdt, d, v))
        tabs.append(t)
    outs = [(torch.empty_like(d), torch.empty_like(v)) for _, d, v in tabs[0]]

    def chunks(t):
        bufs, vals = [], []
        for npdt, d, v in t:
            es = d.element_size()
            for k in range(nb):
                bufs.append(mnr.DeviceBuffer.wrap(ctx, npdt, d.data_ptr() + k * rows_b * es, rows_b, d))
                vals.append(mnr.DeviceBitmask.wrap(ctx, v.data_ptr() + k * rows_b // 8, rows_b, v))
        return bufs, vals
    lb, lv = chunks(tabs[0])
    rb, rv = chunks(tabs[1])
    ob = chunks([(npdt, o, om) for (npdt, _, _), (o, om) in zip(tabs[0], outs)])
    row_bytes = sum(d.element_size() for _, d, _ in tabs[0])          # 24 B/row
    nrows = nb * rows_b
    agg_dev = torch.zeros(len(lb), 4, dtype=torch.int64, device=dev)
    entry("supertable_64x16Mi_4col_sum_min_max_batched", nrows * (row_bytes + 4 / 8),
          lambda: devops.reduce_stats_batch_async(ctx, lb, lv, True, agg_dev.data_ptr()), iters=10)
    entry("supertable_64x16Mi_4col_sum_count_batched", nrows * (row_bytes + 4 / 8),
          lambda: devops.reduce_stats_batch_async(ctx, lb, lv, False, agg_dev.data_ptr()), iters=10)

    def one_by_one():
        for k in range(len(lb)):
            devops.reduce_stats_async(ctx, lb[k], lv[k], True, agg_dev[k].data_ptr())
    entry("supertable_64x16Mi_4col_sum_min_max_per_chunk_launches", nrows * (row_bytes + 4 / 8), one_by_one, iters=5)
    # spot check one column's total against torch
    torch.cuda.synchronize()
    a_host = agg_dev.cpu().numpy()
    d0, v0 = tabs[0][0][1], tabs[0][0][2]
    vb0 = ((v0[: rows_b // 8].to(torch.int32).view(-1, 1) >> torch.arange(8, device=dev, dtype=torch.int32)) & 1).bool().view(-1)
    assert int(a_host[0, 0]) == int((d0[:rows_b].to(torch.int64) * vb0).sum()) and int(a_host[0, 3]) == int(vb0.sum())

    def table_mul():
        for k in range(len(lb)):
            devops.ew_binary_into(ctx, A.Multiply, lb[k], rb[k], lv[k], rv[k], mnr.MaskMode.Or, ob[0][k], ob[1][k])
    entry("supertable_64x16Mi_4col_table_mul_table_per_chunk_launches", nrows * (3 * row_bytes + 4 * 3 / 8), table_mul, iters=5)
    plan = devops.EwBatchPlan(lb, rb, lv, rv, ob[0], ob[1])
    entry("supertable_64x16Mi_4col_table_mul_table_batched", nrows * (3 * row_bytes + 4 * 3 / 8),
          lambda: devops.ew_binary_batch_into(ctx, A.Multiply, lb, rb, lv, rv, mnr.MaskMode.Or, ob[0], ob[1], plan), iters=10)
    # spot check: i32 chunk 0 of the product against torch (wrapping multiply, OR-union validity as the SuperArray route)
    torch.cuda.synchronize()
    l0, r0 = tabs[0][0][1][:1 << 20], tabs[1][0][1][:1 << 20]
    vo = tabs[0][0][2][: 1 << 17] | tabs[1][0][2][: 1 << 17]
    vbo = ((vo.to(torch.int32).view(-1, 1) >> torch.arange(8, device=dev, dtype=torch.int32)) & 1).bool().view(-1)
    assert torch.equal(outs[0][0][:1 << 20], torch.where(vbo, l0 * r0, torch.zeros((), dtype=torch.int32, device=dev)))
    assert torch.equal(outs[0][1][: 1 << 17], vo)
    scal = [3, 3, 2.5, 2.5]
    half = 2 * nb       # integer columns: + 3 ; float columns: * 2.5   (one typed scalar per column)
    entry("supertable_64x16Mi_4col_scalar_broadcast_batched", nrows * (2 * row_bytes + 4 * 2 / 8),
          lambda: (devops.ew_scalar_batch_into(ctx, A.Add, lb[:half], [3] * half, False, lv[:half], ob[0][:half], ob[1][:half]),
                   devops.ew_scalar_batch_into(ctx, A.Multiply, lb[half:], [2.5] * half, False, lv[half:], ob[0][half:], ob[1][half:])),
          iters=10)
    del lb, lv, rb, rv, ob, tabs, outs
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus} (one rank per GPU)")
        raise SystemExit(f"WORLD_SIZE={world} but --gpus {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device. minarrow_b200 has no CPU fallback (use --impl reference for the CPU arm).")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import minarrow_b200 as mnr
    devops = mnr.device_ops
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = mnr.Context(local, stream=stream.cuda_stream)   # kernels launch on torch's current stream
    peak, peak_src = peaks()

    rows = args.rows if args.scaling == "weak" else (args.rows // world + (1 if rank < args.rows % world else 0))
    total_rows = rows * world if args.scaling == "weak" else args.rows
    data, bits, exp_sum, exp_cnt = gen_column(torch, rows, 1000 + rank, dev)
    buf = mnr.DeviceBuffer.wrap(ctx, np.int64, data.data_ptr(), rows, data)
    val = mnr.DeviceBitmask.wrap(ctx, bits.data_ptr(), rows, bits)
    partial = torch.zeros(4, dtype=torch.int64, device=dev)           # mnr_agg image: sum, min, max, count
    gathered = torch.zeros(world, 4, dtype=torch.int64, device=dev)

    fused = world > 1 and args.exchange == "fused"
    fx = None
    fused_note = None
    if fused:
        # The mailboxes need CUDA IPC between the ranks' processes.  Every rank must take the same path, so the outcome
        # is agreed with one all-reduce; if any rank cannot map its peers the run says so and uses the NCCL exchange.
        from minarrow_b200.sharded import FusedExchange
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            fx = FusedExchange(ctx)
        except Exception as e:  # noqa: BLE001
            fused_note = repr(e)[:200]
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0:
            fused, fx = False, None
            fused_note = fused_note or "a peer rank could not map the mailboxes"
    total = torch.zeros(4, dtype=torch.int64, device=dev)             # fused path: the combined aggregate on every rank

    def reduce_step():
        if fused:
            fx.reduce_stats_async(buf, val, False, total.data_ptr())
        else:
            devops.reduce_stats_async(ctx, buf, val, False, partial.data_ptr())

    def step():
        reduce_step()
        if world > 1 and not fused:
            dist.all_gather_into_tensor(gathered.view(-1), partial)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()

    K = args.steps
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local)
    launches0 = ctx.launch_count
    barrier()
    clocks.start()
    t0.record()
    for k in range(K):
        kev[k][0].record()
        reduce_step()
        kev[k][1].record()
        if world > 1 and not fused:
            dist.all_gather_into_tensor(gathered.view(-1), partial)
    t1.record()
    barrier()
    clocks.stop()
    launches = ctx.launch_count - launches0
    ms_total = t0.elapsed_time(t1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / K
    tmax = torch.tensor([ms_total, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms_max = float(tmax[0]), float(tmax[1])

    # result of the last step: per-GPU partials combined in rank order (integer sums wrap; order-free)
    parts = (total.view(1, 4) if fused else gathered if world > 1 else partial.view(1, 4)).cpu().numpy()
    tot_sum = int(np.sum(parts[:, 0].astype(np.uint64), dtype=np.uint64).astype(np.int64))
    tot_cnt = int(parts[:, 3].sum())
    exp = torch.tensor([exp_sum, exp_cnt], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(exp)            # int64 all-reduce wraps like the kernel does
    assert (tot_sum, tot_cnt) == (int(exp[0]), int(exp[1])), f"sum/count mismatch: {(tot_sum, tot_cnt)} vs {exp.tolist()}"
    avg = tot_sum / tot_cnt

    ms_step = ms_total / K
    value = total_rows * BYTES_PER_ROW / (ms_step * 1e-3) / 1e9
    achieved = rows * BYTES_PER_ROW / (kernel_ms_max * 1e-3) / 1e9

    # ---- e2e: host-slice C ABI (mnr_stats_host), pinned host buffers, copies inside the timed region ----------
    e2e = None
    host_data = host_bits = None
    if not args.no_e2e:
        # page-locking is bounded by what the box has: at most a quarter of MemAvailable over all ranks
        try:
            with open("/proc/meminfo") as f:
                avail = next(int(l.split()[1]) * 1024 for l in f if l.startswith("MemAvailable"))
        except Exception:  # noqa: BLE001
            avail = 64 << 30
        if rows * 8.125 * world > avail / 4:
            raise SystemExit(f"bench.py: {rows} rows x {world} ranks of pinned host memory do not fit a quarter of "
                             f"MemAvailable ({avail >> 30} GiB); rerun with --no-e2e or fewer --rows")
        host_data = torch.empty(rows, dtype=torch.int64, pin_memory=True)
        host_bits = torch.empty(bits.numel(), dtype=torch.uint8, pin_memory=True)
        host_data.copy_(data)
        host_bits.copy_(bits)
        torch.cuda.synchronize()
        agg = mnr._lib.Agg()
        hp, vp = C.c_void_p(host_data.data_ptr()), C.c_void_p(host_bits.data_ptr())
        hpart = torch.zeros(4, dtype=torch.int64, pin_memory=True)

        def e2e_step():
            mnr.core.check(ctx.lib.mnr_stats_host(ctx.h, 2, hp, rows, vp, 0, C.byref(agg)))   # 2 = MNR_I64
            if world > 1:
                hpart[0], hpart[3] = agg.sum.i64, agg.count
                partial.copy_(hpart, non_blocking=True)
                dist.all_gather_into_tensor(gathered.view(-1), partial)
                return gathered.cpu()
            return None

        e2e_step()
        # the PCIe roofline of this step: the same bytes as one plain pinned H2D copy
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        data.copy_(host_data, non_blocking=True)
        bits.copy_(host_bits, non_blocking=True)
        c1.record()
        c1.synchronize()
        h2d_gbs = (rows * 8 + bits.numel()) / (c0.elapsed_time(c1) * 1e-3) / 1e9
        barrier()
        l0 = ctx.launch_count
        w0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            g = e2e_step()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        e2e_launches = ctx.launch_count - l0
        barrier()
        if world > 1:
            assert int(g[:, 3].sum()) == tot_cnt
        else:
            assert (agg.sum.i64, agg.count) == (tot_sum, tot_cnt), "e2e sum mismatch"
        tsec = torch.tensor([(w1 - w0) / args.e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tsec, op=dist.ReduceOp.MAX)
        nchunks = (rows + (1 << 22) - 1) // (1 << 22)
        e2e = {"value": round(total_rows * BYTES_PER_ROW / float(tsec[0]) / 1e9, 3), "unit": UNIT,
               "h2d_bytes_per_step": int(rows * 8 + (rows + 7) // 8) * world, "d2h_bytes_per_step": 32 * nchunks * world,
               "ms_per_step": round(float(tsec[0]) * 1e3, 3), "steps": args.e2e_steps,
               "api": "mnr_stats_host (C ABI, pinned host column + validity -> 32-byte aggregate)",
               "timer": "host wall clock around synchronous calls, max over ranks",
               "gpu_launches": int(e2e_launches), "pcie_h2d_copy_GBps_per_gpu": round(h2d_gbs, 2),
               "frac_of_pcie_copy": round(rows * BYTES_PER_ROW / float(tsec[0]) / 1e9 / h2d_gbs, 4),
               "note": "bound by the host->device link: the same bytes as one plain pinned cudaMemcpy take 1/frac of this"}

    # ---- CPU baseline: bounded sample on this box's host cores (rank 0, N = 1 only) ------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        orc.build()
        srows = min(args.cpu_rows, rows) // 64 * 64
        if host_data is not None:
            sd, sb = host_data[:srows].numpy(), host_bits[:srows // 8].numpy()
        else:
            sd, sb = data[:srows].cpu().numpy(), bits[:srows // 8].cpu().numpy()
        threads = host_threads()
        ts, (cs, cc) = time_cpu(orc, sd, sb, srows, threads, 5, 200, 12.0)
        # the same sample through the CUDA path must agree bit for bit
        g_s, g_c = devops.reduce_sum(ctx, buf.slice(0, srows), mnr.DeviceBitmask.wrap(ctx, bits.data_ptr(), srows, bits))
        assert (cs, cc) == (g_s, g_c), f"oracle vs CUDA on the CPU sample: {(cs, cc)} vs {(g_s, g_c)}"
        mean = sum(ts) / len(ts)
        cpu = {"value": round(srows * BYTES_PER_ROW / mean / 1e9, 3), "unit": UNIT, "cores": threads, "kind": "port",
               "best": round(srows * BYTES_PER_ROW / min(ts) / 1e9, 3),
               "sample": f"first {srows} rows of the workload column, {len(ts)} passes (mean), OpenMP over 2^20-row chunks",
               "host_cpus": os.cpu_count(), "parity_with_cuda_on_sample": True}

    sec = None
    if rank == 0 and world == 1 and not args.no_secondary:
        sec = secondary(torch, mnr, ctx, dev, peak, buf, rows, not args.no_supertable)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms_step, 5), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": "configs[1]: 1B-row IntegerArray<i64> null-aware sum/avg, 10% nulls, "
                                   "SuperArray shards over GPUs + NCCL all-gather of 32-byte partials",
                       "exchange": ("fused kernel: reduce + P2P mailbox all-gather over NVLink + rank-order combine" if fused
                                    else ("NCCL all-gather of 32-byte partials" + (f" (fused exchange unavailable: {fused_note})" if fused_note else ""))
                                    if world > 1 else "none (1 GPU)"),
                       "rows_per_gpu": rows, "total_rows": total_rows, "bytes_per_row": BYTES_PER_ROW,
                       "l2": "inputs (8.1 GB per GPU) far larger than the 126 MB L2; no flush needed",
                       "values": "i64 uniform in [-2^31, 2^31), seeded per rank", "p_valid": P_VALID},
            "rows_per_s": round(total_rows / (ms_step * 1e-3), 1),
            "result": {"sum": tot_sum, "count": tot_cnt, "avg": avg},
            "frac_of_8TBps_nominal_per_gpu": round(value / world / 8000.0, 4),
            "roofline": {"bound": "hbm", "kernel": "reduce_stats_kernel<i64, V16, masked, no-minmax>",
                         "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": ncu_traffic("reduce_stats_kernel_i64_masked"), "peak_source": peak_src,
                         "kernel_ms": round(kernel_ms_max, 5), "algorithmic_bytes_per_launch": int(rows * BYTES_PER_ROW),
                         "timing": "CUDA events around each launch inside the timed region, mean of K, max over ranks"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks.summary(),
            "cpu_baseline": cpu, "secondary": sec,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
