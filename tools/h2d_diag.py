#!/usr/bin/env python
"""Why does per-GPU host->device bandwidth drop when all GPUs of the box copy at once (55 -> 23 GB/s at 8)?
Run under torchrun.  Each rank pins 1 GiB (optionally bound to its GPU's NUMA node, --bind) and copies it to its GPU:
alone (ranks take turns), then in growing groups {0..k-1} at once.  Rank 0 prints one JSON line with the matrix plus what
the host looks like (NUMA nodes, each GPU's node)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from bench import numa_bind_to_gpu
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bind = "--bind" in sys.argv
    numa = numa_bind_to_gpu(torch, local) if bind else {"bound": False}
    if not bind:
        try:
            p = torch.cuda.get_device_properties(local)
            pci = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
            numa["gpu_node"] = int(open(f"/sys/bus/pci/devices/{pci}/numa_node").read())
            numa["nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
        except Exception as e:  # noqa: BLE001
            numa["note"] = repr(e)[:100]
    dist.init_process_group("nccl", device_id=dev)
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    d.copy_(h)
    torch.cuda.synchronize()

    def timed(active):
        dist.barrier()
        torch.cuda.synchronize()
        gbs = 0.0
        if active:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(4):
                d.copy_(h, non_blocking=True)
            b.record()
            b.synchronize()
            gbs = 4 * n / (a.elapsed_time(b) * 1e-3) / 1e9
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [round(float(x), 1) for x in out]

    res = {"solo": [timed(rank == r)[r] for r in range(world)]}
    k = 2
    while k <= world:
        res[f"first_{k}_together"] = timed(rank < k)[:k]
        k *= 2
    if world > 2:
        res["odd_ranks_together"] = [x for i, x in enumerate(timed(rank % 2 == 1)) if i % 2 == 1]
    nodes = [None] * world
    dist.all_gather_object(nodes, numa)
    if rank == 0:
        print(json.dumps({"h2d_GBps": res, "bind": bind, "numa_per_rank": nodes, "host_cpus": os.cpu_count(),
                          "affinity": len(os.sched_getaffinity(0))}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
