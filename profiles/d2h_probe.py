#!/usr/bin/env python
"""Raw pinned device->host bandwidth with 1 .. N GPUs of one box copying CONCURRENTLY, with and without binding each
rank (and therefore its pinned buffer, first-touched by the rank) to the CPU cores next to its GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/d2h_probe.py

This is the ceiling of pyitd_decompose_host at N GPUs: the call returns ~86 bytes for every 8-byte input sample, so
the aggregate D2H rate / 86 B is the most input samples per second any host-buffer API can decompose.
Rank 0 prints one JSON line.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as ex:
        return f"<{ex}>"


def main():
    import torch
    import torch.distributed as dist

    from pyitd_b200 import shard

    rank, world, local = shard.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("gloo")
    nbytes = int(os.environ.get("PROBE_BYTES", str(2 << 30)))
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    src.zero_()
    out = {"world": world, "bytes_per_rank": nbytes, "modes": {}}
    mask0 = os.sched_getaffinity(0)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def measure(active_ranks, dst, reps=3):
        """ranks < active_ranks copy concurrently; returns aggregate GB/s by the slowest rank's time"""
        dst.copy_(src, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        if rank < active_ranks:
            for _ in range(reps):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0 if rank < active_ranks else 0.0
        t = torch.tensor([dt], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return active_ranks * reps * nbytes / float(t.item()) / 1e9

    for mode in ("unbound", "bound"):
        cpus = None
        if mode == "bound":
            cpus = shard.bind_to_gpu_numa_node(local)
        dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()          # first-touched by this (possibly bound) thread
        res = {}
        n = 1
        while n <= world:
            res[str(n)] = round(measure(n, dst), 2)
            n *= 2
        info = {"aggregate_GBps_by_active_gpus": res,
                "cpus": (f"{cpus[0]}-{cpus[-1]} ({len(cpus)})" if cpus else None)}
        allinfo = [None] * world
        if world > 1:
            dist.all_gather_object(allinfo, info["cpus"])
            info["cpus_by_rank"] = allinfo
        out["modes"][mode] = info
        del dst
    os.sched_setaffinity(0, mask0)
    if rank == 0:
        out["topology"] = {"nvidia_smi_topo": sh("nvidia-smi topo -m | head -20"),
                           "lscpu_numa": sh("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"),
                           "gpu_numa_nodes": sh("for d in /sys/bus/pci/devices/*; do if [ \"$(cat $d/class 2>/dev/null)\" = 0x030200 ]; then "
                                                "echo $(basename $d) node=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist); fi; done"),
                           "affinity_at_start": f"{min(mask0)}-{max(mask0)} ({len(mask0)})"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
