"""Channel sharding across the GPUs of one box.

Channels are independent -- there is no cross-channel term anywhere in ``ITD.py:79-121`` -- so N GPUs
run N shards of the channel axis with NO collective on the data path.  One process per GPU
(``torchrun``); ``torch.distributed`` is used only to (a) agree on the shard boundaries, (b) gather
the small per-signal integers (rows, stop kind, status, knot counts) when a caller wants the global
picture, and (c) reduce timings with MAX.  Works with the ``nccl`` backend on GPUs and with ``gloo``
on CPU tensors (the CPU test-suite runs it at world_size 2).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_range(n_signals: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block ``[start, stop)`` of the channel axis owned by ``rank``; block sizes differ
    by at most one and concatenate to ``range(n_signals)`` in rank order."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if n_signals < 0:
        raise ValueError("n_signals must be >= 0")
    base, extra = divmod(n_signals, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def chunk_ranges(start: int, stop: int, chunk: int):
    """Split a shard into chunks of at most ``chunk`` channels (config 5: 8 192-65 536 channels per
    GPU do not fit HBM with all their output rows, so a shard is walked in chunks that recycle one
    output buffer)."""
    if chunk < 1:
        raise ValueError("chunk must be >= 1")
    c0 = start
    while c0 < stop:
        c1 = min(c0 + chunk, stop)
        yield c0, c1
        c0 = c1


def env_rank_world() -> tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


@dataclass
class ShardSummary:
    """Per-signal integers of the WHOLE batch, in global channel order (small: a few ints per signal)."""
    n_rows: torch.Tensor          # [S] int32
    stop_kind: torch.Tensor       # [S] int32
    status: torch.Tensor          # [S] int32
    knot_counts: torch.Tensor     # [S, rows] int32
    owner: torch.Tensor           # [S] int32: rank that decomposed the signal


def _all_gather_ragged(t: torch.Tensor, sizes: list[int]) -> torch.Tensor:
    """all_gather of tensors whose first dimension differs per rank (pads to the maximum)."""
    world = dist.get_world_size()
    mx = max(sizes) if sizes else 0
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)


def gather_summary(n_signals: int, n_rows: torch.Tensor, stop_kind: torch.Tensor, status: torch.Tensor,
                   knot_counts: torch.Tensor) -> ShardSummary:
    """Concatenate every rank's per-signal integers in global channel order.  The only collective in
    the multi-GPU path, and it is off the data path (it moves 4 * (3 + rows) bytes per signal)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        owner = torch.zeros(n_signals, dtype=torch.int32, device=n_rows.device)
        return ShardSummary(n_rows, stop_kind, status, knot_counts, owner)
    world = dist.get_world_size()
    sizes = [shard_range(n_signals, r, world)[1] - shard_range(n_signals, r, world)[0] for r in range(world)]
    if n_rows.shape[0] != sizes[dist.get_rank()]:
        raise ValueError("local result does not match this rank's shard size")
    owner = torch.cat([torch.full((n,), r, dtype=torch.int32) for r, n in enumerate(sizes)]).to(n_rows.device)
    return ShardSummary(_all_gather_ragged(n_rows, sizes), _all_gather_ragged(stop_kind, sizes),
                        _all_gather_ragged(status, sizes), _all_gather_ragged(knot_counts, sizes), owner)


def gpu_numa_cpus(device_index: int) -> Optional[list[int]]:
    """CPU cores local to the GPU's PCIe root (``/sys/bus/pci/devices/<bdf>/local_cpulist``), or None when the
    platform does not say (single-socket box, container without sysfs)."""
    bdf = None
    try:
        pr = torch.cuda.get_device_properties(device_index)
        if all(hasattr(pr, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
    except Exception:
        bdf = None
    if not bdf:
        try:
            import subprocess
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device_index)],
                                 capture_output=True, text=True, timeout=20).stdout.strip()
            bdf = out.splitlines()[0].strip() if out else None
        except Exception:
            bdf = None
    if not bdf:
        return None
    bdf = bdf.lower()
    if len(bdf.split(":")[0]) == 8:            # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
        bdf = bdf[4:]
    try:
        text = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
    except OSError:
        return None
    cpus: list[int] = []
    for part in text.split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus or None


def bind_to_gpu_numa_node(device_index: int) -> Optional[list[int]]:
    """Pin this process to the CPU cores next to its GPU.  Pinned host buffers allocated afterwards are first-touched
    on that NUMA node, so the D2H stream of every rank stays on its own socket instead of all ranks sharing the node
    the launcher happened to start them on (SCALE_r01: eight D2H streams saturated one node's memory path).  Returns
    the CPU list used, or None when nothing was changed."""
    cpus = gpu_numa_cpus(device_index)
    if not cpus:
        return None
    try:
        allowed = os.sched_getaffinity(0)
        use = sorted(set(cpus) & set(allowed))
        if not use:
            return None
        os.sched_setaffinity(0, use)
        return use
    except (AttributeError, OSError):
        return None


def max_over_ranks(value: float, device=None) -> float:
    """MAX-reduce a timing over the ranks (every multi-GPU number is the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def decompose_sharded(make_shard: Callable[[int, int], torch.Tensor], n_signals: int, max_iteration: int = 11,
                      min_extrema: int = 2, decompose_fn: Optional[Callable] = None, gather: bool = True,
                      **kw):
    """Decompose this rank's block of a batch of ``n_signals`` channels.

    ``make_shard(start, stop)`` returns the ``[stop - start, N]`` tensor of this rank's channels
    (already on this rank's GPU: data never crosses ranks).  ``decompose_fn`` defaults to
    :func:`pyitd_b200.decompose`.  Returns ``(local_result, summary)`` where ``summary`` is the global
    :class:`ShardSummary` (or ``None`` with ``gather=False``)."""
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    start, stop = shard_range(n_signals, rank, world)
    if decompose_fn is None:
        from .itd import decompose as decompose_fn          # the CUDA path; raises without a GPU
    x = make_shard(start, stop)
    if x.shape[0] != stop - start:
        raise ValueError(f"make_shard returned {x.shape[0]} channels for the block [{start}, {stop})")
    if stop == start:
        # fewer signals than ranks: this rank owns nothing, but it must still enter the gather below (the other ranks
        # would block in all_gather forever if it raised on the empty batch instead)
        from .itd import ITDResult
        rows = max_iteration + 2
        i32 = dict(dtype=torch.int32, device=x.device)
        res = ITDResult(torch.empty((0, rows, x.shape[1]), dtype=x.dtype, device=x.device), torch.empty(0, **i32),
                        torch.empty((0, rows), **i32), torch.empty(0, **i32), torch.empty(0, **i32), torch.empty(0, **i32))
    else:
        res = decompose_fn(x, max_iteration=max_iteration, min_extrema=min_extrema, **kw)
    summary = None
    if gather:
        summary = gather_summary(n_signals, res.n_rows, res.stop_kind, res.status, res.knot_counts)
    return res, summary
