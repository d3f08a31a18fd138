"""Host-side logic of the multi-GPU path, run on CPU with the gloo backend at
world_size 2: shard boundaries, the off-data-path summary gather and the MAX timing reduction.  The
per-shard compute is played by the CPU oracle here (tests may use it); on the GPU box the same driver
code calls the CUDA path (tests/test_gpu_parity.py::test_sharded_driver_on_gpu)."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import itd_oracle as o
from pyitd_b200 import shard


def test_shard_ranges_cover_the_batch():
    for n in (0, 1, 7, 8, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def test_chunk_ranges():
    assert list(shard.chunk_ranges(3, 11, 4)) == [(3, 7), (7, 11)]
    assert list(shard.chunk_ranges(0, 9, 4)) == [(0, 4), (4, 8), (8, 9)]
    assert list(shard.chunk_ranges(5, 5, 4)) == []


def _oracle_decompose(x, max_iteration=11, min_extrema=2, **kw):
    rot, n_rows, counts, status, _ = o.c_decompose_batch(x.numpy(), max_iteration, min_extrema)
    return SimpleNamespace(rotations=torch.from_numpy(rot), n_rows=torch.from_numpy(n_rows),
                           knot_counts=torch.from_numpy(counts), status=torch.from_numpy(status),
                           stop_kind=torch.where(torch.from_numpy(n_rows) == max_iteration + 2, 2, 1).int())


def _batch(n_signals, n):
    rng = np.random.default_rng(11)
    x = rng.standard_normal((n_signals, n))
    if n_signals > 1:
        x[1] = np.cumsum(x[1])
    if n_signals > 2:
        x[2] = np.arange(n, dtype=np.float64)      # monotone: one zero row
    return torch.from_numpy(x)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_signals, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = _batch(n_signals, n)
        res, summ = shard.decompose_sharded(lambda a, b: full[a:b], n_signals, max_iteration=5,
                                            decompose_fn=_oracle_decompose)
        a, b = shard.shard_range(n_signals, rank, world)
        assert res.rotations.shape[0] == b - a
        slowest = shard.max_over_ranks(10.0 + rank)
        torch.save({"n_rows": summ.n_rows, "status": summ.status, "counts": summ.knot_counts,
                    "owner": summ.owner, "slowest": slowest, "local_rows": res.n_rows, "block": (a, b)},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_signals", [1, 5, 8])      # 1: rank 1 owns an EMPTY block and must still enter the gather
def test_two_rank_sharded_driver_matches_single_process(tmp_path, n_signals):
    n, world = 600, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_signals, n, str(tmp_path)), nprocs=world, join=True)
    single = _oracle_decompose(_batch(n_signals, n), max_iteration=5)
    outs = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    for r, out in enumerate(outs):
        # every rank sees the same global summary, in channel order, equal to the unsharded run
        assert torch.equal(out["n_rows"], single.n_rows)
        assert torch.equal(out["status"], single.status)
        assert torch.equal(out["counts"], single.knot_counts)
        a, b = out["block"]
        assert torch.equal(out["local_rows"], single.n_rows[a:b])
        assert torch.equal(out["owner"][a:b], torch.full((b - a,), r, dtype=torch.int32))
        assert out["slowest"] == 10.0 + (world - 1)            # MAX over ranks
    assert outs[0]["block"][1] == outs[1]["block"][0]


def test_single_process_paths_need_no_process_group():
    full = _batch(4, 300)
    res, summ = shard.decompose_sharded(lambda a, b: full[a:b], 4, max_iteration=3, decompose_fn=_oracle_decompose)
    assert torch.equal(summ.n_rows, res.n_rows) and int(summ.owner.max()) == 0
    assert shard.max_over_ranks(3.5) == 3.5


def test_numa_binding_helper_is_safe_without_a_gpu():
    """bind_to_gpu_numa_node only narrows the affinity mask to cores the platform lists next to the GPU; with no GPU
    (or no sysfs entry) it changes nothing and says so."""
    before = os.sched_getaffinity(0)
    got = shard.bind_to_gpu_numa_node(0)
    after = os.sched_getaffinity(0)
    if got is None:
        assert after == before
    else:
        assert set(got) == after and after <= before
        os.sched_setaffinity(0, before)
