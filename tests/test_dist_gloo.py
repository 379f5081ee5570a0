"""N>1 host logic on CPU: volumes are sharded round-robin with no data-path collective, and per-volume detections
are assembled with one padded all_gather (gloo here, NCCL on the GPUs)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, n_vol, q):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from roi3d_b200.parallel import gather_detections, shard_indices
    mine = shard_indices(n_vol)
    g = torch.Generator().manual_seed(1234)
    all_dets = [torch.rand((int(torch.randint(0, 9, (1,), generator=g)), 7), generator=g) for _ in range(n_vol)]
    all_labels = [torch.randint(0, 3, (d.shape[0],), generator=g) for d in all_dets]
    out = gather_detections([all_dets[i] for i in mine], [all_labels[i] for i in mine], mine)
    ok = sorted(out.keys()) == list(range(n_vol))
    for i in range(n_vol):
        ok = ok and torch.equal(out[i][0], all_dets[i]) and torch.equal(out[i][1], all_labels[i])
    q.put((rank, mine, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_vol", [5, 1])
def test_shard_and_gather_world2(n_vol):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200) + n_vol
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_vol, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = {r: mine for r, mine, _ in res}
    assert sorted(shards[0] + shards[1]) == list(range(n_vol))
    assert set(shards[0]).isdisjoint(shards[1])
    assert all(ok for _, _, ok in res)


def test_shard_indices_single_process():
    from roi3d_b200.parallel import gather_detections, shard_indices
    assert shard_indices(7, rank=1, world_size=3) == [1, 4]
    assert shard_indices(2, rank=3, world_size=4) == []
    d = [torch.zeros(2, 7)]
    out = gather_detections(d, [torch.zeros(2, dtype=torch.int64)], [4])
    assert list(out.keys()) == [4]
