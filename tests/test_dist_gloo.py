"""CPU, world_size 2 (gloo): the rank-0 ingest/egress collectives of the frame loop (scatter inputs, gather frames)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, nb, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vtoonify_b200.frame_loop import gather_frames, scatter_batches, shard_indices
    shape = (2, 3, 4, 5)
    batches = [torch.full(shape, float(i)) + torch.arange(5.0) for i in range(nb)] if rank == 0 else None
    mine = scatter_batches(batches, nb, shape, torch.float32, "cpu")
    idx = shard_indices(nb, rank, world)
    ok = len(mine) == len(idx) and all(torch.equal(m, torch.full(shape, float(i)) + torch.arange(5.0)) for m, i in zip(mine, idx))
    # "synthesis": a per-frame function, then gather uint8 results in order on rank 0
    outs = [(m[:, :, :, :3] * 2).to(torch.uint8) for m in mine]
    ordered = gather_frames(outs, nb, (2, 3, 4, 3), torch.uint8, "cpu")
    if rank == 0:
        ok = ok and len(ordered) == nb and all(
            torch.equal(o, ((torch.full(shape, float(i)) + torch.arange(5.0))[:, :, :, :3] * 2).to(torch.uint8))
            for i, o in enumerate(ordered))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("nb", [5, 4])
def test_scatter_gather_world2(nb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nb, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}
