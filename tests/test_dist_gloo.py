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


def _loop_worker(rank, world, port, nb, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vtoonify_b200.frame_loop import ShardedFrameLoop, shard_indices
    in_shape, out_shape = (2, 6, 5, 3), (2, 12, 10, 3)

    def batch(i):
        return (torch.arange(2 * 6 * 5 * 3, dtype=torch.float32).reshape(in_shape) + 100.0 * i) % 251

    def fn(x):                                  # stands for assemble + synthesize: any per-batch function of the inputs
        return x.repeat_interleave(2, 1).repeat_interleave(2, 2).to(torch.uint8)

    got, staged = {}, []
    loop = ShardedFrameLoop(fn, in_shape, torch.float32, out_shape, torch.uint8, "cpu")

    def stage(i):
        staged.append(i)
        return batch(i)

    def sink(i, buf, ready):
        ready()
        got[i] = buf.clone()

    mine = loop.run(nb, stage=stage if rank == 0 else None, sink=sink if rank == 0 else None)
    ok = mine == len(shard_indices(nb, rank, world))
    if rank == 0:
        ok = ok and staged == list(range(nb)) and sorted(got) == list(range(nb))
        ok = ok and all(torch.equal(got[i], fn(batch(i))) for i in range(nb))
        ok = ok and loop.scatter_bytes > 0 and loop.gather_bytes > 0
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("nb", [1, 4, 7])
def test_sharded_frame_loop_world2(nb):
    """rank-0 clip -> per-round scatter -> per-rank function -> gather, double-buffered, results in clip order"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_loop_worker, args=(r, 2, port, nb, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}
