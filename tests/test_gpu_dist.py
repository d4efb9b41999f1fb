"""e: the rank-0 scatter / per-rank synthesis / gather loop of the frame path over NCCL (``ShardedFrameLoop``).
world = min(2, visible GPUs): on a one-GPU box the same code path runs as a single-rank NCCL group (degenerate collectives,
same streams / events / buffer hand-over), on a multi-GPU box as two ranks exchanging over NVLink."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, nb, q):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        torch.set_grad_enabled(False)
        from vtoonify_b200.frame_loop import FramePipeline, ShardedFrameLoop, shard_indices
        from vtoonify_b200.vtoonify import VToonify
        from vtoonify_b200.weights import det_inputs, det_state_dict
        m = VToonify(backbone="toonify").eval()
        m.load_state_dict(det_state_dict(m, seed=0), strict=True)
        m.to(dev)
        B, H, W = 1, 32, 32
        style = det_inputs(B, H, W, seed=0)[1]
        pipe = FramePipeline(m, style, d_s=0.5, device=dev, copy=True)
        loop = ShardedFrameLoop(lambda t: pipe.synthesize(pipe.assemble(t)), (B, 22, H, W), torch.float32,
                                (B, 4 * H, 4 * W, 3), torch.uint8, dev)
        hosts = [det_inputs(B, H, W, seed=100 + i)[0].pin_memory() for i in range(nb)] if rank == 0 else None
        results = [torch.empty((B, 4 * H, 4 * W, 3), dtype=torch.uint8).pin_memory() for _ in range(nb)] if rank == 0 else None
        d2h = torch.cuda.Stream(dev)

        def stage(i):                                     # runs on the loop's side stream: pinned host -> device
            return hosts[i].to(dev, non_blocking=True)

        def sink(i, buf, ready):
            with torch.cuda.stream(d2h):
                ready()
                results[i].copy_(buf, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(d2h)
            return ev

        mine = loop.run(nb, stage=stage if rank == 0 else None, sink=sink if rank == 0 else None)
        torch.cuda.synchronize()
        ok = mine == len(shard_indices(nb, rank, world))
        if rank == 0:
            single = list(pipe.run(hosts))                # the same batches through the one-GPU pipeline
            ok = ok and all(torch.equal(a, b) for a, b in zip(results, single))
            ok = ok and (world == 1 or (loop.scatter_bytes > 0 and loop.gather_bytes > 0))
        q.put((rank, bool(ok), ""))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))
        raise


@pytest.mark.parametrize("nb", [5])
def test_sharded_frame_loop_nccl(nb):
    world = min(2, torch.cuda.device_count())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nb, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert all(ok for _, ok, _ in res), "\n".join(msg for _, _, msg in res)
