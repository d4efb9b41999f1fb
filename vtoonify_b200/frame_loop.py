"""Frame loop (the hot loop of style_transfer.py:99-183) for batches of frames, B200-style.

Reference behaviour per batch: stack frames -> ``inputs = cat(x, x_p/16)`` -> ``y = vtoonify(inputs, s_w.repeat(B,1,1),
d_s)`` -> ``clamp(-1,1)`` -> per frame ``tensor2cv2(y[k].cpu())`` (style_transfer.py:160-179).  There the H2D copy is
pageable and per-frame, the D2H is a blocking fp32 ``.cpu()`` per frame and nothing overlaps.

Here a :class:`FramePipeline` owns pinned staging buffers and two side streams: batch i+1's host->device copy and
batch i-1's device->host copy overlap batch i's synthesis; the clamp + uint8 + RGB->BGR conversion
(``util.tensor2cv2``) runs on the device so 4x fewer bytes cross PCIe.

Multi-GPU: frames are independent units, so ranks take round-robin shards of the frame batches with no
collective inside the forward.  :func:`scatter_batches` / :func:`gather_frames` implement the rank-0 ingest/egress of
the reference's single-decoder layout over ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
from typing import Iterable, Iterator, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import ops


def shard_indices(num_batches: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership of frame batches: batch i belongs to rank i % world (BASELINE.json configs[3])."""
    return list(range(rank, num_batches, world))


def merge_order(num_batches: int, world: int) -> List[tuple]:
    """(rank, local_index) of every global batch, in output order."""
    return [(i % world, i // world) for i in range(num_batches)]


class FramePipeline:
    """``run(batches)`` yields one uint8 ``[B, 4H, 4W, 3]`` (BGR, like cv2 frames) host tensor per input batch.

    ``batches``: iterable of pinned (or pageable) host fp32 tensors ``[B, 22, H, W]`` — the ``inputs`` of
    style_transfer.py:174 — or tuples ``(frames_u8 [B,H,W,3] RGB, parsing [B,19,H,W] fp32)``.
    """

    def __init__(self, model, style: torch.Tensor, d_s: Optional[float] = 0.5, device: Optional[torch.device] = None,
                 output: str = "u8"):
        self.model = model
        self.device = device or next(model.parameters()).device
        self.style = style.to(self.device)
        self.d_s = d_s
        self.output = output
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._host_out = {}

    def _upload(self, item):
        """Host -> device on the h2d stream; returns (x_dev, ready_event)."""
        with torch.cuda.stream(self.h2d):
            if isinstance(item, (tuple, list)):
                frames, parsing = item
                fr = frames.to(self.device, non_blocking=True)
                pr = parsing.to(self.device, non_blocking=True)
                B, H, W, _ = frames.shape
                x = torch.empty((B, 22, H, W), device=self.device, dtype=torch.float32)
                ops.frames_u8_to_f32(fr, out=x)           # channels 0..2 (ToTensor + Normalize on device)
                x[:, 3:] = pr / 16.0                      # style_transfer.py:174
                self.h2d_bytes += frames.numel() + parsing.numel() * 4
            else:
                x = item.to(self.device, non_blocking=True)
                self.h2d_bytes += item.numel() * 4
            ev = torch.cuda.Event()
            ev.record(self.h2d)
        return x, ev

    def _host_buffer(self, shape, dtype, slot):
        key = (tuple(shape), dtype, slot)
        if key not in self._host_out:
            self._host_out[key] = torch.empty(shape, dtype=dtype, pin_memory=True)
        return self._host_out[key]

    def run(self, batches: Iterable) -> Iterator[torch.Tensor]:
        main = torch.cuda.current_stream(self.device)
        it = iter(batches)
        nxt = next(it, None)
        pending = None  # (host_tensor, event) of the previous batch's download
        slot = 0
        up = self._upload(nxt) if nxt is not None else None
        with torch.no_grad():
            while up is not None:
                x, ev = up
                nxt = next(it, None)
                up = self._upload(nxt) if nxt is not None else None   # prefetch the next batch while this one computes
                main.wait_event(ev)
                x.record_stream(main)
                B = x.shape[0]
                y = self.model(x, self.style.expand(B, -1, -1) if self.style.shape[0] == 1 else self.style, d_s=self.d_s)
                out_dev = ops.f32_to_frames_u8(y, swap_rb=True) if self.output == "u8" else y.clamp(-1, 1)
                done = torch.cuda.Event()
                done.record(main)
                host = self._host_buffer(out_dev.shape, out_dev.dtype, slot)
                slot ^= 1
                with torch.cuda.stream(self.d2h):
                    self.d2h.wait_event(done)
                    host.copy_(out_dev, non_blocking=True)
                    out_dev.record_stream(self.d2h)
                    dl = torch.cuda.Event()
                    dl.record(self.d2h)
                self.d2h_bytes += out_dev.numel() * out_dev.element_size()
                if pending is not None:
                    pending[1].synchronize()
                    yield pending[0]
                pending = (host, dl)
            if pending is not None:
                pending[1].synchronize()
                yield pending[0]


# ----------------------------------------------------------------------------------------------
# rank-0 ingest / egress over torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests)
# ----------------------------------------------------------------------------------------------
def scatter_batches(batches: Optional[Sequence[torch.Tensor]], num_batches: int, example_shape, dtype, device,
                    src: int = 0, group=None) -> List[torch.Tensor]:
    """Rank ``src`` holds all ``num_batches`` input batches; every rank returns its round-robin shard.
    One ``dist.scatter`` per round of ``world`` batches (the only input-side collective of the path)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = []
    rounds = (num_batches + world - 1) // world
    for r in range(rounds):
        recv = torch.empty(example_shape, dtype=dtype, device=device)
        if rank == src:
            chunk = []
            for k in range(world):
                i = r * world + k
                chunk.append(batches[i].to(device) if i < num_batches else torch.zeros(example_shape, dtype=dtype, device=device))
            dist.scatter(recv, chunk, src=src, group=group)
        else:
            dist.scatter(recv, None, src=src, group=group)
        if r * world + rank < num_batches:
            mine.append(recv)
    return mine


def gather_frames(local_outputs: Sequence[torch.Tensor], num_batches: int, example_shape, dtype, device, dst: int = 0,
                  group=None) -> Optional[List[torch.Tensor]]:
    """Inverse of :func:`scatter_batches` for the uint8 output frames: returns the ordered list on ``dst``."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    rounds = (num_batches + world - 1) // world
    ordered = [] if rank == dst else None
    for r in range(rounds):
        have = r * world + rank < num_batches
        send = local_outputs[r].to(device) if have else torch.zeros(example_shape, dtype=dtype, device=device)
        if rank == dst:
            bufs = [torch.empty(example_shape, dtype=dtype, device=device) for _ in range(world)]
            dist.gather(send, bufs, dst=dst, group=group)
            for k in range(world):
                if r * world + k < num_batches:
                    ordered.append(bufs[k])
        else:
            dist.gather(send, None, dst=dst, group=group)
    return ordered
