"""a11 / f2: the frame loop (style_transfer.py:160-179) through ``FramePipeline``: uint8 frames out of the pipeline against
``oracle.tensor2frame_u8(oracle.vtoonify_forward(...))`` for every input form, and the buffer-ownership contract."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def setup():
    from vtoonify_b200.bisenet import BiSeNet
    from vtoonify_b200.vtoonify import VToonify
    from vtoonify_b200.weights import det_state_dict
    m = VToonify(backbone="dualstylegan").eval()
    sd = det_state_dict(m, seed=0)
    m.load_state_dict(sd, strict=True)
    p = BiSeNet(19).eval()
    psd = det_state_dict(p, seed=21)
    p.load_state_dict(psd, strict=True)
    return m.cuda(), sd, p.cuda(), psd


def _levels(out_u8, ref_u8, what):
    assert out_u8.dtype == torch.uint8 and tuple(out_u8.shape) == tuple(ref_u8.shape), (out_u8.shape, ref_u8.shape)
    d = (out_u8.to(torch.int16) - ref_u8.to(torch.int16)).abs()
    frac = (d > 0).float().mean().item()
    print(f"{what}: max level diff {int(d.max())}, pixels differing {100 * frac:.3f} %")
    assert int(d.max()) <= 1, f"{what}: a pixel is {int(d.max())} levels off"
    assert frac <= 0.02, f"{what}: {100 * frac:.2f} % of the values differ by one level (truncation ties only expected)"


def test_pipeline_fp32_inputs_vs_oracle(setup):
    from oracle import vt_oracle as O
    from vtoonify_b200.frame_loop import FramePipeline
    from vtoonify_b200.weights import det_inputs
    m, sd, _, _ = setup
    batches, style = [], None
    for i in range(3):
        x, s = det_inputs(2, 32, 40, seed=40 + i)
        batches.append(x.pin_memory())
        style = s if style is None else style
    pipe = FramePipeline(m, style[:1], d_s=0.5)
    outs = list(pipe.run(batches))                              # copy=True: the caller owns every result
    assert len(outs) == 3 and len({o.data_ptr() for o in outs}) == 3
    for i, (x, o) in enumerate(zip(batches, outs)):
        ref = O.tensor2frame_u8(O.vtoonify_forward(sd, x, style[:1].repeat(2, 1, 1), 0.5, "dualstylegan"))
        _levels(o, ref, f"FramePipeline batch {i} (fp32 [B,22,H,W] inputs)")
    assert pipe.h2d_bytes == 3 * batches[0].numel() * 4 and pipe.d2h_bytes == 3 * outs[0].numel()


def test_pipeline_borrowed_buffers(setup):
    """copy=False: results are views of a ring of pinned buffers; ring=3 keeps the previous result intact while the current
    one is consumed; ring=2 reuses the buffer of batch i for batch i+2."""
    from vtoonify_b200.frame_loop import FramePipeline
    from vtoonify_b200.weights import det_inputs
    m, _, _, _ = setup
    xs = [det_inputs(1, 32, 32, seed=60 + i)[0].pin_memory() for i in range(5)]
    style = det_inputs(1, 32, 32, seed=60)[1]
    owned = list(FramePipeline(m, style, d_s=0.5, copy=True).run(xs))
    for ring in (2, 3):
        pipe = FramePipeline(m, style, d_s=0.5, copy=False, ring=ring)
        ptrs, prev = [], None
        for i, o in enumerate(pipe.run(xs)):
            assert o.is_pinned()
            assert torch.equal(o, owned[i])
            if ring == 3 and prev is not None:
                assert torch.equal(prev, owned[i - 1]), "ring=3: the previous result must survive the next yield"
            ptrs.append(o.data_ptr())
            prev = o
        assert len(set(ptrs)) == ring and all(ptrs[i] == ptrs[i + ring] for i in range(len(ptrs) - ring))
    with pytest.raises(ValueError):
        FramePipeline(m, style, ring=1)


def test_pipeline_uint8_frames(setup):
    """(uint8 RGB frames, parsing maps) tuples and uint8 frames alone with the parsing computed on the device."""
    from oracle import vt_oracle as O
    from vtoonify_b200.frame_loop import FramePipeline
    from vtoonify_b200.weights import det_inputs
    m, sd, pnet, psd = setup
    g = torch.Generator().manual_seed(77)
    frames = [torch.randint(0, 256, (2, 48, 64, 3), generator=g, dtype=torch.uint8) for _ in range(3)]
    style = det_inputs(1, 48, 64, seed=3)[1]
    refs, tuples = [], []
    for fr in frames:
        rgb = O.frame_u8_to_f32(fr)
        x_p = O.parsing_for_vtoonify(psd, rgb)                                          # style_transfer.py:171-172
        inputs = torch.cat([rgb, x_p / 16.0], dim=1)                                    # :174
        refs.append(O.tensor2frame_u8(O.vtoonify_forward(sd, inputs, style.repeat(2, 1, 1), 0.5, "dualstylegan")))
        tuples.append((fr.pin_memory(), x_p.contiguous().pin_memory()))
    outs = list(FramePipeline(m, style, d_s=0.5).run(tuples))
    for i, (o, r) in enumerate(zip(outs, refs)):
        _levels(o, r, f"FramePipeline batch {i} (uint8 frames + host parsing maps)")
    pipe = FramePipeline(m, style, d_s=0.5, parsing_net=pnet)
    outs = list(pipe.run([f.pin_memory() for f in frames]))
    for i, (o, r) in enumerate(zip(outs, refs)):
        _levels(o, r, f"FramePipeline batch {i} (uint8 frames, BiSeNet parsing on the device)")
    assert pipe.h2d_bytes == sum(f.numel() for f in frames)
    with pytest.raises(ValueError):
        list(FramePipeline(m, style).run([frames[0]]))


def test_pipeline_cuda_graph(setup):
    """graph=True: one captured CUDA graph per input geometry, bit-identical frames, a handful of launches per batch"""
    from vtoonify_b200 import _lib
    from vtoonify_b200.frame_loop import FramePipeline
    from vtoonify_b200.weights import det_inputs
    m, _, pnet, _ = setup
    xs = [det_inputs(2, 32, 40, seed=80 + i)[0].pin_memory() for i in range(4)]
    style = det_inputs(1, 32, 40, seed=80)[1]
    eager = list(FramePipeline(m, style, d_s=0.5).run(xs))
    pipe = FramePipeline(m, style, d_s=0.5, graph=True)
    first = list(pipe.run(xs[:1]))                           # captures
    n0 = _lib.launch_count()
    outs = first + list(pipe.run(xs[1:]))
    replay_launches = _lib.launch_count() - n0
    for a, b in zip(outs, eager):
        assert torch.equal(a, b)
    assert replay_launches == 0, f"replays must not go through the host launch path ({replay_launches} launches counted)"
    assert len(pipe._graphs) == 1
    # a second geometry gets its own graph; uint8 frames + on-device parsing are capturable too
    g = torch.Generator().manual_seed(3)
    frames = [torch.randint(0, 256, (1, 48, 64, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(3)]
    eager_u8 = list(FramePipeline(m, style, d_s=0.5, parsing_net=pnet).run(frames))
    pipe_u8 = FramePipeline(m, style, d_s=0.5, parsing_net=pnet, graph=True)
    for a, b in zip(pipe_u8.run(frames), eager_u8):
        assert torch.equal(a, b)
