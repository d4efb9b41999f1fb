"""f3: device pre-filter + resize of high-resolution frames (style_transfer.py:124-130, 151-156) — bit-exact against the oracle
(which is pinned to OpenCV's own outputs) and against the OpenCV goldens themselves."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def test_goldens_bit_exact(golden):
    from vtoonify_b200 import ops
    g = golden("frame_prep")
    for i in range(int(g["n_cases"])):
        f = torch.from_numpy(g[f"c{i}_frame"]).unsqueeze(0).cuda()
        scale, w, h, top, bottom, left, right = g[f"c{i}_params"]
        n_blur = int(scale <= 0.75) + int(scale <= 0.375)
        out = ops.frame_prefilter_resize(f, n_blur, (int(w), int(h)), (int(top), int(bottom), int(left), int(right)))
        assert np.array_equal(out[0].cpu().numpy(), g[f"c{i}_out"]), f"case {i} differs from OpenCV"


@pytest.mark.parametrize("H,W,scale", [(270, 480, 0.6), (216, 384, 0.3), (97, 131, 0.8), (60, 80, 1.25), (1080, 1920, 0.35)])
def test_random_frames_vs_oracle(H, W, scale):
    from oracle import vt_oracle as O
    from vtoonify_b200 import ops
    rng = np.random.default_rng(H * 7 + W)
    B = 2
    frames = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    w, h = max(2, int(W * scale)), max(2, int(H * scale))
    crop = (h // 10, h - h // 12, w // 9, w - w // 11)
    n_blur = int(scale <= 0.75) + int(scale <= 0.375)
    out = ops.frame_prefilter_resize(torch.from_numpy(frames).cuda(), n_blur, (w, h), crop).cpu().numpy()
    for b in range(B):
        ref = O.prefilter_resize_crop(frames[b], scale, (w, h), crop)
        assert out[b].shape == ref.shape
        assert np.array_equal(out[b], ref), f"frame {b}: {np.abs(out[b].astype(int) - ref.astype(int)).max()} levels off"


def test_pipeline_with_prefilter():
    """uint8 full-resolution frames in, blur + resize + crop + parsing + synthesis on the device"""
    from oracle import vt_oracle as O
    from vtoonify_b200.bisenet import BiSeNet
    from vtoonify_b200.frame_loop import FramePipeline
    from vtoonify_b200.vtoonify import VToonify
    from vtoonify_b200.weights import det_inputs, det_state_dict
    m = VToonify(backbone="toonify").eval()
    sd = det_state_dict(m, seed=0)
    m.load_state_dict(sd, strict=True)
    p = BiSeNet(19).eval()
    psd = det_state_dict(p, seed=21)
    p.load_state_dict(psd, strict=True)
    rng = np.random.default_rng(5)
    raw = [torch.from_numpy(rng.integers(0, 256, (1, 150, 200, 3), dtype=np.uint8)) for _ in range(2)]
    scale, size, crop = 0.4, (80, 60), (4, 52, 8, 72)             # -> 48 x 64 frames
    style = det_inputs(1, 48, 64, seed=2)[1]
    pipe = FramePipeline(m.cuda(), style, d_s=0.5, parsing_net=p.cuda(), prefilter=(1, size, crop))
    outs = list(pipe.run([r.pin_memory() for r in raw]))
    for r, o in zip(raw, outs):
        fr = torch.from_numpy(O.prefilter_resize_crop(r[0].numpy(), scale, size, crop)).unsqueeze(0)
        rgb = O.frame_u8_to_f32(fr)
        inputs = torch.cat([rgb, O.parsing_for_vtoonify(psd, rgb) / 16.0], dim=1)
        ref = O.tensor2frame_u8(O.vtoonify_forward(sd, inputs, style, 0.5, "toonify"))
        d = (o.to(torch.int16) - ref.to(torch.int16)).abs()
        assert int(d.max()) <= 1 and (d > 0).float().mean().item() <= 0.02
