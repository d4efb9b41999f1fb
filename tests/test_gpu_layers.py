"""GPU parity of the drop-in modules against the reference outputs in tests/golden (same deterministic weights)."""
import json

import numpy as np
import pytest
import torch

from tests.shapes import layer_state_dict

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)

# max-abs tolerance as a fraction of the reference output's rms, per precision mode
TOL = {"fp32": 1e-4, "bf16x3": 4e-4, "tf32": 1.5e-2}


def T(a):
    return torch.from_numpy(np.asarray(a))


def check(y, ref, prec, what):
    assert tuple(y.shape) == tuple(ref.shape), f"{what}: shape {tuple(y.shape)} vs {tuple(ref.shape)}"
    err = (y.cpu().double() - ref.double()).abs().max().item()
    rms = ref.pow(2).mean().sqrt().item()
    print(f"{what} [{prec}]: max err {err:.3e}  ref rms {rms:.3f}  err/rms {err / rms:.2e}")
    assert err <= TOL[prec] * max(rms, 1e-3), f"{what} [{prec}]: err {err:.3e} > {TOL[prec]} * rms {rms:.3f}"


@pytest.fixture(params=["fp32", "bf16x3", "tf32"])
def prec(request):
    from vtoonify_b200 import ops
    ops.set_precision(request.param)
    yield request.param
    ops.set_precision(ops.DEFAULT_PRECISION)


@pytest.mark.parametrize("name,args", [("sc_plain", (32, 64, False)), ("sc_up", (64, 32, True)), ("sc_plain512", (512, 512, False))])
def test_styled_conv(golden, prec, name, args):
    from vtoonify_b200.stylegan import StyledConv
    g = golden("layers")
    cin, cout, up = args
    m = StyledConv(cin, cout, 3, 512, upsample=up)
    m.load_state_dict(layer_state_dict("StyledConv", name), strict=True); m.cuda()
    x, s, nz = T(g[name + "_x"]).cuda(), T(g[name + "_s"]).cuda(), T(g[name + "_noise"]).cuda()
    check(m(x, s, noise=nz), T(g[name + "_y"]), prec, name)
    check(m.conv(x, s), T(g[name + "_yconv"]), prec, name + ".conv")
    # channels_last input takes the zero-copy path and must agree
    check(m(x.contiguous(memory_format=torch.channels_last), s, noise=nz), T(g[name + "_y"]), prec, name + " (channels_last in)")


def test_to_rgb(golden, prec):
    from vtoonify_b200.stylegan import ToRGB
    g = golden("layers")
    m = ToRGB(64, 512)
    m.load_state_dict(layer_state_dict("ToRGB", "rgb"), strict=True); m.cuda()
    x, s, skip = T(g["rgb_x"]).cuda(), T(g["rgb_s"]).cuda(), T(g["rgb_skip"]).cuda()
    # CUDA-core fp32 kernel; in tf32 mode only its *input* is TF32-rounded (activations are stored rounded)
    check(m(x, s, skip), T(g["rgb_y"]), prec, "ToRGB+skip")
    check(m(x, s), T(g["rgb_y_noskip"]), prec, "ToRGB")


def test_modconv_down(golden, prec):
    from vtoonify_b200.stylegan import ModulatedConv2d
    g = golden("layers")
    m = ModulatedConv2d(32, 32, 3, 512, downsample=True)
    m.load_state_dict(layer_state_dict("ModulatedConv2dDown", "mcd"), strict=True); m.cuda()
    check(m(T(g["mcd_x"]).cuda(), T(g["mcd_s"]).cuda()), T(g["mcd_y"]), prec, "ModulatedConv2d(down)")


def test_adares_and_fusion(golden, prec):
    from vtoonify_b200.dualstylegan import AdaResBlock
    from vtoonify_b200.vtoonify import Fusion
    g = golden("layers")
    m = AdaResBlock(64, dilation=2)
    m.load_state_dict(layer_state_dict("AdaResBlock", "ada"), strict=True); m.cuda()
    check(m(T(g["ada_x"]).cuda(), T(g["ada_s"]).cuda(), 0.6), T(g["ada_y"]), prec, "AdaResBlock(dil 2)")
    f = Fusion(32, 32, 32)
    f.load_state_dict(layer_state_dict("Fusion", "fus"), strict=True); f.cuda()
    fo, me = f(T(g["fus_fg"]).cuda(), T(g["fus_fe"]).cuda(), 0.5)
    check(me, T(g["fus_m"]), prec, "Fusion mask")
    check(fo, T(g["fus_out"]), prec, "Fusion out")


def test_generator32(golden, prec):
    from vtoonify_b200.stylegan import Generator
    g = golden("generator32")
    m = Generator(32, 512, 2)
    m.load_state_dict(layer_state_dict("Generator32", "gen"), strict=True); m.cuda()
    img, _ = m([T(g["latent"]).cuda()], input_is_latent=True, randomize_noise=False)
    check(img, T(g["y"]), prec, "Generator(32) from latent")
    img2, lat = m([T(g["z"]).cuda()], randomize_noise=False, return_latents=True)
    check(img2, T(g["y_from_z"]), prec, "Generator(32) from z")
    assert lat.shape == (2, 8, 512)
