"""oracle/vt_oracle.py — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this file, and only as the checker (or as the timed CPU baseline); nothing under ``vtoonify_b200/`` imports it.

It is a functional, state_dict-driven restatement in fp32 on CPU (torch CPU tensors; the dense contractions go through
``F.conv2d`` / ``F.conv_transpose2d`` / ``F.linear`` exactly where the reference calls them — SURVEY.md §8c: the conv
arithmetic of the reference itself lives in PyTorch/ATen).  Each function cites the reference lines it follows.

Parity pinning: the reference has no tests or golden vectors, so this oracle is pinned against outputs
of the reference's own code (``model/stylegan/op_cpu`` path, imported unmodified in the build container) stored under
``tests/golden/`` by ``tests/golden/make_golden.py``; ``tests/test_oracle_golden.py`` checks every fixture.
"""
import math
import re

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


# ------------------------------------------------------------------------------------------------
# a1: upfirdn2d   (model/stylegan/op_cpu/upfirdn2d.py:7-60; CUDA op upfirdn2d_kernel.cu:49-105)
# ------------------------------------------------------------------------------------------------
def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0)):
    """Direct tap-by-tap evaluation (no conv call): zero-stuff, pad / crop, correlate with the flipped kernel,
    decimate.  Integer index semantics identical to the reference."""
    up_x, up_y = (up, up) if isinstance(up, int) else up
    down_x, down_y = (down, down) if isinstance(down, int) else down
    if len(pad) == 2:
        pad = (pad[0], pad[1], pad[0], pad[1])
    px0, px1, py0, py1 = pad
    B, C, H, W = x.shape
    kh, kw = kernel.shape
    # zero-stuffing
    u = x.new_zeros((B, C, H * up_y, W * up_x))
    u[:, :, ::up_y, ::up_x] = x
    # positive pad, then negative pad == crop (op_cpu/upfirdn2d.py:33-41)
    u = F.pad(u, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    u = u[:, :, max(-py0, 0): u.shape[2] - max(-py1, 0), max(-px0, 0): u.shape[3] - max(-px1, 0)]
    full_h = H * up_y + py0 + py1 - kh + 1
    full_w = W * up_x + px0 + px1 - kw + 1
    out_h = (H * up_y + py0 + py1 - kh + down_y) // down_y
    out_w = (W * up_x + px0 + px1 - kw + down_x) // down_x
    kf = torch.flip(kernel, [0, 1])
    acc = x.new_zeros((B, C, full_h, full_w))
    for ky in range(kh):
        for kx in range(kw):
            acc += kf[ky, kx] * u[:, :, ky: ky + full_h, kx: kx + full_w]
    out = acc[:, :, ::down_y, ::down_x]
    assert out.shape[2] == out_h and out.shape[3] == out_w
    return out.contiguous()


# ------------------------------------------------------------------------------------------------
# a2: fused bias + leaky relu   (model/stylegan/op_cpu/fused_act.py:23-34)
# ------------------------------------------------------------------------------------------------
def fused_leaky_relu(x, bias=None, negative_slope=0.2, scale=SQRT2):
    if bias is not None:
        x = x + bias.view(1, -1, *([1] * (x.ndim - 2)))
    return F.leaky_relu(x, negative_slope) * scale


# ------------------------------------------------------------------------------------------------
# a8: EqualLinear   (model/stylegan/model.py:133-162)
# ------------------------------------------------------------------------------------------------
def equal_linear(x, weight, bias, lr_mul=1.0, activation=False):
    scale = (1 / math.sqrt(weight.shape[1])) * lr_mul
    if activation:
        return fused_leaky_relu(F.linear(x, weight * scale), bias * lr_mul)
    return F.linear(x, weight * scale, bias=bias * lr_mul)


def pixel_norm(x):  # model/stylegan/model.py:13-18
    return x * torch.rsqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)


def make_kernel(k):  # model/stylegan/model.py:21-29
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    return k / k.sum()


# ------------------------------------------------------------------------------------------------
# a3: ModulatedConv2d   (model/stylegan/model.py:227-306).  Restated in the "un-fused" algebra (model.py:230-257):
#     conv(x * s, scale*W) * d  with  d = rsqrt(sum((scale*W*s)^2) + 1e-8)  — one shared-weight conv per batch.
# ------------------------------------------------------------------------------------------------
def modulated_conv2d(x, style, sd, prefix, demodulate=True, upsample=False, downsample=False):
    W = sd[prefix + "weight"][0]                       # [Cout, Cin, k, k]
    Cout, Cin, k, _ = W.shape
    scale = 1 / math.sqrt(Cin * k * k)
    s = equal_linear(style, sd[prefix + "modulation.weight"], sd[prefix + "modulation.bias"])  # [B, Cin]
    w = scale * W
    B = x.shape[0]
    xs = x * s.view(B, Cin, 1, 1)
    if upsample:
        out = F.conv_transpose2d(xs, w.transpose(0, 1), padding=0, stride=2)
        kern = sd[prefix + "blur.kernel"]
        p = (kern.shape[0] - 2) - (k - 1)
        out = upfirdn2d(out, kern, pad=((p + 1) // 2 + 1, p // 2 + 1))
    elif downsample:
        kern = sd[prefix + "blur.kernel"]
        p = (kern.shape[0] - 2) + (k - 1)
        out = F.conv2d(upfirdn2d(xs, kern, pad=((p + 1) // 2, p // 2)), w, padding=0, stride=2)
    else:
        out = F.conv2d(xs, w, padding=k // 2)
    if demodulate:
        d = torch.rsqrt((w.unsqueeze(0) * s.view(B, 1, Cin, 1, 1)).square().sum((2, 3, 4)) + 1e-8)
        out = out * d.view(B, Cout, 1, 1)
    return out


def styled_conv(x, style, sd, prefix, noise=None, upsample=False):
    """model/stylegan/model.py:364-370: conv -> + noise_weight*noise -> FusedLeakyReLU."""
    out = modulated_conv2d(x, style, sd, prefix + "conv.", upsample=upsample)
    if noise is not None:
        out = out + sd[prefix + "noise.weight"] * noise
    return fused_leaky_relu(out, sd[prefix + "activate.bias"])


def to_rgb(x, style, sd, prefix, skip=None):
    """model/stylegan/model.py:383-392."""
    out = modulated_conv2d(x, style, sd, prefix + "conv.", demodulate=False) + sd[prefix + "bias"]
    if skip is not None:
        out = out + upfirdn2d(skip, sd[prefix + "upsample.kernel"], up=2, pad=(2, 1))
    return out


def generator_forward(sd, latent, noises, prefix="", log_size=None):
    """model/stylegan/model.py:566-590 with input_is_latent=True and explicit noise list."""
    B = latent.shape[0]
    out = sd[prefix + "input.input"].repeat(B, 1, 1, 1)
    out = styled_conv(out, latent[:, 0], sd, prefix + "conv1.", noises[0])
    skip = to_rgb(out, latent[:, 1], sd, prefix + "to_rgb1.")
    n_levels = len([k for k in sd if re.fullmatch(re.escape(prefix) + r"to_rgbs\.\d+\.bias", k)])
    i = 1
    for lv in range(n_levels):
        out = styled_conv(out, latent[:, i], sd, f"{prefix}convs.{2 * lv}.", noises[1 + 2 * lv], upsample=True)
        out = styled_conv(out, latent[:, i + 1], sd, f"{prefix}convs.{2 * lv + 1}.", noises[2 + 2 * lv])
        skip = to_rgb(out, latent[:, i + 2], sd, f"{prefix}to_rgbs.{lv}.", skip)
        i += 2
    return skip


# ------------------------------------------------------------------------------------------------
# a7: AdaIN / AdaResBlock   (model/dualstylegan.py:6-45)
# ------------------------------------------------------------------------------------------------
def adain(x, style, sd, prefix):
    gb = F.linear(style, sd[prefix + "style.weight"], sd[prefix + "style.bias"]).unsqueeze(2).unsqueeze(3)
    gamma, beta = gb.chunk(2, 1)
    return gamma * F.instance_norm(x, eps=1e-5) + beta


def conv_layer(x, sd, prefix, dilation=1):
    """ConvLayer = EqualConv2d(no bias) + FusedLeakyReLU(bias)  (model/stylegan/model.py:593-637)."""
    W = sd[prefix + "0.weight"]
    scale = 1 / math.sqrt(W.shape[1] * W.shape[2] ** 2)
    out = F.conv2d(x, W * scale, padding=W.shape[2] // 2 + dilation - 1, dilation=dilation)
    return fused_leaky_relu(out, sd[prefix + "1.bias"])


def ada_res_block(x, s, w, sd, prefix, dilation):
    if w == 0:
        return x
    out = conv_layer(adain(x, s, sd, prefix + "norm."), sd, prefix + "conv.", dilation)
    out = conv_layer(adain(out, s, sd, prefix + "norm2."), sd, prefix + "conv2.", dilation)
    return out * w + x


# ------------------------------------------------------------------------------------------------
# a6: VToonify.forward   (model/vtoonify.py:210-277)
# ------------------------------------------------------------------------------------------------
def vtoonify_forward(sd, x, style, d_s=None, backbone="dualstylegan", in_size=256, return_mask=False):
    D = backbone == "dualstylegan"
    gp = "generator.generator." if D else "generator."
    n_latent = 18
    # styles (:212-224)
    if style.ndim < 3:
        style = style.unsqueeze(1).repeat(1, n_latent, 1)
    nB, nL, nD = style.shape
    adastyles = style
    if D:
        t = pixel_norm(style.reshape(nB * nL, nD))
        for i in (1, 2):
            t = equal_linear(t, sd[f"generator.style.{i}.weight"], sd[f"generator.style.{i}.bias"], 0.01, True)
        resstyles = t.reshape(nB, nL, nD)
        adastyles = adastyles.clone()
        for i in range(7, n_latent):
            adastyles[:, i] = equal_linear(adastyles[:, i], sd[f"generator.res.{i}.weight"], sd[f"generator.res.{i}.bias"])

    def conv(t, key, stride=1, padding=1):
        return F.conv2d(t, sd[key + ".weight"], sd[key + ".bias"], stride=stride, padding=padding)

    # encoder (:160-176, :230-242)
    n_blocks = int(math.log2(in_size)) - 4          # 4 for in_size 256
    feat = x
    feats = []
    for bi in range(n_blocks):
        feat = F.leaky_relu(conv(feat, f"encoder.{bi}.0", stride=1 if bi == 0 else 2), 0.2)
        feat = F.leaky_relu(conv(feat, f"encoder.{bi}.2"), 0.2)
        feats.append(feat)
    feats = feats[::-1]
    dil = {1: 4, 2: 4, 3: 2, 4: 2, 5: 1, 6: 1}
    for ii in range(6):
        p = f"encoder.{n_blocks}.{ii}."
        out = F.leaky_relu(conv(feat, p + "conv"), 0.2)
        out = F.leaky_relu(conv(out, p + "conv2"), 0.2)
        feat = (out + feat) / math.sqrt(2)
        if D:
            feat = ada_res_block(feat, resstyles[:, ii + 1], d_s, sd, f"res.{ii + 1}.", dil[ii + 1])
    out = feat
    skip = conv(feat, f"encoder.{n_blocks + 1}", padding=0)

    # generator tail with fusion (:249-272)
    m_Es = []
    idx = 1
    for lv in range(5):
        if 2 ** (5 + ((idx - 1) // 2)) <= in_size:
            fi = (idx - 1) // 2
            f_E = feats[fi]
            if D:
                fp = f"fusion_out.{fi}."
                label = torch.zeros(out.shape[0], 1, device=out.device) + d_s
                label = F.leaky_relu(F.linear(label, sd[fp + "linear.0.weight"], sd[fp + "linear.0.bias"]), 0.2)
                label = F.leaky_relu(F.linear(label, sd[fp + "linear.2.weight"], sd[fp + "linear.2.bias"]), 0.2)
                cat = torch.cat([out, (out - f_E).abs()], dim=1)
                m_E = torch.tanh(F.relu(conv(adain(cat, label, sd, fp + "norm."), fp + "conv2")))
                out = conv(torch.cat([out, f_E * m_E], dim=1), fp + "conv")
                skip = conv(torch.cat([skip, f_E * m_E], dim=1), f"fusion_skip.{fi}")
                m_Es.append(m_E)
            else:
                out = conv(torch.cat([out, f_E], dim=1), f"fusion_out.{fi}")
                skip = conv(torch.cat([skip, f_E], dim=1), f"fusion_skip.{fi}")
        # noise is all-zero in the reference (:266-267), i.e. absent
        out = styled_conv(out, adastyles[:, idx + 6], sd, f"{gp}convs.{6 + 2 * lv}.", None, upsample=True)
        out = styled_conv(out, adastyles[:, idx + 7], sd, f"{gp}convs.{7 + 2 * lv}.", None)
        skip = to_rgb(out, adastyles[:, idx + 8], sd, f"{gp}to_rgbs.{3 + lv}.", skip)
        idx += 2
    if return_mask and D:
        return skip, m_Es
    return skip


# ------------------------------------------------------------------------------------------------
# a11: frame transforms   (style_transfer.py:57-60,160; util.py:190-192)
# ------------------------------------------------------------------------------------------------
def frame_u8_to_f32(frames_u8):
    """uint8 [B,H,W,3] -> fp32 [B,3,H,W]: ToTensor (v/255) then Normalize(0.5, 0.5)."""
    t = frames_u8.permute(0, 3, 1, 2).to(torch.float32).div(255)
    return (t - 0.5) / 0.5


def tensor2frame_u8(img, swap_rb=True):
    """clamp(-1,1) (style_transfer.py:177) then util.tensor2cv2: ((x+1)*127.5).astype(uint8), RGB->BGR."""
    t = ((img.clamp(-1, 1).permute(0, 2, 3, 1) + 1.0) * 127.5).to(torch.uint8)  # truncation like numpy astype
    return t.flip(-1) if swap_rb else t


# ------------------------------------------------------------------------------------------------
# f3: pre-filter + resize of high-resolution frames   (style_transfer.py:97, 124-130, 151-156)
# The arithmetic lives in OpenCV (cv2.sepFilter2D / cv2.resize on uint8 frames, opencv-python 4.x; the reference pins no
# version): restated here in integer numpy and pinned bit-exactly to cv2 outputs in tests/golden/frame_prep.npz
# (tests/golden/make_golden_frames.py).
# ------------------------------------------------------------------------------------------------
def sep_filter_1331_u8(frame):
    """cv2.sepFilter2D(frame, -1, k, k) with k = [[0.125],[0.375],[0.375],[0.125]] (style_transfer.py:97, 127) on a uint8 HxWxC
    frame: anchor = 2 (taps -2..+1), BORDER_REFLECT_101, exact value / 64 rounded half-to-even (cvRound), saturated."""
    import numpy as np
    f = np.asarray(frame)
    H, W = f.shape[:2]

    def refl(i, n):
        i = np.asarray(i)
        if n == 1:
            return np.zeros_like(i)
        i = np.where(i < 0, -i, i)
        return np.where(i >= n, 2 * n - 2 - i, i)
    w = (1, 3, 3, 1)
    acc = np.zeros(f.shape, dtype=np.int64)
    ys, xs = np.arange(H), np.arange(W)
    for a in range(4):
        rows = f[refl(ys + a - 2, H)].astype(np.int64)
        for b in range(4):
            acc += w[a] * w[b] * rows[:, refl(xs + b - 2, W)]
    v = (acc + 31 + ((acc >> 6) & 1)) >> 6
    return np.minimum(v, 255).astype(np.uint8)


def _resize_coefs(dst, src, clamp):
    import numpy as np
    scale = np.float64(src) / np.float64(dst)
    ofs, c0, c1 = [], [], []
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        sidx = int(np.floor(f))
        f = np.float32(f - np.float32(sidx))
        if clamp and sidx < 0:
            sidx, f = 0, np.float32(0)
        if clamp and sidx >= src - 1:
            sidx, f = src - 1, np.float32(0)
        ofs.append(sidx)
        c0.append(int(np.rint(np.float32(1.0 - f) * np.float32(2048))))
        c1.append(int(np.rint(f * np.float32(2048))))
    return np.array(ofs), np.array(c0, dtype=np.int64), np.array(c1, dtype=np.int64)


def resize_linear_u8(frame, w, h):
    """cv2.resize(frame, (w, h)) (INTER_LINEAR) on a uint8 HxWxC frame: 11-bit fixed-point coefficients, horizontal pass in
    int32, vertical pass (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2  (cv::resize, 8-bit linear path)."""
    import numpy as np
    f = np.asarray(frame)
    sh, sw = f.shape[:2]
    xo, xa0, xa1 = _resize_coefs(w, sw, True)
    yo, yb0, yb1 = _resize_coefs(h, sh, False)
    S = f.astype(np.int64)
    x1 = np.minimum(xo + 1, sw - 1)
    hor = S[:, xo] * xa0[None, :, None] + S[:, x1] * xa1[None, :, None]
    r0, r1 = hor[np.clip(yo, 0, sh - 1)], hor[np.clip(yo + 1, 0, sh - 1)]
    out = (((yb0[:, None, None] * (r0 >> 4)) >> 16) + ((yb1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def prefilter_resize_crop(frame, scale, size, crop):
    """style_transfer.py:124-130 / 151-156: blur once if scale <= 0.75, twice if scale <= 0.375, resize to (w, h), crop
    [top:bottom, left:right]."""
    if scale <= 0.75:
        frame = sep_filter_1331_u8(frame)
    if scale <= 0.375:
        frame = sep_filter_1331_u8(frame)
    top, bottom, left, right = crop
    return resize_linear_u8(frame, size[0], size[1])[top:bottom, left:right]


# ------------------------------------------------------------------------------------------------
# a10: pSp GradualStyleEncoder (IR-SE-50)   (model/encoder/encoders/psp_encoders.py:35-116, helpers.py:56-119)
# ------------------------------------------------------------------------------------------------
def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], False, 0.0, 1e-5)


def psp_forward(sd, x, n_styles=18):
    """GradualStyleEncoder(50, 'ir_se').forward in eval mode, driven by its state_dict."""
    h = F.prelu(_bn(F.conv2d(x, sd["input_layer.0.weight"], padding=1), sd, "input_layer.1."), sd["input_layer.2.weight"])
    feats = {}
    n_units = len({k.split(".")[1] for k in sd if k.startswith("body.")})
    for i in range(n_units):
        p = f"body.{i}."
        w2 = sd[p + "res_layer.3.weight"]
        has_sc_conv = (p + "shortcut_layer.0.weight") in sd
        # the first unit of every stage has stride 2 (helpers.py:25-26): unit 0 (64->64, MaxPool2d(1,2) shortcut) and the
        # units whose depth changes (conv shortcut)
        stride = 2 if (has_sc_conv or i == 0) else 1
        if has_sc_conv:
            sc = _bn(F.conv2d(h, sd[p + "shortcut_layer.0.weight"], stride=stride), sd, p + "shortcut_layer.1.")
        else:
            sc = F.max_pool2d(h, 1, stride)
        t = _bn(h, sd, p + "res_layer.0.")
        t = F.prelu(F.conv2d(t, sd[p + "res_layer.1.weight"], padding=1), sd[p + "res_layer.2.weight"])
        t = _bn(F.conv2d(t, w2, stride=stride, padding=1), sd, p + "res_layer.4.")
        g = F.adaptive_avg_pool2d(t, 1)
        g = torch.sigmoid(F.conv2d(F.relu(F.conv2d(g, sd[p + "res_layer.5.fc1.weight"])), sd[p + "res_layer.5.fc2.weight"]))
        h = t * g + sc
        feats[i] = h
    c1, c2, c3 = feats[6], feats[20], feats[23]

    def head(j, f):
        t, k = f, 0
        while f"styles.{j}.convs.{k}.weight" in sd:
            t = F.leaky_relu(F.conv2d(t, sd[f"styles.{j}.convs.{k}.weight"], sd[f"styles.{j}.convs.{k}.bias"], stride=2, padding=1), 0.01)
            k += 2
        return equal_linear(t.view(-1, 512), sd[f"styles.{j}.linear.weight"], sd[f"styles.{j}.linear.bias"])

    lat = [head(j, c3) for j in range(3)]
    p2 = F.interpolate(c3, size=c2.shape[2:], mode="bilinear", align_corners=True) + F.conv2d(c2, sd["latlayer1.weight"], sd["latlayer1.bias"])
    lat += [head(j, p2) for j in range(3, 7)]
    p1 = F.interpolate(p2, size=c1.shape[2:], mode="bilinear", align_corners=True) + F.conv2d(c1, sd["latlayer2.weight"], sd["latlayer2.bias"])
    lat += [head(j, p1) for j in range(7, n_styles)]
    return torch.stack(lat, dim=1)


# ------------------------------------------------------------------------------------------------
# f (next row): BiSeNet face parsing as the frame loop uses it   (model/bisenet/model.py:17-259, resnet.py:20-83,
# style_transfer.py:165-174)
# ------------------------------------------------------------------------------------------------
def _cbr(x, sd, p, stride=1, padding=1):
    """ConvBNReLU (model/bisenet/model.py:17-34)."""
    return F.relu(_bn(F.conv2d(x, sd[p + "conv.weight"], stride=stride, padding=padding), sd, p + "bn."))


def _basic_block(x, sd, p):
    """resnet.py:20-47: relu(shortcut + bn2(conv2(relu(bn1(conv1 x))))); a 1x1 stride-2 conv + BN shortcut exists in the
    state_dict exactly for the blocks that change channels and resolution."""
    has_ds = (p + "downsample.0.weight") in sd
    stride = 2 if has_ds else 1      # Resnet18 (resnet.py:58-61): the first block of layers 2-4 halves the resolution
    r = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"], stride=stride, padding=1), sd, p + "bn1."))
    r = _bn(F.conv2d(r, sd[p + "conv2.weight"], padding=1), sd, p + "bn2.")
    sc = x
    if has_ds:
        sc = _bn(F.conv2d(x, sd[p + "downsample.0.weight"], stride=stride), sd, p + "downsample.1.")
    return F.relu(sc + r)


def _arm(x, sd, p):
    """AttentionRefinementModule (model.py:63-84)."""
    feat = _cbr(x, sd, p + "conv.")
    att = F.avg_pool2d(feat, feat.shape[2:])
    att = torch.sigmoid(_bn(F.conv2d(att, sd[p + "conv_atten.weight"]), sd, p + "bn_atten."))
    return feat * att


def bisenet_forward(sd, x):
    """BiSeNet(n_classes).forward in eval mode -> feat_out only (the frame loop uses output [0]); x: [B,3,H,W]."""
    H, W = x.shape[2:]
    r = "cp.resnet."
    h = F.relu(_bn(F.conv2d(x, sd[r + "conv1.weight"], stride=2, padding=3), sd, r + "bn1."))
    h = F.max_pool2d(h, 3, 2, 1)
    feats = []
    for layer in (1, 2, 3, 4):
        for blk in (0, 1):
            h = _basic_block(h, sd, f"{r}layer{layer}.{blk}.")
        feats.append(h)
    feat8, feat16, feat32 = feats[1], feats[2], feats[3]
    avg = _cbr(F.avg_pool2d(feat32, feat32.shape[2:]), sd, "cp.conv_avg.", padding=0)
    f32 = _arm(feat32, sd, "cp.arm32.") + F.interpolate(avg, feat32.shape[2:], mode="nearest")
    f32_up = _cbr(F.interpolate(f32, feat16.shape[2:], mode="nearest"), sd, "cp.conv_head32.")
    f16 = _arm(feat16, sd, "cp.arm16.") + f32_up
    f16_up = _cbr(F.interpolate(f16, feat8.shape[2:], mode="nearest"), sd, "cp.conv_head16.")
    # FeatureFusionModule(feat_sp = feat8, feat_cp8 = f16_up)  (model.py:187-213, 237-239)
    feat = _cbr(torch.cat([feat8, f16_up], 1), sd, "ffm.convblk.", padding=0)
    att = F.avg_pool2d(feat, feat.shape[2:])
    att = torch.sigmoid(F.conv2d(F.relu(F.conv2d(att, sd["ffm.conv1.weight"])), sd["ffm.conv2.weight"]))
    fuse = feat * att + feat
    out = F.conv2d(_cbr(fuse, sd, "conv_out.conv."), sd["conv_out.conv_out.weight"])
    return F.interpolate(out, (H, W), mode="bilinear", align_corners=True)


def parsing_for_vtoonify(sd, frames):
    """style_transfer.py:171-174: parsing logits of the 2x-upsampled frame, taken back to frame size by nearest sampling;
    the caller concatenates ``cat(frames, x_p / 16)``.  frames: [B,3,H,W] in [-1,1]."""
    x2 = 2 * F.interpolate(frames, scale_factor=2, mode="bilinear", align_corners=False)
    return F.interpolate(bisenet_forward(sd, x2), scale_factor=0.5, recompute_scale_factor=False)
