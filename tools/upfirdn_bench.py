import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import ops
from vtoonify_b200.op import upfirdn2d, fused_leaky_relu
k = (torch.tensor([1., 3., 3., 1.])[:, None] * torch.tensor([1., 3., 3., 1.])[None, :] / 64 * 4).cuda()
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
x = torch.randn((8, 32, 1025, 1025), device="cuda")          # BASELINE configs[2]: blur [8,32,1025,1025]
s = torch.randn((8, 3, 512, 512), device="cuda")
from vtoonify_b200 import _lib
for sep in (2, 1, 0):
    _lib.load().vt_set_option(b"upfirdn_tiled", sep)
    ms = t(lambda: upfirdn2d(x, k, pad=(1, 1)))
    gb = (x.numel() + 8 * 32 * 1024 * 1024) * 4 / 1e9
    ms2 = t(lambda: upfirdn2d(s, k, up=2, pad=(2, 1)))
    gb2 = (s.numel() * 5) * 4 / 1e9
    kd = k / 4
    ms3 = t(lambda: upfirdn2d(x, kd, down=2, pad=(1, 1)))
    gb3 = (x.numel() * 1.25) * 4 / 1e9
    print(f"tiled={sep}: blur 4x4 [8,32,1025,1025] {ms:.3f} ms {gb / ms * 1e3:.0f} GB/s | upsample x2 [8,3,512,512] {ms2:.3f} ms {gb2 / ms2 * 1e3:.0f} GB/s | downsample /2 {ms3:.3f} ms {gb3 / ms3 * 1e3:.0f} GB/s")
y = torch.randn((8, 32, 1024, 1024), device="cuda"); b = torch.randn(32, device="cuda")
ms = t(lambda: fused_leaky_relu(y, b))
print(f"fused_leaky_relu [8,32,1024,1024] {ms:.3f} ms {y.numel() * 8 / 1e9 / ms * 1e3:.0f} GB/s")
