"""One call of each StyleGAN upfirdn2d instance at the BASELINE configs[2] sizes (blur [8,32,1025,1025], upsample x2, downsample /2),
for ncu:  ncu --set full --clock-control none -k regex:upfirdn2d_stream -c 3 -o gpurun_out/ncu_upfirdn python tools/upfirdn_case.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200.op import upfirdn2d
k1 = torch.tensor([1., 3., 3., 1.])
k = (k1[:, None] * k1[None, :] / 64).cuda()
x = torch.randn((8, 32, 1025, 1025), device="cuda")
s = torch.randn((8, 32, 512, 512), device="cuda")
upfirdn2d(x, k, pad=(1, 1)); upfirdn2d(s, k * 4, up=2, pad=(2, 1)); upfirdn2d(x, k, down=2, pad=(1, 1))
torch.cuda.synchronize()
print("done")
