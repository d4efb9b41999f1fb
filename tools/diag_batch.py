"""diagnostic: batch independence of the 576x1024 forward under the routing / caching options"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import ops
from vtoonify_b200.vtoonify import VToonify
from vtoonify_b200.weights import det_inputs, det_state_dict
torch.set_grad_enabled(False)
m = VToonify(backbone="dualstylegan").eval()
m.load_state_dict(det_state_dict(m, seed=0), strict=True)
m.cuda()
x, style = det_inputs(4, 576, 1024, seed=0)
x, style = x.cuda(), style.cuda()
for rs in (True, False):
    ops.set_option("rs_conv", rs)
    y = m(x, style, d_s=0.5)
    y1 = m(x[2:3], style[2:3], d_s=0.5)
    y2 = m(x[2:4], style[2:4], d_s=0.5)
    print(f"rs_conv={rs}: |y[2]-y1| {float((y[2:3]-y1).abs().max()):.3e}  |y[2:4]-y2| {float((y[2:4]-y2).abs().max()):.3e}  "
          f"|y1 - y2[0]| {float((y1 - y2[:1]).abs().max()):.3e}")
    if rs:
        yr = y.clone()
    else:
        print(f"rs vs tap-by-tap: {float((yr - y).abs().max()):.3e}")
