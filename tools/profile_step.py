"""One VToonify-D step (B=4, 576x1024) inside a cudaProfilerStart/Stop window, for ncu:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import ops  # noqa: E402
from vtoonify_b200.vtoonify import VToonify  # noqa: E402
from vtoonify_b200.weights import det_inputs, det_state_dict  # noqa: E402

B = int(os.environ.get("VT_B", 4)); H = int(os.environ.get("VT_H", 576)); W = int(os.environ.get("VT_W", 1024))
backbone = os.environ.get("VT_BACKBONE", "dualstylegan")
dev = torch.device("cuda:0")
with torch.no_grad():
    m = VToonify(backbone=backbone).eval()
    m.load_state_dict(det_state_dict(m, seed=0)); m.to(dev)
    x, s = det_inputs(B, H, W)
    x, s = x.to(dev), s.to(dev)
    for _ in range(2):
        m(x, s, d_s=0.5).clamp_(-1, 1)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    m(x, s, d_s=0.5).clamp_(-1, 1)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done")
