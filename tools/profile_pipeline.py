"""One uint8-wire FramePipeline step (B=4, 576x1024: frames -> BiSeNet parsing -> VToonify-D -> uint8 BGR) inside a
cudaProfilerStart/Stop window, for ncu:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pipe.csv python tools/profile_pipeline.py
Without ncu it prints CUDA-event times of the stages (parsing / synthesis)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200.bisenet import BiSeNet  # noqa: E402
from vtoonify_b200.frame_loop import FramePipeline  # noqa: E402
from vtoonify_b200.vtoonify import VToonify  # noqa: E402
from vtoonify_b200.weights import det_inputs, det_state_dict  # noqa: E402

B = int(os.environ.get("VT_B", 4)); H = int(os.environ.get("VT_H", 576)); W = int(os.environ.get("VT_W", 1024))
dev = torch.device("cuda:0")
with torch.no_grad():
    m = VToonify(backbone="dualstylegan").eval()
    m.load_state_dict(det_state_dict(m, seed=0)); m.to(dev)
    pnet = BiSeNet(19).eval()
    pnet.load_state_dict(det_state_dict(pnet, seed=21), strict=True); pnet.to(dev)
    _, s = det_inputs(B, H, W)
    pipe = FramePipeline(m, s[:1], d_s=0.5, device=dev, parsing_net=pnet)
    frames = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3)).to(dev)
    for _ in range(2):
        pipe.process(frames)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.cudart().cudaProfilerStart()
    ev[0].record()
    x = pipe.assemble(frames)
    ev[1].record()
    y = pipe.synthesize(x)
    ev[2].record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(f"assemble (u8 -> f32 + BiSeNet parsing): {ev[0].elapsed_time(ev[1]):.3f} ms; synthesis + u8: {ev[1].elapsed_time(ev[2]):.3f} ms")
