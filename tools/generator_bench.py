"""BASELINE configs[2]: 1024x1024 StyleGAN2 generator-only synthesis, batch 8 — per-layer timing of the modulated
convolutions (conv_tc launches) and the aggregate against the tensor roofline.   python tools/generator_bench.py [precision]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import _lib, ops
from vtoonify_b200.stylegan import Generator
from vtoonify_b200.weights import det_state_dict

prec = sys.argv[1] if len(sys.argv) > 1 else ops.DEFAULT_PRECISION
ops.set_precision(prec)
B = int(os.environ.get("VT_B", 8))
with torch.no_grad():
    g = Generator(1024, 512, 8).eval()
    g.load_state_dict(det_state_dict(g, seed=3), strict=True)
    g.cuda()
    latent = torch.randn((B, g.n_latent, 512), generator=torch.Generator().manual_seed(7)).cuda()
    run = lambda: g([latent], input_is_latent=True, randomize_noise=False)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    if os.environ.get("VT_PROFILE_WINDOW"):
        torch.cuda.cudart().cudaProfilerStart(); run(); torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
        sys.exit(0)
    prof = []
    ops.set_tc_profile(prof)
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    e0.record()
    for _ in range(steps):
        run()
    e1.record(); torch.cuda.synchronize()
    ops.set_tc_profile(None)
    ms = e0.elapsed_time(e1) / steps
    per = {}
    for a, b, f, nb, label, *_ in prof:
        d = per.setdefault(label, [0.0, 0.0, 0]); d[0] += a.elapsed_time(b); d[1] += f; d[2] += 1
    tc_ms = sum(v[0] for v in per.values()) / steps
    tc_fl = sum(v[1] for v in per.values()) / steps
    peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
    print(f"Generator(1024) B={B} [{prec}]: {ms:.2f} ms/step = {B / ms * 1e3:.1f} images/s; conv_tc {tc_ms:.2f} ms "
          f"({100 * tc_ms / ms:.0f}% of step), {tc_fl / tc_ms / 1e9:.0f} TF/s algorithmic; launches/step {(_lib.launch_count() - n0) / steps:.0f}")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][0]):
        print(f"{v[0] / steps:8.3f} ms  x{v[2] / steps:4.1f}  {v[1] / (v[0] * 1e-3) / 1e12:6.1f} TF/s  {k}")
